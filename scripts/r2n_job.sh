VARIANTS="prev:-@rvq_tc_prev.cu cur:- nocnt:RVQ_NO_WIDE_COUNTERS owner:RVQ_WIDE_OWNER both:RVQ_NO_WIDE_COUNTERS,RVQ_WIDE_OWNER" bash scripts/run_variants.sh
