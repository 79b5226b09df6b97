VARIANTS="cur:- l1:RVQ_L1_PREFETCH cur2:- l12:RVQ_L1_PREFETCH" bash scripts/run_variants.sh 2>&1 | grep -v "train variant"
