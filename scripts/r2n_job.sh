VARIANTS="cur:- pref:RVQ_SCORE_PREFETCH cur2:- pref2:RVQ_SCORE_PREFETCH" bash scripts/run_variants.sh
