VARIANTS="cur:- u2:RVQ_SCORE_UNROLL2 cur2:- u22:RVQ_SCORE_UNROLL2" bash scripts/run_variants.sh 2>&1 | grep -v "train variant"
