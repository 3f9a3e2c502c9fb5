for d in 0 3 4 5 6 7; do echo "== dense_ctas $d (overlap 0)"; OV=0 DENSE_CTAS=$d python scripts/profile_insitu.py 2>&1 | grep -E "k_dense_update|plic  "; done
for d in 8 10 12 14; do echo "== dense_ctas $d x128 thr (overlap 0)"; OV=0 DENSE_CTAS=$d DENSE_THREADS=128 python scripts/profile_insitu.py 2>&1 | grep -E "k_dense_update"; done
