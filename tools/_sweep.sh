for c in 0 74 0 74 100; do echo -n "slice ctas $c: "; STF_B200_SLICE_CTAS=$c timeout 120 python tools/split_sweep.py 2>&1 | tail -1; done
