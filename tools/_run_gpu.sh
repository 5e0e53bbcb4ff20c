for g in 1 2 4 6 8 12; do echo "FIN_GROUP max $g"; STF_B200_FIN_GROUP=$g STF_B200_PRECISION=fp32 timeout 300 python tools/bench_ops.py --only linear 2>&1 | grep -E "fc2|proj|sum"; done
