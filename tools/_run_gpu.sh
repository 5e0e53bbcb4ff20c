for skip in 8 9 10 11 12 15; do
  echo "== DEBUG_SKIP=$skip (1: no A loads, 2: no weight loads, 4: no stores; 8: traced specialised instance)"
  STF_B200_DEBUG_SKIP=$skip STF_B200_PRECISION=fp32 timeout 300 python tools/bench_ops.py --only linear --stages 0,2 2>&1 | grep -E "qkv|fc1|fc2"
done
