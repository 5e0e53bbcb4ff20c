timeout 600 python -m pytest tests/test_gpu_swin.py -x -q -m gpu > gpurun_out/t60.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t60.log
tail -5 gpurun_out/t60.log
STF_B200_PRECISION=fp32 timeout 300 python tools/bench_ops.py --only linear --stages 2,3 > gpurun_out/ops60_fp32_pair.log 2>&1
STF_B200_PAIR=0 STF_B200_PRECISION=fp32 timeout 300 python tools/bench_ops.py --only linear --stages 2,3 > gpurun_out/ops60_fp32_nopair.log 2>&1
STF_B200_PRECISION=tf32 timeout 300 python tools/bench_ops.py --only linear --stages 2,3 > gpurun_out/ops60_tf32_pair.log 2>&1
paste <(tail -9 gpurun_out/ops60_fp32_pair.log | cut -c1-75) <(tail -9 gpurun_out/ops60_fp32_nopair.log | cut -c50-75) <(tail -9 gpurun_out/ops60_tf32_pair.log | cut -c50-75)
