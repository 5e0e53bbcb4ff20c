timeout 600 python -m pytest tests/test_gpu_swin.py -x -q -m gpu > gpurun_out/t72.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t72.log; tail -2 gpurun_out/t72.log
for w in 1 0; do echo "WARP_WAIT=$w"; STF_B200_WARP_WAIT=$w STF_B200_PRECISION=fp32 timeout 300 python tools/bench_ops.py --only linear 2>&1 | tail -17; done
STF_B200_WARP_WAIT=1 STF_B200_PRECISION=tf32 timeout 300 python tools/bench_ops.py --only linear 2>&1 | tail -1
