set -x
timeout 900 python -m pytest tests/test_gpu_swin.py -x -q -m gpu > gpurun_out/t39.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t39.log
STF_B200_PRECISION=fp32 timeout 300 python tools/bench_ops.py --only linear > gpurun_out/ops39_fp32.log 2>&1
tail -2 gpurun_out/t39.log; tail -18 gpurun_out/ops39_fp32.log
