set -x
timeout 900 python -m pytest tests/test_gpu_swin.py tests/test_gpu_entropy.py -x -q -m gpu > gpurun_out/t28.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t28.log
STF_B200_PRECISION=tf32 timeout 300 python tools/bench_ops.py > gpurun_out/ops28_tf32.log 2>&1
STF_B200_PRECISION=fp32 timeout 300 python tools/bench_ops.py --only linear > gpurun_out/ops28_fp32.log 2>&1
tail -3 gpurun_out/t28.log
