timeout 600 python -m pytest tests/test_gpu_swin.py tests/test_gpu_train.py -x -q -m gpu > gpurun_out/t67.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t67.log; tail -3 gpurun_out/t67.log
STF_B200_PRECISION=fp32 timeout 300 python tools/bench_ops.py --only linear 2>&1 | tail -17
STF_B200_KB_GROUP=2 STF_B200_PRECISION=fp32 timeout 300 python tools/bench_ops.py --only linear 2>&1 | tail -1
STF_B200_KB_GROUP=1 STF_B200_PRECISION=fp32 timeout 300 python tools/bench_ops.py --only linear 2>&1 | tail -1
STF_B200_PRECISION=tf32 timeout 300 python tools/bench_ops.py --only linear 2>&1 | tail -1
