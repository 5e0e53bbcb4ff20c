set -x
timeout 900 python -m pytest tests/test_gpu_swin.py tests/test_gpu_codec.py -x -q -m gpu -s > gpurun_out/t24.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t24.log
STF_B200_PRECISION=fp32 timeout 300 python tools/bench_ops.py > gpurun_out/ops24_fp32.log 2>&1
STF_B200_PRECISION=tf32 timeout 300 python tools/bench_ops.py --only linear > gpurun_out/ops24_tf32.log 2>&1
timeout 600 python bench.py --precision fp32 --no-cpu-baseline > gpurun_out/bench24_fp32.json 2> gpurun_out/bench24_fp32.err
timeout 600 python bench.py --precision tf32 --no-cpu-baseline > gpurun_out/bench24_tf32.json 2> gpurun_out/bench24_tf32.err
tail -3 gpurun_out/t24.log
