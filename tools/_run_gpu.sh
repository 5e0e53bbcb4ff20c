set -x
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t33.log
STF_B200_CUDNN_BENCHMARK=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench33_cudnnbench.json 2> gpurun_out/bench33_cudnnbench.err
tail -3 gpurun_out/t33.log; cat gpurun_out/bench33_cudnnbench.json | head -c 400
