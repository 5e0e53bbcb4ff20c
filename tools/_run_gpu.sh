set -x
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -m gpu -s > gpurun_out/t30.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t30.log
tail -40 gpurun_out/t30.log
