timeout 900 python -m pytest tests/test_gpu_train.py -x -q -m gpu -s -k wacnn > gpurun_out/t58.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t58.log
tail -12 gpurun_out/t58.log
