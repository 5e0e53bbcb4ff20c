timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_swin.py -x -q -m gpu > gpurun_out/t59.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t59.log
tail -8 gpurun_out/t59.log
