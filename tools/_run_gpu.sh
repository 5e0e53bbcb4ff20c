timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_configs.py -x -q -m gpu -s > gpurun_out/t47.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t47.log
timeout 600 python tools/bench_train.py > gpurun_out/train47_n1.json 2> gpurun_out/train47_n1.err
tail -3 gpurun_out/t47.log; grep -E "training step|linearity" gpurun_out/t47.log; cat gpurun_out/train47_n1.json | head -c 300
