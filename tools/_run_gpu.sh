timeout 900 python -m pytest tests/test_gpu_swin.py tests/test_gpu_codec.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/t53.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t53.log
timeout 900 python tools/bench_wacnn.py > gpurun_out/wacnn53.json 2> gpurun_out/wacnn53.err
tail -3 gpurun_out/t53.log; cat gpurun_out/wacnn53.json; tail -c 300 gpurun_out/wacnn53.err
