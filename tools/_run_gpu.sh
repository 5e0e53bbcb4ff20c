timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t55.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t55.log
timeout 600 python tools/bench_forward.py > gpurun_out/fwd55.json 2> gpurun_out/fwd55.err
tail -3 gpurun_out/t55.log; cat gpurun_out/fwd55.json | head -c 400; tail -c 300 gpurun_out/fwd55.err
