set -x
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/t35.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t35.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench35.json 2> gpurun_out/bench35.err
tail -3 gpurun_out/t35.log; grep -E "fp32 strict|symbol flips" gpurun_out/t35.log; head -c 300 gpurun_out/bench35.json
