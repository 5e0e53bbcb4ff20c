timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t66.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t66.log
timeout 900 python bench.py > gpurun_out/bench66.json 2> gpurun_out/bench66.err
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke66.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke66.log
tail -3 gpurun_out/t66.log; tail -2 gpurun_out/smoke66.log; python -c "
import json
d=json.load(open('gpurun_out/bench66.json')); print(round(d['value'],1), round(d['e2e']['value'],1), d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])"
