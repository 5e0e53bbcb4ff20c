timeout 300 python tools/split_times.py --batch 32
STF_B200_PIPELINE_MIN_BATCH=1000 timeout 300 python tools/split_times.py --batch 32
timeout 300 python tools/split_times.py --batch 16
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread"
