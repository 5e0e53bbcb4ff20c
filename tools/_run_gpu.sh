nproc; free -g | head -2 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench56_n8.json 2> gpurun_out/bench56_n8.err
grep "^{" gpurun_out/bench56_n8.json | head -c 500; tail -c 400 gpurun_out/bench56_n8.err
