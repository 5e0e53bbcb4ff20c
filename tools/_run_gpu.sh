set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench42_n2.json 2> gpurun_out/bench42_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/bench_train.py --gpus 2 > gpurun_out/train42_n2.json 2> gpurun_out/train42_n2.err
tail -c 300 gpurun_out/bench42_n2.err; head -c 300 gpurun_out/bench42_n2.json; echo; cat gpurun_out/train42_n2.json; tail -c 500 gpurun_out/train42_n2.err
