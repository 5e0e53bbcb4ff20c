timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29554 tools/bench_train.py --gpus 2 > gpurun_out/train50_n2_ddp.json 2> gpurun_out/train50_n2_ddp.err
grep "^{" gpurun_out/train50_n2_ddp.json | head -c 330; tail -c 600 gpurun_out/train50_n2_ddp.err
