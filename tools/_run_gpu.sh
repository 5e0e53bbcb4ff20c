for kb in 72 48 36 24 18; do echo "tile_kb $kb"; STF_B200_ATTN_TILE_KB=$kb timeout 300 python tools/bench_ops.py --only attention 2>&1 | grep attention; done
timeout 600 python -m pytest tests/test_gpu_swin.py -x -q -m gpu -k pair 2>&1 | tail -2
