timeout 600 python tools/gap_trace.py --batch 32 2>&1 | tail -32
