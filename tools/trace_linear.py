"""In-kernel timeline of the linear kernel's roles for CTA 0 (developer tool).
   STF_B200_DEBUG_SKIP=8 python tools/trace_linear.py [stage] [op]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["STF_B200_DEBUG_SKIP"] = str(int(os.environ.get("STF_B200_DEBUG_SKIP", "0")) | 8)
import torch  # noqa: E402

from stf_b200 import _C, ops  # noqa: E402

NAMES = ["issue:tile_start", "issue:all_kb_issued", "fin:first_landed", "fin:tile_published", "mma:accEmpty_ok",
         "mma:first_full_ok", "mma:last_commit", "epi:accFull_ok", "epi:phase1_done", "epi:stores_done", "epi:ptrs_ready", "mma:kb1_begin", "mma:kb1_fullA_ok", "mma:kb1_fullB_ok", "mma:kb1_mma_issued", "mma:kb1_committed"]


def main():
    st = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    op = sys.argv[2] if len(sys.argv) > 2 else "qkv"
    B, C = 8, 48 << st
    H, W = 256 >> st, 384 >> st
    T = B * H * W
    dev = "cuda"
    x = torch.randn(T, C, device=dev)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    if op == "qkv":
        lin = ops.PackedLinear(torch.randn(3 * C, C, device=dev) * C ** -0.5, torch.zeros(3 * C, device=dev), (g, b, 1e-5))
        fn = lambda: ops.linear(x, lin, rows=_C.ROWS_WINDOW, epilogue=_C.EPI_QKV, q_cols=C, q_scale=0.25, geom=(B, H, W, 4, 2))
    elif op == "fc1":
        lin = ops.PackedLinear(torch.randn(4 * C, C, device=dev) * C ** -0.5, torch.zeros(4 * C, device=dev), (g, b, 1e-5))
        fn = lambda: ops.linear(x, lin, epilogue=_C.EPI_GELU)
    else:
        h = torch.randn(T, 4 * C, device=dev)
        lin = ops.PackedLinear(torch.randn(C, 4 * C, device=dev) * (4 * C) ** -0.5, torch.zeros(C, device=dev))
        fn = lambda: ops.linear(h, lin, epilogue=_C.EPI_RESIDUAL, residual=x, x_is_tf32=os.environ.get("TRACE_LITE", "0") == "1")
    fn()
    fn()
    torch.cuda.synchronize()
    L = _C.lib()
    n_ev, n_t = 16, 24
    buf = (ctypes.c_longlong * (n_ev * n_t))()
    L.stf_debug_read_trace.restype = ctypes.c_int
    L.stf_debug_read_trace.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
    L.stf_debug_read_trace(buf, n_ev * n_t)
    t = [[buf[e * n_t + i] for i in range(n_t)] for e in range(len(NAMES))]
    t0 = t[0][0]
    print(f"stage {st} op {op}: cycles relative to the first tile's start (CTA 0), tiles 0..{n_t - 1}")
    for e, name in enumerate(NAMES):
        print(f"{name:22s}", " ".join(f"{(v - t0):7d}" for v in t[e][:12]))
    kb = (ctypes.c_longlong * (8 * 32))()
    L.stf_debug_read_ktrace.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
    L.stf_debug_read_ktrace(kb)
    kn = ["issue:copies_issued", "pub:landed_ok", "pub:arrived", "mma:full_ok", "mma:committed", "fwd:B_arrived", "load:B_issue", "mma:completed(exp)"]
    k0 = min(v for v in kb[:32] if v > 0) if any(v > 0 for v in kb[:32]) else 0
    print("per-k-block events of the third tile (cycles from its first issue):")
    for e, name in enumerate(kn):
        print(f"{name:22s}", " ".join(f"{(kb[e * 32 + i] - k0):6d}" for i in range(14)))
    per_tile = (t[8][n_t - 1] - t[8][4]) / (n_t - 1 - 4)
    print(f"steady-state cycles per tile (store_issued deltas): {per_tile:.0f}")


if __name__ == "__main__":
    main()
