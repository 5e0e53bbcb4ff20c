"""GPU bring-up probe (developer tool, not a test): runs each check in its own process so that a
trapped kernel cannot poison the others.   python tools/bringup.py [check ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHECKS = {}


def check(f):
    CHECKS[f.__name__] = f
    return f


def _rna(t):
    import torch
    i = t.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF          # round-to-nearest, ties away (magnitude), like cvt.rna.tf32.f32
    return i.view(torch.float32)


def _linear_case(M, N, K, flags, seed=0):
    import torch
    from stf_b200 import ops
    torch.manual_seed(seed)
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    lin = ops.PackedLinear(w, b)
    y = ops.linear(x, lin, debug_flags=flags)
    torch.cuda.synchronize()
    ref = (_rna(x).double() @ _rna(w).double().t() + b.double()).float()
    err = (y - ref).abs().max().item()
    ref32 = x @ w.t() + b
    err32 = (y - ref32).abs().max().item()
    return err, err32, ref.abs().max().item()


@check
def linear_flag0():
    for (M, N, K) in ((128, 16, 16), (128, 48, 48), (256, 144, 48), (1000, 192, 192), (4096, 1152, 384)):
        print("flag0", (M, N, K), "err_vs_tf32ref %.3e err_vs_fp32 %.3e refmax %.2f" % _linear_case(M, N, K, 0))


@check
def linear_flag1():
    for (M, N, K) in ((128, 16, 16), (128, 48, 48), (256, 144, 48)):
        print("flag1", (M, N, K), "err_vs_tf32ref %.3e err_vs_fp32 %.3e refmax %.2f" % _linear_case(M, N, K, 1))


@check
def entropy_smoke():
    import numpy as np
    import torch
    from stf_b200 import ops
    g = np.load(os.path.join(ROOT, "tests/golden/entropy_ops.npz"))
    from oracle import entropy as OE
    table = OE.scale_table()
    idx = ops.build_indexes(torch.from_numpy(g["bi_scales"]).cuda(), table)
    print("build_indexes mismatches:", int((idx.cpu().numpy() != g["bi_indexes"]).sum()), "of", idx.numel())


@check
def attention_smoke():
    import numpy as np
    import torch
    from stf_b200 import layers
    from stf_b200.synth import synthetic_state_dict
    g = np.load(os.path.join(ROOT, "tests/golden/swin_ops.npz"))
    wa = layers.WindowAttention(48, (4, 4), 3).eval()
    spec = {k: (tuple(v.shape), v.dtype) for k, v in wa.state_dict().items()}
    wa.load_state_dict(synthetic_state_dict(spec, 14), strict=False)
    wa = wa.cuda()
    with torch.no_grad():
        y = wa(torch.from_numpy(g["wa_x"]).cuda())
        ym = wa(torch.from_numpy(g["wa_x"]).cuda(), torch.from_numpy(g["mask_8_12_4_2"]).cuda())
    print("wa nomask maxerr %.3e  mask maxerr %.3e  refmax %.2f" % (
        (y.cpu() - torch.from_numpy(g["wa_y_nomask"])).abs().max().item(),
        (ym.cpu() - torch.from_numpy(g["wa_y_mask"])).abs().max().item(), float(np.abs(g["wa_y_mask"]).max())))


if __name__ == "__main__":
    names = sys.argv[1:] or list(CHECKS)
    if len(names) == 1 and os.environ.get("BRINGUP_CHILD"):
        CHECKS[names[0]]()
        sys.exit(0)
    for n in names:
        print(f"=== {n}", flush=True)
        r = subprocess.run([sys.executable, __file__, n], env=dict(os.environ, BRINGUP_CHILD="1"), timeout=600)
        print(f"=== {n} exit {r.returncode}", flush=True)
