"""GPU: the TMA / tcgen05 implicit-GEMM convolution (stf_conv2d) against torch's conv2d in float64.

Covers what the slice loop, the hyperprior and end_conv ask of it (stf.py:466-548): 3x3 / 5x5 / 1x1, stride 1 / 2,
1-3 channel-concatenated sources incl. channel counts that are not multiples of the 32-float k-block, bias + exact GELU,
PixelShuffle(2) folded into the store, ragged feature maps, channel-slice inputs / outputs, both precision modes, and the
property the decoder depends on: results are bit-identical whatever the batch size."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _tf32_exact(t):
    """Values with <= 10 mantissa bits: the TF32 tensor core reads them exactly."""
    return (t.contiguous().view(torch.int32) & -8192).view(torch.float32)


def _case(B, H, W, chans, N, k, stride=1, shuffle=0, act=False, prec="tf32", exact_inputs=False, seed=0, pad_ld=0):
    from stf_b200 import ops
    g = torch.Generator().manual_seed(seed)
    ctot = sum(chans)
    srcs = []
    for c in chans:
        t = torch.randn(B, H, W, c + pad_ld, generator=g)
        if exact_inputs:
            t = _tf32_exact(t)
        srcs.append(t.cuda()[..., :c] if pad_ld else t.cuda())
    w = torch.randn(N, ctot, k, k, generator=g) / (ctot * k * k) ** 0.5
    b = torch.randn(N, generator=g) * 0.1
    if exact_inputs:
        w = _tf32_exact(w)
    w, b = w.cuda(), b.cuda()
    pc = ops.PackedConv(w, b, chans, stride=stride, pixel_shuffle=shuffle, prec=ops._PRECISIONS[prec])
    y = ops.conv2d(srcs, pc, act=act)
    x64 = torch.cat([s.double() for s in srcs], dim=3).permute(0, 3, 1, 2)
    ref = F.conv2d(x64, w.double(), b.double(), stride=stride, padding=k // 2)
    if act:
        ref = F.gelu(ref)
    if shuffle:
        ref = F.pixel_shuffle(ref, shuffle)
    ref = ref.permute(0, 2, 3, 1).contiguous()
    assert y.shape == ref.shape, (y.shape, ref.shape)
    err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
    return y, err


@pytest.mark.parametrize("prec,tol", [("tf32", 2e-3), ("fp32", 1e-4)])
@pytest.mark.parametrize("cfg", [
    dict(B=1, H=32, W=48, chans=(64,), N=32, k=3),
    dict(B=2, H=32, W=48, chans=(176,), N=128, k=3, act=True),                 # 176 = 5.5 k-blocks: TMA zero-fills the half
    dict(B=3, H=32, W=48, chans=(384, 96, 32), N=224, k=3, act=True),           # cat([latent, support, y_hat_i])
    dict(B=2, H=32, W=48, chans=(336,), N=288, k=3, stride=2, act=True),        # h_a conv3 (stf.py:478)
    dict(B=2, H=8, W=12, chans=(240,), N=1152, k=3, shuffle=2, act=True),       # subpel_conv3x3(240, 288, 2)
    dict(B=40, H=8, W=12, chans=(240,), N=1152, k=3, shuffle=2, act=True),      # ... column tiles overhanging N
    dict(B=5, H=1, W=1, chans=(192,), N=240, k=3, act=True),                    # 1 x 1 feature map (z of a 64 x 64 image)
    dict(B=5, H=4, W=4, chans=(336,), N=288, k=3, stride=2, act=True),
    dict(B=1, H=64, W=96, chans=(48,), N=192, k=5, shuffle=2),                  # end_conv[0] + PixelShuffle (stf.py:466-467)
    dict(B=2, H=30, W=44, chans=(64, 32), N=64, k=3, act=True),                 # ragged: tiles overhang right / bottom
    dict(B=1, H=22, W=32, chans=(192,), N=192, k=3),                            # WACNN hyper size
    dict(B=2, H=16, W=16, chans=(128,), N=336, k=1),                            # 1x1, N tiled with an overhanging last tile
    dict(B=1, H=32, W=48, chans=(64, 32), N=32, k=3, pad_ld=16),                # sources are channel slices of wider tensors
])
def test_conv_matches_float64_reference(cfg, prec, tol):
    _, err = _case(prec=prec, **cfg)
    assert err <= tol, err


@pytest.mark.parametrize("cfg", [
    dict(B=2, H=32, W=48, chans=(96, 32), N=64, k=3, act=False),
    dict(B=1, H=17, W=23, chans=(32,), N=16, k=5, act=False),
    dict(B=2, H=12, W=20, chans=(64,), N=64, k=3, stride=2),
    dict(B=1, H=8, W=12, chans=(32,), N=128, k=3, shuffle=2),
])
def test_conv_index_math_exact_on_tf32_inputs(cfg):
    """Inputs with <= 10 mantissa bits make the single-pass products exact: only fp32 accumulation order is left, so a wrong
    tap / halo / swizzle / shuffle index shows up as an O(1) error instead of hiding under the TF32 tolerance."""
    _, err = _case(prec="tf32", exact_inputs=True, **cfg)
    assert err <= 1e-5, err


@pytest.mark.parametrize("prec", ["tf32", "fp32"])
def test_conv_is_batch_invariant_bit_for_bit(prec):
    """Image b of a batch gives exactly the bits of a batch-1 call (and of another batch size): what lets a stream that was
    encoded inside a batch be decoded alone (stf.py:767: the decoder rebuilds the encoder's indexes)."""
    from stf_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, H, W = 7, 32, 48
    a = torch.randn(B, H, W, 384, generator=g).cuda()
    s = torch.randn(B, H, W, 64, generator=g).cuda()
    w = (torch.randn(224, 448, 3, 3, generator=g) / 60).cuda()
    bias = torch.randn(224, generator=g).cuda()
    pc = ops.PackedConv(w, bias, (384, 64), prec=ops._PRECISIONS[prec])
    full = ops.conv2d([a, s], pc, act=True)
    for lo, hi in ((0, 1), (3, 4), (6, 7), (2, 5)):
        part = ops.conv2d([a[lo:hi].contiguous(), s[lo:hi].contiguous()], pc, act=True)
        assert torch.equal(part, full[lo:hi]), (lo, hi)
    again = ops.conv2d([a, s], pc, act=True)
    assert torch.equal(again, full)                                       # run-to-run deterministic


def test_conv_output_into_channel_slice():
    from stf_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 32, 48, 64, generator=g).cuda()
    w = (torch.randn(32, 64, 3, 3, generator=g) / 24).cuda()
    pc = ops.PackedConv(w, None, (64,))
    wide = torch.full((2, 32, 48, 96), 7.0, device="cuda")
    ops.conv2d([x], pc, out=wide[..., 32:64])
    ref = ops.conv2d([x], pc)
    assert torch.equal(wide[..., 32:64], ref)
    assert bool((wide[..., :32] == 7).all()) and bool((wide[..., 64:] == 7).all())


def test_conv_argument_errors():
    from stf_b200 import ops
    w = torch.randn(32, 64, 3, 3).cuda()
    pc = ops.PackedConv(w, None, (64,))
    with pytest.raises(ValueError):
        ops.conv2d([torch.randn(1, 8, 8, 32).cuda()], pc)                 # wrong channel count
    with pytest.raises(RuntimeError):
        ops.conv2d([torch.randn(1, 8, 8, 64)], pc)                        # CPU tensor: no CPU path
    with pytest.raises(ValueError):
        ops.PackedConv(torch.randn(30, 64, 3, 3).cuda(), None, (64,)).packed is None or ops.conv2d(
            [torch.randn(1, 8, 8, 64).cuda()], ops.PackedConv(torch.randn(30, 64, 3, 3).cuda(), None, (64,)))  # N % 16


@pytest.mark.parametrize("prec,tol", [("tf32", 2e-3), ("fp32", 1e-4)])
def test_conv_lrp_epilogue_in_place(prec, tol):
    """The LRP tail (stf.py:631-633): slot <- slot + 0.5 * tanh(conv(cat([latent, support, slot])) + bias), written in
    place into a 32-channel slot of the wide y_hat buffer that is also the convolution's third source."""
    from stf_b200 import ops
    g = torch.Generator().manual_seed(11)
    B, H, W = 2, 32, 48
    latent = torch.randn(B, H, W, 64, generator=g).cuda()
    y_hat = torch.randn(B, H, W, 384, generator=g).cuda()
    w = (torch.randn(32, 64 + 64 + 32, 3, 3, generator=g) / 30).cuda()
    bias = (torch.randn(32, generator=g) * 0.1).cuda()
    pc = ops.PackedConv(w, bias, (64, 64, 32), prec=ops._PRECISIONS[prec])
    slot = y_hat[..., 224:256]
    before = y_hat.clone()
    x64 = torch.cat([latent, y_hat[..., :64], slot], dim=3).double().permute(0, 3, 1, 2)
    ref = slot.double() + 0.5 * torch.tanh(F.conv2d(x64, w.double(), bias.double(), padding=1)).permute(0, 2, 3, 1)
    ops.conv2d([latent, y_hat[..., :64], slot], pc, act="lrp", out=slot)
    err = (y_hat[..., 224:256].double() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= tol, err
    assert torch.equal(y_hat[..., :224], before[..., :224]) and torch.equal(y_hat[..., 256:], before[..., 256:])
