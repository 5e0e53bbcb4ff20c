"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's window-attention path.

Functional (state_dict + key prefix) restatement in plain torch fp32 ops, in the reference's
operation order, of:
  WindowAttention.forward            compressai/models/stf.py:90-121 (= layers/win_attention.py:84-115)
  relative_position_index            stf.py:69-80
  shifted-window mask                stf.py:316-334, layers/win_attention.py:159-179
  SwinTransformerBlock.forward       stf.py:149-199 (eval: DropPath = identity)
  Mlp.forward                        stf.py:34-40
  PatchMerging / PatchSplit          stf.py:209-235 / 251-260
  BasicLayer.forward                 stf.py:308-347
  WinBasedAttention.forward          layers/win_attention.py:153-207
"""
import math

import torch
import torch.nn.functional as F

MASK_VALUE = -100.0  # stf.py:334


def to_windows(x, ws):
    """stf.py:42-46: (B,H,W,C) -> (B*nW, ws, ws, C), windows row-major inside each image."""
    B, H, W, C = x.shape
    x = x.reshape(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C)


def from_windows(w, ws, H, W):
    """stf.py:49-53."""
    B = w.shape[0] // ((H // ws) * (W // ws))
    x = w.reshape(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


def relative_position_index(ws):
    """stf.py:69-80: idx[n,m] = (h_n-h_m+ws-1)*(2ws-1) + (w_n-w_m+ws-1)."""
    r = torch.arange(ws)
    hh, ww = torch.meshgrid(r, r, indexing="ij")
    hh, ww = hh.reshape(-1), ww.reshape(-1)
    return (hh[:, None] - hh[None, :] + ws - 1) * (2 * ws - 1) + (ww[:, None] - ww[None, :] + ws - 1)


def shift_mask(Hp, Wp, ws, shift):
    """stf.py:316-334: region ids from 3x3 slices on the (already shifted) frame -> {0,-100}."""
    img = torch.zeros(1, Hp, Wp, 1)
    edges = (slice(0, -ws), slice(-ws, -shift), slice(-shift, None))
    k = 0
    for hs in edges:
        for wsl in edges:
            img[:, hs, wsl, :] = k
            k += 1
    ids = to_windows(img, ws).reshape(-1, ws * ws)
    diff = ids[:, None, :] - ids[:, :, None]
    return torch.where(diff != 0, torch.tensor(MASK_VALUE), torch.tensor(0.0))


def window_attention(sd, pfx, x, num_heads, ws, mask=None):
    """stf.py:90-121.  x: (B_, N, C); mask: (nW, N, N) or None."""
    B_, N, C = x.shape
    d = C // num_heads
    qkv = F.linear(x, sd[pfx + "qkv.weight"], sd[pfx + "qkv.bias"])
    qkv = qkv.reshape(B_, N, 3, num_heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (d ** -0.5)
    attn = q @ k.transpose(-2, -1)
    table = sd[pfx + "relative_position_bias_table"]
    bias = table[relative_position_index(ws).reshape(-1)].reshape(N, N, num_heads).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.reshape(B_ // nW, nW, num_heads, N, N) + mask[None, :, None]
        attn = attn.reshape(-1, num_heads, N, N)
    attn = torch.softmax(attn, dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(out, sd[pfx + "proj.weight"], sd[pfx + "proj.bias"])


def mlp(sd, pfx, x):
    """stf.py:34-40 (dropout p=0)."""
    h = F.gelu(F.linear(x, sd[pfx + "fc1.weight"], sd[pfx + "fc1.bias"]))
    return F.linear(h, sd[pfx + "fc2.weight"], sd[pfx + "fc2.bias"])


def layer_norm(sd, pfx, x):
    return F.layer_norm(x, (x.shape[-1],), sd[pfx + "weight"], sd[pfx + "bias"], 1e-5)


def swin_block(sd, pfx, x, H, W, num_heads, ws, shift, mask):
    """stf.py:149-199 in eval mode.  x: (B, H*W, C)."""
    B, L, C = x.shape
    assert L == H * W
    shortcut = x
    h = layer_norm(sd, pfx + "norm1.", x).reshape(B, H, W, C)
    pad_r = (ws - W % ws) % ws
    pad_b = (ws - H % ws) % ws
    h = F.pad(h, (0, 0, 0, pad_r, 0, pad_b))
    Hp, Wp = h.shape[1], h.shape[2]
    if shift > 0:
        h = torch.roll(h, shifts=(-shift, -shift), dims=(1, 2))
    win = to_windows(h, ws).reshape(-1, ws * ws, C)
    win = window_attention(sd, pfx + "attn.", win, num_heads, ws, mask if shift > 0 else None)
    h = from_windows(win.reshape(-1, ws, ws, C), ws, Hp, Wp)
    if shift > 0:
        h = torch.roll(h, shifts=(shift, shift), dims=(1, 2))
    h = h[:, :H, :W, :].reshape(B, H * W, C)
    x = shortcut + h
    return x + mlp(sd, pfx + "mlp.", layer_norm(sd, pfx + "norm2.", x))


def patch_merging(sd, pfx, x, H, W):
    """stf.py:209-235."""
    B, L, C = x.shape
    x = x.reshape(B, H, W, C)
    if H % 2 or W % 2:
        x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    x = x.reshape(B, -1, 4 * C)
    x = layer_norm(sd, pfx + "norm.", x)
    return F.linear(x, sd[pfx + "reduction.weight"])


def patch_split(sd, pfx, x, H, W):
    """stf.py:251-260: LN -> Linear(C,2C) -> PixelShuffle(2) in token layout."""
    B, L, C = x.shape
    x = F.linear(layer_norm(sd, pfx + "norm.", x), sd[pfx + "reduction.weight"])
    x = x.permute(0, 2, 1).reshape(B, 2 * C, H, W)
    x = F.pixel_shuffle(x, 2)
    return x.permute(0, 2, 3, 1).reshape(B, 4 * L, -1)


def basic_layer(sd, pfx, x, H, W, depth, num_heads, ws, resample):
    """stf.py:308-347.  resample in (None, 'merge', 'split').  Returns (x, H', W')."""
    shift = ws // 2
    Hp = int(math.ceil(H / ws)) * ws
    Wp = int(math.ceil(W / ws)) * ws
    mask = shift_mask(Hp, Wp, ws, shift)
    for i in range(depth):
        x = swin_block(sd, f"{pfx}blocks.{i}.", x, H, W, num_heads, ws, 0 if i % 2 == 0 else shift, mask)
    if resample == "merge":
        return patch_merging(sd, pfx + "downsample.", x, H, W), (H + 1) // 2, (W + 1) // 2
    if resample == "split":
        return patch_split(sd, pfx + "downsample.", x, H, W), H * 2, W * 2
    return x, H, W


def win_based_attention(sd, pfx, x, num_heads, ws, shift):
    """layers/win_attention.py:153-207.  x: (B,C,H,W) -> same; residual included, no LN/MLP."""
    B, C, H, W = x.shape
    h = x.permute(0, 2, 3, 1)
    mask = shift_mask(H, W, ws, shift) if shift > 0 else None
    if shift > 0:
        h = torch.roll(h, shifts=(-shift, -shift), dims=(1, 2))
    win = to_windows(h, ws).reshape(-1, ws * ws, C)
    win = window_attention(sd, pfx + "attn.", win, num_heads, ws, mask)
    h = from_windows(win.reshape(-1, ws, ws, C), ws, H, W)
    if shift > 0:
        h = torch.roll(h, shifts=(shift, shift), dims=(1, 2))
    return x + h.permute(0, 3, 1, 2)
