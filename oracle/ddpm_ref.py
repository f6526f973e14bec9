"""ORACLE (test infrastructure, never on the product path).

CPU fp32 restatement of the reference's DDPM U-Net forward pass as a pure function of a
state_dict.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module.  Pinned against the reference module itself by
tests/golden/make_golden.py (which imports /root/reference/src/models/ddpm/diffusion.py unmodified
in the build container) -> tests/golden/*.pt, checked by tests/test_oracle_golden.py.

Each function cites the reference lines it restates (paths relative to /root/reference/src).
"""
import math

import torch
import torch.nn.functional as F


def timestep_embedding(t, dim):
    """models/ddpm/diffusion.py:783-804 (get_timestep_embedding): [sin, cos], divisor half-1."""
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half, dtype=torch.float32) * -emb)
    emb = t.float()[:, None] * emb[None, :]
    return torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)


def swish(x):
    """models/ddpm/diffusion.py:806-808 (nonlinearity)."""
    return x * torch.sigmoid(x)


def norm(sd, p, x, eps):
    """models/ddpm/diffusion.py:810-811 (Normalize = GroupNorm(32, eps=1e-6, affine))."""
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def conv(sd, p, x, stride=1, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def resnet_block(sd, p, x, temb, eps):
    """models/ddpm/diffusion.py:893-912 (ResnetBlock.forward, dropout = 0)."""
    h = conv(sd, p + ".conv1", swish(norm(sd, p + ".norm1", x, eps)), padding=1)
    h = h + F.linear(swish(temb), sd[p + ".temb_proj.weight"], sd[p + ".temb_proj.bias"])[:, :, None, None]
    h = conv(sd, p + ".conv2", swish(norm(sd, p + ".norm2", h, eps)), padding=1)
    if (p + ".nin_shortcut.weight") in sd:
        x = conv(sd, p + ".nin_shortcut", x)
    return x + h


def attn_block(sd, p, x, eps):
    """models/ddpm/diffusion.py:941-966 (AttnBlock.forward): single head, scale C^-0.5."""
    h_ = norm(sd, p + ".norm", x, eps)
    q, k, v = conv(sd, p + ".q", h_), conv(sd, p + ".k", h_), conv(sd, p + ".v", h_)
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + conv(sd, p + ".proj_out", h_)


def cross_attn_block(sd, p, x, ctx, heads, eps):
    """Cross-attention sub-block of the text-conditioned stand-in (no reference restatement exists:
    diffusers' `Attention` processor computes softmax(q k^T / sqrt(d)) v with q = to_q(hidden),
    k = to_k(encoder_hidden_states), v = to_v(encoder_hidden_states), then to_out; parity unpinned at
    the network level).  x [B,C,H,W], ctx [n_tok, ctx_dim]."""
    b, c, h, w = x.shape
    d = c // heads
    q = conv(sd, p + ".q2", norm(sd, p + ".norm2", x, eps)).reshape(b, heads, d, h * w).transpose(2, 3)   # [b,hd,T,d]
    kv = F.linear(ctx, sd[p + ".kv2.weight"], sd[p + ".kv2.bias"])                                       # [n_tok, 2c]
    k = kv[:, :c].reshape(-1, heads, d).transpose(0, 1)                                                   # [hd,n,d]
    v = kv[:, c:].reshape(-1, heads, d).transpose(0, 1)
    wgt = F.softmax(torch.einsum("bhtd,hnd->bhtn", q, k) * d ** -0.5, dim=-1)
    o = torch.einsum("bhtn,hnd->bhtd", wgt, v).transpose(2, 3).reshape(b, c, h, w)
    return x + conv(sd, p + ".proj_out2", o)


def unet_forward(sd, arch, x, t, cond=None, ctx=None):
    """models/ddpm/diffusion.py:145-200 (PullBackDDPM.forward with op=None).

    x: [B,3,R,R]; t: 0-dim or [1] tensor / float (shared by the batch, as in the reference where
    temb has batch 1 and broadcasts).
    cond (optional, [4*ch]): conditioning embedding added to the timestep embedding before the
    blocks' SiLU + projection -- the stand-in conditional U-Net eps(x, t, c) used to pin the
    Edit-class logic of the T-LOCO twins (the SD / IF networks themselves are diffusers
    code that is not in /root/reference)."""
    ch, mult, nrb = arch["ch"], tuple(arch["ch_mult"]), arch["num_res_blocks"]
    attn_res, eps = tuple(arch["attn_resolutions"]), arch.get("gn_eps", 1e-6)
    L = len(mult)
    heads = arch.get("ctx_heads", 1)
    cross = arch.get("ctx_dim", 0) > 0
    if cross:
        assert ctx is not None, "this architecture has cross-attention layers: pass ctx [n_tok, ctx_dim]"
        _attn = attn_block

        def attn_block_(sd_, p_, h_, eps_):
            return cross_attn_block(sd_, p_, _attn(sd_, p_, h_, eps_), ctx, heads, eps_)
    else:
        attn_block_ = attn_block
    t = torch.as_tensor(t, dtype=torch.float32).reshape(-1)[:1]
    temb = timestep_embedding(t, ch)
    temb = F.linear(temb, sd["temb.dense.0.weight"], sd["temb.dense.0.bias"])
    temb = swish(temb)
    temb = F.linear(temb, sd["temb.dense.1.weight"], sd["temb.dense.1.bias"])
    if cond is not None:
        temb = temb + cond.reshape(1, -1).to(temb.dtype)

    cur = arch["resolution"]
    hs = [conv(sd, "conv_in", x, padding=1)]
    for l in range(L):
        for b in range(nrb):
            h = resnet_block(sd, f"down.{l}.block.{b}", hs[-1], temb, eps)
            if cur in attn_res:
                h = attn_block_(sd, f"down.{l}.attn.{b}", h, eps)
            hs.append(h)
        if l != L - 1:
            # Downsample: models/ddpm/diffusion.py:846-850, pad (0,1,0,1) then stride-2 conv
            hs.append(conv(sd, f"down.{l}.downsample.conv", F.pad(hs[-1], (0, 1, 0, 1)), stride=2))
            cur //= 2
    h = hs[-1]
    h = resnet_block(sd, "mid.block_1", h, temb, eps)
    h = attn_block_(sd, "mid.attn_1", h, eps)
    h = resnet_block(sd, "mid.block_2", h, temb, eps)
    for l in reversed(range(L)):
        for b in range(nrb + 1):
            h = resnet_block(sd, f"up.{l}.block.{b}", torch.cat([h, hs.pop()], dim=1), temb, eps)
            if cur in attn_res:
                h = attn_block_(sd, f"up.{l}.attn.{b}", h, eps)
        if l != 0:
            # Upsample: models/ddpm/diffusion.py:826-832, nearest x2 then 3x3 conv
            h = conv(sd, f"up.{l}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"), padding=1)
            cur *= 2
    h = swish(norm(sd, "norm_out", h, eps))
    return conv(sd, "conv_out", h, padding=1)


class RefUNet:
    """Callable with the reference's `unet(x, t)` protocol (src/modules/edit.py:2375)."""

    def __init__(self, arch, sd, dtype=torch.float32):
        self.arch = dict(arch)
        self.sd = {k: v.to(dtype) for k, v in sd.items()}

    def __call__(self, x, t, cond=None, ctx=None):
        return unet_forward(self.sd, self.arch, x, t, cond=cond, ctx=ctx)
