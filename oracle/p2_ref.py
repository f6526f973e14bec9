"""ORACLE (test infrastructure, never on the product path).

CPU fp32 restatement of the reference's P2 / guided-diffusion U-Net forward pass
(models/guided_diffusion/unet.py:398-684, configuration P2_DICT of script_util.py:166-190) as a
pure function of a state_dict.  Pinned against the unmodified reference `create_model(...)` by
tests/golden/make_golden_p2.py -> tests/golden/p2_tiny.pt, p2_pullback_tiny.pt
(tests/test_oracle_golden.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this.
Citations are relative to /root/reference/src/models/guided_diffusion.
"""
import math

import torch
import torch.nn.functional as F


def timestep_embedding(t, dim):
    """nn.py:103-121: [cos, sin], freqs = exp(-ln(1e4) * i / half)."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def gn(sd, p, x, eps):
    """nn.py:17-19, 93-100: GroupNorm32(32, C), default eps 1e-5."""
    return F.group_norm(x.float(), 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def res_block(sd, p, x, emb, eps, up=False, down=False):
    """unet.py:238-258 (ResBlock.forward, use_scale_shift_norm = True, dropout 0)."""
    h = F.silu(gn(sd, p + ".in_layers.0", x, eps))
    if up:                                                      # Upsample(use_conv=False): nearest x2
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    elif down:                                                  # Downsample(use_conv=False): avg-pool 2
        h = F.avg_pool2d(h, 2, 2)
        x = F.avg_pool2d(x, 2, 2)
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    emb_out = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    scale, shift = torch.chunk(emb_out[:, :, None, None], 2, dim=1)
    h = gn(sd, p + ".out_layers.0", h, eps) * (1 + scale) + shift
    h = F.conv2d(F.silu(h), sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if (p + ".skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def attention_block(sd, p, x, head_ch, eps):
    """unet.py:301-356 (AttentionBlock + QKVAttentionLegacy: heads split before q/k/v)."""
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(gn(sd, p + ".norm", xf, eps), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    n_heads = c // head_ch
    length = qkv.shape[-1]
    q, k, v = qkv.reshape(b * n_heads, head_ch * 3, length).split(head_ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(head_ch))
    weight = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    weight = torch.softmax(weight.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", weight, v).reshape(b, -1, length)
    h = F.conv1d(a, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xf + h).reshape(b, c, hh, ww)


def p2_layout(arch):
    """Module layout of UNetModel.__init__ (unet.py:470-618) for resblock_updown = True:
    list of input blocks / output blocks, each a list of ('res'|'attn'|'down'|'up', cin, cout)."""
    ch0, mult, nrb = arch["ch"], tuple(arch["ch_mult"]), arch["num_res_blocks"]
    attn_ds = tuple(arch["resolution"] // r for r in arch["attn_resolutions"])   # script_util.py:411-413
    ch = int(mult[0] * ch0)
    inputs = [[("conv_in", 3, ch)]]
    chans = [ch]
    ds = 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            layers = [("res", ch, int(m * ch0))]
            ch = int(m * ch0)
            if ds in attn_ds:
                layers.append(("attn", ch, ch))
            inputs.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            inputs.append([("down", ch, ch)])
            chans.append(ch)
            ds *= 2
    middle = [("res", ch, ch), ("attn", ch, ch), ("res", ch, ch)]
    outputs = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, int(ch0 * m))]
            ch = int(ch0 * m)
            if ds in attn_ds:
                layers.append(("attn", ch, ch))
            if level and i == nrb:
                layers.append(("up", ch, ch))
                ds //= 2
            outputs.append(layers)
    return inputs, middle, outputs, ch


def unet_forward(sd, arch, x, t):
    """unet.py:636-684 (UNetModel.forward): eps = first half of the 6 output channels."""
    eps_gn = arch.get("gn_eps", 1e-5)
    head_ch = arch["head_ch"]
    t = torch.as_tensor(t, dtype=torch.float32).reshape(-1)[:1]
    emb = timestep_embedding(t, arch["ch"])
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    inputs, middle, outputs, _ = p2_layout(arch)

    def run(layers, prefix, h):
        for j, (kind, _, _) in enumerate(layers):
            p = f"{prefix}.{j}"
            if kind == "conv_in":
                h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
            elif kind == "res":
                h = res_block(sd, p, h, emb, eps_gn)
            elif kind == "down":
                h = res_block(sd, p, h, emb, eps_gn, down=True)
            elif kind == "up":
                h = res_block(sd, p, h, emb, eps_gn, up=True)
            elif kind == "attn":
                h = attention_block(sd, p, h, head_ch, eps_gn)
        return h

    hs = []
    h = x
    for i, layers in enumerate(inputs):
        h = run(layers, f"input_blocks.{i}", h)
        hs.append(h)
    h = run(middle, "middle_block", h)
    for i, layers in enumerate(outputs):
        h = run(layers, f"output_blocks.{i}", torch.cat([h, hs.pop()], dim=1))
    h = F.silu(gn(sd, "out.0", h, eps_gn))
    h = F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)
    return h[:, : h.shape[1] // 2]            # et, logvar_learned = split(h, 3)


class RefP2UNet:
    def __init__(self, arch, sd):
        self.arch = dict(arch)
        self.sd = {k: v.float() for k, v in sd.items()}

    def __call__(self, x, t):
        return unet_forward(self.sd, self.arch, x, t)
