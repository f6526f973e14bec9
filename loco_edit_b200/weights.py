"""Architecture descriptors and the synthetic-weight factory.

All BASELINE configs run on random-init weights (there is no network for checkpoints).  Parameter
names and shapes are those of the reference module tree `DDPM` (src/models/ddpm/diffusion.py:24-126),
i.e. the state_dict a `celeba_hq.ckpt` / converted `google/ddpm-ema-celebahq-256` checkpoint holds,
so real checkpoints load through the same path.
"""
import zlib

import torch

# src/configs/custom_celeba_ddpm.yml:21-30 (== google/ddpm-ema-celebahq-256 / church-256)
DDPM256 = dict(ch=128, ch_mult=(1, 1, 2, 2, 4, 4), num_res_blocks=2, attn_resolutions=(16,),
               resolution=256, in_ch=3, out_ch=3, gn_eps=1e-6)


def tiny_arch(resolution=32, ch_mult=(1, 2), attn_resolutions=(16,), num_res_blocks=1, ch=128):
    """Reduced-depth variant of the same architecture for fast parity tests."""
    return dict(ch=ch, ch_mult=tuple(ch_mult), num_res_blocks=num_res_blocks,
                attn_resolutions=tuple(attn_resolutions), resolution=resolution, in_ch=3, out_ch=3,
                gn_eps=1e-6)


def ddpm_param_shapes(arch):
    """Ordered {name: shape} of DDPM(arch).state_dict() (reference: ddpm/diffusion.py:24-126)."""
    ch, mult, nrb = arch["ch"], tuple(arch["ch_mult"]), arch["num_res_blocks"]
    attn_res, res = tuple(arch["attn_resolutions"]), arch["resolution"]
    temb = 4 * ch
    out = {}

    def conv(p, cin, cout, k):
        out[p + ".weight"] = (cout, cin, k, k)
        out[p + ".bias"] = (cout,)

    def norm(p, c):
        out[p + ".weight"] = (c,)
        out[p + ".bias"] = (c,)

    def resblock(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cin, cout, 3)
        out[p + ".temb_proj.weight"] = (cout, temb)
        out[p + ".temb_proj.bias"] = (cout,)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".nin_shortcut", cin, cout, 1)

    def attn(p, c):
        norm(p + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv(p + "." + n, c, c, 1)

    out["temb.dense.0.weight"] = (temb, ch)
    out["temb.dense.0.bias"] = (temb,)
    out["temb.dense.1.weight"] = (temb, temb)
    out["temb.dense.1.bias"] = (temb,)
    conv("conv_in", arch["in_ch"], ch, 3)
    in_mult = (1,) + mult
    cur = res
    block_in = ch
    L = len(mult)
    for l in range(L):
        block_in = ch * in_mult[l]
        block_out = ch * mult[l]
        for b in range(nrb):
            resblock(f"down.{l}.block.{b}", block_in, block_out)
            block_in = block_out
            if cur in attn_res:
                attn(f"down.{l}.attn.{b}", block_in)
        if l != L - 1:
            conv(f"down.{l}.downsample.conv", block_in, block_in, 3)
            cur //= 2
    resblock("mid.block_1", block_in, block_in)
    attn("mid.attn_1", block_in)
    resblock("mid.block_2", block_in, block_in)
    for l in reversed(range(L)):
        block_out = ch * mult[l]
        skip_in = ch * mult[l]
        for b in range(nrb + 1):
            if b == nrb:
                skip_in = ch * in_mult[l]
            resblock(f"up.{l}.block.{b}", block_in + skip_in, block_out)
            block_in = block_out
            if cur in attn_res:
                attn(f"up.{l}.attn.{b}", block_in)
        if l != 0:
            conv(f"up.{l}.upsample.conv", block_in, block_in, 3)
            cur *= 2
    norm("norm_out", block_in)
    conv("conv_out", block_in, arch["out_ch"], 3)
    return out


def random_state_dict(arch, seed=1234, perturb_norm=0.0):
    """Seeded random-init weights with torch's default init distributions
    (Conv2d/Linear: U(-1/sqrt(fan_in), 1/sqrt(fan_in)); GroupNorm: weight 1, bias 0).
    Each tensor has its own generator keyed by (seed, name), so the values do not depend on
    enumeration order.  perturb_norm > 0 randomises the GroupNorm affine (tests only)."""
    sd = {}
    for name, shape in ddpm_param_shapes(arch).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
        is_norm = ".norm" in name or name.startswith("norm_out")
        if is_norm:
            base = torch.ones(shape) if name.endswith(".weight") else torch.zeros(shape)
            if perturb_norm > 0:
                base = base + perturb_norm * torch.randn(shape, generator=g)
            sd[name] = base
            continue
        if name.endswith(".weight"):
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
        else:
            wshape = ddpm_param_shapes_cache(arch)[name[:-5] + ".weight"]
            fan_in = 1
            for s in wshape[1:]:
                fan_in *= s
        bound = 1.0 / (fan_in ** 0.5)
        sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


_shape_cache = {}


def ddpm_param_shapes_cache(arch):
    key = repr(sorted(arch.items()))
    if key not in _shape_cache:
        _shape_cache[key] = ddpm_param_shapes(arch)
    return _shape_cache[key]
