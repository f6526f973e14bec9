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


def tiny_arch(resolution=32, ch_mult=(1, 2), attn_resolutions=(16,), num_res_blocks=1, ch=128, ctx_dim=0,
              ctx_heads=1):
    """Reduced-depth variant of the same architecture for fast parity tests (ctx_dim > 0: with a
    cross-attention sub-block in every AttnBlock)."""
    a = dict(ch=ch, ch_mult=tuple(ch_mult), num_res_blocks=num_res_blocks,
             attn_resolutions=tuple(attn_resolutions), resolution=resolution, in_ch=3, out_ch=3,
             gn_eps=1e-6)
    if ctx_dim > 0:
        a.update(ctx_dim=ctx_dim, ctx_heads=ctx_heads)
    return a


# The DDPM-256 U-Net with cross-attention to a 77 x 768 prompt embedding (CLIP ViT-L/14 text states, the
# conditioning of Stable Diffusion 1.x) in each of its six AttnBlocks, 8 heads of 64 channels: the
# text-conditioned stand-in of BASELINE configs 4-5 (the SD / IF networks are diffusers code that is not
# under /root/reference)
DDPM256_TEXT = dict(DDPM256, ctx_dim=768, ctx_heads=8)


def if_standin_arch(resolution=64):
    """Text-conditioned stand-in at the DeepFloyd-IF stage-I size (BASELINE config 5: 64 x 64 pixels): four
    levels 64 -> 8, self- and cross-attention (4 heads) at 16^2 and 8^2, prompt embedding 77 x 768."""
    return dict(DDPM256, resolution=resolution, ch_mult=(1, 2, 4, 4), attn_resolutions=(16, 8), ctx_dim=768,
                ctx_heads=4)


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
        if arch.get("ctx_dim", 0) > 0:
            # cross-attention sub-block to a prompt embedding [n_tok, ctx_dim] (text-conditioned twins):
            # GroupNorm, q projection, one Linear for K_c | V_c, output projection
            norm(p + ".norm2", c)
            conv(p + ".q2", c, c, 1)
            out[p + ".kv2.weight"] = (2 * c, arch["ctx_dim"])
            out[p + ".kv2.bias"] = (2 * c,)
            conv(p + ".proj_out2", c, c, 1)

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


# src/models/guided_diffusion/script_util.py:166-190 (P2_DICT: FFHQ_P2 / AFHQ_P2 / Flower_P2 / ...)
P2_256 = dict(kind="p2", ch=128, ch_mult=(1, 1, 2, 2, 4, 4), num_res_blocks=1,
              attn_resolutions=(16,), head_ch=64, resolution=256, in_ch=3, out_ch=3, gn_eps=1e-5)


def tiny_p2_arch(resolution=32, ch_mult=(1, 2), attn_resolutions=(16,), num_res_blocks=1, ch=128,
                 head_ch=64):
    """Reduced-depth variant of the P2 / guided-diffusion architecture for fast parity tests."""
    return dict(kind="p2", ch=ch, ch_mult=tuple(ch_mult), num_res_blocks=num_res_blocks,
                attn_resolutions=tuple(attn_resolutions), head_ch=head_ch, resolution=resolution,
                in_ch=3, out_ch=3, gn_eps=1e-5)


def p2_layout(arch):
    """Module layout of UNetModel.__init__ (guided_diffusion/unet.py:470-618) for
    resblock_updown = True: (input blocks, middle, output blocks), each block a list of
    ('conv_in'|'res'|'attn'|'down'|'up', cin, cout)."""
    ch0, mult, nrb = arch["ch"], tuple(arch["ch_mult"]), arch["num_res_blocks"]
    attn_ds = tuple(arch["resolution"] // r for r in arch["attn_resolutions"])
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
    return inputs, middle, outputs


def p2_param_shapes(arch):
    """Ordered {name: shape} of the reference's guided-diffusion UNetModel.state_dict() built by
    create_model(**P2_DICT) (script_util.py:379-435; learn_sigma => 6 output channels)."""
    ch0 = arch["ch"]
    temb = 4 * ch0
    out = {}

    def res(p, cin, cout):
        out[p + ".in_layers.0.weight"] = (cin,)
        out[p + ".in_layers.0.bias"] = (cin,)
        out[p + ".in_layers.2.weight"] = (cout, cin, 3, 3)
        out[p + ".in_layers.2.bias"] = (cout,)
        out[p + ".emb_layers.1.weight"] = (2 * cout, temb)
        out[p + ".emb_layers.1.bias"] = (2 * cout,)
        out[p + ".out_layers.0.weight"] = (cout,)
        out[p + ".out_layers.0.bias"] = (cout,)
        out[p + ".out_layers.3.weight"] = (cout, cout, 3, 3)
        out[p + ".out_layers.3.bias"] = (cout,)
        if cin != cout:
            out[p + ".skip_connection.weight"] = (cout, cin, 1, 1)
            out[p + ".skip_connection.bias"] = (cout,)

    def attn(p, c):
        out[p + ".norm.weight"] = (c,)
        out[p + ".norm.bias"] = (c,)
        out[p + ".qkv.weight"] = (3 * c, c, 1)
        out[p + ".qkv.bias"] = (3 * c,)
        out[p + ".proj_out.weight"] = (c, c, 1)
        out[p + ".proj_out.bias"] = (c,)

    def block(prefix, layers):
        for j, (kind, cin, cout) in enumerate(layers):
            p = f"{prefix}.{j}"
            if kind == "conv_in":
                out[p + ".weight"] = (cout, cin, 3, 3)
                out[p + ".bias"] = (cout,)
            elif kind == "attn":
                attn(p, cin)
            else:
                res(p, cin, cout)

    out["time_embed.0.weight"] = (temb, ch0)
    out["time_embed.0.bias"] = (temb,)
    out["time_embed.2.weight"] = (temb, temb)
    out["time_embed.2.bias"] = (temb,)
    inputs, middle, outputs = p2_layout(arch)
    for i, layers in enumerate(inputs):
        block(f"input_blocks.{i}", layers)
    block("middle_block", middle)
    for i, layers in enumerate(outputs):
        block(f"output_blocks.{i}", layers)
    c_last = outputs[-1][0][2]
    out["out.0.weight"] = (c_last,)
    out["out.0.bias"] = (c_last,)
    out["out.2.weight"] = (6, c_last, 3, 3)
    out["out.2.bias"] = (6,)
    return out


# Decoder half of the Stable Diffusion 1.x VAE (diffusers AutoencoderKL config of CompVis/stable-diffusion-v1-*:
# block_out_channels (128, 256, 512, 512), layers_per_block 2, latent_channels 4, one mid-block attention):
# latent [4, 64, 64] -> image [3, 512, 512].  `resolution` is the LATENT resolution.
SD_VAE_DECODER = dict(kind="vae_decoder", ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, attn_resolutions=(),
                      resolution=64, in_ch=4, out_ch=3, gn_eps=1e-6)


def tiny_vae_decoder_arch(resolution=8, ch_mult=(1, 2), num_res_blocks=1, ch=128, in_ch=4):
    """Reduced-depth VAE decoder for fast parity tests (image = resolution << (levels - 1))."""
    return dict(kind="vae_decoder", ch=ch, ch_mult=tuple(ch_mult), num_res_blocks=num_res_blocks,
                attn_resolutions=(), resolution=resolution, in_ch=in_ch, out_ch=3, gn_eps=1e-6)


def latent_unet_arch(resolution=16, ch_mult=(1, 2), attn_resolutions=(8,), num_res_blocks=1, ch=128, ctx_dim=64,
                     ctx_heads=2):
    """Text-conditioned stand-in U-Net over 4-channel latents (the shape of the Stable Diffusion U-Net's
    interface: src/modules/edit.py:655-658)."""
    a = tiny_arch(resolution, ch_mult, attn_resolutions, num_res_blocks, ch, ctx_dim, ctx_heads)
    a.update(in_ch=4, out_ch=4)
    return a


def sd_standin_unet_arch(resolution=64):
    """Text-conditioned stand-in at the Stable Diffusion 1.x interface (BASELINE config 4): latents [4, 64, 64],
    four levels 64 -> 8, self- and cross-attention (8 heads) at 16^2 and 8^2 (<= 256 tokens: the fused tcgen05
    attention kernels), prompt embedding 77 x 768."""
    return dict(DDPM256, resolution=resolution, ch_mult=(1, 2, 4, 4), attn_resolutions=(resolution // 4, resolution // 8),
                ctx_dim=768, ctx_heads=8, in_ch=4, out_ch=4)


def vae_decoder_param_shapes(arch):
    """Ordered {name: shape} of the decoder half of a latent-diffusion VAE: `post_quant_conv` + the CompVis
    `Decoder` module tree that diffusers' AutoencoderKL restates (conv_in, mid.{block_1, attn_1, block_2},
    up.{l}.block.{b} [+ up.{l}.upsample.conv], norm_out, conv_out; ResnetBlock / AttnBlock / Upsample are the
    modules of src/models/ddpm/diffusion.py:816-966 without the timestep projection)."""
    ch, mult, nrb, zc = arch["ch"], tuple(arch["ch_mult"]), arch["num_res_blocks"], arch.get("in_ch", 4)
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
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".nin_shortcut", cin, cout, 1)

    conv("post_quant_conv", zc, zc, 1)
    L = len(mult)
    block_in = ch * mult[L - 1]
    conv("decoder.conv_in", zc, block_in, 3)
    resblock("decoder.mid.block_1", block_in, block_in)
    norm("decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv("decoder.mid.attn_1." + n, block_in, block_in, 1)
    resblock("decoder.mid.block_2", block_in, block_in)
    for l in reversed(range(L)):
        block_out = ch * mult[l]
        for b in range(nrb + 1):
            resblock(f"decoder.up.{l}.block.{b}", block_in, block_out)
            block_in = block_out
        if l != 0:
            conv(f"decoder.up.{l}.upsample.conv", block_in, block_in, 3)
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", block_in, arch.get("out_ch", 3), 3)
    return out


def param_shapes(arch):
    """{name: shape} of the reference state_dict for either architecture family."""
    if arch.get("kind") == "vae_decoder":
        return vae_decoder_param_shapes(arch)
    return p2_param_shapes(arch) if arch.get("kind") == "p2" else ddpm_param_shapes(arch)


def _is_norm_name(arch, name):
    if arch.get("kind") == "p2":
        return (".in_layers.0." in name or ".out_layers.0." in name or ".norm." in name
                or name.startswith("out.0."))
    return ".norm" in name or name.startswith("norm_out")


def _is_zero_init(arch, name):
    """Tensors the reference zero-initialises (zero_module, guided_diffusion/nn.py:68-74;
    unet.py:212-214, 296, 617).  With random weights they would make eps == 0, so the synthetic
    factory re-randomises them with N(0, 0.02)."""
    return arch.get("kind") == "p2" and (".out_layers.3." in name or ".proj_out." in name
                                         or name.startswith("out.2."))


def random_state_dict(arch, seed=1234, perturb_norm=0.0):
    """Seeded random-init weights with torch's default init distributions
    (Conv2d/Linear: U(-1/sqrt(fan_in), 1/sqrt(fan_in)); GroupNorm: weight 1, bias 0).
    Each tensor has its own generator keyed by (seed, name), so the values do not depend on
    enumeration order.  perturb_norm > 0 randomises the GroupNorm affine (tests only)."""
    sd = {}
    for name, shape in param_shapes(arch).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
        is_norm = _is_norm_name(arch, name)
        if _is_zero_init(arch, name):
            sd[name] = 0.02 * torch.randn(shape, generator=g)
            continue
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
        _shape_cache[key] = param_shapes(arch)
    return _shape_cache[key]


# ------------------------------------------------------------------------------------------------
# Hugging Face `UNet2DModel` checkpoints (google/ddpm-ema-{celebahq,church,bedroom}-256)
# ------------------------------------------------------------------------------------------------
_HF_RES = {"norm1": "norm1", "conv1": "conv1", "time_emb_proj": "temb_proj", "norm2": "norm2",
           "conv2": "conv2", "conv_shortcut": "nin_shortcut"}
# diffusers 0.11 (the reference's pin, requirements.txt:4) names the attention projections
# query/key/value/proj_attn; later releases to_q/to_k/to_v/to_out.0.  Both are Linear [C, C].
_HF_ATTN = {"group_norm": "norm", "query": "q", "key": "k", "value": "v", "proj_attn": "proj_out",
            "to_q": "q", "to_k": "k", "to_v": "v", "to_out.0": "proj_out"}


def hf_unet2d_to_ddpm(sd, arch=DDPM256):
    """Rename a diffusers `UNet2DModel.state_dict()` of the `*_HF` models the reference loads with
    `DDIMPipeline.from_pretrained` (src/utils/utils.py:93-98, 122-125) to the `DDPM.state_dict()` names
    this package packs (src/models/ddpm/diffusion.py:24-126; SURVEY 8c: same architecture).  The
    attention projections are Linear [C, C] there and 1x1 Conv [C, C, 1, 1] here.

    diffusers is not installed in this image, so the name table follows the published module tree of
    `UNet2DModel` (down_blocks / mid_block / up_blocks, resnets / attentions / downsamplers /
    upsamplers) and is checked for completeness against `ddpm_param_shapes(arch)`: every expected
    parameter must be produced exactly once with the right shape, anything else raises."""
    L = len(tuple(arch["ch_mult"]))
    out = {}
    for name, w in sd.items():
        parts = name.split(".")
        leaf = parts[-1]                      # weight | bias
        new = None
        if parts[0] == "time_embedding":
            new = "temb.dense.%d.%s" % ({"linear_1": 0, "linear_2": 1}[parts[1]], leaf)
        elif parts[0] in ("conv_in", "conv_out"):
            new = name
        elif parts[0] == "conv_norm_out":
            new = "norm_out." + leaf
        elif parts[0] in ("down_blocks", "up_blocks", "mid_block"):
            if parts[0] == "mid_block":
                kind, idx, rest = parts[1], int(parts[2]), parts[3:-1]
                prefix = None
            else:
                lvl, kind, idx, rest = int(parts[1]), parts[2], int(parts[3]), parts[4:-1]
                prefix = "down.%d" % lvl if parts[0] == "down_blocks" else "up.%d" % (L - 1 - lvl)
            sub = ".".join(rest)
            if kind == "resnets":
                tgt = ("mid.block_%d" % (idx + 1)) if prefix is None else "%s.block.%d" % (prefix, idx)
                new = "%s.%s.%s" % (tgt, _HF_RES[sub], leaf)
            elif kind == "attentions":
                tgt = "mid.attn_1" if prefix is None else "%s.attn.%d" % (prefix, idx)
                new = "%s.%s.%s" % (tgt, _HF_ATTN[sub], leaf)
                if sub != "group_norm" and leaf == "weight" and w.dim() == 2:
                    w = w[:, :, None, None]
            elif kind == "downsamplers":
                new = "%s.downsample.conv.%s" % (prefix, leaf)
            elif kind == "upsamplers":
                new = "%s.upsample.conv.%s" % (prefix, leaf)
        if new is None:
            raise KeyError("hf_unet2d_to_ddpm: unexpected parameter '%s'" % name)
        if new in out:
            raise KeyError("hf_unet2d_to_ddpm: '%s' maps to '%s' twice" % (name, new))
        out[new] = w
    want = ddpm_param_shapes(arch)
    missing = [k for k in want if k not in out]
    extra = [k for k in out if k not in want]
    if missing or extra:
        raise KeyError("hf_unet2d_to_ddpm: missing %s, unexpected %s" % (missing[:3], extra[:3]))
    for k, shp in want.items():
        if tuple(out[k].shape) != tuple(shp):
            raise ValueError("hf_unet2d_to_ddpm: %s has shape %s, expected %s" % (k, tuple(out[k].shape), shp))
    return out


def is_hf_unet2d_state_dict(sd):
    return any(k.startswith(("down_blocks.", "time_embedding.")) for k in sd)


# ------------------------------------------------------------------------------------------------
# Hugging Face `AutoencoderKL` checkpoints (the `vae` of a Stable Diffusion pipeline, src/utils/utils.py:217)
# ------------------------------------------------------------------------------------------------
def hf_autoencoderkl_to_decoder(sd, arch=SD_VAE_DECODER):
    """diffusers `AutoencoderKL.state_dict()` -> the decoder-half names of `vae_decoder_param_shapes`
    (encoder.* and quant_conv.* are dropped: the path only decodes, src/modules/edit.py:770).
    up_blocks are numbered in execution order (0 = lowest resolution), the CompVis `up.{l}` by level;
    the mid-block attention projections are Linear [C, C] (query/key/value/proj_attn in diffusers 0.11,
    to_q/to_k/to_v/to_out.0 later) and become 1x1 convs."""
    L = len(tuple(arch["ch_mult"]))
    out = {}
    for name, w in sd.items():
        parts = name.split(".")
        leaf = parts[-1]
        if parts[0] in ("encoder", "quant_conv"):
            continue
        new = None
        if parts[0] == "post_quant_conv":
            new = name
        elif parts[0] == "decoder":
            if parts[1] in ("conv_in", "conv_out"):
                new = name
            elif parts[1] == "conv_norm_out":
                new = "decoder.norm_out." + leaf
            elif parts[1] == "mid_block":
                kind, idx, sub = parts[2], int(parts[3]), ".".join(parts[4:-1])
                if kind == "resnets":
                    new = "decoder.mid.block_%d.%s.%s" % (idx + 1, _HF_RES[sub], leaf)
                elif kind == "attentions":
                    new = "decoder.mid.attn_1.%s.%s" % (_HF_ATTN[sub], leaf)
                    if sub != "group_norm" and leaf == "weight" and w.dim() == 2:
                        w = w[:, :, None, None]
            elif parts[1] == "up_blocks" and len(parts) >= 6 and parts[4].isdigit():
                lvl, kind, idx, sub = L - 1 - int(parts[2]), parts[3], int(parts[4]), ".".join(parts[5:-1])
                if kind == "resnets":
                    new = "decoder.up.%d.block.%d.%s.%s" % (lvl, idx, _HF_RES[sub], leaf)
                elif kind == "upsamplers":
                    new = "decoder.up.%d.upsample.conv.%s" % (lvl, leaf)
        if new is None:
            raise KeyError("hf_autoencoderkl_to_decoder: unexpected parameter '%s'" % name)
        if new in out:
            raise KeyError("hf_autoencoderkl_to_decoder: '%s' maps to '%s' twice" % (name, new))
        out[new] = w
    want = vae_decoder_param_shapes(arch)
    missing = [k for k in want if k not in out]
    extra = [k for k in out if k not in want]
    if missing or extra:
        raise KeyError("hf_autoencoderkl_to_decoder: missing %s, unexpected %s" % (missing[:3], extra[:3]))
    for k, shp in want.items():
        if tuple(out[k].shape) != tuple(shp):
            raise ValueError("hf_autoencoderkl_to_decoder: %s has shape %s, expected %s" % (k, tuple(out[k].shape), shp))
    return out


def is_hf_autoencoderkl_state_dict(sd):
    return any(k.startswith(("decoder.up_blocks.", "decoder.mid_block.")) for k in sd)
