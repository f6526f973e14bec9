"""ORACLE (test infrastructure, never on the product path).

CPU fp32 restatement of the decoder half of a latent-diffusion VAE -- the `self.vae.decode(z).sample` of the
Stable-Diffusion twins of the hot path (src/modules/edit.py:764-774) -- as a pure function of a state_dict,
and of the Edit-class arithmetic around it (x0_hat in pixel space, the latent-space power method).

PARITY UNPINNED at the network level: the reference calls diffusers' `AutoencoderKL`, a third-party
dependency (`diffusers==0.11.0`, requirements.txt:4) that is neither under /root/reference nor installed
here, and no checkpoint can be downloaded.  What is restated is its published architecture:
decode(z) = Decoder(post_quant_conv(z)), the CompVis latent-diffusion `Decoder` (conv_in -> mid ResnetBlock,
single-head AttnBlock, ResnetBlock -> per level (num_res_blocks + 1) ResnetBlocks [+ nearest x2 Upsample
with a 3x3 conv] -> GroupNorm(32, eps 1e-6) + SiLU -> conv_out), whose ResnetBlock / AttnBlock / Upsample
are the same modules as src/models/ddpm/diffusion.py:816-966 with temb_channels = 0.  The Edit-class logic
around the networks IS pinned: tests/golden/make_golden_sd.py runs the UNMODIFIED reference class
`EditStableDiffusion` (get_x0, local_encoder_decoder_pullback_zt, get_delta_zt_via_grad) with this decoder
and the stand-in U-Net plugged in as `self.vae` / `self.unet`.

Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this module.
"""
import types

import torch
import torch.nn.functional as F

from .ddpm_ref import attn_block, conv, norm, swish


def resnet_block_notemb(sd, p, x, eps):
    """models/ddpm/diffusion.py:893-912 with temb = None (latent-diffusion Decoder: temb_channels = 0)."""
    h = conv(sd, p + ".conv1", swish(norm(sd, p + ".norm1", x, eps)), padding=1)
    h = conv(sd, p + ".conv2", swish(norm(sd, p + ".norm2", h, eps)), padding=1)
    if (p + ".nin_shortcut.weight") in sd:
        x = conv(sd, p + ".nin_shortcut", x)
    return x + h


def decoder_forward(sd, arch, z):
    """AutoencoderKL.decode: z [B, zc, R, R] -> image [B, 3, R << (L-1), ...]."""
    mult, nrb, eps = tuple(arch["ch_mult"]), arch["num_res_blocks"], arch.get("gn_eps", 1e-6)
    L = len(mult)
    h = conv(sd, "post_quant_conv", z)
    h = conv(sd, "decoder.conv_in", h, padding=1)
    h = resnet_block_notemb(sd, "decoder.mid.block_1", h, eps)
    h = attn_block(sd, "decoder.mid.attn_1", h, eps)
    h = resnet_block_notemb(sd, "decoder.mid.block_2", h, eps)
    for l in reversed(range(L)):
        for b in range(nrb + 1):
            h = resnet_block_notemb(sd, f"decoder.up.{l}.block.{b}", h, eps)
        if l != 0:
            # Upsample: models/ddpm/diffusion.py:826-832, nearest x2 then 3x3 conv
            h = conv(sd, f"decoder.up.{l}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"), padding=1)
    h = swish(norm(sd, "decoder.norm_out", h, eps))
    return conv(sd, "decoder.conv_out", h, padding=1)


class RefVAE:
    """`vae.decode(z).sample` protocol of diffusers' AutoencoderKL (src/modules/edit.py:770)."""

    def __init__(self, arch, sd, dtype=torch.float32):
        self.arch = dict(arch)
        self.sd = {k: v.to(dtype) for k, v in sd.items()}

    def decode(self, z):
        return types.SimpleNamespace(sample=decoder_forward(self.sd, self.arch, z))


VAE_SCALE = 0.18215          # src/modules/edit.py:769 (`z0_hat = 1 / 0.18215 * z0_hat`)


def x0_hat_pixels(eps_fn, vae, zt, at, mask=None):
    """src/modules/edit.py:757-781 (EditStableDiffusion.get_x0): eps_fn(z) is the guided noise prediction;
    z0_hat = (zt - eps sqrt(1-at)) / sqrt(at), decoded after the 1 / 0.18215 rescale, mask-selected."""
    z0 = (zt - eps_fn(zt) * (1 - at) ** 0.5) / at ** 0.5
    x0 = vae.decode(z0 / VAE_SCALE).sample
    return x0[:, mask] if mask is not None else x0


def power_iteration_zt(eps_fn, vae, zt, at, V, mask=None):
    """One pass of the subspace iteration of src/modules/edit.py:872-900 from V [k, d_z]: returns
    (u [k, l_o], s [k] = svdvals(u^T J), V_new [k, d_z]); J = d(mask o decode(z0_hat(z)))/dz."""
    k = V.shape[0]
    f = lambda z: x0_hat_pixels(eps_fn, vae, z, at, mask).reshape(-1)
    us = [torch.func.jvp(f, (zt,), (V[i].reshape(zt.shape),))[1] for i in range(k)]
    u = torch.stack(us, 0)
    z = zt.detach().clone().requires_grad_(True)
    out = f(z)
    w = torch.stack([torch.autograd.grad(out, z, u[i], retain_graph=True)[0].reshape(-1) for i in range(k)], 0)
    _, s, Vh = torch.linalg.svd(w, full_matrices=False)
    return u, s, Vh
