"""GPU parity of the fp16 storage / tcgen05 kind::f16 variant of the Jacobian-free U-Net programs
(the DDIM inversion and denoising loops, `self.unet(xt, t)` at src/modules/edit.py:2151, 2572).

fp16 carries the same 10-bit mantissa as the tf32 operands of the default path, so the tolerances
are the TF32 ones:
  * convolutions on fp16-representable inputs: products exact, fp32 accumulation, fp16 rounding of the
    stored result -> 1e-3 relative (half an fp16 ulp is 4.9e-4);
  * U-Net eps against the fp32 CPU oracle / the unmodified reference: relative L2 < 5e-3;
  * the final DDIM stage (59 steps, eta = 1 from index 79) against the oracle: PSNR >= 40 dB
    (north_star) -- the CPU experiment profiles/r2_precision_experiment_64.txt predicts ~66 dB.
"""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from gpu_util import nchw, nhwc, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


def _ref(kind, x, w):
    xd, wd = x.double(), w.double()
    if kind == 0:
        return F.conv2d(xd, wd, padding=1)
    if kind == 1:
        return F.conv2d(xd, wd)
    if kind == 2:
        return F.conv2d(F.pad(xd, (0, 1, 0, 1)), wd, stride=2)
    if kind == 3:
        return F.conv_transpose2d(xd, wd, padding=1)
    N, _, h, w_ = xd.shape
    return F.conv_transpose2d(xd, wd, stride=2)[:, :, : 2 * h, : 2 * w_]


CASES = [
    # kind, N, H, W, Cin, Cout -- one-tile kernel, split-K, two-tile, halo / CTA-pair variants
    (0, 1, 16, 16, 128, 128), (0, 3, 8, 8, 256, 128), (0, 2, 64, 64, 256, 256), (1, 6, 32, 32, 256, 128),
    (1, 2, 16, 16, 128, 384), (2, 4, 32, 32, 128, 128), (3, 5, 32, 32, 128, 256), (4, 3, 16, 16, 128, 128),
    (0, 6, 128, 128, 128, 128), (3, 5, 128, 128, 128, 128), (1, 6, 128, 128, 256, 128), (0, 5, 128, 128, 256, 256),
    (0, 1, 256, 256, 128, 128), (0, 8, 256, 256, 128, 128), (0, 1, 8, 8, 512, 512),
]


@pytest.mark.parametrize("kind,N,H,W,Cin,Cout", CASES)
@pytest.mark.parametrize("io", ["h->h", "h->f", "f->h"])
def test_conv_fp16_variants(dev, kind, N, H, W, Cin, Cout, io):
    """fp16 operands on kind::f16 (64 channels per swizzle row) and the mixed boundaries the U-Net
    programs use around the fp32 attention core: fp16 in -> fp32 out (q|k|v), fp32 in -> fp16 out
    (proj_out + residual)."""
    from loco_edit_b200 import ops
    tin = torch.float16 if io[0] == "h" else torch.float32
    tout = torch.float16 if io[-1] == "h" else torch.float32
    g = torch.Generator().manual_seed(kind * 1000 + N * 100 + H)
    ksz = 1 if kind == 1 else 3
    w = (torch.randn(Cout, Cin, ksz, ksz, generator=g) / (Cin * ksz * ksz) ** 0.5).half().float().to(dev)
    cx = Cin if kind in (0, 1, 2) else Cout
    x = torch.randn(N, cx, H, W, generator=g).half().float().to(dev)      # fp16- (hence tf32-) representable
    ref = _ref(kind, x, w)
    add = None
    if kind in (0, 1) and io != "h->f":
        add = torch.randn(ref.shape, generator=g).half().to(dev)
        ref = ref + add.double()
    y = ops.conv2d_nhwc_typed(kind, nhwc(x).to(tin), w, addend=None if add is None else nhwc(add).to(tout),
                              out_dtype=tout)
    torch.cuda.synchronize()
    assert y.dtype == tout
    e = rel_err(nchw(y.float()), ref)
    print(f"conv {io} kind={kind} N={N} {H}x{W} {Cin}->{Cout}: rel_err={e:.3e}")
    assert e < (1e-3 if tout == torch.float16 else 1e-5)


@pytest.mark.parametrize("name", ["two_level_attn16", "three_level_attn8"])
def test_unet_forward_fp16_matches_oracle(dev, name):
    from test_gpu_unet import ARCHS
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import random_state_dict, tiny_arch
    from oracle import ddpm_ref
    arch = tiny_arch(**ARCHS[name])
    sd = random_state_dict(arch, seed=1234, perturb_norm=0.1)
    net = B200UNet(arch, sd, device=dev)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 3, arch["resolution"], arch["resolution"], generator=g)
    t = torch.tensor(595.3636)
    with torch.no_grad():
        eref = ddpm_ref.RefUNet(arch, sd)(x, t)
    e16 = net.plan(3, half=True).forward(x.to(dev), float(t))
    e32 = net.plan(3, half=False).forward(x.to(dev), float(t))
    torch.cuda.synchronize()
    a, b = rel_err(e16.cpu(), eref), rel_err(e32.cpu(), eref)
    print(f"{name}: eps rel_err fp16 plan {a:.3e}, tf32 plan {b:.3e}, fp16 vs tf32 {rel_err(e16, e32):.3e}")
    assert torch.isfinite(e16).all() and a < 5e-3


@pytest.mark.parametrize("name", ["two_level_attn16", "three_level_attn8"])
def test_unet_jvp_vjp_fp16_match_oracle(dev, name):
    """Fused primal + k-tangent JVP and k-cotangent VJP on fp16 plans (rows range-scaled by a power of
    two inside the library) against the fp32 oracle (torch.func.jvp / autograd), plus the adjoint
    identity <J v, g> = <v, J^T g> between the two CUDA passes.  Tangents of very different magnitude
    (1e-3 .. 1e+2) in one batch: the scale is per pass, so each row keeps ~11 bits only if the rows of a
    pass are of similar size -- the power method's rows are (orthonormal V, U = J V)."""
    from test_gpu_unet import ARCHS
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import random_state_dict, tiny_arch
    from oracle import ddpm_ref
    arch = tiny_arch(**ARCHS[name])
    sd = random_state_dict(arch, seed=1234, perturb_norm=0.1)
    net = B200UNet(arch, sd, device=dev)
    net.jac_half = True
    ref = ddpm_ref.RefUNet(arch, sd)
    R = arch["resolution"]
    g = torch.Generator().manual_seed(1)
    k = 3
    x = torch.randn(1, 3, R, R, generator=g)
    t = torch.tensor(595.3636)
    for scale in (1.0, 1e-3, 50.0):
        V = torch.randn(k, 3, R, R, generator=g) * scale
        G = torch.randn(k, 3, R, R, generator=g) * scale
        eps, deps = net.jvp(x.to(dev), t, V.to(dev))
        gx = net.vjp(k, G.to(dev))
        torch.cuda.synchronize()
        assert net.plan(1, k, k).half
        ref_d, ref_g = [], []
        for j in range(k):
            _, d = torch.func.jvp(lambda xx: ref(xx, t), (x,), (V[j:j + 1],))
            ref_d.append(d)
            xg = x.clone().requires_grad_(True)
            (ref(xg, t) * G[j:j + 1]).sum().backward()
            ref_g.append(xg.grad)
        ej = rel_err(deps.cpu(), torch.cat(ref_d))
        ev = rel_err(gx.cpu(), torch.cat(ref_g))
        lhs = (deps.double() * G.to(dev).double()).sum()
        rhs = (V.to(dev).double() * gx.double()).sum()
        # random v, g are nearly orthogonal to J^T g, J v (|<Jv,g>| ~ 1e-3..1e-2 of the norms), so the gap is
        # measured against |Jv| |g| (profiles/diag_fp16_jac.py: the tf32 programs give the same numbers)
        adj = abs(float(lhs - rhs)) / float(deps.double().norm() * G.double().norm())
        print(f"{name} scale {scale:g}: JVP rel_err {ej:.3e}, VJP rel_err {ev:.3e}, adjoint gap {adj:.2e} of |Jv||g|")
        assert ej < 5e-3 and ev < 5e-3 and adj < 1e-4


def test_p2_forward_fp16_matches_oracle(dev):
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import random_state_dict, tiny_p2_arch
    from oracle import p2_ref
    arch = tiny_p2_arch(resolution=32, ch_mult=(1, 2), attn_resolutions=(16,), num_res_blocks=1)
    sd = random_state_dict(arch, seed=4321, perturb_norm=0.1)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 32, 32, generator=g)
    net = B200UNet(arch, sd, device=dev)
    e16 = net.plan(2, half=True).forward(x.to(dev), 198.1818).cpu()
    with torch.no_grad():
        ref = p2_ref.RefP2UNet(arch, sd)(x, torch.tensor(198.1818))
    print(f"P2 eps rel_err fp16 plan {rel_err(e16, ref):.3e}")
    assert rel_err(e16, ref) < 5e-3


def test_full256_forward_fp16_matches_reference(dev, golden_dir):
    """The 113.7 M-parameter DDPM-256 U-Net at 256 x 256 (every layer shape of the real configuration:
    CTA-pair halo convs with fused shortcuts, split-K layers, fused statistics) in the fp16 plan
    against eps of the UNMODIFIED reference (tests/golden/make_golden_full.py)."""
    from test_gpu_full256 import _inputs
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import DDPM256, P2_256, random_state_dict
    g = torch.load(os.path.join(golden_dir, "full256_ddpm.pt"), weights_only=False)
    _, xt, _ = _inputs(g["input_seed"])
    net = B200UNet(DDPM256, random_state_dict(DDPM256, seed=g["weights_seed"]), device=dev)
    ref = g["eps"].float()
    e1 = net.plan(1, half=True).forward(xt.to(dev), float(g["t"])).cpu()
    e3 = net.plan(3, half=True).forward(torch.cat([xt + 0.3, xt, -xt]).to(dev), float(g["t"]))[1:2].cpu()
    print(f"DDPM-256 eps (fp16 plan) vs reference: B=1 {rel_err(e1, ref):.3e}, row of B=3 {rel_err(e3, ref):.3e}")
    assert rel_err(e1, ref) < 5e-3 and rel_err(e3, ref) < 5e-3
    del net
    torch.cuda.empty_cache()
    g2 = torch.load(os.path.join(golden_dir, "full256_p2.pt"), weights_only=False)
    net2 = B200UNet(P2_256, random_state_dict(P2_256, seed=g2["weights_seed"]), device=dev)
    e = net2.plan(1, half=True).forward(xt.to(dev), float(g2["t"])).cpu()
    print(f"P2-256 eps (fp16 plan) vs reference: {rel_err(e, g2['eps'].float()):.3e}")
    assert rel_err(e, g2["eps"].float()) < 5e-3


def test_final_stage_fp16_psnr_vs_reference_driver(dev, golden_dir, tmp_path):
    """north_star bar for the fp16 programs: the 59-step final stage at 256 x 256 from the reference's
    x_t, its -vT.pt and its eta = 1 noise stream, run on fp16 plans, against the images the
    UNMODIFIED reference driver produced (driver_full256.pt): PSNR >= 40 dB."""
    import types
    from loco_edit_b200.edit import EditUncondDiffusion
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import DDPM256, random_state_dict
    g = torch.load(os.path.join(golden_dir, "driver_full256.pt"), weights_only=False)
    d = g["xt"].numel()
    torch.manual_seed(g["seed"])
    torch.randn(d, 5); torch.randn(d, 5)                    # the two V0 draws precede the noise draws
    noises = [torch.randn(5, 3, 256, 256) for _ in range(20)]
    net = B200UNet(DDPM256, random_state_dict(DDPM256, seed=g["weights_seed"]), device=dev)
    net.fwd_half = True
    args = types.SimpleNamespace(
        device=dev, dtype=torch.float32, seed=11, model_name="CelebA_HQ_HF", dataset_name="CelebA_HQ_mask",
        image_size=256, for_steps=100, inv_steps=100, edit_t=0.6, performance_boosting_t=0.2,
        x_space_guidance_edit_step=1.0, x_space_guidance_scale=0.5, x_space_guidance_num_step=16,
        result_folder=str(tmp_path), sample_idx=7, choose_sem="hair", mask_index=0, sampling_mode=False,
        vT_path="", vT1_path="", verbose=False, save_images=False, noise_schedule=None)
    e = EditUncondDiffusion(args, unet=net)
    vref = [v for n, v in g["files"].items() if n.endswith("pc_000-vT.pt")][0]
    batch = e.build_edit_batch(g["xt"].to(dev), vref[0].to(dev), 2)
    e.noise_fn = lambda i, x: noises[i - 79].to(dev)
    img = e.DDIMforwardsteps(batch, t_start_idx=40, t_end_idx=-1, save_image=False, performance_boosting=True).cpu()
    ref = g["finals"][0].float()
    mse = float(((img.double() - ref.double()) ** 2).mean())
    p = 10 * math.log10(4.0 / mse)
    print(f"final stage on fp16 plans vs the reference driver's images: PSNR {p:.1f} dB")
    assert net.plan(5).half and p >= 40.0
