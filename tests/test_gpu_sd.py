"""GPU parity of the latent-space (Stable-Diffusion) twin of the hot path: the VAE-decoder Jacobian
 and the `EditStableDiffusion` logic around it, against the UNMODIFIED reference class run
on the same stand-in networks (tests/golden/make_golden_sd.py -> sd_tiny.pt) and against the CPU oracle.

Tolerances: north_star (singular values 1e-3 relative, principal angles < 1 degree); 5e-3 relative L2 for
fields (10-bit operand mantissas on the tensor cores, as in the U-Net tests); guidance modes that are
differences of two predictions are measured against |guided eps| (see test_gpu_t2i.py)."""
import os
import types

import pytest
import torch

from gpu_util import principal_angles_deg, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "sd_tiny.pt"), weights_only=False)


@pytest.fixture(scope="module")
def nets(dev, golden):
    from loco_edit_b200.t2i import TextB200UNet
    from loco_edit_b200.unet import B200UNet, B200VAEDecoder
    from loco_edit_b200.weights import random_state_dict
    g = golden
    sd = random_state_dict(g["arch"], seed=1234, perturb_norm=0.1)
    vsd = random_state_dict(g["vae_arch"], seed=4321, perturb_norm=0.1)
    return TextB200UNet(B200UNet(g["arch"], sd, device=dev)), B200VAEDecoder(g["vae_arch"], vsd, device=dev), sd, vsd


@pytest.fixture(scope="module")
def setup(dev, golden, nets, tmp_path_factory):
    from loco_edit_b200.sd import EditStableDiffusion
    from loco_edit_b200.t2i import synthetic_prompt_embedding
    g = golden
    unet, vae, _, _ = nets
    embs = [synthetic_prompt_embedding(p, g["ntok"], g["dim"]) for p in g["prompts"]]
    args = types.SimpleNamespace(device=dev, dtype=torch.float32, seed=3, for_steps=100, edit_t=0.4,
                                 guidance_scale=g["g"], guidance_scale_edit=g["g_edit"],
                                 x_space_guidance_edit_step=1.0, x_space_guidance_scale=0.5, x_space_guidance_num_step=4,
                                 result_folder=str(tmp_path_factory.mktemp("sd")), for_prompt="a photo of a dog",
                                 edit_prompt="a dog with glasses", tilda_v_score_type="null+(for-null)+(edit-null)")
    return g, EditStableDiffusion(args, unet, vae, *embs), embs


@pytest.mark.parametrize("half", [False, True])
def test_vae_decoder_forward_jvp_vjp_match_oracle(dev, golden, nets, half):
    """decode(z), J_dec dZ and J_dec^T G of the CUDA decoder program against the oracle restatement
    (torch.func.jvp / autograd of oracle/vae_ref.py), for the tf32 and the fp16 programs."""
    from oracle import vae_ref
    _, vae, _, vsd = nets
    g = golden
    k = 3
    gen = torch.Generator().manual_seed(11)
    z = g["z2"][:1]
    dZ = torch.randn(k, *z.shape[1:], generator=gen)
    f = lambda zz: vae_ref.decoder_forward(vsd, g["vae_arch"], zz)
    x_ref = f(z)
    dX_ref = torch.cat([torch.func.jvp(f, (z,), (dZ[i:i + 1],))[1] for i in range(k)], 0)
    G = torch.randn(k, *x_ref.shape[1:], generator=gen)
    zr = z.clone().requires_grad_(True)
    xr = f(zr)
    gz_ref = torch.cat([torch.autograd.grad(xr, zr, G[i:i + 1], retain_graph=True)[0] for i in range(k)], 0)
    plan = vae.plan(1, k, k, half=half)
    out = plan.forward(torch.cat([z, dZ], 0).to(dev).contiguous(), 0.0)
    gz = plan.vjp(G.to(dev).contiguous())
    torch.cuda.synchronize()
    e_x, e_dx, e_gz = rel_err(out[:1].cpu(), x_ref), rel_err(out[1:].cpu(), dX_ref), rel_err(gz.cpu(), gz_ref)
    # adjoint identity on the CUDA side alone: <J dZ, G> == <dZ, J^T G>
    lhs = (out[1:].double().cpu() * G.double()).sum(dim=(1, 2, 3))
    rhs = (dZ.double() * gz.double().cpu()).sum(dim=(1, 2, 3))
    adj = float(((lhs - rhs).abs() / (lhs.abs() + 1e-12)).max())
    print(f"VAE decoder half={half}: decode {e_x:.2e}, JVP {e_dx:.2e}, VJP {e_gz:.2e}, adjoint identity {adj:.2e}")
    assert out.shape == (1 + k, 3, 32, 32) and gz.shape == (k, 4, 16, 16)
    assert e_x < 5e-3 and e_dx < 5e-3 and e_gz < 5e-3 and adj < 1e-2
    # batch decode (forward-only program) against the reference-side stand-in output
    dec = vae.decode(g["z2"].to(dev)).cpu()
    assert rel_err(dec, g["decode"]) < 5e-3


def test_latent_unet_four_channel_edges(dev, golden, nets):
    """eps(z, t, prompt) of the 4-channel U-Net (padded tensor-core edge convolutions) against the oracle."""
    from loco_edit_b200.t2i import synthetic_prompt_embedding
    from oracle import ddpm_ref
    unet, _, sd, _ = nets
    g = golden
    emb = synthetic_prompt_embedding(g["prompts"][0], g["ntok"], g["dim"])
    ref = ddpm_ref.unet_forward(sd, g["arch"], g["z2"], g["t"], ctx=emb[0])
    out = unet.eps(g["z2"].to(dev), float(g["t"]), emb).cpu()
    print(f"4-channel U-Net eps: rel err {rel_err(out, ref):.2e}")
    assert out.shape == ref.shape and rel_err(out, ref) < 5e-3


def test_sd_scheduler_matches_reference_monkey_patch(setup):
    g, e, _ = setup
    e.scheduler.set_timesteps(100, device="cpu")
    assert torch.equal(e.scheduler.timesteps.cpu(), g["timesteps"])
    assert torch.equal(e.scheduler.alphas_cumprod.cpu(), g["alphas_cumprod"])


def test_sd_guidance_and_x0_match_reference(setup, dev):
    g, e, embs = setup
    z2, t = g["z2"].to(dev), float(g["t"])
    scale = float(g["cfg"]["null+(for-null)"].norm())
    for mode, ref in g["cfg"].items():
        out = e._classifer_free_guidance(z2, t, *embs, mode=mode, do_classifier_free_guidance=True).cpu()
        err = float((out - ref).norm())
        print(f"SD CFG mode {mode}: rel_err {rel_err(out, ref):.3e}, error / |guided eps| {err / scale:.3e}")
        # yardstick: the larger of the field's own norm and |guided eps| (difference modes cancel)
        assert out.shape == ref.shape and min(rel_err(out, ref), err / scale) < 5e-3
    off = e._classifer_free_guidance(z2, t, *embs, mode="null+(for-null)", do_classifier_free_guidance=False).cpu()
    assert rel_err(off, g["cfg_off"]) < 5e-3
    zt = g["zt"].to(dev)
    a = e.get_x0(zt, t, g["t_idx"], *embs, mask=g["mask"].to(dev), mode="null+(for-null)").cpu()
    b = e.get_x0(zt, t, g["t_idx"], *embs, mask=None, mode="null+(for-null)+(edit-null)", flatten=True).cpu()
    print(f"pixel-space x0_hat masked {rel_err(a, g['x0_masked']):.3e}, flat {rel_err(b, g['x0_flat']):.3e}")
    assert a.shape == g["x0_masked"].shape and rel_err(a, g["x0_masked"]) < 1e-2
    assert b.shape == g["x0_flat"].shape and rel_err(b, g["x0_flat"]) < 1e-2


def test_latent_power_method_through_the_decoder_matches_reference(setup, dev):
    """local_encoder_decoder_pullback_zt (src/modules/edit.py:830-915): J through U-Net, PMP and VAE decoder,
    from the reference's V0 draw (seed 7), mask and ~mask over the decoded image."""
    g, e, embs = setup
    zt, t = g["zt"].to(dev), float(g["t"])
    torch.manual_seed(7)
    v0, _ = torch.linalg.qr(torch.randn(zt.numel(), 2))
    for (mode, mname, n_iter), ref in g["pullback"].items():
        m = g["mask"] if mname == "mask" else ~g["mask"]
        u, s, vT = e.local_encoder_decoder_pullback_zt(zt, t, g["t_idx"], *embs, pca_rank=2, min_iter=10 ** 6,
                                                       max_iter=n_iter, mask=m.to(dev), mode=mode, v0=v0.T.contiguous())
        torch.cuda.synchronize()
        srel = float(((s.cpu() - ref["s"]).abs() / ref["s"]).max())
        ang = float(principal_angles_deg(vT, ref["vT"]).max())
        uang = float(principal_angles_deg(u.T, ref["u"].T).max())
        print(f"latent power method {mode} {mname} N={n_iter}: s rel {srel:.2e}, vT {ang:.3f} deg, u {uang:.3f} deg")
        assert u.shape == ref["u"].shape and vT.shape == ref["vT"].shape
        assert srel < 1e-3 and ang < 1.0 and uang < 1.0


def test_latent_text_supervised_direction_matches_reference(setup, dev):
    g, e, embs = setup
    zt, t = g["zt"].to(dev), float(g["t"])
    v = e.get_delta_zt_via_grad(zt, t, g["t_idx"], *embs, mask=g["mask"].to(dev), mode="null+(for-null)+(edit-null)").cpu()
    ref = g["delta_masked"]
    c = float((v.double() * ref.double()).sum() / (v.double().norm() * ref.double().norm()))
    a = float(torch.rad2deg(torch.acos(torch.tensor(min(1.0, abs(c))))))
    print(f"get_delta_zt_via_grad: angle to the reference direction {a:.3f} deg (cos {c:+.6f}), |v| {float(v.norm()):.6f}")
    assert v.shape == ref.shape and c > 0 and a < 1.0 and abs(float(v.norm()) - 1) < 1e-4


def test_latent_ddim_loop_and_decode_match_reference(setup, dev):
    g, e, embs = setup
    kw = dict(for_prompt_emb=embs[0], edit_prompt_emb=embs[1], null_prompt_emb=embs[2])
    lat, u8 = e.DDIMforwardsteps(g["z2"].to(dev), t_start_idx=90, t_end_idx=-1, mode="null+(for-null)", **kw)
    ref = g["ddim_final"]
    diff = (u8.cpu().int() - ref["u8"].int()).abs()
    print(f"9 guided latent DDIM steps + decode: latents {rel_err(lat.cpu(), ref['latents']):.2e}, uint8 images differ by at most "
          f"{int(diff.max())} level(s), mean {float(diff.float().mean()):.4f}")
    assert u8.shape == ref["u8"].shape and u8.dtype == torch.uint8
    assert rel_err(lat.cpu(), ref["latents"]) < 5e-3 and int(diff.max()) <= 2 and float(diff.float().mean()) < 0.1
    zt, t, i = e.DDIMforwardsteps(g["z2"].to(dev), t_start_idx=88, t_end_idx=92, mode="null+(for-null)+(edit-null)", **kw)
    assert i == g["ddim_mid"]["idx"] == 92 and float(t) == float(g["ddim_mid"]["t"])
    assert rel_err(zt.cpu(), g["ddim_mid"]["zt"]) < 5e-3


def test_sd_driver_writes_the_reference_files(setup, dev):
    """run_edit_null_space_projection_zt (src/modules/edit.py:918-1043): basis file names / shapes, projected
    direction orthogonal to the null basis, decoded uint8 images."""
    from loco_edit_b200.masks import save_masks
    g, e, embs = setup
    save_masks(e.result_folder, g["mask"][:1])
    e.zT = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(9))
    orig = e.local_encoder_decoder_pullback_zt
    e.local_encoder_decoder_pullback_zt = lambda *a, **k: orig(*a, **dict(k, min_iter=0, max_iter=2))
    try:
        r = e.run_edit_null_space_projection_zt(op="mid", block_idx=0, vis_num=2, mask_index=0, vis_num_pc=1, pca_rank=2,
                                                null_space_projection=True, pca_rank_null=3)
    finally:
        e.local_encoder_decoder_pullback_zt = orig
    d = os.path.join(e.result_folder, "basis", f"local_basis-{e.edit_t}T-pca-rank-2-select-mask0")
    assert sorted(os.listdir(d)) == ["u-modify.pt", "u-null-null_space_rank_3.pt", "vT-modify.pt", "vT-null-null_space_rank_3.pt"]
    assert torch.load(os.path.join(d, "vT-modify.pt")).shape == (2, 1024)
    assert torch.load(os.path.join(d, "u-modify.pt")).shape == (int(g["mask"].sum()), 2)
    assert torch.load(os.path.join(d, "vT-null-null_space_rank_3.pt")).shape == (3, 1024)
    vT, vn = r["vT"].double().cpu(), r["vT_null"].double().cpu()
    assert float((vT @ vn.T).abs().max()) < 1e-4 and float((vT.norm(dim=1) - 1).abs().max()) < 1e-5
    assert r["images"].shape == (5, 32, 32, 3) and r["images"].dtype == torch.uint8


def test_sd_semantic_driver(setup, dev):
    """run_edit_null_space_projection_zt_semantic (src/modules/edit.py:1045-1175): the text-supervised direction
    projected off the null basis of ~mask; file names of the reference."""
    from loco_edit_b200.masks import save_masks
    g, e, embs = setup
    save_masks(e.result_folder, g["mask"][:1])
    e.zT = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(9))
    e.edit_prompt = "a dog with glasses"
    orig = e.local_encoder_decoder_pullback_zt
    e.local_encoder_decoder_pullback_zt = lambda *a, **k: orig(*a, **dict(k, min_iter=0, max_iter=2))
    try:
        r = e.run_edit_null_space_projection_zt_semantic(op="mid", block_idx=0, vis_num=1, mask_index=0, vis_num_pc=1, pca_rank=1,
                                                         null_space_projection=True, pca_rank_null=2)
    finally:
        e.local_encoder_decoder_pullback_zt = orig
    d = os.path.join(e.result_folder, "basis", f'local_basis-{e.edit_t}T-"a dog with glasses"-pca-rank-1-select-mask0')
    assert sorted(os.listdir(d)) == ["u-null-null_space_rank_2.pt", "vT-modify.pt", "vT-null-null_space_rank_2.pt"]
    assert torch.load(os.path.join(d, "vT-modify.pt")).shape == (1, 1024)
    vT, vn = r["vT"].double().cpu(), r["vT_null"].double().cpu()
    assert vT.shape == (1, 1024) and float((vT @ vn.T).abs().max()) < 1e-4 and abs(float(vT.norm()) - 1) < 1e-5
    assert r["images"].shape == (3, 32, 32, 3) and r["images"].dtype == torch.uint8


def test_sd_shaped_vae_decoder_full_size(dev):
    """The decoder at the Stable Diffusion 1.x size (latent 4 x 64 x 64 -> 3 x 512 x 512, mid-block attention over
    4096 tokens on the CUDA-core path): decode against the CPU oracle; JVP linearity and the adjoint identity
    <J dz, g> = <dz, J^T g> on the CUDA side (size-independent properties)."""
    from loco_edit_b200.unet import B200VAEDecoder
    from loco_edit_b200.weights import SD_VAE_DECODER, random_state_dict
    from oracle import vae_ref
    vsd = random_state_dict(SD_VAE_DECODER, seed=4321)
    vae = B200VAEDecoder(SD_VAE_DECODER, vsd, device=dev)
    gen = torch.Generator().manual_seed(3)
    z = torch.randn(1, 4, 64, 64, generator=gen)
    with torch.no_grad():
        ref = vae_ref.decoder_forward(vsd, SD_VAE_DECODER, z)
    out = vae.decode(z.to(dev)).cpu()
    print(f"SD-shaped decoder 64^2 -> 512^2: decode rel err {rel_err(out, ref):.2e}")
    assert out.shape == (1, 3, 512, 512) and rel_err(out, ref) < 5e-3
    k = 2
    dZ = torch.randn(k, 4, 64, 64, generator=gen).to(dev)
    G = torch.randn(k, 3, 512, 512, generator=gen).to(dev)
    x, dX = vae.jvp(z.to(dev), dZ)
    gz = vae.vjp(k, G)
    torch.cuda.synchronize()
    assert rel_err(x.cpu(), ref) < 5e-3
    lhs = (dX.double() * G.double()).sum(dim=(1, 2, 3))
    rhs = (dZ.double() * gz.double()).sum(dim=(1, 2, 3))
    adj = float(((lhs - rhs).abs() / (dX.double().flatten(1).norm(dim=1) * G.double().flatten(1).norm(dim=1))).max())
    # linearity: the tangent of (dz_0 + 2 dz_1) is dX_0 + 2 dX_1
    _, dX2 = vae.jvp(z.to(dev), torch.stack([dZ[0] + 2 * dZ[1], dZ[1]], 0))
    lin = rel_err(dX2[0], dX[0] + 2 * dX[1])
    print(f"adjoint identity (relative to |J dz||g|) {adj:.2e}, linearity {lin:.2e}")
    assert adj < 1e-3 and lin < 5e-3
    vae.release_plans()
