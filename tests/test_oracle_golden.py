"""CPU: the oracle restatement against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  Tolerances are fp32 round-off of two implementations of the same
arithmetic on the same CPU (different op fusion / summation order only)."""
import json
import os

import pytest
import torch

from loco_edit_b200.weights import DDPM256, ddpm_param_shapes, random_state_dict
from oracle import ddpm_ref, pullback_ref


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_param_shapes_match_reference(golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "ddpm_param_shapes.json")))
    mine = ddpm_param_shapes(DDPM256)
    assert sorted(mine.keys()) == sorted(ref.keys())
    for k in ref:
        assert list(mine[k]) == ref[k], k
    n = 0
    for s in mine.values():
        c = 1
        for d in s:
            c *= d
        n += c
    assert n == 113673219


def test_unet_forward_matches_reference(golden_dir):
    g = _load(golden_dir, "unet_tiny.pt")
    sd = random_state_dict(g["arch"], seed=g["seed"], perturb_norm=g["perturb_norm"])
    with torch.no_grad():
        eps = ddpm_ref.unet_forward(sd, g["arch"], g["x"], g["t"])
    assert torch.allclose(eps, g["eps"], atol=2e-5, rtol=1e-5), float((eps - g["eps"]).abs().max())


def test_scheduler_matches_reference(golden_dir):
    g = _load(golden_dir, "scheduler.pt")
    s = pullback_ref.RefScheduler()
    s.set_timesteps(100)
    assert torch.equal(s.timesteps, g["timesteps"])
    assert torch.equal(s.timesteps_next, g["timesteps_next"])
    assert torch.equal(s.alphas_cumprod, g["alphas_cumprod"])
    for idx, o in g["steps"].items():
        t = s.timesteps[idx]
        xn, x0 = s.step(g["et"], t, g["xt"], eta=0)
        assert torch.equal(xn, o["eta0"]) and torch.equal(x0, o["x0"])
        xn1, _ = s.step(g["et"], t, g["xt"], eta=1, noise=o["noise"])
        assert torch.equal(xn1, o["eta1"])
    s.set_timesteps(100, is_inversion=True)
    assert torch.equal(s.timesteps, g["inv_timesteps"])
    assert torch.equal(s.timesteps_next, g["inv_timesteps_next"])
    for idx, o in g["inv_steps"].items():
        xn, _ = s.step(g["et"], s.timesteps[idx], g["xt"], eta=0)
        assert torch.equal(xn, o)


def _principal_angle_deg(A, B):
    qa, _ = torch.linalg.qr(A.double().T)
    qb, _ = torch.linalg.qr(B.double().T)
    sv = torch.linalg.svdvals(qa.T @ qb).clamp(max=1.0)
    return float(torch.rad2deg(torch.acos(sv.min())))


@pytest.mark.parametrize("case", ["mask_k2", "notmask_k3", "nomask_k2", "noise_k2"])
def test_local_basis_matches_reference(golden_dir, case):
    g = _load(golden_dir, "pullback_tiny.pt")
    sd = random_state_dict(g["arch"], seed=g["seed"], perturb_norm=g["perturb_norm"])
    unet = ddpm_ref.RefUNet(g["arch"], sd)
    sched = pullback_ref.RefScheduler()
    sched.set_timesteps(100)
    kw = {"mask_k2": dict(mask=g["mask"], k=2), "notmask_k3": dict(mask=~g["mask"], k=3),
          "nomask_k2": dict(mask=None, k=2), "noise_k2": dict(mask=g["mask"], k=2, noise=True)}[case]
    k = kw.pop("k")
    d = g["xt"].numel()
    torch.manual_seed(g["v0_seed"])                      # modules/edit.py:2435-2437
    v0, _ = torch.linalg.qr(torch.randn(d, k))
    for n_iter in (1, 3):
        u, s, vT = pullback_ref.local_basis(unet, sched, g["xt"], g["t"], v0.T, n_iter, **kw)
        ref = g["cases"][case][n_iter]
        assert torch.allclose(s, ref["s"], rtol=1e-4), (s, ref["s"])
        assert _principal_angle_deg(vT, ref["vT"]) < 0.05
        # rows equal up to sign
        dots = (vT * ref["vT"]).sum(1).abs()
        assert float((1 - dots).abs().max()) < 1e-3, dots
        uref = ref["u"].reshape(ref["u"].shape[0], -1)
        assert u.shape == uref.shape
        assert torch.allclose(u, uref, atol=1e-4 * float(uref.abs().max()) + 1e-6)


def test_driver_pipeline_matches_reference_driver(golden_dir):
    """The oracle's restatement of run_edit_null_space_projection (inversion -> forward -> bases ->
    projection -> edit batch -> eta=1 DDIM) against the unmodified reference driver run
    (tests/golden/make_golden.py section 5), replaying the reference's RNG stream (seed 11)."""
    import math
    from loco_edit_b200.weights import tiny_arch
    g = _load(golden_dir, "driver_tiny.pt")
    arch = tiny_arch(resolution=32, ch_mult=(1, 2), attn_resolutions=(16,), num_res_blocks=1)
    unet = ddpm_ref.RefUNet(arch, random_state_dict(arch, seed=1234, perturb_norm=0.1))
    sched = pullback_ref.RefScheduler()
    torch.manual_seed(g["seed"])
    d = g["x0"].numel()
    v0a, _ = torch.linalg.qr(torch.randn(d, 2))
    v0b, _ = torch.linalg.qr(torch.randn(d, 3))
    noises = [torch.randn(5, 3, 32, 32) for _ in range(40)]
    xT = pullback_ref.ddim_inversion(unet, sched, g["x0"])
    xt, t, idx = pullback_ref.ddim_forward(unet, sched, xT, 0, g["edit_t_idx"])
    assert idx == 40
    _, _, vm = pullback_ref.local_basis(unet, sched, xt, t, v0a.T, 2, mask=g["mask"])
    _, _, vn = pullback_ref.local_basis(unet, sched, xt, t, v0b.T, 2, mask=~g["mask"])   # 2 iterations
    vT = pullback_ref.nullspace_project(vm, vn, 3)
    files = g["files"]
    base = "basis/local_basis-0.6T-select-mask-hair/"
    for mine, name in [(vm, "vT-modify-pca-rank-2.pt"), (vn, "vT-null-3.pt")]:
        # subspaces agree; individual rows of a near-degenerate cluster may rotate inside it
        # (flat spectrum of random-init weights), which the projection below is invariant to
        assert _principal_angle_deg(mine, files[base + name]) < 0.2, name
    for pc in range(2):
        ref_v = [v for k, v in files.items() if k.endswith("pc_%03d-vT.pt" % pc)][0]
        assert float(1 - (vT[pc] * ref_v[0]).sum().abs()) < 1e-3
        # edit with the reference's own direction (removes the sign ambiguity), replayed noise
        batch = pullback_ref.edit_batch(xt, ref_v[0], 0.5, 4, 2)
        nz = {g["boost_idx"] + i: noises[20 * pc + i] for i in range(20)}
        img = pullback_ref.ddim_forward(unet, sched, batch, 40, -1, boost_idx=g["boost_idx"], noises=nz)
        mse = float(((img - g["finals"][pc]) ** 2).mean())
        assert mse < 1e-8, mse      # identical arithmetic on the same CPU


# ---- P2 / guided-diffusion U-Net family (BASELINE config 2; tests/golden/make_golden_p2.py) ----
def test_p2_param_shapes_match_reference(golden_dir):
    from loco_edit_b200.weights import P2_256, p2_param_shapes
    ref = json.load(open(os.path.join(golden_dir, "p2_param_shapes.json")))
    mine = p2_param_shapes(P2_256)
    assert list(mine.keys()) == list(ref.keys())
    n = 0
    for k in ref:
        assert list(mine[k]) == ref[k], k
        c = 1
        for d in ref[k]:
            c *= d
        n += c
    assert n == 93563910


def test_p2_unet_forward_matches_reference(golden_dir):
    from oracle import p2_ref
    g = _load(golden_dir, "p2_tiny.pt")
    sd = random_state_dict(g["arch"], seed=g["seed"], perturb_norm=g["perturb_norm"])
    with torch.no_grad():
        eps = p2_ref.unet_forward(sd, g["arch"], g["x"], g["t"])
    assert torch.allclose(eps, g["eps"], atol=2e-5, rtol=1e-5), float((eps - g["eps"]).abs().max())


@pytest.mark.parametrize("case", ["mask_k3", "notmask_k2"])
def test_p2_local_basis_matches_reference(golden_dir, case):
    from oracle import p2_ref
    g = _load(golden_dir, "p2_pullback_tiny.pt")
    sd = random_state_dict(g["arch"], seed=g["seed"], perturb_norm=g["perturb_norm"])
    unet = p2_ref.RefP2UNet(g["arch"], sd)
    sched = pullback_ref.RefScheduler()
    sched.set_timesteps(100)
    mask, k = {"mask_k3": (g["mask"], 3), "notmask_k2": (~g["mask"], 2)}[case]
    d = g["xt"].numel()
    torch.manual_seed(g["v0_seed"])
    v0, _ = torch.linalg.qr(torch.randn(d, k))
    for n_iter in (1, 2):
        u, s, vT = pullback_ref.local_basis(unet, sched, g["xt"], g["t"], v0.T, n_iter, mask=mask)
        ref = g["cases"][case][n_iter]
        assert torch.allclose(s, ref["s"], rtol=1e-4), (s, ref["s"])
        assert _principal_angle_deg(vT, ref["vT"]) < 0.05
        dots = (vT * ref["vT"]).sum(1).abs()
        assert float((1 - dots).abs().max()) < 1e-3, dots


def test_full_size_forward_oracles_match_reference(golden_dir):
    """256 x 256, full-depth configurations (tests/golden/make_golden_full.py): one forward each."""
    from loco_edit_b200.weights import P2_256
    from oracle import p2_ref
    gen = torch.Generator().manual_seed(0)
    (0.5 * torch.randn(1, 3, 256, 256, generator=gen)).clamp(-1, 1)
    xt = torch.randn(1, 3, 256, 256, generator=gen)
    for name, arch, fwd in (("full256_ddpm.pt", DDPM256, ddpm_ref.unet_forward),
                            ("full256_p2.pt", P2_256, p2_ref.unet_forward)):
        g = _load(golden_dir, name)
        sd = random_state_dict(arch, seed=g["weights_seed"])
        with torch.no_grad():
            eps = fwd(sd, arch, xt, g["t"])
        ref = g["eps"].float()          # stored in fp16
        err = float((eps - ref).norm() / ref.norm())
        assert err < 1e-3, (name, err)


def test_latent_space_oracle_matches_reference_stable_diffusion_class(golden_dir):
    """oracle/vae_ref.py (pixel-space x0_hat through the VAE decoder, one pass of the latent-space power
    method) against the unmodified `EditStableDiffusion` run on the same stand-in networks
    (tests/golden/make_golden_sd.py)."""
    from loco_edit_b200.scheduler import YHCustomScheduler
    from loco_edit_b200.t2i import synthetic_prompt_embedding
    from oracle import vae_ref
    g = _load(golden_dir, "sd_tiny.pt")
    sd = random_state_dict(g["arch"], seed=1234, perturb_norm=0.1)
    vsd = random_state_dict(g["vae_arch"], seed=4321, perturb_norm=0.1)
    vae = vae_ref.RefVAE(g["vae_arch"], vsd)
    with torch.no_grad():
        assert torch.allclose(vae.decode(g["z2"]).sample, g["decode"], atol=1e-5)
    # the scheduler mirror holds Stable Diffusion's table, bit for bit
    import types
    sched = YHCustomScheduler(types.SimpleNamespace(device="cpu", dtype=torch.float32, t_max=999, noise_schedule="scaled_linear"),
                              device="cpu")
    assert torch.equal(sched.alphas_cumprod, g["alphas_cumprod"])
    sched.set_timesteps(100, device="cpu")
    assert torch.equal(sched.timesteps, g["timesteps"])
    embs = [synthetic_prompt_embedding(p, g["ntok"], g["dim"]) for p in g["prompts"]]
    G = g["g"]
    t = g["t"]
    at = float(g["alphas_cumprod"][int(float(t))])

    def eps_fn(z):      # "null+(for-null)": e_null + g (e_for - e_null), src/modules/edit.py:664-666
        ef = ddpm_ref.unet_forward(sd, g["arch"], z, t, ctx=embs[0][0])
        en = ddpm_ref.unet_forward(sd, g["arch"], z, t, ctx=embs[2][0])
        return en + G * (ef - en)

    with torch.no_grad():
        x0m = vae_ref.x0_hat_pixels(eps_fn, vae, g["zt"], at, g["mask"])
    assert torch.allclose(x0m, g["x0_masked"], atol=2e-4, rtol=1e-4), float((x0m - g["x0_masked"]).abs().max())
    torch.manual_seed(7)
    v0, _ = torch.linalg.qr(torch.randn(g["zt"].numel(), 2))
    u, s, vT = vae_ref.power_iteration_zt(eps_fn, vae, g["zt"], at, v0.T.contiguous(), g["mask"])
    ref = g["pullback"][("null+(for-null)", "mask", 1)]
    assert torch.allclose(s.sqrt(), ref["s"], rtol=1e-4), (s.sqrt(), ref["s"])
    assert _principal_angle_deg(vT, ref["vT"]) < 0.05


def test_vae_decoder_param_shapes_are_the_sd_vae():
    """The decoder half of the SD 1.x VAE has 49.49 M parameters (diffusers AutoencoderKL: 83.65 M in total,
    34.16 M of them in the encoder + quant_conv)."""
    from loco_edit_b200.weights import SD_VAE_DECODER, vae_decoder_param_shapes
    shapes = vae_decoder_param_shapes(SD_VAE_DECODER)
    n = 0
    for v in shapes.values():
        c = 1
        for d in v:
            c *= d
        n += c
    assert shapes["decoder.conv_in.weight"] == (512, 4, 3, 3) and shapes["decoder.conv_out.weight"] == (3, 128, 3, 3)
    assert shapes["decoder.up.1.block.0.nin_shortcut.weight"] == (256, 512, 1, 1)
    assert n == 49490199, n
