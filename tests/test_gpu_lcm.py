"""GPU parity of the latent-consistency twin (`EditLatentConsistency`, src/modules/edit.py:42-480) against the UNMODIFIED
reference class run on stand-ins for everything diffusers provides (tests/golden/make_golden_lcm.py -> lcm_tiny.pt).
Tolerances as in test_gpu_sd.py."""
import os
import types

import pytest
import torch

from gpu_util import principal_angles_deg, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(golden_dir, tmp_path_factory):
    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    from loco_edit_b200.sd import EditLatentConsistency, LCMB200UNet
    from loco_edit_b200.t2i import synthetic_prompt_embedding
    from loco_edit_b200.unet import B200UNet, B200VAEDecoder
    from loco_edit_b200.weights import random_state_dict
    g = torch.load(os.path.join(golden_dir, "lcm_tiny.pt"), weights_only=False)
    unet = LCMB200UNet(B200UNet(g["arch"], random_state_dict(g["arch"], seed=1234, perturb_norm=0.1), device=dev), w_dim=g["wdim"])
    vae = B200VAEDecoder(g["vae_arch"], random_state_dict(g["vae_arch"], seed=4321, perturb_norm=0.1), device=dev)
    args = types.SimpleNamespace(device=dev, dtype=torch.float32, seed=3, for_steps=100, guidance_scale=g["guidance"],
                                 num_inference_steps=g["steps"], edit_t_idx=g["t_idx"], time_cond_proj_dim=g["wdim"],
                                 x_space_guidance_edit_step=1.0, x_space_guidance_scale=0.5, x_space_guidance_num_step=4,
                                 result_folder=str(tmp_path_factory.mktemp("lcm")), for_prompt=g["prompts"][0],
                                 edit_prompt=g["prompts"][1])
    e = EditLatentConsistency(args, unet, vae, encode_prompt=lambda p: synthetic_prompt_embedding(p, g["ntok"], g["dim"]))
    return g, e, dev


def test_lcm_scheduler_grid(setup):
    g, e, _ = setup
    assert torch.equal(e.scheduler.timesteps.cpu(), g["timesteps"])


def test_lcm_x0_matches_reference(setup):
    g, e, dev = setup
    t = int(g["t"])
    a = e.get_x0(g["zt"].to(dev), g["prompts"][0], t, g["t_idx"], mask=g["mask"].to(dev)).cpu()
    b = e.get_x0(g["z2"].to(dev), g["prompts"][1], t, g["t_idx"], flatten=True).cpu()
    print(f"LCM pixel-space x0_hat masked {rel_err(a, g['x0_masked']):.3e}, flat {rel_err(b, g['x0_flat']):.3e}")
    assert a.shape == g["x0_masked"].shape and rel_err(a, g["x0_masked"]) < 1e-2
    assert b.shape == g["x0_flat"].shape and rel_err(b, g["x0_flat"]) < 1e-2


def test_lcm_power_method_matches_reference(setup):
    g, e, dev = setup
    zt, t = g["zt"].to(dev), int(g["t"])
    torch.manual_seed(7)
    v0, _ = torch.linalg.qr(torch.randn(zt.numel(), 2))
    for (mname, n_iter), ref in g["pullback"].items():
        m = g["mask"] if mname == "mask" else ~g["mask"]
        u, s, vT = e.local_encoder_decoder_pullback_zt(zt, t, g["t_idx"], g["prompts"][0], pca_rank=2, min_iter=10 ** 6,
                                                       max_iter=n_iter, mask=m.to(dev), v0=v0.T.contiguous())
        torch.cuda.synchronize()
        srel = float(((s.cpu() - ref["s"]).abs() / ref["s"]).max())
        ang = float(principal_angles_deg(vT, ref["vT"]).max())
        uang = float(principal_angles_deg(u.T, ref["u"].T).max())
        print(f"LCM power method {mname} N={n_iter}: s rel {srel:.2e}, vT {ang:.3f} deg, u {uang:.3f} deg")
        assert u.shape == ref["u"].shape and vT.shape == ref["vT"].shape
        assert srel < 1e-3 and ang < 1.0 and uang < 1.0


def test_lcm_text_supervised_direction_matches_reference(setup):
    g, e, dev = setup
    v = e.get_delta_zt_via_grad(g["zt"].to(dev), int(g["t"]), g["t_idx"], g["prompts"][0], g["prompts"][1],
                                mask=g["mask"].to(dev)).cpu()
    ref = g["delta_masked"]
    c = float((v.double() * ref.double()).sum() / (v.double().norm() * ref.double().norm()))
    a = float(torch.rad2deg(torch.acos(torch.tensor(min(1.0, abs(c))))))
    print(f"LCM get_delta_zt_via_grad: angle to the reference direction {a:.3f} deg (cos {c:+.6f})")
    assert v.shape == ref.shape and c > 0 and a < 1.0 and abs(float(v.norm()) - 1) < 1e-4


def test_lcm_sampler_matches_reference(setup):
    """LCMforwardsteps with the reference run's noise draws injected: four consistency steps + decode."""
    g, e, dev = setup
    e.step_noise = g["noise_list"]
    try:
        lat, u8 = e.LCMforwardsteps(g["z2"].to(dev), g["prompts"][0], t_start_idx=0, t_end_idx=-1)
        ref = g["lcm_final"]
        diff = (u8.cpu().int() - ref["u8"].int()).abs()
        print(f"LCM 4-step sampler: latents {rel_err(lat.cpu(), ref['latents']):.2e}, uint8 images differ by at most "
              f"{int(diff.max())} level(s), mean {float(diff.float().mean()):.4f}")
        assert u8.shape == ref["u8"].shape and rel_err(lat.cpu(), ref["latents"]) < 1e-2
        assert int(diff.max()) <= 3 and float(diff.float().mean()) < 0.2
        zt, t, i = e.LCMforwardsteps(g["z2"].to(dev), g["prompts"][0], t_start_idx=0, t_end_idx=2)
        assert i == g["lcm_mid"]["idx"] == 2 and int(t) == int(g["lcm_mid"]["t"])
        assert rel_err(zt.cpu(), g["lcm_mid"]["zt"]) < 1e-2
    finally:
        e.step_noise = None


def test_lcm_driver(setup):
    from loco_edit_b200.masks import save_masks
    g, e, dev = setup
    save_masks(e.result_folder, g["mask"][:1])
    e.zT = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(9))
    orig = e.local_encoder_decoder_pullback_zt
    e.local_encoder_decoder_pullback_zt = lambda *a, **k: orig(*a, **dict(k, min_iter=0, max_iter=2))
    try:
        r = e.run_edit_null_space_projection_zt(op="mid", block_idx=0, vis_num=1, pca_rank=1, null_space_projection=True,
                                                pca_rank_null=2)
    finally:
        e.local_encoder_decoder_pullback_zt = orig
    vT, vn = r["vT"].double().cpu(), r["vT_null"].double().cpu()
    assert vT.shape == (1, 1024) and float((vT @ vn.T).abs().max()) < 1e-4 and abs(float(vT.norm()) - 1) < 1e-5
    assert r["images"].shape == (3, 32, 32, 3) and r["images"].dtype == torch.uint8
