"""GPU parity of the T-LOCO (text-conditioned) Edit-class logic against the UNMODIFIED reference class
`EditDeepFloydIF` run on the same stand-in conditional U-Net (tests/golden/make_golden_t2i.py ->
t2i_tiny.pt; the network itself is a stand-in on both sides, see loco_edit_b200/t2i.py).

Tolerances: north_star (singular values 1e-3 relative, principal angles < 1 degree) and the 5e-3
relative L2 of the U-Net tests for fields; guidance modes that are DIFFERENCES of two predictions are
measured against guidance_scale * |eps| (the difference itself is a cancellation of two ~equal
fields, so its own norm is not a meaningful yardstick for 10-bit operands)."""
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


# two stand-in networks, each pinned by the unmodified reference class run on the same network:
#   "temb":  conditioning = pooled prompt embedding added to the timestep embedding (t2i_tiny.pt)
#   "cross": every AttnBlock cross-attends to the prompt tokens (t2i_cross_tiny.pt)
@pytest.fixture(scope="module", params=["temb", "cross"])
def setup(request, dev, golden_dir, tmp_path_factory):
    from loco_edit_b200.t2i import CondB200UNet, EditDeepFloydIF, TextB200UNet, synthetic_prompt_embedding
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import random_state_dict
    cross = request.param == "cross"
    g = torch.load(os.path.join(golden_dir, "t2i_cross_tiny.pt" if cross else "t2i_tiny.pt"), weights_only=False)
    sd = random_state_dict(g["arch"], seed=1234, perturb_norm=0.1)
    base = B200UNet(g["arch"], sd, device=dev)
    net = TextB200UNet(base) if cross else CondB200UNet(base, g["dim"])
    embs = [synthetic_prompt_embedding(p, g["ntok"], g["dim"]) for p in g["prompts"]]
    args = types.SimpleNamespace(device=dev, dtype=torch.float32, seed=3, for_steps=100, edit_t=0.4,
                                 guidance_scale=g["g"], guidance_scale_edit=g["g_edit"], image_size=32,
                                 x_space_guidance_edit_step=1.0, x_space_guidance_scale=0.5, x_space_guidance_num_step=4,
                                 result_folder=str(tmp_path_factory.mktemp("t2i")), for_prompt="a photo of a dog",
                                 edit_prompt="a dog with glasses", tilda_v_score_type="null+(for-null)+(edit-null)")
    e = EditDeepFloydIF(args, net, *embs)
    return g, e, embs


def test_if_scheduler_grid_matches_reference_monkey_patch(setup):
    """get_deepfloyd_if_scheduler (src/utils/utils.py:159-185): t_max = 990 grid, the scheduler's own table."""
    g, e, _ = setup
    e.scheduler.set_timesteps(100, device="cpu")
    assert torch.equal(e.scheduler.timesteps.cpu(), g["timesteps"])
    assert torch.equal(e.scheduler.alphas_cumprod.cpu(), g["alphas_cumprod"])
    assert e.edit_t_idx == int((g["timesteps"] - 400.0).abs().argmin())


def test_classifier_free_guidance_modes_match_reference(setup, dev):
    g, e, embs = setup
    x2, t = g["x2"].to(dev), float(g["t"])
    scale = g["g"] * float(g["cfg"]["null+(for-null)"][:, :3].norm()) / g["g"]          # |guided eps|
    for mode, ref in g["cfg"].items():
        out = e._classifer_free_guidance(x2, t, *embs, mode=mode, do_classifier_free_guidance=True).cpu()
        err = float((out - ref).norm())
        print(f"CFG mode {mode}: rel_err {rel_err(out, ref):.3e}, error / |guided eps| {err / scale:.3e}")
        assert err / scale < 5e-3
        if mode.startswith("null+"):
            assert rel_err(out, ref) < 5e-3
    off = e._classifer_free_guidance(x2, t, *embs, mode="null+(for-null)", do_classifier_free_guidance=False).cpu()
    assert rel_err(off, g["cfg_off"][:, :3]) < 5e-3
    with pytest.raises(NotImplementedError):
        e._classifer_free_guidance(x2, t, *embs, mode="edit-proj[for](edit)", do_classifier_free_guidance=True)


def test_get_x0_matches_reference(setup, dev):
    g, e, embs = setup
    xt, t = g["xt"].to(dev), float(g["t"])
    a = e.get_x0(xt, t, g["t_idx"], *embs, mask=g["mask"].to(dev), mode="null+(for-null)").cpu()
    b = e.get_x0(xt, t, g["t_idx"], *embs, mask=None, mode="null+(for-null)+(edit-null)", flatten=True).cpu()
    print(f"x0_hat masked {rel_err(a, g['x0_masked']):.3e}, flat {rel_err(b, g['x0_flat']):.3e}")
    assert a.shape == g["x0_masked"].shape and rel_err(a, g["x0_masked"]) < 5e-3
    assert b.shape == g["x0_flat"].shape and rel_err(b, g["x0_flat"]) < 5e-3


def test_guided_power_method_matches_reference(setup, dev):
    """local_encoder_decoder_pullback_xt under classifier-free guidance (src/modules/edit.py:1589-1676):
    J = sum_i w_i J_i over the 2-3 conditionings, from the reference's V0 draw (seed 7)."""
    g, e, embs = setup
    xt, t = g["xt"].to(dev), float(g["t"])
    torch.manual_seed(7)
    v0, _ = torch.linalg.qr(torch.randn(xt.numel(), 2))
    for (mode, n_iter), ref in g["pullback"].items():
        u, s, vT = e.local_encoder_decoder_pullback_xt(xt, t, g["t_idx"], *embs, pca_rank=2, min_iter=10 ** 6,
                                                       max_iter=n_iter, mask=g["mask"].to(dev), mode=mode,
                                                       v0=v0.T.contiguous())
        torch.cuda.synchronize()
        srel = float(((s.cpu() - ref["s"]).abs() / ref["s"]).max())
        ang = float(principal_angles_deg(vT, ref["vT"]).max())
        uang = float(principal_angles_deg(u.T, ref["u"].T).max())
        print(f"guided power method {mode} N={n_iter}: s rel {srel:.2e}, vT {ang:.3f} deg, u {uang:.3f} deg")
        assert u.shape == ref["u"].shape and srel < 1e-3 and ang < 1.0 and uang < 1.0


def test_text_supervised_direction_matches_reference(setup, dev):
    """get_delta_xt_via_grad / get_v_modify (src/modules/edit.py:1680-1741): one transposed pass with the
    data-dependent cotangent x0_hat_after - x0_hat."""
    g, e, embs = setup
    xt, t = g["xt"].to(dev), float(g["t"])

    def angle(a, b):
        c = float((a.double() * b.double()).sum() / (a.double().norm() * b.double().norm()))
        return float(torch.rad2deg(torch.acos(torch.tensor(min(1.0, abs(c)))))), c

    for name, mask in (("delta_masked", g["mask"].to(dev)), ("delta_nomask", None)):
        v = e.get_delta_xt_via_grad(xt, t, g["t_idx"], *embs, mask=mask, mode="null+(for-null)+(edit-null)").cpu()
        a, c = angle(v, g[name])
        print(f"{name}: angle to the reference direction {a:.3f} deg (cos {c:+.6f}), |v| {float(v.norm()):.6f}")
        assert v.shape == g[name].shape and c > 0 and a < 1.0 and abs(float(v.norm()) - 1) < 1e-4
    vj = e.get_v_modify(xt, t, g["t_idx"], *embs, mask=g["mask"].to(dev), jacobian=True).cpu()
    a, c = angle(vj, g["v_modify_jacobian"])
    assert c > 0 and a < 1.0
    for mode, ref in g["v_modify"].items():
        v = e.get_v_modify(xt, t, g["t_idx"], *embs, mask=g["mask"].to(dev), mode=mode, jacobian=False).cpu()
        a, c = angle(v, ref)
        print(f"get_v_modify {mode}: angle {a:.3f} deg, norm ratio {float(v.norm() / ref.norm()):.5f}")
        assert v.shape == ref.shape and c > 0 and a < 1.0 and abs(float(v.norm() / ref.norm()) - 1) < 5e-3


def test_guided_ddim_loop_matches_reference(setup, dev):
    g, e, embs = setup
    kw = dict(for_prompt_emb=embs[0], edit_prompt_emb=embs[1], null_prompt_emb=embs[2])
    u8 = e.DDPMforwardsteps(g["x2"].to(dev), t_start_idx=90, t_end_idx=-1, mode="null+(for-null)", **kw).cpu()
    ref = g["ddpm_final_u8"]
    diff = (u8.int() - ref.int()).abs()
    print(f"9 guided DDIM steps: uint8 images differ by at most {int(diff.max())} level(s), mean {float(diff.float().mean()):.4f}")
    assert u8.shape == ref.shape and u8.dtype == torch.uint8 and int(diff.max()) <= 2 and float(diff.float().mean()) < 0.1
    xt, t, i = e.DDPMforwardsteps(g["x2"].to(dev), t_start_idx=88, t_end_idx=92, mode="null+(for-null)+(edit-null)", **kw)
    assert i == g["ddpm_mid"]["idx"] == 92 and float(t) == float(g["ddpm_mid"]["t"])
    assert rel_err(xt.cpu(), g["ddpm_mid"]["xt"]) < 5e-3


def test_t2i_drivers_write_the_reference_files_and_compose_three_directions(setup, dev):
    """Non-semantic driver (src/modules/edit.py:1745-1871), semantic driver with the Jacobian-based
    text-supervised direction + null-space projection (:1874-2018), and the composition of THREE saved
    directions (BASELINE config 5; the reference's group edit composes two, :2171-2212)."""
    from loco_edit_b200.masks import save_masks
    g, e, embs = setup
    save_masks(e.result_folder, torch.stack([g["mask"][0], ~g["mask"][0]], 0))
    gen = torch.Generator().manual_seed(9)
    e.xT = torch.randn(1, 3, 32, 32, generator=gen)
    d = 3 * 32 * 32
    out = e.run_edit_null_space_projection_xt(op="mid", block_idx=0, vis_num=2, mask_index=0, vis_num_pc=1, pca_rank=2,
                                              null_space_projection=True, pca_rank_null=3)
    torch.cuda.synchronize()
    base = os.path.join(e.result_folder, "basis", "local_basis-0.4T-pca-rank-2-select-mask0")
    for f, shp in (("u-modify.pt", (int(g["mask"].sum()), 2)), ("vT-modify.pt", (2, d)),
                   ("u-null-null_space_rank_3.pt", (d - int(g["mask"].sum()), 3)), ("vT-null-null_space_rank_3.pt", (3, d))):
        assert tuple(torch.load(os.path.join(base, f)).shape) == shp, f
    vT, vn = out["vT"].double().cpu(), out["vT_null"].double().cpu()
    assert float((vT @ vn.T).abs().max()) < 1e-4 and float((vT.norm(dim=1) - 1).abs().max()) < 1e-5
    assert out["images"].shape == (5, 32, 32, 3) and out["images"].dtype == torch.uint8
    # second call: the cached basis files are loaded (same direction)
    out2 = e.run_edit_null_space_projection_xt(op="mid", block_idx=0, vis_num=2, mask_index=0, vis_num_pc=1, pca_rank=2,
                                               null_space_projection=True, pca_rank_null=3)
    assert float((out2["vT"] - out["vT"]).abs().max()) < 1e-6
    # semantic driver, Jacobian direction
    sem = e.run_edit_null_space_projection_xt_semantic(op="mid", block_idx=0, vis_num=2, mask_index=0, vis_num_pc=1,
                                                      null_space_projection=True, pca_rank_null=3, jacobian=True)
    name = ("Semantic_Edit_xt-edit-0.4T-edit_prompt-a dog with glasses-select_mask0-null_space_projection_True_"
            "null_space_rank_3_null+(for-null)+(edit-null)-pc_000-vT.pt")
    p0 = os.path.join(e.result_folder, "basis", name)
    assert os.path.exists(p0) and tuple(torch.load(p0).shape) == (1, d)
    assert sem["images"].shape == (5, 32, 32, 3)
    # three directions composed cumulatively
    paths = [p0]
    for i in range(2):
        p = os.path.join(e.result_folder, f"dir{i}-vT.pt")
        torch.save(out["vT"][i:i + 1].cpu(), p)
        paths.append(p)
    grp = e.group_edit_null_space_projection(paths, mask_index=0)
    lat = grp["latents"].cpu()
    step = 0.5 * 4
    vs = [torch.load(p, map_location="cpu").reshape(1, 3, 32, 32).float() for p in paths]
    want = [lat[0:1]]
    for v in vs:
        want.append(want[-1] + step * v)
    assert lat.shape == (4, 3, 32, 32) and torch.equal(lat, torch.cat(want, 0))      # axpy: bit-exact
    assert grp["images"].shape == (4, 32, 32, 3)
