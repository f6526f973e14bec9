"""GPU: the drop-in driver `EditUncondDiffusion.run_edit_null_space_projection` end to end against
the fixture produced by the UNMODIFIED reference driver (tests/golden/make_golden.py section 5):
same weights, image, mask, V0 draws and eta=1 noise draws (seed 11).

Checked: basis files written with the reference's names/shapes; vT-modify / vT-null / projected vT
equal to the reference's up to row sign (principal angles < 1 deg, north_star); the inversion /
forward chain against the oracle; the edit itself from the reference's x_t and -vT.pt file
(transfer-edit workflow of the README) with PSNR >= 40 dB."""
import math
import os
import types

import pytest
import torch

from gpu_util import principal_angles_deg

pytestmark = pytest.mark.gpu


class _Dataset:
    def __init__(self, x0, mask):
        self.x0, self.mask = x0, mask

    def __getitem__(self, idx):
        return self.x0

    def getmask(self, idx, choose_sem):
        return self.mask


def _make(dev, tmp, g, vT_path="", fuse=False):
    from loco_edit_b200.edit import EditUncondDiffusion
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import random_state_dict, tiny_arch
    arch = tiny_arch(resolution=32, ch_mult=(1, 2), attn_resolutions=(16,), num_res_blocks=1)
    sd = random_state_dict(arch, seed=1234, perturb_norm=0.1)
    net = B200UNet(arch, sd, device=dev)
    args = types.SimpleNamespace(
        device=dev, dtype=torch.float32, seed=11, model_name="CelebA_HQ_HF", dataset_name="CelebA_HQ_mask",
        image_size=32, for_steps=100, inv_steps=100, edit_t=0.6, performance_boosting_t=0.2,
        x_space_guidance_edit_step=1.0, x_space_guidance_scale=0.5, x_space_guidance_num_step=4,
        result_folder=str(tmp), sample_idx=7, choose_sem="hair", mask_index=0, sampling_mode=False,
        vT_path=vT_path, vT1_path="", verbose=False, save_images=False, noise_schedule=None)
    e = EditUncondDiffusion(args, unet=net, dataset=_Dataset(g["x0"], g["mask"]))
    orig = e.local_encoder_decoder_pullback_xt

    def capped(**kw):          # the fixture was generated with the same 2-iteration cap
        kw["max_iter"] = 2
        return orig(**kw)

    e.local_encoder_decoder_pullback_xt = capped
    orig_pair = e.local_encoder_decoder_pullback_xt_pair

    def capped_pair(**kw):
        kw["max_iter"] = 2
        return orig_pair(**kw)

    e.local_encoder_decoder_pullback_xt_pair = capped_pair
    e.fuse_bases = fuse
    return e


def _psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 10 * math.log10(4.0 / mse)      # images live in [-1, 1]


def test_driver_matches_reference_driver(golden_dir, tmp_path):
    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    g = torch.load(os.path.join(golden_dir, "driver_tiny.pt"), weights_only=False)
    assert g["edit_t_idx"] == 40 and g["boost_idx"] == 79
    # replay the reference's RNG stream: V0 (d x 2), V0 (d x 3), then 20 noise draws per direction
    torch.manual_seed(g["seed"])
    d = g["x0"].numel()
    v0a, _ = torch.linalg.qr(torch.randn(d, 2))
    v0b, _ = torch.linalg.qr(torch.randn(d, 3))
    noises = [torch.randn(5, 3, 32, 32) for _ in range(40)]

    # ---- pass 1: compute the bases, write the files ----
    e = _make(dev, tmp_path / "a", g)
    assert e.edit_t_idx == 40 and e.performance_boosting_t_idx == 79
    e.v0 = {2: v0a.T.contiguous().to(dev), 3: v0b.T.contiguous().to(dev)}
    it = iter(noises)
    e.noise_fn = lambda i, x: next(it).to(dev)
    e.run_edit_null_space_projection(idx=7, vis_num=2, vis_num_pc=2, pca_rank=2, pca_rank_null=3,
                                     null_space_projection=True, use_mask=True)
    torch.cuda.synchronize()
    got = {}
    for root, _, fs in os.walk(e.result_folder):
        for f in fs:
            if f.endswith(".pt"):
                got[os.path.relpath(os.path.join(root, f), e.result_folder)] = torch.load(os.path.join(root, f)).cpu()
    assert sorted(got) == sorted(g["files"]), (sorted(got), sorted(g["files"]))
    for name, ref in g["files"].items():
        mine = got[name]
        assert mine.shape == ref.shape and mine.dtype == ref.dtype, name
        ang = float(principal_angles_deg(mine, ref).max())
        print(f"{os.path.basename(name)}: max principal angle {ang:.3f} deg (bases at the driver's own x_t)")
        # these bases sit at the END of the chaotic x0 -> xT -> xt chain (x_t itself is ~5 % away from
        # the reference's, see pass 2), so only a loose bound is meaningful here; the 1-degree bar is
        # asserted below from the reference's x_t
        assert ang < 5.0
    assert len(e.last_images) == 2 and e.last_images[0].shape == (5, 3, 32, 32)

    # ---- pass 2: the chain x0 -> xT -> xt against the CPU oracle (== reference, bit for bit) ----
    # With random-init weights the DDIM inversion/forward chain is expanding: a relative input
    # perturbation of 1e-5 grows to 7e-5 at x_T and 3.6e-4 at x_t in the fp32 oracle itself
    # (35-70x, measured), so the per-step TF32 noise (5e-4 on eps) ends at the percent level at x_t.
    from loco_edit_b200.weights import random_state_dict, tiny_arch
    from oracle import ddpm_ref, pullback_ref
    arch = tiny_arch(resolution=32, ch_mult=(1, 2), attn_resolutions=(16,), num_res_blocks=1)
    ref_unet = ddpm_ref.RefUNet(arch, random_state_dict(arch, seed=1234, perturb_norm=0.1))
    rs = pullback_ref.RefScheduler()
    xT_ref = pullback_ref.ddim_inversion(ref_unet, rs, g["x0"])
    xt_ref, _, _ = pullback_ref.ddim_forward(ref_unet, rs, xT_ref, 0, 40)
    e2 = _make(dev, tmp_path / "b", g)
    xT = e2.run_DDIMinversion(7)
    xt, _, t_idx = e2.DDIMforwardsteps(xT, 0, e2.edit_t_idx)
    torch.cuda.synchronize()
    rel_T = float((xT.cpu() - xT_ref).norm() / xT_ref.norm())
    rel_t = float((xt.cpu() - xt_ref).norm() / xt_ref.norm())
    print(f"x_T rel err {rel_T:.3e}, x_t rel err {rel_t:.3e} (chaotic chain, see comment)")
    assert t_idx == 40 and rel_T < 0.05 and rel_t < 0.15

    # ---- pass 2b: both local bases + projection from the reference's x_t (north_star: < 1 deg) ----
    base = "basis/local_basis-0.6T-select-mask-hair/"
    t40 = e2.scheduler.timesteps[40]
    _, _, vm = e2.local_encoder_decoder_pullback_xt(x=xt_ref.to(dev), t=t40, pca_rank=2, min_iter=10, max_iter=50,
                                                    convergence_threshold=1e-4, mask=g["mask"].to(dev),
                                                    v0=v0a.T.contiguous().to(dev))
    _, _, vn = e2.local_encoder_decoder_pullback_xt(x=xt_ref.to(dev), t=t40, pca_rank=3, min_iter=10, max_iter=50,
                                                    convergence_threshold=1e-4, mask=~g["mask"].to(dev),
                                                    v0=v0b.T.contiguous().to(dev))
    from loco_edit_b200 import ops
    vproj = ops.nullspace_project(vm, vn, project=True)
    for mine, name in [(vm, "vT-modify-pca-rank-2.pt"), (vn, "vT-null-3.pt")]:
        ang = float(principal_angles_deg(mine, g["files"][base + name]).max())
        print(f"{name} from the reference x_t: max principal angle {ang:.3f} deg")
        assert ang < 1.0
    for pc in range(2):
        ref_v = [v for k_, v in g["files"].items() if k_.endswith("pc_%03d-vT.pt" % pc)][0]
        ang = float(principal_angles_deg(vproj[pc:pc + 1], ref_v).max())
        print(f"projected direction {pc}: angle {ang:.3f} deg")
        assert ang < 1.0

    # ---- pass 3: the edit itself from the reference's x_t and the reference's -vT.pt file ----
    ref_name = [n for n in g["files"] if n.endswith("pc_000-vT.pt")][0]
    vref = g["files"][ref_name]
    batch = e2.build_edit_batch(xt_ref.to(dev), vref[0].to(dev), 2)
    assert torch.equal(batch.cpu(), pullback_ref.edit_batch(xt_ref, vref[0], 0.5, 4, 2))   # bit-exact
    it2 = iter(noises)
    e2.noise_fn = lambda i, x: next(it2).to(dev)
    img = e2.DDIMforwardsteps(batch, t_start_idx=40, t_end_idx=-1, save_image=False,
                              performance_boosting=True).cpu()
    ref = g["finals"][0]
    p = _psnr(img, ref)
    print(f"edited images vs reference driver: PSNR {p:.1f} dB, max abs diff {float((img - ref).abs().max()):.2e}, "
          f"rel {float((img - ref).norm() / ref.norm()):.2e}")
    assert p >= 40.0


def test_group_edit_composes_directions_like_the_reference(golden_dir, tmp_path):
    """group_edit_null_space_projection (src/modules/edit.py:2171-2212): two saved directions are
    applied cumulatively with step scale * num_step; the 3-latent batch (bit-exact) and the
    eta = 1 DDIM stage from the driver's own x_t against the oracle with the same injected noise."""
    from oracle import ddpm_ref, pullback_ref
    from loco_edit_b200.weights import random_state_dict, tiny_arch
    dev = torch.device("cuda:0")
    g = torch.load(os.path.join(golden_dir, "driver_tiny.pt"), weights_only=False)
    d = g["x0"].numel()
    gen = torch.Generator().manual_seed(21)
    v = torch.randn(2, 1, d, generator=gen)
    v = v / v.norm(dim=2, keepdim=True)
    paths = []
    for i in range(2):
        p = str(tmp_path / f"dir{i}-vT.pt")
        torch.save(v[i], p)
        paths.append(p)
    e = _make(dev, tmp_path / "g", g, vT_path=paths[0])
    e.vT1_path = paths[1]
    noises = [torch.randn(3, 3, 32, 32, generator=gen) for _ in range(20)]
    it = iter(noises)
    e.noise_fn = lambda i, x: next(it).to(dev)
    xt = e.group_edit_null_space_projection(idx=7)
    torch.cuda.synchronize()
    assert len(e.last_images) == 1 and e.last_images[0].shape == (3, 3, 32, 32)
    # the composed latents: xt, xt + s*n*v0, xt + s*n*(v0 + v1)   (s = 0.5, n = 4 in _make)
    xt_c = xt.cpu()
    step = 0.5 * 4
    b1 = xt_c + step * v[0].reshape(1, 3, 32, 32)
    batch = torch.cat([xt_c, b1, b1 + step * v[1].reshape(1, 3, 32, 32)], 0)
    arch = tiny_arch(resolution=32, ch_mult=(1, 2), attn_resolutions=(16,), num_res_blocks=1)
    ref_unet = ddpm_ref.RefUNet(arch, random_state_dict(arch, seed=1234, perturb_norm=0.1))
    rs = pullback_ref.RefScheduler()
    nz = {g["boost_idx"] + i: noises[i] for i in range(20)}
    ref = pullback_ref.ddim_forward(ref_unet, rs, batch, 40, -1, boost_idx=g["boost_idx"], noises=nz)
    p = _psnr(e.last_images[0].cpu(), ref)
    print(f"group edit (2 directions) vs oracle from the same x_t: PSNR {p:.1f} dB")
    assert p >= 40.0


def test_driver_fused_power_methods_write_the_same_bases(golden_dir, tmp_path):
    """run_edit_null_space_projection with the edit and the null power method advanced in one fused pass per
    iteration (the default) writes the files of the two-separate-loops driver: same names, shapes, subspaces."""
    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    g = torch.load(os.path.join(golden_dir, "driver_tiny.pt"), weights_only=False)
    torch.manual_seed(g["seed"])
    d = g["x0"].numel()
    v0a, _ = torch.linalg.qr(torch.randn(d, 2))
    v0b, _ = torch.linalg.qr(torch.randn(d, 3))
    noises = [torch.randn(5, 3, 32, 32) for _ in range(40)]
    got = {}
    for fuse in (False, True):
        e = _make(dev, tmp_path / ("fused" if fuse else "separate"), g, fuse=fuse)
        e.v0 = {2: v0a.T.contiguous().to(dev), 3: v0b.T.contiguous().to(dev)}
        it = iter(noises)
        e.noise_fn = lambda i, x: next(it).to(dev)
        e.run_edit_null_space_projection(idx=7, vis_num=2, vis_num_pc=2, pca_rank=2, pca_rank_null=3,
                                         null_space_projection=True, use_mask=True)
        torch.cuda.synchronize()
        files = {}
        for root, _, fs in os.walk(e.result_folder):
            for f in fs:
                if f.endswith(".pt"):
                    files[os.path.relpath(os.path.join(root, f), e.result_folder)] = torch.load(os.path.join(root, f)).cpu()
        got[fuse] = files
    assert sorted(got[True]) == sorted(got[False]) == sorted(g["files"])
    for name, ref in got[False].items():
        mine = got[True][name]
        assert mine.shape == ref.shape and mine.dtype == ref.dtype, name
        # both runs sit at the end of their own (noise-level different) x0 -> xT -> xt chain
        assert float(principal_angles_deg(mine, ref).max()) < 5.0, name


def test_converging_pair_equals_two_separate_loops(golden_dir):
    """local_basis_pair_converging == two local_basis calls with the reference's stopping rule, including the
    hand-over when one loop stops first (forced here with a loose threshold on a fast-converging rank-1 basis)."""
    from loco_edit_b200.edit import local_basis, local_basis_pair_converging
    from loco_edit_b200.scheduler import YHCustomScheduler
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import random_state_dict
    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    g = torch.load(os.path.join(golden_dir, "pullback_tiny.pt"), weights_only=False)
    sd = random_state_dict(g["arch"], seed=g["seed"], perturb_norm=g["perturb_norm"])
    net = B200UNet(g["arch"], sd, device=dev)
    sched = YHCustomScheduler(device=dev)
    d = g["xt"].numel()
    gen = torch.Generator().manual_seed(5)
    va, _ = torch.linalg.qr(torch.randn(d, 1, generator=gen))
    vb, _ = torch.linalg.qr(torch.randn(d, 3, generator=gen))
    va, vb = va.T.contiguous().to(dev), vb.T.contiguous().to(dev)
    mask, xt = g["mask"].to(dev), g["xt"].to(dev)
    for thr, max_iter in ((1e-4, 6), (3e-2, 40)):
        kw = dict(min_iter=2, max_iter=max_iter, convergence_threshold=thr, verbose=False)
        _, s1, v1 = local_basis(net, sched, xt, g["t"], 1, v0=va, mask=mask, **kw)
        _, s2, v2 = local_basis(net, sched, xt, g["t"], 3, v0=vb, mask=~mask, **kw)
        pv1, ps1, pv2, ps2 = local_basis_pair_converging(net, sched, xt, g["t"], 1, 3, mask, v0=va, v0_null=vb,
                                                         min_iter=2, max_iter=max_iter, convergence_threshold=thr)
        torch.cuda.synchronize()
        assert float(((ps1 - s1).abs() / s1).max()) < 2e-3 and float(((ps2 - s2).abs() / s2).max()) < 2e-3
        assert float(principal_angles_deg(pv1, v1).max()) < 0.5
        assert float(principal_angles_deg(pv2, v2).max()) < 0.5
