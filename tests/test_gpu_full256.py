"""GPU parity at BASELINE's full sizes against the UNMODIFIED reference (fixtures produced on CPU
by tests/golden/make_golden_full.py): the 113.7 M-parameter DDPM-256 U-Net and the 93.6 M-parameter
P2 U-Net at 256 x 256 with the bench weights (seed 1234).  These runs go through every layer shape
of the real configuration (CTA-pair halo convs with fused shortcuts, split-K small layers, fused
GroupNorm statistics), which the reduced-depth fixtures cannot reach.

Tolerances (north_star): eps relative L2 < 5e-3 (TF32 tensor-core convs vs fp32 CPU; fixtures are
stored in fp16, 5e-4); singular values 1e-3 relative; principal angles < 1 degree."""
import os

import pytest
import torch

from gpu_util import principal_angles_deg, rel_err

pytestmark = pytest.mark.gpu


def _inputs(seed=0):
    g = torch.Generator().manual_seed(seed)
    x = (0.5 * torch.randn(1, 3, 256, 256, generator=g)).clamp(-1, 1)
    xt = torch.randn(1, 3, 256, 256, generator=g)
    mask = torch.zeros(3, 256, 256, dtype=torch.bool)
    mask[:, 96:160, 64:192] = True
    return x, xt, mask


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def test_ddpm256_forward_and_power_iteration_match_reference(dev, golden_dir):
    from loco_edit_b200.edit import local_basis
    from loco_edit_b200.scheduler import YHCustomScheduler
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import DDPM256, random_state_dict
    g = torch.load(os.path.join(golden_dir, "full256_ddpm.pt"), weights_only=False)
    _, xt, mask = _inputs(g["input_seed"])
    net = B200UNet(DDPM256, random_state_dict(DDPM256, seed=g["weights_seed"]), device=dev)
    # forward through the B = 1 plan (fused GroupNorm statistics) and as row 1 of a B = 3 batch
    e1 = net(xt.to(dev), g["t"]).cpu()
    e3 = net(torch.cat([xt + 0.3, xt, -xt]).to(dev), g["t"])[1:2].cpu()
    ref = g["eps"].float()
    print(f"DDPM-256 eps vs reference: B=1 {rel_err(e1, ref):.3e}, row of B=3 {rel_err(e3, ref):.3e}")
    assert rel_err(e1, ref) < 5e-3 and rel_err(e3, ref) < 5e-3
    # one rank-2 power iteration from the reference's V0 draw (seed 7, edit.py:2435-2437)
    torch.manual_seed(g["v0_seed"])
    v0, _ = torch.linalg.qr(torch.randn(xt.numel(), 2))
    sched = YHCustomScheduler(device=dev)
    sched.set_timesteps(100)
    _, s, vT = local_basis(net, sched, xt.to(dev), g["t"], 2, v0=v0.T.contiguous().to(dev),
                           min_iter=10 ** 6, max_iter=1, mask=mask.to(dev), verbose=False)
    torch.cuda.synchronize()
    srel = float(((s.cpu() - g["s"]).abs() / g["s"]).max())
    ang = float(principal_angles_deg(vT, g["vT"]).max())
    print(f"DDPM-256 rank-2 power iteration vs reference: s rel {srel:.2e}, max principal angle {ang:.3f} deg")
    assert srel < 1e-3 and ang < 1.0


def test_p2_256_forward_matches_reference(dev, golden_dir):
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import P2_256, random_state_dict
    g = torch.load(os.path.join(golden_dir, "full256_p2.pt"), weights_only=False)
    _, xt, _ = _inputs(g["input_seed"])
    net = B200UNet(P2_256, random_state_dict(P2_256, seed=g["weights_seed"]), device=dev)
    e1 = net(xt.to(dev), g["t"]).cpu()
    ref = g["eps"].float()
    print(f"P2-256 eps vs reference: {rel_err(e1, ref):.3e}")
    assert rel_err(e1, ref) < 5e-3


def _psnr(a, b):
    import math
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 10 * math.log10(4.0 / mse)      # images live in [-1, 1]


@pytest.mark.parametrize("arith", ["tf32", "fp16"])
def test_headline_config_rank5_n12_matches_reference_driver(dev, golden_dir, tmp_path, arith):
    """BASELINE config 1 at its own size and depth against the UNMODIFIED reference driver
    (tests/golden/make_golden_n12.py -> driver_full256.pt: `run_edit_null_space_projection`,
    src/modules/edit.py:2216-2366, on the 113.7 M-parameter DDPM-256 U-Net at 256 x 256, rank 5 + null
    rank 5, N = 12 power iterations each, 59-step final stage with eta = 1 from index 79).

    From the reference's own x_t and its RNG stream (seed 11: V0 edit, V0 null, 20 noise draws):
      * singular values after 4 / 8 / 12 iterations: 1e-3 relative (north_star);
      * principal angles of the edit basis after 4 / 8 / 12 iterations, of the null basis after 12 and
        of the projected direction: < 1 degree (north_star);
      * the 5 edited images of the final stage, from the reference's -vT.pt (transfer edit) AND from
        this library's own direction (full edit): PSNR >= 40 dB (north_star)."""
    import types
    from loco_edit_b200 import ops
    from loco_edit_b200.edit import EditUncondDiffusion, local_basis
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import DDPM256, random_state_dict
    g = torch.load(os.path.join(golden_dir, "driver_full256.pt"), weights_only=False)
    assert g["n_iter"] == 12 and g["edit_t_idx"] == 40 and g["boost_idx"] == 79
    _, _, mask = _inputs(g["input_seed"])
    xt_ref = g["xt"]
    d = xt_ref.numel()
    # the reference's RNG stream (edit.py:2435 twice, then utils.py:371 once per eta = 1 step)
    torch.manual_seed(g["seed"])
    v0a, _ = torch.linalg.qr(torch.randn(d, 5))
    v0b, _ = torch.linalg.qr(torch.randn(d, 5))
    noises = [torch.randn(5, 3, 256, 256) for _ in range(20)]

    net = B200UNet(DDPM256, random_state_dict(DDPM256, seed=g["weights_seed"]), device=dev)
    # "tf32": fp32 storage + tcgen05 kind::tf32 everywhere; "fp16": fp16 storage + kind::f16 for the DDIM
    # programs AND the Jacobian programs (tangent / cotangent rows range-scaled inside the library)
    net.fwd_half = net.jac_half = (arith == "fp16")
    args = types.SimpleNamespace(
        device=dev, dtype=torch.float32, seed=11, model_name="CelebA_HQ_HF", dataset_name="CelebA_HQ_mask",
        image_size=256, for_steps=100, inv_steps=100, edit_t=0.6, performance_boosting_t=0.2,
        x_space_guidance_edit_step=1.0, x_space_guidance_scale=0.5, x_space_guidance_num_step=16,
        result_folder=str(tmp_path), sample_idx=7, choose_sem="hair", mask_index=0, sampling_mode=False,
        vT_path="", vT1_path="", verbose=False, save_images=False, noise_schedule=None)
    e = EditUncondDiffusion(args, unet=net)
    assert e.edit_t_idx == 40 and e.performance_boosting_t_idx == 79
    assert net.plan(1, 5, 5).half == (arith == "fp16")
    t40 = e.scheduler._ts_host[40]
    base = "basis/local_basis-0.6T-select-mask-hair/"
    files = g["files"]
    trace = g["svd_trace"]

    def run(v0, m, tr, name):
        V = v0.T.contiguous().to(dev)
        worst_s, out = 0.0, None
        for it in (4, 8, 12):
            _, s, V = local_basis(net, e.scheduler, xt_ref.to(dev), t40, 5, v0=V, min_iter=10 ** 6, max_iter=4,
                                  mask=m.to(dev), verbose=False)
            sref = tr["s"][it - 1]
            srel = float(((s.cpu() - sref).abs() / sref).max())
            worst_s = max(worst_s, srel)
            msg = f"{name}: N={it:2d} s rel {srel:.2e}"
            if it in tr["V"]:
                ang = float(principal_angles_deg(V, tr["V"][it].float()).max())
                msg += f", max principal angle {ang:.3f} deg"
                assert ang < 1.0, msg
            print(msg)
            assert srel < 1e-3, msg
        return s, V

    s_m, vm = run(v0a, mask, trace[0], arith + " edit basis")
    ang_m = float(principal_angles_deg(vm, files[base + "vT-modify-pca-rank-5.pt"]).max())
    s_n, vn = run(v0b, ~mask, trace[1], arith + " null basis")
    ang_n = float(principal_angles_deg(vn, files[base + "vT-null-5.pt"]).max())
    print(f"after N=12: edit basis {ang_m:.3f} deg, null basis {ang_n:.3f} deg vs the reference's files")
    assert ang_m < 1.0 and ang_n < 1.0
    vproj = ops.nullspace_project(vm, vn, project=True)
    ref_name = [n for n in files if n.endswith("pc_000-vT.pt")][0]
    vref = files[ref_name]
    ang_p = float(principal_angles_deg(vproj[0:1], vref).max())
    print(f"projected direction 0: {ang_p:.3f} deg")
    assert ang_p < 1.0

    # final stage: transfer edit (reference's direction) and full edit (our direction)
    ref_imgs = g["finals"][0].float()
    for label, v in (("transfer edit (reference -vT.pt)", vref[0].to(dev)),
                     ("full edit (own direction)", vproj[0] * torch.sign((vproj[0].cpu() * vref[0]).sum()).to(dev))):
        batch = e.build_edit_batch(xt_ref.to(dev), v.contiguous(), 2)
        assert batch.shape == (5, 3, 256, 256)
        e.noise_fn = lambda i, x: noises[i - 79].to(dev)
        img = e.DDIMforwardsteps(batch, t_start_idx=40, t_end_idx=-1, save_image=False,
                                 performance_boosting=True).cpu()
        p = _psnr(img, ref_imgs)
        print(f"{label}: PSNR {p:.1f} dB vs the reference driver's images, "
              f"max abs diff {float((img - ref_imgs).abs().max()):.2e}")
        assert p >= 40.0
