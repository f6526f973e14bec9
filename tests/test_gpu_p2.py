"""GPU parity of the P2 / guided-diffusion U-Net family (BASELINE config 2: FFHQ_P2 / AFHQ_P2 ...;
reference models/guided_diffusion/unet.py with script_util.P2_DICT) against the CPU oracle
(oracle/p2_ref.py, bit-exact against the reference) and the reference-generated goldens
(tests/golden/make_golden_p2.py).  Same tolerances as tests/test_gpu_unet.py (TF32 tensor-core
convolutions vs exact fp32 on the CPU)."""
import os

import pytest
import torch

from gpu_util import principal_angles_deg, rel_err

pytestmark = pytest.mark.gpu

ARCHS = {
    "two_level_attn16": dict(resolution=32, ch_mult=(1, 2), attn_resolutions=(16,), num_res_blocks=1),
    "three_level_attn8": dict(resolution=32, ch_mult=(1, 1, 2), attn_resolutions=(8,), num_res_blocks=2),
}


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _setup(name, dev, perturb=0.1, seed=4321):
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import random_state_dict, tiny_p2_arch
    from oracle import p2_ref
    arch = tiny_p2_arch(**ARCHS[name])
    sd = random_state_dict(arch, seed=seed, perturb_norm=perturb)
    return arch, sd, B200UNet(arch, sd, device=dev), p2_ref.RefP2UNet(arch, sd)


def _legacy_attn(qkv, head_ch):
    """QKVAttentionLegacy on [N, T, 3C] (guided_diffusion/unet.py:339-356)."""
    n, t, c3 = qkv.shape
    heads = c3 // (3 * head_ch)
    x = qkv.transpose(1, 2).reshape(n * heads, 3 * head_ch, t)
    q, k, v = x.split(head_ch, dim=1)
    scale = head_ch ** -0.25
    w = torch.softmax(torch.einsum("bct,bcs->bts", q * scale, k * scale), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(n, -1, t)
    return a.transpose(1, 2)


@pytest.mark.parametrize("T,C,hc", [(64, 128, 64), (256, 512, 64)])
def test_multihead_attention_fwd_jvp_vjp(dev, T, C, hc):
    from loco_edit_b200 import ops
    g = torch.Generator().manual_seed(T + C)
    k = 2
    tf32 = lambda x: ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    qkv = tf32(torch.randn(1, T, 3 * C, generator=g)).to(dev)
    dq = tf32(torch.randn(k, T, 3 * C, generator=g)).to(dev)
    f = lambda z: _legacy_attn(z, hc)
    oref = f(qkv.double())
    dref = torch.cat([torch.func.jvp(f, (qkv.double(),), (dq[j:j + 1].double(),))[1] for j in range(k)], 0)
    o, S = ops.attention_fwd(torch.cat([qkv, dq], 0).contiguous(), 1, head_ch=hc)
    torch.cuda.synchronize()
    assert rel_err(o[:1], oref) < 5e-4 and rel_err(o[1:], dref) < 1e-3
    go = tf32(torch.randn(k, T, C, generator=g)).to(dev)
    qd = qkv.double().requires_grad_(True)
    od = f(qd)
    gref = torch.cat([torch.autograd.grad(od, qd, go[j:j + 1].double(), retain_graph=True)[0] for j in range(k)], 0)
    gq = ops.attention_vjp(go, qkv, S[0].contiguous(), head_ch=hc)
    torch.cuda.synchronize()
    e = rel_err(gq, gref)
    print(f"legacy attention T={T} C={C} heads={C // hc}: vjp rel_err {e:.2e}")
    assert e < 1e-3


def test_p2_forward_matches_reference_golden(dev, golden_dir):
    """eps of the unmodified reference UNetModel (CPU fp32) on the same weights / inputs."""
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import random_state_dict
    g = torch.load(os.path.join(golden_dir, "p2_tiny.pt"), weights_only=False)
    sd = random_state_dict(g["arch"], seed=g["seed"], perturb_norm=g["perturb_norm"])
    net = B200UNet(g["arch"], sd, device=dev)
    e = net(g["x"].to(dev), g["t"])
    torch.cuda.synchronize()
    err = rel_err(e.cpu(), g["eps"])
    print(f"P2 forward vs reference golden: rel_err {err:.3e}")
    assert torch.isfinite(e).all() and err < 5e-3


@pytest.mark.parametrize("name", list(ARCHS))
def test_p2_jvp_vjp_match_oracle(dev, name):
    arch, sd, net, ref = _setup(name, dev)
    R = arch["resolution"]
    g = torch.Generator().manual_seed(1)
    k = 3
    x = torch.randn(1, 3, R, R, generator=g)
    V = torch.randn(k, 3, R, R, generator=g)
    G = torch.randn(k, 3, R, R, generator=g)
    t = torch.tensor(198.1818)
    f = lambda z: ref(z, t)
    dref = []
    for j in range(k):
        e0, de = torch.func.jvp(f, (x,), (V[j:j + 1],))
        dref.append(de)
    dref = torch.cat(dref, 0)
    xg = x.clone().requires_grad_(True)
    out = f(xg)
    gref = torch.cat([torch.autograd.grad(out, xg, G[j:j + 1], retain_graph=True)[0] for j in range(k)], 0)
    eps, deps = net.jvp(x.to(dev), t, V.to(dev))
    gx = net.vjp(k, G.to(dev))
    e3 = net(torch.cat([x, x + 0.1, x - 0.2]).to(dev), t)      # forward-only plan (fused GN statistics)
    torch.cuda.synchronize()
    e_p, e_t, e_g = rel_err(eps.cpu(), e0), rel_err(deps.cpu(), dref), rel_err(gx.cpu(), gref)
    print(f"P2 {name}: primal {e_p:.3e} jvp {e_t:.3e} vjp {e_g:.3e}")
    assert e_p < 5e-3 and e_t < 5e-3 and e_g < 5e-3
    with torch.no_grad():
        assert rel_err(e3.cpu(), ref(torch.cat([x, x + 0.1, x - 0.2]), t)) < 5e-3
    lhs = (deps.double() * G.to(dev).double()).sum(dim=(1, 2, 3))
    rhs = (V.to(dev).double() * gx.double()).sum(dim=(1, 2, 3))
    scale = deps.double().flatten(1).norm(dim=1) * G.to(dev).double().flatten(1).norm(dim=1)
    print("adjoint gap / (|Jv||g|):", ((lhs - rhs).abs() / scale).tolist())
    assert float(((lhs - rhs).abs() / scale).max()) < 2e-3


@pytest.mark.parametrize("case", ["mask_k3", "notmask_k2"])
def test_p2_power_iteration_matches_reference_golden(dev, golden_dir, case):
    from loco_edit_b200.edit import local_basis
    from loco_edit_b200.scheduler import YHCustomScheduler
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import random_state_dict
    g = torch.load(os.path.join(golden_dir, "p2_pullback_tiny.pt"), weights_only=False)
    sd = random_state_dict(g["arch"], seed=g["seed"], perturb_norm=g["perturb_norm"])
    net = B200UNet(g["arch"], sd, device=dev)
    mask, k = {"mask_k3": (g["mask"], 3), "notmask_k2": (~g["mask"], 2)}[case]
    d = g["xt"].numel()
    torch.manual_seed(g["v0_seed"])
    v0, _ = torch.linalg.qr(torch.randn(d, k))
    sched = YHCustomScheduler(device=dev)
    for n_iter in (1, 2):
        ref = g["cases"][case][n_iter]
        u, s, vT = local_basis(net, sched, g["xt"].to(dev), g["t"], k, v0=v0.T.contiguous().to(dev),
                               min_iter=10 ** 6, max_iter=n_iter, mask=mask.to(dev), verbose=False)
        torch.cuda.synchronize()
        srel = float(((s.cpu() - ref["s"]).abs() / ref["s"]).max())
        ang = float(principal_angles_deg(vT, ref["vT"]).max())
        print(f"P2 {case} N={n_iter}: s rel {srel:.2e}, max principal angle {ang:.3f} deg")
        assert srel < 1e-3
        assert ang < 1.0


def test_full_size_p2_256_properties(dev):
    """BASELINE config-2 model (P2_DICT at 256x256, 93.6 M parameters): size-independent properties."""
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import P2_256, random_state_dict
    sd = random_state_dict(P2_256, seed=1234)
    net = B200UNet(P2_256, sd, device=dev)
    g = torch.Generator().manual_seed(0)
    k = 3
    x = (0.5 * torch.randn(1, 3, 256, 256, generator=g)).clamp(-1, 1).to(dev)
    V = torch.randn(k, 3, 256, 256, generator=g).to(dev)
    G = torch.randn(k, 3, 256, 256, generator=g).to(dev)
    t = 198.1818
    eps, deps = net.jvp(x, t, V)
    gx = net.vjp(k, G)
    e1 = net(x, t)
    torch.cuda.synchronize()
    assert torch.isfinite(eps).all() and torch.isfinite(deps).all() and torch.isfinite(gx).all()
    assert float(eps.abs().mean()) > 1e-3
    print("P2-256 fused-vs-plain primal rel err:", rel_err(eps, e1))
    assert rel_err(eps, e1) < 3e-3
    V2 = V.clone()
    V2[2] = 2 * V[0] - V[1]
    _, d2 = net.jvp(x, t, V2)
    assert rel_err(d2[2], 2 * deps[0] - deps[1]) < 2e-3
    lhs = (deps.double() * G.double()).sum(dim=(1, 2, 3))
    rhs = (V.double() * gx.double()).sum(dim=(1, 2, 3))
    scale = deps.double().flatten(1).norm(dim=1) * G.double().flatten(1).norm(dim=1)
    print("P2-256 adjoint gap:", ((lhs - rhs).abs() / scale).tolist())
    assert float(((lhs - rhs).abs() / scale).max()) < 2e-3


def test_p2_driver_ffhq_script_settings(dev, tmp_path):
    """Drop-in driver on the P2 family with the FFHQ_P2 launch-script settings
    (src/scripts/main_hf_null_space_projection_FFHQ_P2.sh: edit_t 0.2, pca_rank 3, pca_rank_null 5, scale 12, 1 step): basis
    files with the reference's names / shapes, orthonormal rows, the edit basis equal to the
    oracle's power method from the same x_t (< 1 degree), finite edited images."""
    import types
    from loco_edit_b200.edit import EditUncondDiffusion
    from loco_edit_b200.scheduler import YHCustomScheduler
    from oracle import pullback_ref
    arch, sd, net, ref = _setup("two_level_attn16", dev)
    R = arch["resolution"]
    g = torch.Generator().manual_seed(0)
    x0 = (0.5 * torch.randn(1, 3, R, R, generator=g)).clamp(-1, 1)
    mask = torch.zeros(3, R, R, dtype=torch.bool)
    mask[:, 12:20, 8:24] = True

    class _DS:
        def __getitem__(self, idx):
            return x0

        def getmask(self, idx, choose_sem):
            return mask

    args = types.SimpleNamespace(
        device=dev, dtype=torch.float32, seed=3, model_name="FFHQ_P2", dataset_name="FFHQ",
        image_size=R, for_steps=100, inv_steps=100, edit_t=0.2, performance_boosting_t=0.2,
        x_space_guidance_edit_step=1.0, x_space_guidance_scale=12.0, x_space_guidance_num_step=1,
        result_folder=str(tmp_path), sample_idx=4, choose_sem="hair", mask_index=0, sampling_mode=False,
        vT_path="", vT1_path="", verbose=False, save_images=False, noise_schedule=None)
    e = EditUncondDiffusion(args, unet=net, dataset=_DS())
    # FFHQ-style datasets take their mask from the cached mask/mask.pt (src/modules/edit.py:2252-2265;
    # the reference fills it with SAM): row 1 of a 2-mask file, selected with --mask_index
    from loco_edit_b200.masks import save_masks
    save_masks(e.result_folder, torch.stack([~mask[0], mask[0]], 0))
    e.args.mask_index = 1
    d = x0.numel()
    v0a, _ = torch.linalg.qr(torch.randn(d, 3, generator=g))
    v0b, _ = torch.linalg.qr(torch.randn(d, 5, generator=g))
    e.v0 = {3: v0a.T.contiguous().to(dev), 5: v0b.T.contiguous().to(dev)}
    orig = e.local_encoder_decoder_pullback_xt
    xts = []

    def capped(**kw):
        kw["max_iter"] = 2
        xts.append(kw["x"].detach().clone())
        return orig(**kw)

    e.local_encoder_decoder_pullback_xt = capped
    e.fuse_bases = False          # this test follows the two separate loops (the fused loop: test_gpu_driver.py)
    e.run_edit_null_space_projection(idx=4, vis_num=1, vis_num_pc=1, pca_rank=3, pca_rank_null=5,
                                     null_space_projection=True, use_mask=True)
    torch.cuda.synchronize()
    files = {}
    for root, _, fs in os.walk(e.result_folder):
        for f in fs:
            if f.endswith(".pt"):
                files[f] = torch.load(os.path.join(root, f)).cpu()
    assert files["vT-modify-pca-rank-3.pt"].shape == (3, d) and files["vT-null-5.pt"].shape == (5, d)
    pcs = [k for k in files if k.endswith("pc_000-vT.pt")]
    assert len(pcs) == 1 and "edit_0.2T_null_proj_True_rank5_scale_12.0" in pcs[0] and files[pcs[0]].shape == (1, d)
    for k in ("vT-modify-pca-rank-3.pt", "vT-null-5.pt"):
        v = files[k].double()
        assert float((v @ v.T - torch.eye(v.shape[0], dtype=torch.float64)).abs().max()) < 1e-4
    # projected direction is orthogonal to the null basis and has unit norm
    vp = files[pcs[0]].double()
    assert abs(float(vp.norm()) - 1) < 1e-5 and float((files["vT-null-5.pt"].double() @ vp.T).abs().max()) < 1e-4
    # the edit basis against the oracle's power method from the driver's own x_t
    sched = pullback_ref.RefScheduler()
    sched.set_timesteps(100)
    assert e.edit_t_idx == 79          # |t - 0.2 * 1000| is smallest at timesteps[79] = 201.8
    _, _, vref = pullback_ref.local_basis(ref, sched, xts[0].cpu(), sched.timesteps[79], v0a.T, 2, mask=mask)
    ang = float(principal_angles_deg(files["vT-modify-pca-rank-3.pt"], vref).max())
    print(f"P2 driver edit basis vs oracle from the same x_t: {ang:.3f} deg")
    assert ang < 1.0
    assert len(e.last_images) == 1 and e.last_images[0].shape == (3, 3, R, R)
    assert torch.isfinite(e.last_images[0]).all()
