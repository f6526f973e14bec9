"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference; the GPU box does not have it).  The three
third-party imports the reference needs for unrelated lines (diffusers, matplotlib, skimage) are
stubbed in sys.modules; nothing of the reference is copied.  Usage:

    python tests/golden/make_golden.py            # writes tests/golden/*.pt / *.json

What is pinned
  ddpm_param_shapes.json   DDPM(arch).state_dict() names/shapes for the 256x256 config
  unet_tiny.pt             eps = PullBackDDPM(x, t) on a reduced-depth arch, seeded weights
  scheduler.pt             YHCustomScheduler timesteps / alphas / step() outputs (eta = 0 and 1)
  pullback_tiny.pt         EditUncondDiffusion.local_encoder_decoder_pullback_xt after N = 1, 2, 3
                           iterations from a seeded V0 (mask / ~mask / no mask / noise=True)
  driver_tiny.pt           EditUncondDiffusion.run_edit_null_space_projection end to end (inversion,
                           forward to t, two bases, projection, basis files, edited images)
"""
import argparse
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference/src"
sys.path.insert(0, ROOT)


def import_reference():
    for name in ["diffusers", "diffusers.utils", "matplotlib", "matplotlib.pyplot", "skimage"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            sys.modules[name] = m
    d = sys.modules["diffusers"]
    for n in ["DDIMScheduler", "DDIMPipeline", "StableDiffusionPipeline", "DiffusionPipeline",
              "LCMScheduler", "DDPMScheduler"]:
        setattr(d, n, type(n, (), {}))
    sys.modules["diffusers.utils"].pt_to_pil = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    import models.ddpm.diffusion as ddpm        # noqa
    import utils.utils as uu                    # noqa
    import modules.edit as edit                 # noqa
    return ddpm, uu, edit


def ns(**kw):
    return types.SimpleNamespace(**kw)


def ref_unet(ddpm, arch, sd):
    cfg = ns(model=ns(ch=arch["ch"], out_ch=3, ch_mult=list(arch["ch_mult"]),
                      num_res_blocks=arch["num_res_blocks"],
                      attn_resolutions=list(arch["attn_resolutions"]), dropout=0.0, in_channels=3,
                      resamp_with_conv=True),
             data=ns(image_size=arch["resolution"]))
    args = ns(config=cfg, device=torch.device("cpu"), dtype=torch.float32)
    m = ddpm.PullBackDDPM(args)
    missing = m.load_state_dict(sd, strict=True)
    m.eval()
    return m


def make_edit_obj(edit, uu, unet, res, tmpdir, **extra):
    args = ns(noise_schedule=None, device=torch.device("cpu"), dtype=torch.float32, sample_idx=7,
              choose_sem="hair", mask_index=0, sampling_mode=False)
    e = object.__new__(edit.EditUncondDiffusion)
    e.unet = unet
    e.scheduler = uu.YHCustomScheduler(args)
    e.device = torch.device("cpu")
    e.dtype = torch.float32
    e.for_steps = 100
    e.inv_steps = 100
    e.use_yh_custom_scheduler = True
    e.buffer_device = "cpu"
    e.memory_bound = 50
    e.image_size = res
    e.c_in = 3
    e.args = args
    e.result_folder = tmpdir
    e.obs_folder = tmpdir
    for k, v in extra.items():
        setattr(e, k, v)
    return e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-driver", action="store_true")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    ddpm, uu, edit = import_reference()
    from loco_edit_b200.weights import DDPM256, tiny_arch, random_state_dict

    # ---- 1. parameter names/shapes of the 256 config ----
    big = ref_unet.__wrapped__ if hasattr(ref_unet, "__wrapped__") else None
    cfg = ns(model=ns(ch=128, out_ch=3, ch_mult=[1, 1, 2, 2, 4, 4], num_res_blocks=2,
                      attn_resolutions=[16], dropout=0.0, in_channels=3, resamp_with_conv=True),
             data=ns(image_size=256))
    with torch.device("meta"):
        m = ddpm.DDPM(cfg)
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    json.dump(shapes, open(os.path.join(HERE, "ddpm_param_shapes.json"), "w"), indent=0)
    print("params:", len(shapes), sum(int(torch.tensor(s).prod()) for s in shapes.values()))

    # ---- 2. tiny U-Net forward ----
    arch = tiny_arch(resolution=32, ch_mult=(1, 2), attn_resolutions=(16,), num_res_blocks=1)
    sd = random_state_dict(arch, seed=1234, perturb_norm=0.1)
    unet = ref_unet(ddpm, arch, sd)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 32, 32, generator=g)
    t = torch.tensor(595.3636)
    with torch.no_grad():
        eps = unet(x, t)
    torch.save({"arch": arch, "seed": 1234, "perturb_norm": 0.1, "x": x, "t": t, "eps": eps},
               os.path.join(HERE, "unet_tiny.pt"))
    print("unet_tiny eps", eps.shape, float(eps.abs().mean()))

    # ---- 3. scheduler ----
    sargs = ns(noise_schedule=None, device=torch.device("cpu"), dtype=torch.float32)
    sch = uu.YHCustomScheduler(sargs)
    sch.set_timesteps(100)
    ts_f, tsn_f = sch.timesteps.clone(), sch.timesteps_next.clone()
    g = torch.Generator().manual_seed(1)
    xt = torch.randn(2, 3, 8, 8, generator=g)
    et = torch.randn(2, 3, 8, 8, generator=g)
    steps = {}
    for idx in (0, 40, 79, 98):
        tt = sch.timesteps[idx]
        o = sch.step(et, tt, xt, eta=0)
        torch.manual_seed(123)
        o1 = sch.step(et, tt, xt, eta=1)
        torch.manual_seed(123)
        nz = torch.randn_like(xt)
        steps[idx] = {"eta0": o.prev_sample, "x0": o.x0, "eta1": o1.prev_sample, "noise": nz}
    sch.set_timesteps(100, is_inversion=True)
    inv = {}
    for idx in (0, 50, 97):
        tt = sch.timesteps[idx]
        inv[idx] = sch.step(et, tt, xt, eta=0).prev_sample
    torch.save({"timesteps": ts_f, "timesteps_next": tsn_f, "inv_timesteps": sch.timesteps.clone(),
                "inv_timesteps_next": sch.timesteps_next.clone(),
                "alphas_cumprod": sch.alphas_cumprod.clone(), "xt": xt, "et": et, "steps": steps,
                "inv_steps": inv}, os.path.join(HERE, "scheduler.pt"))

    # ---- 4. power-method local basis on the tiny arch ----
    import tempfile
    tmp = tempfile.mkdtemp()
    e = make_edit_obj(edit, uu, unet, 32, tmp)
    e.scheduler.set_timesteps(100)
    t40 = e.scheduler.timesteps[40]
    g = torch.Generator().manual_seed(3)
    xt = torch.randn(1, 3, 32, 32, generator=g)
    mask = torch.zeros(3, 32, 32, dtype=torch.bool)
    mask[:, 12:20, 8:24] = True
    cases = {}
    for name, kw in {"mask_k2": dict(mask=mask, pca_rank=2),
                     "notmask_k3": dict(mask=~mask, pca_rank=3),
                     "nomask_k2": dict(mask=None, pca_rank=2),
                     "noise_k2": dict(mask=mask, pca_rank=2, noise=True)}.items():
        res = {}
        for n_iter in (1, 2, 3):
            torch.manual_seed(7)
            u, s, vT = e.local_encoder_decoder_pullback_xt(
                x=xt, t=t40, op="mid", block_idx=0, min_iter=10 ** 6, max_iter=n_iter,
                convergence_threshold=1e-4, **kw)
            res[n_iter] = {"u": u.clone(), "s": s.clone(), "vT": vT.clone()}
        cases[name] = res
        print(name, "s(N=3) =", res[3]["s"].tolist())
    torch.save({"arch": arch, "seed": 1234, "perturb_norm": 0.1, "xt": xt, "t": t40, "mask": mask,
                "v0_seed": 7, "cases": cases}, os.path.join(HERE, "pullback_tiny.pt"))

    if a.skip_driver:
        return
    # ---- 5. driver end to end (CPU; `.to(cuda:0)` redirected to cpu, nothing else touched) ----
    orig_to = torch.Tensor.to

    def to_nocuda(self, *args, **kw):
        args = tuple(torch.device("cpu") if isinstance(x, torch.device) and x.type == "cuda" else x
                     for x in args)
        if isinstance(kw.get("device"), torch.device) and kw["device"].type == "cuda":
            kw["device"] = torch.device("cpu")
        return orig_to(self, *args, **kw)

    class FakeDataset:
        def __init__(self):
            g = torch.Generator().manual_seed(0)
            self.x0 = (0.5 * torch.randn(1, 3, 32, 32, generator=g)).clamp(-1, 1)

        def __getitem__(self, idx):
            return self.x0

        def getmask(self, idx, choose_sem):
            return mask

    drv_dir = tempfile.mkdtemp()
    e = make_edit_obj(edit, uu, unet, 32, drv_dir, dataset=FakeDataset(),
                      dataset_name="CelebA_HQ_mask", edit_t=0.6, vT_path="", vT1_path="",
                      x_space_guidance_edit_step=1.0, x_space_guidance_scale=0.5,
                      x_space_guidance_num_step=4)
    e.scheduler.set_timesteps(100)
    e.edit_t_idx = (e.scheduler.timesteps - 0.6 * 1000).abs().argmin()
    e.performance_boosting_t_idx = (e.scheduler.timesteps - 0.2 * 1000).abs().argmin()
    finals = []
    orig_fwd = e.DDIMforwardsteps

    def rec_fwd(*args, **kw):
        out = orig_fwd(*args, **kw)
        if kw.get("performance_boosting", False):
            finals.append(out.clone())
        return out

    e.DDIMforwardsteps = rec_fwd
    # the driver hard-codes min_iter=10,max_iter=50; cap iterations to keep the fixture cheap and
    # well-posed (flat spectrum never converges): wrap the method, same code underneath.
    orig_pb = e.local_encoder_decoder_pullback_xt

    def capped(**kw):
        kw["max_iter"] = 2
        return orig_pb(**kw)

    e.local_encoder_decoder_pullback_xt = capped
    torch.Tensor.to = to_nocuda
    try:
        torch.manual_seed(11)
        e.run_edit_null_space_projection(idx=7, vis_num=2, vis_num_pc=2, pca_rank=2, pca_rank_null=3,
                                         null_space_projection=True, use_mask=True)
    finally:
        torch.Tensor.to = orig_to
    files = {}
    for root, _, fs in os.walk(drv_dir):
        for f in fs:
            if f.endswith(".pt"):
                files[os.path.relpath(os.path.join(root, f), drv_dir)] = torch.load(os.path.join(root, f))
    print("driver files:", sorted(files))
    torch.save({"x0": e.dataset.x0, "mask": mask, "files": files, "finals": finals, "seed": 11,
                "edit_t_idx": int(e.edit_t_idx), "boost_idx": int(e.performance_boosting_t_idx)},
               os.path.join(HERE, "driver_tiny.pt"))


if __name__ == "__main__":
    main()
