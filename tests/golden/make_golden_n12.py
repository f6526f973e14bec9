"""Headline-configuration golden from the UNMODIFIED reference driver, at full size and depth
(build container only; see make_golden.py for the import stubs; about 50 minutes on 8 CPU threads):

    python tests/golden/make_golden_n12.py

BASELINE config 1: the 113.7 M-parameter DDPM-256 U-Net (bench weights, seed 1234) at 256 x 256,
`EditUncondDiffusion.run_edit_null_space_projection` (src/modules/edit.py:2216-2366) end to end:
DDIM inversion (98 steps), forward to t = 0.6T (40 steps), edit basis (mask, rank 5) and null basis
(~mask, rank 5) by `local_encoder_decoder_pullback_xt` (src/modules/edit.py:2406-2504) with
N = 12 power iterations each, null-space projection, basis files, and the 59-step final stage
(eta = 1 from index 79) of the first projected direction.

The driver hard-codes min_iter=10, max_iter=50 (src/modules/edit.py:2294-2297); with random-init
weights the spectrum is flat and the method never converges, so the method is
wrapped to cap max_iter at 12 = the reference's minimum iteration count.  Nothing else is touched:
`.to(cuda:0)` is redirected to the CPU, `torch.linalg.svd` is wrapped by a recorder that returns the
untouched result, so the singular values of every iteration and the basis after iterations 4 and 8
are kept for diagnosis.

Written to tests/golden/driver_full256.pt:
  x0, mask seeds, xt (the reference's own x_t at index 40), files{...} (vT-modify [5,d], vT-null
  [5,d], pc_000-vT [1,d], fp32), finals[0] (5 edited images, fp16), svd_trace (s per iteration for
  both bases; V after iterations 4, 8 of the edit basis in fp16).
"""
import os
import sys
import tempfile
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

N_ITER = 12
SEED = 11
RES = 256
TINY = bool(int(os.environ.get("LOCO_GOLDEN_TINY", "0")))     # dry run of this script at 32 x 32
if TINY:
    N_ITER, RES = 2, 32


def inputs(seed=0):
    g = torch.Generator().manual_seed(seed)
    x0 = (0.5 * torch.randn(1, 3, RES, RES, generator=g)).clamp(-1, 1)
    mask = torch.zeros(3, RES, RES, dtype=torch.bool)
    mask[:, 3 * RES // 8:5 * RES // 8, RES // 4:3 * RES // 4] = True
    return x0, mask


def main():
    torch.set_num_threads(os.cpu_count())
    ddpm, uu, edit = mg.import_reference()
    from loco_edit_b200.weights import DDPM256, random_state_dict, tiny_arch

    x0, mask = inputs()
    arch = tiny_arch(resolution=32, ch_mult=(1, 2), attn_resolutions=(16,), num_res_blocks=1) if TINY else DDPM256
    unet = mg.ref_unet(ddpm, arch, random_state_dict(arch, seed=1234))

    class FakeDataset:
        def __getitem__(self, idx):
            return x0

        def getmask(self, idx, choose_sem):
            return mask

    drv_dir = tempfile.mkdtemp()
    e = mg.make_edit_obj(edit, uu, unet, RES, drv_dir, dataset=FakeDataset(),
                         dataset_name="CelebA_HQ_mask", edit_t=0.6, vT_path="", vT1_path="",
                         x_space_guidance_edit_step=1.0, x_space_guidance_scale=0.5,
                         x_space_guidance_num_step=16)
    e.scheduler.set_timesteps(100)
    e.edit_t_idx = (e.scheduler.timesteps - 0.6 * 1000).abs().argmin()
    e.performance_boosting_t_idx = (e.scheduler.timesteps - 0.2 * 1000).abs().argmin()

    finals, xts = [], []
    orig_fwd = e.DDIMforwardsteps

    def rec_fwd(*args, **kw):
        t0 = time.time()
        out = orig_fwd(*args, **kw)
        print("DDIMforwardsteps", time.time() - t0, flush=True)
        if kw.get("performance_boosting", False):
            finals.append(out.clone())
        else:
            xts.append(out[0].clone() if isinstance(out, tuple) else out.clone())
        return out

    e.DDIMforwardsteps = rec_fwd
    orig_pb = e.local_encoder_decoder_pullback_xt
    trace = []

    def capped(**kw):
        kw["max_iter"] = N_ITER
        trace.append({"s": [], "V": {}})
        t0 = time.time()
        out = orig_pb(**kw)
        print("local basis", time.time() - t0, out[1].tolist(), flush=True)
        return out

    e.local_encoder_decoder_pullback_xt = capped
    orig_svd = torch.linalg.svd

    def rec_svd(a, *args, **kw):
        out = orig_svd(a, *args, **kw)
        if trace and a.dim() == 2 and a.shape[1] == 3 * RES * RES:
            rec = trace[-1]
            rec["s"].append(out[1].sqrt().clone())
            it = len(rec["s"])
            if len(trace) == 1 and it in (4, 8):
                rec["V"][it] = out[2].half().clone()
        return out

    orig_to = torch.Tensor.to

    def to_nocuda(self, *args, **kw):
        args = tuple(torch.device("cpu") if isinstance(x, torch.device) and x.type == "cuda" else x
                     for x in args)
        if isinstance(kw.get("device"), torch.device) and kw["device"].type == "cuda":
            kw["device"] = torch.device("cpu")
        return orig_to(self, *args, **kw)

    torch.Tensor.to = to_nocuda
    torch.linalg.svd = rec_svd
    t0 = time.time()
    try:
        torch.manual_seed(SEED)
        e.run_edit_null_space_projection(idx=7, vis_num=2, vis_num_pc=1, pca_rank=5, pca_rank_null=5,
                                         null_space_projection=True, use_mask=True)
    finally:
        torch.Tensor.to = orig_to
        torch.linalg.svd = orig_svd
    print("driver total", time.time() - t0, flush=True)
    files = {}
    for root, _, fs in os.walk(drv_dir):
        for f in fs:
            if f.endswith(".pt"):
                files[os.path.relpath(os.path.join(root, f), drv_dir)] = torch.load(os.path.join(root, f))
    print("driver files:", {k: tuple(v.shape) for k, v in files.items()})
    # the pc_001.. files are rows of the same projected matrix; keep the first (the one that is edited)
    files = {k: v for k, v in files.items() if "-pc_" not in k or "-pc_000-" in k}
    if TINY:
        print([f.shape for f in finals], [x.shape for x in xts], [r["s"] for r in trace])
        return
    torch.save({"input_seed": 0, "weights_seed": 1234, "seed": SEED, "n_iter": N_ITER,
                "xt": xts[-1], "files": files, "finals": [f.half() for f in finals],
                "svd_trace": [{"s": torch.stack(r["s"]), "V": r["V"]} for r in trace],
                "edit_t_idx": int(e.edit_t_idx), "boost_idx": int(e.performance_boosting_t_idx)},
               os.path.join(HERE, "driver_full256.pt"))


if __name__ == "__main__":
    main()
