"""Full-size (BASELINE config-1 / config-2 shapes, 256 x 256) golden fixtures from the UNMODIFIED
reference (build container only; see make_golden.py for the import stubs):

    python tests/golden/make_golden_full.py

  full256_ddpm.pt   eps = PullBackDDPM(x, t) of the 113.7 M-parameter DDPM-256 U-Net (weights seed
                    1234 = the bench weights) at t = timesteps[40], and one rank-2 power iteration
                    of EditUncondDiffusion.local_encoder_decoder_pullback_xt (config-1 mask, V0 seed 7)
  full256_p2.pt     eps = UNetModel(x, t) of the 93.6 M-parameter P2 U-Net (weights seed 1234) at
                    t = timesteps[79]
Inputs are regenerated from seeds by the tests; only the reference outputs are stored (fp16 for the
eps fields to keep the fixtures small: the GPU tolerance is 5e-3, fp16 rounding 5e-4).
"""
import os
import sys
import tempfile
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from make_golden_p2 import ref_p2_unet  # noqa: E402


def inputs(seed=0):
    g = torch.Generator().manual_seed(seed)
    x = (0.5 * torch.randn(1, 3, 256, 256, generator=g)).clamp(-1, 1)
    xt = torch.randn(1, 3, 256, 256, generator=g)
    mask = torch.zeros(3, 256, 256, dtype=torch.bool)
    mask[:, 96:160, 64:192] = True
    return x, xt, mask


def main():
    torch.set_num_threads(os.cpu_count())
    ddpm, uu, edit = mg.import_reference()
    import models.guided_diffusion.script_util as su
    from loco_edit_b200.weights import DDPM256, P2_256, random_state_dict

    x, xt, mask = inputs()
    # ---- DDPM-256 ----
    sd = random_state_dict(DDPM256, seed=1234)
    unet = mg.ref_unet(ddpm, DDPM256, sd)
    e = mg.make_edit_obj(edit, uu, unet, 256, tempfile.mkdtemp())
    e.scheduler.set_timesteps(100)
    t40 = e.scheduler.timesteps[40]
    t0 = time.time()
    with torch.no_grad():
        eps = unet(xt, t40)
    print("ddpm forward", time.time() - t0, float(eps.abs().mean()))
    torch.manual_seed(7)
    t0 = time.time()
    u, s, vT = e.local_encoder_decoder_pullback_xt(x=xt, t=t40, op="mid", block_idx=0, pca_rank=2,
                                                   min_iter=10 ** 6, max_iter=1, convergence_threshold=1e-4,
                                                   mask=mask)
    print("ddpm power iteration", time.time() - t0, s.tolist())
    torch.save({"t": t40, "eps": eps.half(), "s": s.clone(), "vT": vT.clone(), "v0_seed": 7, "input_seed": 0,
                "weights_seed": 1234}, os.path.join(HERE, "full256_ddpm.pt"))
    del unet, e
    # ---- P2-256 ----
    sd2 = random_state_dict(P2_256, seed=1234)
    unet2 = ref_p2_unet(su, P2_256, sd2)
    sch = uu.YHCustomScheduler(mg.ns(noise_schedule=None, device=torch.device("cpu"), dtype=torch.float32))
    sch.set_timesteps(100)
    t79 = sch.timesteps[79]
    t0 = time.time()
    with torch.no_grad():
        eps2 = unet2(xt, t79)
    print("p2 forward", time.time() - t0, float(eps2.abs().mean()))
    torch.save({"t": t79, "eps": eps2.half(), "input_seed": 0, "weights_seed": 1234},
               os.path.join(HERE, "full256_p2.pt"))


if __name__ == "__main__":
    main()
