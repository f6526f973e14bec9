"""CPU experiment: how much does a 16-bit operand / storage type cost the FINAL
DDIM stage (59 steps from t = 0.6T, eta = 1 from index 79), the only stage whose output the
north_star bounds (edited-image PSNR >= 40 dB)?

The CPU oracle's U-Net is run with every convolution's operands (activations and weights) and every
stored block output rounded to tf32 (what the CUDA path does today), fp16 (same 10-bit mantissa,
5-bit exponent) or bf16 (7-bit mantissa), from identical x_t and identical injected noise; PSNR is
against the fp32 chain.

    python profiles/precision_experiment.py [resolution=64] [batch=5]
"""
import math
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from loco_edit_b200.weights import DDPM256, random_state_dict  # noqa: E402
from oracle import ddpm_ref, pullback_ref  # noqa: E402


def q_tf32(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


QS = {"fp32": lambda x: x, "tf32": q_tf32, "fp16": lambda x: x.half().float(), "bf16": lambda x: x.bfloat16().float()}


def make_unet(arch, sd, q):
    orig_conv, orig_res, orig_attn = ddpm_ref.conv, ddpm_ref.resnet_block, ddpm_ref.attn_block
    sdq = {k: (q(v) if (k.endswith(".weight") and v.dim() == 4) else v) for k, v in sd.items()}

    def conv(sd_, p, x, stride=1, padding=0):
        return q(F.conv2d(q(x), sd_[p + ".weight"], sd_[p + ".bias"], stride=stride, padding=padding))

    def call(x, t):
        ddpm_ref.conv = conv
        ddpm_ref.resnet_block = lambda *a, **k: q(orig_res(*a, **k))
        ddpm_ref.attn_block = lambda *a, **k: q(orig_attn(*a, **k))
        try:
            return ddpm_ref.unet_forward(sdq, arch, x, t)
        finally:
            ddpm_ref.conv, ddpm_ref.resnet_block, ddpm_ref.attn_block = orig_conv, orig_res, orig_attn

    return call


def psnr(a, b):
    return 10 * math.log10(4.0 / float(((a.double() - b.double()) ** 2).mean()))


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    torch.set_num_threads(int(os.environ.get("THREADS", "4")))
    arch = dict(DDPM256, resolution=R)
    sd = random_state_dict(arch, seed=1234)
    g = torch.Generator().manual_seed(0)
    xt = torch.randn(1, 3, R, R, generator=g)
    v = torch.randn(3 * R * R, generator=g)
    v = v / v.norm()
    batch = pullback_ref.edit_batch(xt, v, 0.5, 16, 2)[:B]
    noises = {79 + i: torch.randn(B, 3, R, R, generator=g) for i in range(20)}
    out = {}
    for name, q in QS.items():
        t0 = time.time()
        sched = pullback_ref.RefScheduler()
        out[name] = pullback_ref.ddim_forward(make_unet(arch, sd, q), sched, batch, 40, -1, boost_idx=79, noises=noises)
        msg = f"{name}: {time.time() - t0:.0f} s"
        if name != "fp32":
            msg += (f", PSNR vs fp32 {psnr(out[name], out['fp32']):.1f} dB, rel L2 "
                    f"{float((out[name] - out['fp32']).norm() / out['fp32'].norm()):.2e}")
        print(f"final stage {R}x{R}, {B} latents, 59 steps: " + msg, flush=True)
    print(f"fp16 vs tf32 chain: PSNR {psnr(out['fp16'], out['tf32']):.1f} dB")


if __name__ == "__main__":
    main()
