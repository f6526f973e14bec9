"""Achieved HBM bandwidth of the GroupNorm kernels (the bandwidth-bound family of the U-Net programs)
on the sites of the DDPM-256 bench, fp16 storage, through the typed C-ABI entry points
(`loco_groupnorm_silu_fwd_ex` / `_vjp_ex`, one pass per call).  CUDA events around `REPS` back-to-back
launches per case; the tensors of the large sites exceed the 126 MB L2.
    python profiles/gn_bench.py            # prints one line per (kernel, site)
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from loco_edit_b200 import ops

dev = torch.device("cuda:0")
REPS = int(os.environ.get("REPS", "20"))
PEAK = 6551.4


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REPS * 1e3   # us


def report(name, shape, us, nbytes):
    print(f"{name:28s} {str(shape):24s} {us:8.1f} us  {nbytes/us/1e3:7.0f} GB/s  {nbytes/us/1e3/PEAK:5.2f} of HBM peak")


sites = [(256, 128), (128, 128), (64, 256), (32, 256), (16, 512)]
if os.environ.get("SITES"):      # e.g. SITES=248x128,256x128
    sites = [tuple(int(v) for v in t.split("x")) for t in os.environ["SITES"].split(",")]
dt = torch.float32 if os.environ.get("HALF", "1") == "0" else torch.float16
es = 4 if dt == torch.float32 else 2
for H, C in sites:
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    # forward-only batch (B = 8 and B = 40): apply pass only (statistics come from the conv epilogue)
    for B in (8, 40):
        x = torch.randn(B, H, H, C, device=dev).to(dt)
        y = torch.empty_like(x)
        _, st = ops.groupnorm_silu_fwd_ex(x, B, gamma, beta, 1e-6, True, y=y)
        us = timed(lambda: ops.groupnorm_silu_fwd_ex(x, B, gamma, beta, 1e-6, True, y=y, stats=st, stages=2))
        report("fwd apply", tuple(x.shape), us, 2 * x.numel() * es)
        del x, y
    # JVP rows (1 primal + k tangents), k = 5 and 10
    for k in (5, 10):
        x = torch.randn(1 + k, H, H, C, device=dev).to(dt)
        y = torch.empty_like(x)
        _, st = ops.groupnorm_silu_fwd_ex(x, 1, gamma, beta, 1e-6, True, y=y)
        us = timed(lambda: ops.groupnorm_silu_fwd_ex(x, 1, gamma, beta, 1e-6, True, y=y, stats=st, stages=1))
        report("jvp stats (+memset)", tuple(x.shape), us, x.numel() * es)
        us = timed(lambda: ops.groupnorm_silu_fwd_ex(x, 1, gamma, beta, 1e-6, True, y=y, stats=st, stages=2))
        report("jvp apply", tuple(x.shape), us, 2 * x.numel() * es)
        # VJP rows: k cotangents at the primal point x[:1]
        gy = torch.randn(k, H, H, C, device=dev).to(dt)
        add = torch.randn(k, H, H, C, device=dev).to(dt)
        gx = torch.zeros_like(gy)
        xp = x[:1].contiguous()
        _, st = ops.groupnorm_silu_vjp_ex(xp, gy, gamma, beta, 1e-6, True, gx=gx)
        us = timed(lambda: ops.groupnorm_silu_vjp_ex(xp, gy, gamma, beta, 1e-6, True, gx=gx, stats=st, stages=1))
        report("vjp stats (+primal stats)", tuple(gy.shape), us, (gy.numel() + 2 * xp.numel()) * es)
        us = timed(lambda: ops.groupnorm_silu_vjp_ex(xp, gy, gamma, beta, 1e-6, True, gx=gx, stats=st, stages=2))
        report("vjp apply", tuple(gy.shape), us, (2 * gy.numel() + xp.numel()) * es)
        us = timed(lambda: ops.groupnorm_silu_vjp_ex(xp, gy, gamma, beta, 1e-6, True, addend=add, accumulate=True,
                                                    gx=gx, stats=st, stages=2))
        report("vjp apply +addend +acc", tuple(gy.shape), us, (4 * gy.numel() + xp.numel()) * es)
        del x, y, gy, add, gx
