"""One GroupNorm+SiLU site, back-to-back launches (for ncu captures): [N, H, W, C] with 1 primal row
and N-1 tangent rows (N >= 2), or a plain forward batch (--fwd).
usage: python profiles/gn_one.py N H C [--fwd]"""
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from loco_edit_b200 import ops

N, H, C = [int(a) for a in sys.argv[1:4]]
fwd = "--fwd" in sys.argv
dev = torch.device("cuda:0")
x = torch.randn(N, H, H, C, device=dev)
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
for _ in range(3):
    ops.groupnorm_silu_fwd(x, N if fwd else 1, gamma, beta, 1e-6, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    ops.groupnorm_silu_fwd(x, N if fwd else 1, gamma, beta, 1e-6, True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
byts = x.numel() * 4 * 3          # statistics pass reads once, apply pass reads + writes
print(f"GroupNorm+SiLU {'fwd' if fwd else 'jvp'} [{N},{H},{H},{C}]: {ms*1e3:.1f} us (stats memset + stats + apply), "
      f"{byts/ms/1e6:.0f} GB/s algorithmic")
