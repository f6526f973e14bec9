"""One GroupNorm+SiLU site, a few back-to-back launches of each pass (for ncu captures).
usage: python profiles/gn_one.py N H C [fwd|jvp|vjp] [f32]
  fwd: N primal rows (forward-only programs); jvp: 1 primal + N-1 tangent rows; vjp: N cotangent rows
  (plain, then with addend + accumulate)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from loco_edit_b200 import ops

N, H, C = [int(a) for a in sys.argv[1:4]]
mode = sys.argv[4] if len(sys.argv) > 4 else "jvp"
dt = torch.float32 if "f32" in sys.argv else torch.float16
dev = torch.device("cuda:0")
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
x = torch.randn(N, H, H, C, device=dev).to(dt)
y = torch.empty_like(x)
for _ in range(4):
    if mode == "vjp":
        xp = x[:1].contiguous()
        _, st = ops.groupnorm_silu_vjp_ex(xp, x, gamma, beta, 1e-6, True, gx=y)
        ops.groupnorm_silu_vjp_ex(xp, x, gamma, beta, 1e-6, True, addend=x, accumulate=True, gx=y, stats=st, stages=2)
    else:
        ops.groupnorm_silu_fwd_ex(x, N if mode == "fwd" else 1, gamma, beta, 1e-6, True, y=y)
torch.cuda.synchronize()
print("done")
