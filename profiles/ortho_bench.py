"""Device time of the replicated orthonormalisation (Gram -> Jacobi -> transform, two passes) by rank."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from loco_edit_b200 import ops

dev = torch.device("cuda:0")
d = 3 * 256 * 256
for k in (5, 10, 16, 32, 64):
    g = torch.Generator(device=dev).manual_seed(k)
    W = torch.randn(k, d, device=dev, generator=g)
    Vp = torch.linalg.qr(torch.randn(d, k, device=dev, generator=g))[0].T.contiguous()
    for _ in range(3):
        ops.orthonormalise(W, v_prev=Vp)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        V, s = ops.orthonormalise(W, v_prev=Vp)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 100
    print("k=%2d  orthonormalise %8.1f us  (%.1f GB/s of 3 k d 4 B; Gram FLOPs 3 x 2 k^2 d = %.2f GFLOP)" %
          (k, us, 3 * k * d * 4 / us / 1e3, 3 * 2 * k * k * d / 1e9))
