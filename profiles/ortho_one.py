"""One orthonormalisation at rank K (argv[1], default 64) for an ncu launch list."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from loco_edit_b200 import ops

k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
d = 3 * 256 * 256
g = torch.Generator(device=dev).manual_seed(k)
W = torch.randn(k, d, device=dev, generator=g)
Vp = torch.linalg.qr(torch.randn(d, k, device=dev, generator=g))[0].T.contiguous()
ops.orthonormalise(W, v_prev=Vp)
torch.cuda.synchronize()
