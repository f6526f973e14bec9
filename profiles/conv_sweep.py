"""Conv kernel micro-benchmark sweep (device ms per launch, back-to-back launches, CUDA events)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from loco_edit_b200 import _lib
from loco_edit_b200._lib import check, ptr, stream_ptr

lib = _lib.load()
dev = torch.device("cuda:0")
CASES = [
    # kind N H W Cin Cout
    (1, 1, 8, 16, 32, 128), (1, 1, 16, 16, 512, 512), (1, 6, 16, 16, 512, 512), (1, 6, 16, 16, 512, 1536),
    (0, 1, 16, 16, 512, 512), (0, 6, 16, 16, 512, 512), (0, 6, 16, 16, 1024, 512), (0, 6, 8, 8, 512, 512),
    (0, 1, 8, 8, 512, 512), (0, 6, 32, 32, 256, 256), (0, 1, 32, 32, 256, 256), (0, 6, 64, 64, 256, 256),
    (0, 1, 64, 64, 256, 256), (0, 1, 128, 128, 128, 128), (0, 6, 128, 128, 128, 128), (0, 1, 256, 256, 128, 128),
    (0, 5, 256, 256, 128, 128), (0, 6, 256, 256, 128, 128), (0, 6, 256, 256, 256, 128), (1, 6, 256, 256, 256, 128),
    # batch-edit DDIM passes (B = 8 inversion / forward, B = 40 edited latents)
    (0, 8, 256, 256, 128, 128), (0, 40, 256, 256, 128, 128), (0, 8, 256, 256, 256, 128), (0, 8, 128, 128, 128, 128),
    (0, 40, 128, 128, 128, 128), (0, 8, 128, 128, 256, 256), (0, 8, 64, 64, 256, 256), (0, 40, 64, 64, 256, 256),
    (0, 8, 64, 64, 512, 256), (0, 8, 32, 32, 256, 256), (0, 40, 32, 32, 256, 256), (0, 8, 32, 32, 512, 512),
    (0, 8, 16, 16, 512, 512), (0, 40, 16, 16, 512, 512), (0, 8, 16, 16, 1024, 512), (0, 8, 8, 8, 512, 512),
    (0, 40, 8, 8, 512, 512), (0, 8, 8, 8, 1024, 512), (1, 8, 16, 16, 512, 1536), (1, 40, 16, 16, 512, 512),
]
print("kind N HxW Cin->Cout | ksplit grid ms TFLOP/s | (no split) ms")
for kind, N, H, W, Cin, Cout in CASES:
    ksz = 1 if kind == 1 else 3
    x = torch.randn(N, H, W, Cin, device=dev)
    wp = torch.randn(Cout * Cin * ksz * ksz, device=dev) * 0.01
    y = torch.empty(N, H, W, Cout, device=dev)
    scr = torch.zeros(32 << 20, dtype=torch.uint8, device=dev)
    res = []
    for mk in (16, 1):
        ms, ks, gr = C.c_float(), C.c_int(), C.c_int()
        check(lib.loco_conv_bench(kind, ptr(x), N, H, W, Cin, ptr(wp), Cout, Cin, ptr(y), ptr(scr), scr.numel(),
                                  mk, 50, C.byref(ms), C.byref(ks), C.byref(gr), stream_ptr()), "bench")
        res.append((ks.value, gr.value, ms.value))
    fl = 2.0 * N * H * W * Cout * Cin * ksz * ksz
    print(f"{kind} {N} {H}x{W} {Cin}->{Cout} | {res[0][0]} {res[0][1]} {res[0][2]*1e3:.1f}us {fl/res[0][2]/1e9:.0f} | "
          f"{res[1][2]*1e3:.1f}us {fl/res[1][2]/1e9:.0f}")
