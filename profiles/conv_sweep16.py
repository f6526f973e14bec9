"""fp16 conv micro-benchmark sweep of the <= 64^2 layer shapes (device us per launch, 50 back-to-back
launches, CUDA events, through loco_conv_bench_ex): default split-K vs no split-K.  Run under
LOCO_CONV_NT=1|2 / LOCO_CONV_PAIR=0 to compare the kernel variants."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from loco_edit_b200 import _lib
from loco_edit_b200._lib import check, ptr, stream_ptr

lib = _lib.load()
dev = torch.device("cuda:0")
SHAPES = [(0, 64, 64, 256, 256), (0, 64, 64, 512, 256), (0, 32, 32, 256, 256), (0, 32, 32, 512, 512), (0, 32, 32, 768, 256),
          (0, 16, 16, 512, 512), (0, 16, 16, 1024, 512), (0, 8, 8, 512, 512), (0, 8, 8, 1024, 512),
          (1, 16, 16, 512, 1536), (1, 16, 16, 512, 512), (1, 32, 32, 512, 256)]
NS = [int(v) for v in os.environ.get("NS", "1,6,8,11,40").split(",")]
print("variant cap", os.environ.get("LOCO_CONV_NT", "-"), "pair", os.environ.get("LOCO_CONV_PAIR", "-"))
print("kind HxW Cin->Cout N | ksplit grid us TFLOP/s | (no split) us")
for kind, H, W, Cin, Cout in SHAPES:
    for N in NS:
        ksz = 1 if kind == 1 else 3
        x = torch.randn(N, H, W, Cin, device=dev).half()
        wp = (torch.randn(Cout * Cin * ksz * ksz, device=dev) * 0.01).half()
        y = torch.empty(N, H, W, Cout, device=dev, dtype=torch.float16)
        scr = torch.zeros(32 << 20, dtype=torch.uint8, device=dev)
        res = []
        for mk in (16, 1):
            ms, ks, gr = C.c_float(), C.c_int(), C.c_int()
            check(lib.loco_conv_bench_ex(kind, ptr(x), N, H, W, Cin, ptr(wp), Cout, Cin, ptr(y), ptr(scr), scr.numel(),
                                         mk, 50, 1, 1, None, None, C.byref(ms), C.byref(ks), C.byref(gr),
                                         stream_ptr()), "bench")
            res.append((ks.value, gr.value, ms.value))
        fl = 2.0 * N * H * W * Cout * Cin * ksz * ksz
        print(f"{kind} {H}x{W} {Cin}->{Cout} N={N} | {res[0][0]} {res[0][1]} {res[0][2]*1e3:.1f}us {fl/res[0][2]/1e9:.0f} | "
              f"{res[1][2]*1e3:.1f}us")
