"""One conv shape, back-to-back launches (for ncu captures and LOCO_CONV_DEBUG decompositions).
usage: [IN16=1] [OUT16=1] [ADD=1] [STATS=1] python profiles/conv_one.py kind N H W Cin Cout [reps]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from loco_edit_b200 import _lib
from loco_edit_b200._lib import check, ptr, stream_ptr

kind, N, H, W, Cin, Cout = [int(a) for a in sys.argv[1:7]]
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 50
in16, out16 = int(os.environ.get("IN16", "0")), int(os.environ.get("OUT16", "0"))
lib = _lib.load()
dev = torch.device("cuda:0")
ksz = 1 if kind == 1 else 3
tin = torch.float16 if in16 else torch.float32
tout = torch.float16 if out16 else torch.float32
x = torch.randn(N, H, W, Cin, device=dev).to(tin)
wp = (torch.randn(Cout * Cin * ksz * ksz, device=dev) * 0.01).to(tin)
y = torch.empty(N, H, W, Cout, device=dev, dtype=tout)
add = torch.randn(N, H, W, Cout, device=dev).to(tout) if int(os.environ.get("ADD", "0")) else None
st = torch.zeros(64 * N, dtype=torch.float64, device=dev) if int(os.environ.get("STATS", "0")) else None
scr = torch.zeros(32 << 20, dtype=torch.uint8, device=dev)
ms, ks, gr = C.c_float(), C.c_int(), C.c_int()
check(lib.loco_conv_bench_ex(kind, ptr(x), N, H, W, Cin, ptr(wp), Cout, Cin, ptr(y), ptr(scr), scr.numel(),
                             16, reps, in16, out16, ptr(add), ptr(st), C.byref(ms), C.byref(ks), C.byref(gr),
                             stream_ptr()), "bench")
fl = 2.0 * N * H * W * Cout * Cin * ksz * ksz
print(f"debug={os.environ.get('LOCO_CONV_DEBUG', '0')} nt={os.environ.get('LOCO_CONV_NT', '-')} in16={in16} out16={out16} "
      f"add={int(add is not None)} stats={int(st is not None)} kind={kind} N={N} {H}x{W} {Cin}->{Cout}: "
      f"{ms.value*1e3:.1f} us  {fl/ms.value/1e9:.0f} TFLOP/s  grid={gr.value}")
