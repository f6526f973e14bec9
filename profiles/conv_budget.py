"""Per-layer conv time budget of one DDPM-256 forward at batch N (device ms, CUDA events)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from loco_edit_b200 import _lib
from loco_edit_b200._lib import check, ptr, stream_ptr
lib = _lib.load(); dev = torch.device("cuda:0")
# (kind, res, Cin, Cout, count) of the GEMM convs of one forward
LAYERS = [
    (0, 256, 128, 128, 8), (0, 256, 256, 128, 3), (1, 256, 256, 128, 3),
    (0, 128, 128, 128, 7), (0, 128, 256, 128, 2), (0, 128, 384, 128, 1), (0, 128, 256, 256, 1), (1, 128, 256, 128, 2), (1, 128, 384, 128, 1), (2, 256, 128, 128, 1),
    (0, 64, 256, 256, 7), (0, 64, 512, 256, 2), (0, 64, 384, 256, 1), (0, 64, 128, 256, 1), (1, 64, 512, 256, 2), (1, 64, 384, 256, 1), (1, 64, 128, 256, 1), (2, 128, 128, 128, 1),
    (0, 32, 256, 256, 7), (0, 32, 512, 256, 2), (0, 32, 768, 256, 1), (0, 32, 512, 512, 1), (1, 32, 512, 256, 2), (1, 32, 768, 256, 1), (2, 64, 256, 256, 1),
    (0, 16, 512, 512, 7), (0, 16, 1024, 512, 2), (0, 16, 768, 512, 1), (0, 16, 256, 512, 1), (1, 16, 512, 1536, 5), (1, 16, 512, 512, 5), (1, 16, 1024, 512, 2), (1, 16, 768, 512, 1), (1, 16, 256, 512, 1), (2, 32, 256, 256, 1),
    (0, 8, 512, 512, 11), (0, 8, 1024, 512, 3), (1, 8, 1024, 512, 3), (1, 8, 512, 1536, 1), (1, 8, 512, 512, 1), (2, 16, 512, 512, 1),
]
for N in [int(a) for a in sys.argv[1:]] or [8, 11, 40]:
    tot = {}; totfl = 0.0; tot_ms = 0.0
    for kind, res, Cin, Cout, cnt in LAYERS:
        ksz = 1 if kind == 1 else 3
        x = torch.randn(N, res, res, Cin, device=dev)
        wp = torch.randn(Cout * Cin * ksz * ksz, device=dev) * 0.01
        ro = res // 2 if kind == 2 else res
        y = torch.empty(N, ro, ro, Cout, device=dev)
        scr = torch.zeros(32 << 20, dtype=torch.uint8, device=dev)
        ms, ks, gr = C.c_float(), C.c_int(), C.c_int()
        check(lib.loco_conv_bench(kind, ptr(x), N, res, res, Cin, ptr(wp), Cout, Cin, ptr(y), ptr(scr), scr.numel(),
                                  16, 20, C.byref(ms), C.byref(ks), C.byref(gr), stream_ptr()), "bench")
        fl = 2.0 * N * ro * ro * Cout * Cin * ksz * ksz
        tot.setdefault(ro, [0.0, 0.0]); tot[ro][0] += ms.value * cnt; tot[ro][1] += fl * cnt
        totfl += fl * cnt; tot_ms += ms.value * cnt
        del x, y, wp
    print(f"N={N}: total conv {tot_ms:.2f} ms, {totfl/tot_ms/1e9:.0f} TFLOP/s")
    for r in sorted(tot, reverse=True):
        print(f"   out {r:3d}^2: {tot[r][0]:6.2f} ms ({100*tot[r][0]/tot_ms:4.1f}%)  {tot[r][1]/tot[r][0]/1e9:5.0f} TFLOP/s")
