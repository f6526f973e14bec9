"""Diagnostic ladder for the tcgen05 conv kernel (run on the GPU box; not a pytest file).
Each rung runs in its own process so a trapped kernel does not poison the rest.
    python tests/diag_conv.py            # runs all rungs, prints one line each
"""
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

RUNGS = [
    # name, kind, N, H, W, Cin, Cout, identity
    ("1x1_ident_1tile", 1, 1, 8, 16, 32, 32, True),
    ("1x1_ident_k64", 1, 1, 8, 16, 64, 64, True),
    ("1x1_rand_128", 1, 1, 8, 16, 128, 128, False),
    ("1x1_rand_2tiles_n", 1, 1, 8, 16, 128, 256, False),
    ("1x1_rand_manytiles", 1, 4, 32, 32, 128, 128, False),
    ("3x3_ident", 0, 1, 8, 16, 32, 32, True),
    ("3x3_rand", 0, 2, 16, 16, 128, 128, False),
    ("3x3_rand_8x8_N3", 0, 3, 8, 8, 128, 128, False),
    ("3x3_s2", 2, 2, 32, 32, 128, 128, False),
    ("3x3_dgrad", 3, 2, 16, 16, 128, 128, False),
    ("3x3_s2_dgrad", 4, 2, 16, 16, 128, 128, False),
    ("3x3_big", 0, 6, 256, 256, 128, 128, False),
]


def run_rung(name):
    import torch
    import torch.nn.functional as F
    from gpu_util import nchw, nhwc, tf32_round
    from loco_edit_b200 import ops
    from test_gpu_layers import _conv_ref
    spec = [r for r in RUNGS if r[0] == name][0]
    _, kind, N, H, W, Cin, Cout, ident = spec
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    ksz = 1 if kind == 1 else 3
    cx = Cin if kind in (0, 1, 2) else Cout
    if ident:
        w = torch.zeros(Cout, Cin, ksz, ksz)
        for i in range(min(Cin, Cout)):
            w[i, i, ksz // 2, ksz // 2] = 1.0
        p = torch.arange(H * W, dtype=torch.float32).reshape(1, 1, H, W)
        c = torch.arange(cx, dtype=torch.float32).reshape(1, cx, 1, 1)
        x = (p + c / 64.0).expand(N, cx, H, W).contiguous()
    else:
        w = tf32_round(torch.randn(Cout, Cin, ksz, ksz, generator=g) / (Cin * ksz * ksz) ** 0.5)
        x = tf32_round(torch.randn(N, cx, H, W, generator=g))
    x, w = x.to(dev), w.to(dev)
    ref = _conv_ref(kind, x, w).float()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    y = ops.conv2d_nhwc(kind, nhwc(x), w)
    torch.cuda.synchronize()
    t0.record()
    y = ops.conv2d_nhwc(kind, nhwc(x), w)
    t1.record()
    torch.cuda.synchronize()
    yn = nchw(y)
    err = float((yn - ref).abs().max())
    rel = float((yn - ref).norm() / (ref.norm() + 1e-30))
    print(f"RUNG {name}: max_err={err:.3e} rel={rel:.3e} ms={t0.elapsed_time(t1):.3f} "
          f"finite={bool(torch.isfinite(yn).all())}")
    if rel > 1e-4 and ident:
        # show where rows went: for each output pixel (channel 0) which input pixel value arrived
        got = yn[0, 0].reshape(-1)[:32].tolist()
        print("   first 32 outputs ch0:", [round(v, 2) for v in got])
        got = yn[0, :8, 0, 1].tolist()
        print("   pixel 1, ch0..7:", [round(v, 3) for v in got])


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_rung(sys.argv[1])
        sys.exit(0)
    for r in RUNGS:
        try:
            out = subprocess.run([sys.executable, __file__, r[0]], capture_output=True, text=True, timeout=180)
            txt = (out.stdout + out.stderr).strip().splitlines()
            keep = [l for l in txt if l.startswith("RUNG") or l.startswith("   ")]
            if not keep:
                keep = ["RUNG %s: FAILED rc=%d :: %s" % (r[0], out.returncode, " | ".join(txt[-4:]))]
            print("\n".join(keep), flush=True)
        except subprocess.TimeoutExpired:
            print("RUNG %s: TIMEOUT" % r[0], flush=True)
