"""Device time of the attention core (DDPM-256 site: 256 tokens x 512 channels, one head) through the
C-ABI entry points: forward batch (B rows), forward + k tangent rows, VJP of k cotangent rows.
CUDA events around REPS back-to-back calls (operands L2-resident, as in the U-Net programs where the
q|k|v projection has just been written).  LOCO_ATTN_TC=0 selects the CUDA-core path for comparison.
    python profiles/attn_bench.py
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from loco_edit_b200 import ops

dev = torch.device("cuda:0")
REPS = int(os.environ.get("REPS", "20"))
T, C = int(os.environ.get("T", "256")), int(os.environ.get("C", "512"))
HC = int(os.environ.get("HEAD_CH", "0"))


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
    return e0.elapsed_time(e1) / REPS * 1e3


for rows, npr in [(1, 1), (8, 8), (40, 40), (6, 1), (11, 1)]:
    qkv = torch.randn(rows, T, 3 * C, device=dev)
    us = timed(lambda: ops.attention_fwd(qkv, npr, head_ch=HC))
    fl = 4.0 * T * T * C * (npr + 2 * (rows - npr))
    print(f"fwd rows={rows} primal={npr}: {us:7.1f} us  {fl/us/1e6:6.1f} TFLOP/s")
for k in (5, 10):
    qkv = torch.randn(1, T, 3 * C, device=dev)
    _, S = ops.attention_fwd(qkv, 1, head_ch=HC)
    P0 = (S[0] if HC else S[0]).contiguous()
    go = torch.randn(k, T, C, device=dev)
    us = timed(lambda: ops.attention_vjp(go, qkv, P0, head_ch=HC))
    print(f"vjp k={k}: {us:7.1f} us  {8.0*T*T*C*k/us/1e6:6.1f} TFLOP/s")
