import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import torch
from loco_edit_b200 import ops
dev = torch.device("cuda:0")
def tf32(x): return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
for T, C in [(64, 128), (256, 512)]:
    g = torch.Generator().manual_seed(1)
    qkv = tf32(torch.randn(2, T, 3 * C, generator=g)).to(dev)
    q, k, v = qkv[..., :C].double(), qkv[..., C:2 * C].double(), qkv[..., 2 * C:].double()
    Sref = torch.softmax(torch.bmm(q, k.transpose(1, 2)) * C ** -0.5, 2)
    oref = torch.bmm(Sref, v)
    o, S = ops.attention_fwd(qkv, 2)
    torch.cuda.synchronize()
    print(T, C, "S err", rel(S, Sref), "o err", rel(o, oref), "o absmax", float(o.abs().max()), "S absmax", float(S.abs().max()),
          "S rowsum", float(S.sum(-1).mean()))
    print(" o[0,0,:8]", o[0, 0, :8].tolist()); print(" oref    ", oref[0, 0, :8].tolist())
    print(" o[0,T-1,-4:]", o[0, T - 1, -4:].tolist(), oref[0, T - 1, -4:].tolist())
