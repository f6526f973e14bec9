"""Diagnostic: fp16 vs tf32 Jacobian programs on the tiny architectures (JVP / VJP differences and the
adjoint identity), several seeds."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests")))
import torch
from loco_edit_b200.unet import B200UNet
from loco_edit_b200.weights import random_state_dict, tiny_arch
dev = torch.device("cuda:0")
arch = tiny_arch(resolution=32, ch_mult=(1, 2), attn_resolutions=(16,), num_res_blocks=1)
sd = random_state_dict(arch, seed=1234, perturb_norm=0.1)
net = B200UNet(arch, sd, device=dev)
k = 3
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
for seed in range(4):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, 3, 32, 32, generator=g).to(dev)
    V = torch.randn(k, 3, 32, 32, generator=g).to(dev)
    G = torch.randn(k, 3, 32, 32, generator=g).to(dev)
    res = {}
    for half in (False, True):
        p = net.plan(1, k, k, half=half)
        out = p.forward(torch.cat([x, V]).contiguous(), 595.3636)
        gx = p.vjp(G.contiguous())
        res[half] = (out[:1].clone(), out[1:].clone(), gx.clone())
    e32, d32, g32 = res[False]; e16, d16, g16 = res[True]
    adj = lambda d, gx: abs(float((d.double() * G.double()).sum() - (V.double() * gx.double()).sum())) / abs(float((d.double() * G.double()).sum()))
    print(f"seed {seed}: eps16 vs eps32 {rel(e16, e32):.2e}  jvp16 vs jvp32 {rel(d16, d32):.2e}  vjp16 vs vjp32 {rel(g16, g32):.2e}  "
          f"adjoint32 {adj(d32, g32):.2e} adjoint16 {adj(d16, g16):.2e} mixed(jvp16,vjp32) {adj(d16, g32):.2e} (jvp32,vjp16) {adj(d32, g16):.2e}  "
          f"<Jv,g>/(|Jv||g|) {float((d32.double()*G.double()).sum()/(d32.double().norm()*G.double().norm())):.2e}")
