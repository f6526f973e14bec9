"""GPU parity of single kernels (through the C ABI) against plain fp32/fp64 PyTorch references of
the same op.  Tolerances are stated per test:
  * tensor-core convs on tf32-representable inputs: products are exact, accumulation fp32 ->
    relative error 1e-5 of the output norm;
  * on arbitrary fp32 inputs the tensor core drops 13 mantissa bits of each operand -> 2e-3;
  * CUDA-core kernels: fp32 round-off, 1e-5 relative;
  * integer / index work and the DDIM update: bit-exact.
"""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import max_err, nchw, nhwc, rel_err, tf32_round

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


def _conv_ref(kind, x_nchw, w, bias=None):
    xd, wd = x_nchw.double(), w.double()
    b = bias.double() if bias is not None else None
    if kind == 0:
        return F.conv2d(xd, wd, b, padding=1)
    if kind == 1:
        return F.conv2d(xd, wd, b)
    if kind == 2:
        return F.conv2d(F.pad(xd, (0, 1, 0, 1)), wd, b, stride=2)
    if kind == 3:   # data gradient of a 3x3 pad-1 conv
        return F.conv_transpose2d(xd, wd, padding=1)
    if kind == 4:   # data gradient of pad(0,1,0,1) + stride-2 conv
        N, _, h, w_ = xd.shape
        full = F.conv_transpose2d(xd, wd, stride=2)          # [N, Cin, 2h+1, 2w+1]
        return full[:, :, : 2 * h, : 2 * w_]
    raise ValueError(kind)


CONV_CASES = [
    # kind, N, H, W, Cin, Cout
    (0, 1, 16, 16, 128, 128),
    (0, 6, 32, 32, 128, 128),
    (0, 3, 8, 8, 256, 128),      # two samples per tile, odd batch
    (0, 2, 64, 64, 256, 256),    # two output-channel tiles
    (0, 5, 16, 16, 384, 128),
    (1, 6, 32, 32, 256, 128),
    (1, 2, 16, 16, 128, 384),    # fused q|k|v shape
    (2, 4, 32, 32, 128, 128),
    (2, 1, 16, 16, 256, 256),
    (3, 5, 32, 32, 128, 256),    # dy has Cout=256 channels -> dx has 128
    (3, 2, 8, 8, 256, 128),
    (4, 3, 16, 16, 128, 128),    # dy [3,16,16,128] -> dx [3,32,32,128]
    (4, 5, 8, 8, 256, 256),
    # >= 4 waves of tiles: the operand-swapped "wide" kernel (two pixel tiles per 128x256x8 MMA)
    (0, 6, 128, 128, 128, 128),
    (3, 5, 128, 128, 128, 128),
    (1, 6, 128, 128, 256, 128),
    (0, 5, 128, 128, 256, 256),
]


@pytest.mark.parametrize("kind,N,H,W,Cin,Cout", CONV_CASES)
def test_conv_tcgen05_exact_inputs(dev, kind, N, H, W, Cin, Cout):
    from loco_edit_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(kind * 1000 + N * 100 + H)
    ksz = 1 if kind == 1 else 3
    w = tf32_round(torch.randn(Cout, Cin, ksz, ksz, generator=g) / (Cin * ksz * ksz) ** 0.5).to(dev)
    cx = Cin if kind in (0, 1, 2) else Cout
    x = tf32_round(torch.randn(N, cx, H, W, generator=g)).to(dev)
    ref = _conv_ref(kind, x, w)
    y = ops.conv2d_nhwc(kind, nhwc(x), w)
    torch.cuda.synchronize()
    e = rel_err(nchw(y), ref)
    print(f"conv kind={kind} N={N} {H}x{W} {Cin}->{Cout}: rel_err={e:.3e} max={max_err(nchw(y), ref):.3e}")
    assert e < 1e-5


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(4, 16, 16, 128, 256), (5, 128, 128, 128, 128)])
def test_conv_epilogue_bias_addend_accumulate(dev, N, H, W, Cin, Cout):
    from loco_edit_b200 import ops
    g = torch.Generator().manual_seed(5)
    w = tf32_round(torch.randn(Cout, Cin, 3, 3, generator=g) / 34.0).to(dev)
    x = tf32_round(torch.randn(N, Cin, H, W, generator=g)).to(dev)
    bias = torch.randn(Cout, generator=g).to(dev)
    add = torch.randn(N, Cout, H, W, generator=g).to(dev)
    base = torch.randn(N, Cout, H, W, generator=g).to(dev)
    ref = F.conv2d(x.double(), w.double(), padding=1) + add.double() + base.double()
    ref[:2] += bias.double()[None, :, None, None]                  # bias on the first 2 rows only
    out = nhwc(base).clone()
    ops.conv2d_nhwc(0, nhwc(x), w, bias=bias, bias_rows=2, addend=nhwc(add), accumulate=True, out=out)
    torch.cuda.synchronize()
    assert rel_err(nchw(out), ref) < 1e-5


def test_conv_tcgen05_fp32_inputs_tf32_tolerance(dev):
    from loco_edit_b200 import ops
    g = torch.Generator().manual_seed(9)
    N, H, W, Cin, Cout = 2, 32, 32, 128, 128
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / 34.0).to(dev)
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    ref = F.conv2d(x.double(), w.double(), padding=1)
    y = ops.conv2d_nhwc(0, nhwc(x), w)
    torch.cuda.synchronize()
    e = rel_err(nchw(y), ref)
    print("tf32 conv on raw fp32 inputs: rel_err", e)
    assert e < 2e-3


def _gn_silu(x, gamma, beta, eps, silu):
    y = F.group_norm(x, 32, gamma, beta, eps)
    return y * torch.sigmoid(y) if silu else y


@pytest.mark.parametrize("C,H,silu", [(128, 32, True), (384, 16, True), (512, 8, False), (768, 8, True)])
def test_groupnorm_silu_fwd_jvp(dev, C, H, silu):
    from loco_edit_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    k = 3
    x = (torch.randn(1, C, H, H, generator=g) * 1.5 + 0.3).to(dev)
    dx = torch.randn(k, C, H, H, generator=g).to(dev)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(dev)
    beta = (0.2 * torch.randn(C, generator=g)).to(dev)
    eps = 1e-6
    f = lambda z: _gn_silu(z, gamma.double(), beta.double(), eps, silu)
    yref = f(x.double())
    dys = [torch.func.jvp(f, (x.double(),), (dx[j:j + 1].double(),))[1] for j in range(k)]
    ref = torch.cat([yref] + dys, 0)
    y = ops.groupnorm_silu_fwd(nhwc(torch.cat([x, dx], 0)), 1, gamma, beta, eps, silu)
    torch.cuda.synchronize()
    e0, e1 = rel_err(nchw(y)[:1], ref[:1]), rel_err(nchw(y)[1:], ref[1:])
    print(f"gn C={C} H={H} silu={silu}: primal {e0:.2e} tangent {e1:.2e}")
    assert e0 < 1e-5 and e1 < 1e-5
    # plain forward over a batch of primal rows
    xb = torch.randn(3, C, H, H, generator=g).to(dev)
    yb = ops.groupnorm_silu_fwd(nhwc(xb), 3, gamma, beta, eps, silu)
    assert rel_err(nchw(yb), f(xb.double())) < 1e-5


@pytest.mark.parametrize("C,H,silu", [(128, 32, True), (512, 8, False), (384, 16, True)])
def test_groupnorm_silu_vjp(dev, C, H, silu):
    from loco_edit_b200 import ops
    g = torch.Generator().manual_seed(C * 3 + H)
    k = 3
    x = (torch.randn(1, C, H, H, generator=g) * 1.5 + 0.3).to(dev)
    gy = torch.randn(k, C, H, H, generator=g).to(dev)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(dev)
    beta = (0.2 * torch.randn(C, generator=g)).to(dev)
    eps = 1e-6
    xd = x.double().requires_grad_(True)
    y = _gn_silu(xd, gamma.double(), beta.double(), eps, silu)
    ref = torch.cat([torch.autograd.grad(y, xd, gy[j:j + 1].double(), retain_graph=True)[0] for j in range(k)], 0)
    gx = ops.groupnorm_silu_vjp(nhwc(x), nhwc(gy), gamma, beta, eps, silu)
    torch.cuda.synchronize()
    e = rel_err(nchw(gx), ref)
    print(f"gn vjp C={C} H={H} silu={silu}: {e:.2e}")
    assert e < 1e-5


@pytest.mark.parametrize("C,H,silu", [(128, 32, True), (256, 16, True), (384, 16, True), (512, 8, False),
                                      (768, 8, True), (1024, 8, True)])
def test_groupnorm_silu_fp16_fwd_jvp_vjp(dev, C, H, silu):
    """The fp16-storage GroupNorm kernels (16-byte vectors; forward-only, JVP and VJP variants,
    incl. addend + accumulate) against fp64 on the same fp16-representable inputs: the only error
    left is the fp16 rounding of the stored result (2^-11 relative per element)."""
    from loco_edit_b200 import ops
    g = torch.Generator().manual_seed(C * 7 + H)
    k = 7                                     # > one register chunk of rows
    h16 = lambda t: t.half().to(dev)
    x = h16(torch.randn(1, C, H, H, generator=g) * 1.5 + 0.3)
    dx = h16(torch.randn(k, C, H, H, generator=g))
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(dev)
    beta = (0.2 * torch.randn(C, generator=g)).to(dev)
    eps = 1e-6
    f = lambda z: _gn_silu(z, gamma.double(), beta.double(), eps, silu)
    yref = f(x.double())
    dys = [torch.func.jvp(f, (x.double(),), (dx[j:j + 1].double(),))[1] for j in range(k)]
    ref = torch.cat([yref] + dys, 0)
    y, _ = ops.groupnorm_silu_fwd_ex(nhwc(torch.cat([x, dx], 0)), 1, gamma, beta, eps, silu)
    assert y.dtype == torch.float16
    e0, e1 = rel_err(nchw(y.float())[:1], ref[:1]), rel_err(nchw(y.float())[1:], ref[1:])
    # forward-only batch (the 16-byte kernel with the folded affine)
    xb = h16(torch.randn(5, C, H, H, generator=g) * 2.0 - 0.5)
    yb, _ = ops.groupnorm_silu_fwd_ex(nhwc(xb), 5, gamma, beta, eps, silu)
    e2 = rel_err(nchw(yb.float()), f(xb.double()))
    # VJP, with and without addend / accumulate
    gy = h16(torch.randn(k, C, H, H, generator=g))
    add = h16(torch.randn(k, C, H, H, generator=g))
    acc = h16(torch.randn(k, C, H, H, generator=g))
    xd = x.double().requires_grad_(True)
    yy = _gn_silu(xd, gamma.double(), beta.double(), eps, silu)
    gref = torch.cat([torch.autograd.grad(yy, xd, gy[j:j + 1].double(), retain_graph=True)[0] for j in range(k)], 0)
    gx, _ = ops.groupnorm_silu_vjp_ex(nhwc(x), nhwc(gy), gamma, beta, eps, silu)
    e3 = rel_err(nchw(gx.float()), gref)
    gx2 = nhwc(acc).clone()
    ops.groupnorm_silu_vjp_ex(nhwc(x), nhwc(gy), gamma, beta, eps, silu, addend=nhwc(add), accumulate=True, gx=gx2)
    e4 = rel_err(nchw(gx2.float()), gref + add.double() + acc.double())
    torch.cuda.synchronize()
    print(f"gn fp16 C={C} H={H} silu={silu}: primal {e0:.2e} tangent {e1:.2e} fwd batch {e2:.2e} vjp {e3:.2e} "
          f"vjp+addend+acc {e4:.2e}")
    assert max(e0, e1, e2, e3, e4) < 6e-4


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("C,H,silu", [(128, 32, True), (128, 64, True), (256, 64, True), (512, 16, False), (1024, 8, True)])
def test_groupnorm_small_site_one_launch_kernel(dev, C, H, silu, dtype):
    """`gn_small_kernel` (stages = 4): statistics + apply of a <= 64^2 site in one launch -- primal row,
    tangent rows (JVP), cotangent rows (VJP, incl. addend + accumulate) against fp64, and against the
    two-launch kernels it replaces; the stored statistics of the primal row are what the VJP pass reads."""
    from loco_edit_b200 import ops
    g = torch.Generator().manual_seed(C * 3 + H)
    k = 5
    cast = lambda t: t.to(dtype).to(dev)
    x = cast(torch.randn(1, C, H, H, generator=g) * 1.5 + 0.3)
    dx = cast(torch.randn(k, C, H, H, generator=g))
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(dev)
    beta = (0.2 * torch.randn(C, generator=g)).to(dev)
    eps = 1e-6
    f = lambda z: _gn_silu(z, gamma.double(), beta.double(), eps, silu)
    ref = torch.cat([f(x.double())] + [torch.func.jvp(f, (x.double(),), (dx[j:j + 1].double(),))[1] for j in range(k)], 0)
    xin = nhwc(torch.cat([x, dx], 0))
    y, st = ops.groupnorm_silu_fwd_ex(xin, 1, gamma, beta, eps, silu, stages=4)
    y2, st2 = ops.groupnorm_silu_fwd_ex(xin, 1, gamma, beta, eps, silu, stages=3)
    tol = 6e-4 if dtype == torch.float16 else 2e-5
    e0, e1 = rel_err(nchw(y.float())[:1], ref[:1]), rel_err(nchw(y.float())[1:], ref[1:])
    assert rel_err(st[:64], st2[:64]) < 1e-6                     # primal (sum x, sum x^2); fp32 partials in another order
    assert rel_err(y.float(), y2.float()) < tol
    gy, add, acc = (cast(torch.randn(k, C, H, H, generator=g)) for _ in range(3))
    xd = x.double().requires_grad_(True)
    yy = _gn_silu(xd, gamma.double(), beta.double(), eps, silu)
    gref = torch.cat([torch.autograd.grad(yy, xd, gy[j:j + 1].double(), retain_graph=True)[0] for j in range(k)], 0)
    gx, _ = ops.groupnorm_silu_vjp_ex(nhwc(x), nhwc(gy), gamma, beta, eps, silu, stages=4)
    e2 = rel_err(nchw(gx.float()), gref)
    gx2 = nhwc(acc).clone()
    ops.groupnorm_silu_vjp_ex(nhwc(x), nhwc(gy), gamma, beta, eps, silu, addend=nhwc(add), accumulate=True, gx=gx2, stages=4)
    e3 = rel_err(nchw(gx2.float()), gref + add.double() + acc.double())
    torch.cuda.synchronize()
    print(f"gn small-site {dtype} C={C} H={H}: primal {e0:.2e} tangent {e1:.2e} vjp {e2:.2e} vjp+addend+acc {e3:.2e}")
    assert max(e0, e1, e2, e3) < tol


def _attn_core(qkv):   # qkv [N, T, 3C] -> o [N, T, C]   (reference ddpm/diffusion.py:950-962)
    C = qkv.shape[-1] // 3
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    w = torch.softmax(torch.bmm(q, k.transpose(1, 2)) * (int(C) ** -0.5), dim=2)
    return torch.bmm(w, v)


def _tf32(x):
    """Round to the nearest tf32-representable fp32 value (what the conv epilogues store)."""
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("T,C", [(64, 128), (256, 512), (128, 256), (256, 128), (1024, 128), (2048, 64)])
def test_attention_fwd_jvp_vjp(dev, T, C):
    """Fused tcgen05 attention (kind::tf32; T <= 256) and the long-sequence path (scores in HBM, products on
    mma.sync tf32: the VAE decoder's 4096-token mid block) against fp64 on tf32-representable operands: the
    products are then exact, what remains is the tf32 rounding of the probabilities / of gS that
    feed the second product and of the stored results (2^-11 relative each)."""
    from loco_edit_b200 import ops
    g = torch.Generator().manual_seed(T + C)
    k = 2
    qkv = _tf32(torch.randn(1, T, 3 * C, generator=g)).to(dev)
    dq = _tf32(torch.randn(k, T, 3 * C, generator=g)).to(dev)
    oref = _attn_core(qkv.double())
    dref = torch.cat([torch.func.jvp(_attn_core, (qkv.double(),), (dq[j:j + 1].double(),))[1] for j in range(k)], 0)
    o, S = ops.attention_fwd(torch.cat([qkv, dq], 0).contiguous(), 1)
    torch.cuda.synchronize()
    # the stored o is rounded to tf32 for the following tensor-core projection: 2^-11 relative
    e0, e1 = rel_err(o[:1], oref), rel_err(o[1:], dref)
    print(f"attention T={T} C={C}: primal rel_err {e0:.2e} tangent {e1:.2e}")
    assert e0 < 5e-4 and e1 < 1e-3
    Sref = torch.softmax(torch.bmm(qkv[..., :C].double(), qkv[..., C:2 * C].double().transpose(1, 2)) * C ** -0.5, 2)
    assert rel_err(S[:1], Sref) < 5e-4
    go = _tf32(torch.randn(k, T, C, generator=g)).to(dev)
    qd = qkv.double().requires_grad_(True)
    od = _attn_core(qd)
    gref = torch.cat([torch.autograd.grad(od, qd, go[j:j + 1].double(), retain_graph=True)[0] for j in range(k)], 0)
    gq = ops.attention_vjp(go, qkv, S[0].contiguous())
    torch.cuda.synchronize()
    e = rel_err(gq, gref)
    print(f"attention T={T} C={C}: vjp rel_err {e:.2e} (q {rel_err(gq[..., :C], gref[..., :C]):.2e} "
          f"k {rel_err(gq[..., C:2 * C], gref[..., C:2 * C]):.2e} v {rel_err(gq[..., 2 * C:], gref[..., 2 * C:]):.2e})")
    assert e < 1e-3
    # batch of primal rows only
    qb = _tf32(torch.randn(3, T, 3 * C, generator=g)).to(dev)
    ob, _ = ops.attention_fwd(qb, 3)
    assert rel_err(ob, _attn_core(qb.double())) < 5e-4


@pytest.mark.parametrize("Tq,C,heads,Tk,Tkv", [(256, 512, 8, 128, 77), (1024, 256, 4, 128, 77), (100, 128, 1, 64, 64),
                                                (4096, 128, 2, 256, 200)])
def test_cross_attention_fwd_jvp_vjp(dev, Tq, C, heads, Tk, Tkv):
    """Fused tcgen05 cross-attention (image tokens -> fixed context of Tkv tokens padded to Tk) against
    fp64: forward, tangent rows (d/dq, the context is a constant) and the VJP."""
    from loco_edit_b200 import ops
    g = torch.Generator().manual_seed(Tq + C + Tkv)
    k = 2
    D = C // heads
    q = _tf32(torch.randn(1, Tq, C, generator=g)).to(dev)
    dq = _tf32(torch.randn(k, Tq, C, generator=g)).to(dev)
    kv = torch.zeros(Tk, 2 * C)
    kv[:Tkv] = _tf32(torch.randn(Tkv, 2 * C, generator=g))
    kv = kv.to(dev)

    def f(z):      # z [n, Tq, C] -> [n, Tq, C]
        n = z.shape[0]
        zh = z.reshape(n, Tq, heads, D).transpose(1, 2)                        # [n, h, Tq, D]
        kc = kv[:Tkv, :C].double().reshape(Tkv, heads, D).transpose(0, 1)      # [h, Tkv, D]
        vc = kv[:Tkv, C:].double().reshape(Tkv, heads, D).transpose(0, 1)
        w = torch.softmax(torch.einsum("nhqd,hkd->nhqk", zh, kc) * D ** -0.5, dim=-1)
        return torch.einsum("nhqk,hkd->nhqd", w, vc).transpose(1, 2).reshape(n, Tq, C)

    oref = f(q.double())
    dref = torch.cat([torch.func.jvp(f, (q.double(),), (dq[j:j + 1].double(),))[1] for j in range(k)], 0)
    o, S = ops.cross_attention_fwd(torch.cat([q, dq], 0).contiguous(), kv, 1, heads, Tkv)
    torch.cuda.synchronize()
    e0, e1 = rel_err(o[:1], oref), rel_err(o[1:], dref)
    assert float(S[0, :, :, Tkv:].abs().max()) == 0.0 if Tkv < Tk else True
    go = _tf32(torch.randn(k, Tq, C, generator=g)).to(dev)
    qd = q.double().requires_grad_(True)
    od = f(qd)
    gref = torch.cat([torch.autograd.grad(od, qd, go[j:j + 1].double(), retain_graph=True)[0] for j in range(k)], 0)
    gq = ops.cross_attention_vjp(go, kv, S[0].contiguous(), heads, Tkv)
    torch.cuda.synchronize()
    e2 = rel_err(gq, gref)
    print(f"cross-attention Tq={Tq} C={C} heads={heads} Tk={Tkv}/{Tk}: primal {e0:.2e} tangent {e1:.2e} vjp {e2:.2e}")
    assert e0 < 5e-4 and e1 < 1e-3 and e2 < 1e-3


@pytest.mark.parametrize("k,d", [(1, 3072), (5, 196608), (22, 12288), (64, 49152)])
def test_orthonormalise_matches_svd(dev, k, d):
    from loco_edit_b200 import ops
    g = torch.Generator().manual_seed(k)
    W = torch.randn(k, d, generator=g)
    W = (W * torch.linspace(3.0, 1.0, k)[:, None]).to(dev)
    V, s = ops.orthonormalise(W)
    torch.cuda.synchronize()
    _, sv, vh = torch.linalg.svd(W.double().cpu(), full_matrices=False)
    assert torch.allclose(s.cpu().double(), sv.sqrt(), rtol=1e-4)       # reference returns s.sqrt()
    gram = (V.double() @ V.double().T).cpu()
    assert float((gram - torch.eye(k, dtype=torch.float64)).abs().max()) < 1e-4
    dots = (V.double().cpu() * vh).sum(1).abs()
    assert float((1 - dots).max()) < 1e-3, dots
    # sign alignment against a previous basis
    Vp = (-V).contiguous()
    V2, _ = ops.orthonormalise(W, v_prev=Vp)
    assert float(((V2 * Vp).sum(1)).min()) > 0.99


@pytest.mark.parametrize("k,d,decay", [(50, 196608, 1e-4), (64, 49152, 1e-5), (25, 196608, 1e-3)])
def test_orthonormalise_decaying_spectrum_matches_svd(dev, k, d, decay):
    """The reference default is pca_rank = 50 on a fast-decaying PMP-Jacobian spectrum
    (src/modules/edit.py:2482 runs torch.linalg.svd on the k x d matrix).  The Gram route squares the
    condition number, so the Gram matrix is accumulated in fp64, the transform is applied in fp64 and
    a second (Loewdin) pass re-orthonormalises: V V^T - I and the principal angles of the leading /
    trailing halves must stay at fp32 round-off for singular values spread over `decay`."""
    from gpu_util import principal_angles_deg
    from loco_edit_b200 import ops
    g = torch.Generator().manual_seed(100 + k)
    Q, _ = torch.linalg.qr(torch.randn(d, k, generator=g, dtype=torch.float64))
    R, _ = torch.linalg.qr(torch.randn(k, k, generator=g, dtype=torch.float64))
    sv = torch.logspace(0, torch.log10(torch.tensor(decay)).item(), k, dtype=torch.float64) * 37.0
    W64 = (R * sv[None, :]) @ Q.T                        # k x d with singular values sv
    W = W64.float().to(dev)
    V, s = ops.orthonormalise(W)
    torch.cuda.synchronize()
    _, sref, vh = torch.linalg.svd(W.double().cpu(), full_matrices=False)     # svd of the fp32 input
    srel = float(((s.cpu().double() ** 2 - sref).abs() / sref).max())
    gram = (V.double() @ V.double().T).cpu()
    orth = float((gram - torch.eye(k, dtype=torch.float64)).abs().max())
    h = k // 2
    a_top = float(principal_angles_deg(V[:h], vh[:h]).max())
    a_all = float(principal_angles_deg(V, vh).max())
    dots = (V.double().cpu() * vh).sum(1).abs()
    print(f"k={k} decay={decay:g}: s rel {srel:.2e}, |VV^T-I| {orth:.2e}, top-half angle {a_top:.4f} deg, "
          f"row space {a_all:.4f} deg, min |<v_i, vh_i>| {float(dots.min()):.6f}")
    assert srel < 1e-3            # north_star: singular values to 1e-3 relative
    assert orth < 2e-5
    assert a_top < 0.1 and a_all < 0.1
    assert float(dots.min()) > 0.999


@pytest.mark.parametrize("k,kn,d,project", [(5, 5, 196608, True), (2, 3, 3072, True), (3, 10, 12288, True), (4, 5, 3072, False)])
def test_nullspace_project(dev, k, kn, d, project):
    from loco_edit_b200 import ops
    from oracle import pullback_ref
    g = torch.Generator().manual_seed(k * 7 + kn)
    vm = torch.randn(k, d, generator=g)
    vn, _ = torch.linalg.qr(torch.randn(d, kn, generator=g))
    vn = vn.T.contiguous()
    ref = pullback_ref.nullspace_project(vm, vn, kn, project)
    out = ops.nullspace_project(vm.to(dev), vn.to(dev), project)
    torch.cuda.synchronize()
    assert rel_err(out.cpu(), ref) < 1e-5
    if project:
        assert float((out.cpu().double() @ vn.double().T).abs().max()) < 1e-5


def test_ddim_step_bit_exact(dev):
    from loco_edit_b200 import ops
    from oracle import pullback_ref
    s = pullback_ref.RefScheduler()
    g = torch.Generator().manual_seed(2)
    xt = torch.randn(2, 3, 32, 32, generator=g)
    et = torch.randn(2, 3, 32, 32, generator=g)
    nz = torch.randn(2, 3, 32, 32, generator=g)
    for inv in (False, True):
        s.set_timesteps(100, is_inversion=inv)
        for idx in (0, 17, 40, 79, 98):
            t, tn = s.timesteps[idx], s.timesteps_next[idx]
            at, atn = float(s.alpha(t)), float(s.alpha(tn))
            ref, p = s.step(et, t, xt, eta=0)
            out, x0 = ops.ddim_step(xt.to(dev), et.to(dev), at, atn, 0.0, want_x0=True)
            assert torch.equal(out.cpu(), ref) and torch.equal(x0.cpu(), p), (inv, idx)
            if not inv:
                ref1, _ = s.step(et, t, xt, eta=1, noise=nz)
                out1 = ops.ddim_step(xt.to(dev), et.to(dev), at, atn, 1.0, noise=nz.to(dev))
                assert torch.equal(out1.cpu(), ref1), (inv, idx)
    # PMP and edit step share the op order
    at = float(s.alphas_cumprod[595])
    pm = ops.pmp_forward(xt.to(dev), et.to(dev), at)
    a = s.alphas_cumprod[595]
    assert torch.equal(pm.cpu(), (xt - et * (1 - a).sqrt()) / a.sqrt())
    ed = ops.axpy(xt.to(dev), et.to(dev), 0.5)
    assert torch.equal(ed.cpu(), xt + 0.5 * 1.0 * et)


def test_mask_selection_bit_exact(dev):
    from loco_edit_b200 import ops
    g = torch.Generator().manual_seed(4)
    for shape in [(3, 32, 32), (3, 256, 256), (3, 7, 5)]:
        for mask in [torch.rand(shape, generator=g) < 0.3, torch.zeros(shape, dtype=torch.bool),
                     torch.ones(shape, dtype=torch.bool)]:
            idx = ops.mask_indices(mask.to(dev))
            ref = mask.reshape(-1).nonzero().reshape(-1)
            assert torch.equal(idx.cpu().long(), ref)
            x = torch.randn(4, *shape, generator=g)
            sel = ops.gather_rows(x.reshape(4, -1).to(dev), idx)
            assert torch.equal(sel.cpu(), x[:, mask])                 # P_xt[:, mask] order
            back = ops.scatter_rows(sel, idx, x[0].numel())
            assert torch.equal(back.cpu(), (x * mask).reshape(4, -1))


def test_gram(dev):
    from loco_edit_b200 import ops
    g = torch.Generator().manual_seed(6)
    for ka, kb, d in [(5, 5, 196608), (3, 7, 1000), (40, 64, 8192), (9, 2, 333)]:
        A, B = torch.randn(ka, d, generator=g), torch.randn(kb, d, generator=g)
        G = ops.gram(A.to(dev), B.to(dev))
        ref = A.double() @ B.double().T
        assert float((G.cpu() - ref).abs().max()) < 1e-3 * (d ** 0.5) * 1e-2


@pytest.mark.parametrize("N,H,Cin,C2,Cout", [(6, 128, 128, 256, 128), (3, 256, 128, 256, 128), (8, 64, 256, 384, 256)])
def test_conv_fused_shortcut_and_statistics(dev, N, H, Cin, C2, Cout):
    """conv2 + 1x1 shortcut in one accumulation (halo / CTA-pair variants) and the GroupNorm
    statistics fused into its epilogue, against fp64 PyTorch on tf32-representable inputs."""
    from loco_edit_b200 import ops
    g = torch.Generator().manual_seed(N * 1000 + H)
    w = tf32_round(torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).to(dev)
    w2 = tf32_round(torch.randn(Cout, C2, 1, 1, generator=g) / C2 ** 0.5).to(dev)
    x = tf32_round(torch.randn(N, Cin, H, H, generator=g)).to(dev)
    x2 = tf32_round(torch.randn(N, C2, H, H, generator=g)).to(dev)
    bias = torch.randn(Cout, generator=g).to(dev)
    ref = F.conv2d(x.double(), w.double(), padding=1) + F.conv2d(x2.double(), w2.double())
    ref[:2] += bias.double()[None, :, None, None]
    y, st = ops.conv2d_fused_nhwc(nhwc(x), w, nhwc(x2), w2, bias=bias, bias_rows=2, stat_groups=32)
    torch.cuda.synchronize()
    e = rel_err(nchw(y), ref)
    print(f"fused conv+shortcut N={N} {H}x{H} {Cin}+{C2}->{Cout}: rel_err={e:.3e}")
    assert e < 1e-5
    grp = ref.reshape(N, 32, -1)
    sref = torch.stack([grp.sum(-1), (grp * grp).sum(-1)], -1)
    assert rel_err(st, sref) < 2e-5      # fp32 partial sums per 32-pixel chunk, fp64 across chunks
    # without the shortcut
    y1 = ops.conv2d_fused_nhwc(nhwc(x), w)
    assert rel_err(nchw(y1), F.conv2d(x.double(), w.double(), padding=1)) < 1e-5
