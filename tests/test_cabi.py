"""CPU: the C-ABI library loads, exports every symbol include/loco_b200.h declares, builds the model
registry without a GPU, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import json
import os
import re

import pytest
import torch

from loco_edit_b200 import _lib
from loco_edit_b200.weights import DDPM256, DDPM256_TEXT, ddpm_param_shapes, if_standin_arch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "loco_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(loco_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libloco_b200.so does not export %s" % n
        assert n in _lib.PROTOTYPES, "python binding lacks a prototype for %s" % n
    assert lib.loco_abi_version() == 3


@pytest.mark.parametrize("arch_name", ["DDPM256_TEXT", "if_standin"])
def test_text_conditioned_registry_and_plan_sizing(arch_name):
    """U-Nets with cross-attention layers (ctx_dim > 0): the library's parameter registry == the Python
    shape table (norm2 / q2 / kv2 / proj_out2 per AttnBlock), plan sizing is host logic, and a bad
    head split is refused at plan creation."""
    from loco_edit_b200.unet import _make_arch
    a = DDPM256_TEXT if arch_name == "DDPM256_TEXT" else if_standin_arch(64)
    lib = _lib.load()
    h = C.c_void_p()
    arch = _make_arch(a)
    assert arch.ctx_dim == 768
    _lib.check(lib.loco_unet_create(C.byref(arch), C.byref(h)))
    try:
        buf = C.create_string_buffer(256)
        shape = (C.c_int * 4)()
        nd = C.c_int()
        got = {}
        for i in range(lib.loco_unet_num_params(h)):
            _lib.check(lib.loco_unet_param_info(h, i, buf, 256, shape, C.byref(nd)))
            got[buf.value.decode()] = [shape[j] for j in range(nd.value)]
        want = {k: list(v) for k, v in ddpm_param_shapes(a).items()}
        assert got == want
        n_attn = sum(1 for k in want if k.endswith(".kv2.weight"))
        assert n_attn == (6 if arch_name == "DDPM256_TEXT" else 11) and all(v == [2 * want[k[:-10] + "q2.bias"][0], 768]
                                                                           for k, v in want.items() if k.endswith(".kv2.weight"))
        p = C.c_void_p()
        _lib.check(lib.loco_plan_create(h, 1, 5, 5, C.byref(p)))
        assert lib.loco_plan_workspace_bytes(p) > 1e8
        lib.loco_plan_destroy(p)
    finally:
        lib.loco_unet_destroy(h)
    # 512 channels in 3 heads is not a multiple of 64 per head: refused when the plan is sized
    bad = _make_arch(dict(a, ctx_heads=3))
    _lib.check(lib.loco_unet_create(C.byref(bad), C.byref(h)))
    try:
        p = C.c_void_p()
        assert lib.loco_plan_create(h, 1, 0, 0, C.byref(p)) != 0
        assert b"cross-attention" in lib.loco_last_error()
    finally:
        lib.loco_unet_destroy(h)


def test_model_registry_matches_reference_state_dict(golden_dir):
    """Parameter names/shapes the C library expects == DDPM.state_dict() of the reference."""
    from loco_edit_b200.unet import _make_arch
    lib = _lib.load()
    h = C.c_void_p()
    arch = _make_arch(DDPM256)
    _lib.check(lib.loco_unet_create(C.byref(arch), C.byref(h)))
    try:
        buf = C.create_string_buffer(256)
        shape = (C.c_int * 4)()
        nd = C.c_int()
        got = {}
        for i in range(lib.loco_unet_num_params(h)):
            _lib.check(lib.loco_unet_param_info(h, i, buf, 256, shape, C.byref(nd)))
            got[buf.value.decode()] = [shape[j] for j in range(nd.value)]
        ref = json.load(open(os.path.join(golden_dir, "ddpm_param_shapes.json")))
        assert got == ref
        assert got == {k: list(v) for k, v in ddpm_param_shapes(DDPM256).items()}
        # packed arena: fprop + dgrad packs of every GEMM conv + raw vectors
        assert lib.loco_unet_weight_floats(h) > 200_000_000
        # plan sizing is pure host logic
        p = C.c_void_p()
        _lib.check(lib.loco_plan_create(h, 1, 5, 5, C.byref(p)))
        nbytes = lib.loco_plan_workspace_bytes(p)
        assert 8e9 < nbytes < 40e9
        ff, vf = C.c_double(), C.c_double()
        fo, vo = C.c_int(), C.c_int()
        _lib.check(lib.loco_plan_info(p, C.byref(ff), C.byref(vf), C.byref(fo), C.byref(vo)))
        # conv FLOPs of the tensor-core layers: 6 rows x (F - edge convs - attention)
        assert abs(ff.value / 6 - 0.4953e12) / 0.4953e12 < 0.01
        assert abs(vf.value / 5 - 0.4953e12) / 0.4953e12 < 0.01
        lib.loco_plan_destroy(p)
    finally:
        lib.loco_unet_destroy(h)


def test_bad_arguments_are_reported_not_thrown():
    lib = _lib.load()
    from loco_edit_b200.unet import _make_arch
    h = C.c_void_p()
    bad = _make_arch(dict(DDPM256, ch=96))
    assert lib.loco_unet_create(C.byref(bad), C.byref(h)) != 0
    assert b"multiple of 128" in lib.loco_last_error()
    with pytest.raises(_lib.LocoError):
        _lib.check(lib.loco_unet_create(C.byref(bad), C.byref(h)), "create")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    lib = _lib.load()
    x = torch.zeros(16)
    rc = lib.loco_axpy(C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()), 1.0, 16,
                       C.c_void_p(x.data_ptr()), None)
    assert rc != 0 and b"no CUDA device" in lib.loco_last_error()
    from loco_edit_b200.unet import B200UNet
    with pytest.raises(_lib.LocoError):
        B200UNet(DDPM256, {}, device="cpu")


def test_p2_model_registry_matches_reference_state_dict(golden_dir):
    """kind 1 (P2 / guided diffusion): the parameter registry equals create_model(**P2_DICT)'s
    state_dict (names, order-independent, shapes) and the plan sizing runs without a device."""
    from loco_edit_b200.unet import _make_arch
    from loco_edit_b200.weights import P2_256
    lib = _lib.load()
    h = C.c_void_p()
    arch = _make_arch(P2_256)
    assert arch.kind == 1 and arch.head_ch == 64
    _lib.check(lib.loco_unet_create(C.byref(arch), C.byref(h)))
    try:
        buf = C.create_string_buffer(256)
        shape = (C.c_int * 4)()
        nd = C.c_int()
        got = {}
        for i in range(lib.loco_unet_num_params(h)):
            _lib.check(lib.loco_unet_param_info(h, i, buf, 256, shape, C.byref(nd)))
            got[buf.value.decode()] = [shape[j] for j in range(nd.value)]
        ref = json.load(open(os.path.join(golden_dir, "p2_param_shapes.json")))
        assert got == ref
        p = C.c_void_p()
        _lib.check(lib.loco_plan_create(h, 1, 3, 3, C.byref(p)))
        ff, vf = C.c_double(), C.c_double()
        fo, vo = C.c_int(), C.c_int()
        _lib.check(lib.loco_plan_info(p, C.byref(ff), C.byref(vf), C.byref(fo), C.byref(vo)))
        # 387.5 GFLOP of convolutions per sample-forward, minus the two
        # 3-channel edge convolutions (0.45 + 0.45 GFLOP at 3 output channels) run on CUDA cores
        assert abs(ff.value / 4 - 0.3866e12) / 0.3866e12 < 0.01, ff.value / 4
        assert abs(vf.value / 3 - 0.3866e12) / 0.3866e12 < 0.01
        lib.loco_plan_destroy(p)
    finally:
        lib.loco_unet_destroy(h)


def test_conv_variant_selection_is_stable():
    """The host-side cost model that routes 3x3 layers to the halo / CTA-pair kernels (pure host
    logic, 148 SMs assumed without a device): the layer shapes of the bench must keep their variant."""
    lib = _lib.load()
    el = lambda kind, N, H, W, C: lib.loco_conv_halo_eligible(kind, N, H, W, C)
    # >= 64^2 layers of the fused (1+k)-row passes and of the batched forwards
    for N in (6, 8, 11, 40):
        assert el(0, N, 256, 256, 128) == 1 and el(3, N, 256, 256, 128) == 1
        assert el(0, N, 128, 128, 128) == 1
    assert el(0, 8, 64, 64, 256) == 1 and el(0, 40, 64, 64, 256) == 1 and el(0, 40, 32, 32, 256) == 1
    assert el(0, 1, 256, 256, 128) == 1                  # B = 1: 256 blocks, 2 waves beat 4 x 0.68
    # split-K / one-tile territory
    assert el(0, 6, 16, 16, 512) == 0 and el(0, 6, 8, 8, 512) == 0 and el(0, 1, 64, 64, 256) == 0
    assert el(0, 6, 32, 32, 256) == 0
    # only stride-1 3x3, 128-multiple output channels, 16-multiple images
    assert el(1, 40, 256, 256, 128) == 0 and el(2, 40, 256, 256, 128) == 0
    assert el(0, 40, 256, 256, 64) == 0 and el(0, 40, 24, 24, 128) == 0
