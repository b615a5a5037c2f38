"""The CUDA-core kernels of the ToMP transformer (csrc/transformer_kernels.cuh, SURVEY 8 row T1: position add, LayerNorm, the decoder's
small linear layers, multi-head attention with online softmax and a key-padding mask, the decoder's single-query attention) executed ON
THE CPU: the same source file the CUDA build compiles (`cuobjdump -sass` identical before and after the kernels moved into the header),
built as host code under tests/cpu_emul/cuda_shim.h with the launch shapes of csrc/transformer.cu, against plain PyTorch float64
restatements of `nn.MultiheadAttention` / `nn.LayerNorm` / `nn.Linear` as the reference's `TransformerEncoderLayer.forward_post` and
`TransformerDecoderLayer.forward_post` use them (ltr/models/transformer/transformer.py:173-181, 224-238)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HD = 32                                                            # head dimension of the kernels (d_model 256 / 8 heads)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    out = os.path.join(str(tmp_path_factory.mktemp("tr_emul")), "libtr_emul.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-fno-gnu-unique", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "cpu_emul", "transformer_emul.cpp"), "-o", out], check=True, capture_output=True)
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _attention_ref(q, k, v, mask, H, scale):
    """q [Lq,B,H*32], k / v [L,B,H*32], mask [B,L] bool (True = padded key) -> [Lq,B,H*32], float64."""
    Lq, B, D = q.shape
    L = k.shape[0]
    qh = q.double().reshape(Lq, B, H, HD).permute(1, 2, 0, 3)
    kh = k.double().reshape(L, B, H, HD).permute(1, 2, 0, 3)
    vh = v.double().reshape(L, B, H, HD).permute(1, 2, 0, 3)
    s = (qh @ kh.transpose(-1, -2)) * scale
    if mask is not None:
        s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    return (torch.softmax(s, -1) @ vh).permute(2, 0, 1, 3).reshape(Lq, B, D)


@pytest.mark.parametrize("L,B,H,use_mask", [(72, 2, 2, True), (130, 1, 3, False), (17, 3, 1, True)])
def test_attention_kernel_sources(emul, L, B, H, use_mask):
    g = torch.Generator().manual_seed(L)
    D = H * HD
    qk = torch.randn(L, B, 2 * D, generator=g)                     # the encoder's fused [q | k] projection buffer (row pitch 2 D)
    v = torch.randn(L, B, D, generator=g)
    mask = None
    if use_mask:
        mask = torch.zeros(B, L, dtype=torch.bool)
        mask[B - 1, L // 3: L // 2] = True
    scale = 1.0 / HD ** 0.5
    qkn, vn = np.ascontiguousarray(qk.numpy()), np.ascontiguousarray(v.numpy())
    m8 = np.ascontiguousarray(mask.numpy().astype(np.uint8)) if mask is not None else None
    out = np.full((L, B, D), np.nan, np.float32)
    kptr = C.c_void_p(qkn.ctypes.data + D * 4)
    assert emul.tr_emul_attention(_p(qkn), kptr, _p(vn), _p(m8), _p(out), L, L, B, H, 2 * D, 2 * D, D, D, C.c_float(scale)) == 0
    ref = _attention_ref(qk[..., :D], qk[..., D:], v, mask, H, scale)
    assert _rel(out, ref.numpy()) < 2e-6
    # the decoder's cross attention: one query per batch element against the same keys / values
    q1 = torch.randn(B, D, generator=g)
    o1 = np.full((B, D), np.nan, np.float32)
    assert emul.tr_emul_attention_q1(_p(np.ascontiguousarray(q1.numpy())), kptr, _p(vn), _p(m8), _p(o1), L, B, H, D, 2 * D, D, D, C.c_float(scale)) == 0
    ref1 = _attention_ref(q1[None], qk[..., D:], v, mask, H, scale)[0]
    assert _rel(o1, ref1.numpy()) < 2e-6


def test_layernorm_add_pos_small_linear_kernel_sources(emul):
    g = torch.Generator().manual_seed(1)
    T, D = 37, 256
    x, gam, bet = torch.randn(T, D, generator=g) * 3 + 1, torch.rand(D, generator=g) + 0.5, torch.randn(D, generator=g)
    y = np.full((T, D), np.nan, np.float32)
    assert emul.tr_emul_layernorm(_p(x.numpy()), _p(gam.numpy()), _p(bet.numpy()), _p(y), T, D) == 0
    assert _rel(y, torch.nn.functional.layer_norm(x.double(), (D,), gam.double(), bet.double(), 1e-5).numpy()) < 2e-6
    L, B = 9, 2
    for Bp in (1, 2):
        src, pos = torch.randn(L, B, D, generator=g), torch.randn(L, Bp, D, generator=g)
        out = np.full((L, B, D), np.nan, np.float32)
        assert emul.tr_emul_add_pos(_p(src.numpy()), _p(np.ascontiguousarray(pos.numpy())), _p(out), L, B, Bp, D) == 0
        assert np.array_equal(out, (src + pos).numpy())
    M, K, N = 2, 256, 96
    xm, W, b, res = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / 16, torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    for relu, r in ((0, res), (1, None)):
        out = np.full((M, N), np.nan, np.float32)
        assert emul.tr_emul_small_linear(_p(xm.numpy()), _p(W.numpy()), _p(b.numpy()), _p(r.numpy()) if r is not None else None, _p(out), M, K, N, relu) == 0
        ref = xm.double() @ W.double().t() + b.double()
        ref = torch.relu(ref) if relu else ref
        ref = ref + r.double() if r is not None else ref
        assert _rel(out, ref.numpy()) < 2e-6
