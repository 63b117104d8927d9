"""The tcgen05 + TMA "weights-as-A" GEMM in isolation (b2a_tc_gemm_test, device pointers) against a torch
fp32 reference of the same op: out[N, M] = X[N, K] @ W[M, K]^T.  bf16 inputs, fp32 accumulation: the only
difference from the reference is summation order, so the tolerance is tight (1e-5 relative)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(b2a, W, X, out, M, N, K, bn, epi, split, hilo, ctas):
    lib = C.CDLL(str(b2a._ffi.LIB_PATH))
    fn = lib.b2a_tc_gemm_test
    fn.restype = C.c_int32
    fn.argtypes = [C.c_void_p] * 3 + [C.c_int32] * 8 + [C.c_void_p]
    st = fn(W.data_ptr(), X.data_ptr(), out.data_ptr(), M, N, K, bn, epi, split, hilo, ctas, None)
    torch.cuda.synchronize()
    assert st == 0, b2a._ffi.lib().b2a_last_error()


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.mark.parametrize("M,K,N,ctas,split", [(128, 64, 8, 1, 0), (256, 128, 8, 2, 0), (384, 512, 5, 3, 0),
                                              (1000, 256, 8, 7, 0), (3072, 3072, 8, 148, 1), (5120, 3072, 8, 148, 1),
                                              (3072, 8192, 8, 148, 1), (640, 1024, 16, 148, 1)])
def test_decode_tile_store_and_streamk(b2a, M, K, N, ctas, split):
    g = torch.Generator(device="cuda").manual_seed(M + K)
    W = (torch.randn(M, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    X = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.zeros(N, M, device="cuda", dtype=torch.float32)
    _run(b2a, W, X, out, M, N, K, 16, 0, split, 0, ctas)
    ref = X.float() @ W.float().T
    assert _rel(out, ref) < 1e-5, _rel(out, ref)


def test_hilo_gives_fp32_activation_accuracy(b2a):
    M, K, B = 1024, 2048, 8
    g = torch.Generator(device="cuda").manual_seed(0)
    W = (torch.randn(M, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    x = torch.randn(B, K, device="cuda", generator=g)
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    X = torch.cat([hi, lo]).contiguous()
    out = torch.zeros(B, M, device="cuda", dtype=torch.float32)
    _run(b2a, W, X, out, M, B, K, 16, 0, 1, 1, 148)
    ref = x.double() @ W.double().T
    assert _rel(out, ref) < 2e-5, _rel(out, ref)
    only_hi = hi.float() @ W.float().T
    assert _rel(only_hi, ref) > 1e-3          # what bf16-rounded activations would have cost


def test_swiglu_epilogue_hilo_out(b2a):
    I, K, B = 512, 1024, 8
    g = torch.Generator(device="cuda").manual_seed(1)
    Wg = (torch.randn(I, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    Wu = (torch.randn(I, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    W = torch.stack([Wg, Wu], dim=1).reshape(2 * I, K).contiguous()        # rows interleaved gate/up
    x = torch.randn(B, K, device="cuda", generator=g)
    hi = x.to(torch.bfloat16)
    X = torch.cat([hi, (x - hi.float()).to(torch.bfloat16)]).contiguous()
    out = torch.zeros(16, I, device="cuda", dtype=torch.bfloat16)
    _run(b2a, W, X, out, 2 * I, B, K, 16, 2, 0, 1, 8)
    gte, up = x.double() @ Wg.double().T, x.double() @ Wu.double().T
    ref = torch.nn.functional.silu(gte) * up
    got = out[:8].double() + out[8:].double()
    assert _rel(got, ref) < 5e-5, _rel(got, ref)


@pytest.mark.parametrize("M,K,N,ctas", [(256, 128, 128, 2), (512, 256, 300, 4), (1024, 3072, 512, 37)])
def test_prefill_tile_bn128(b2a, M, K, N, ctas):
    g = torch.Generator(device="cuda").manual_seed(N)
    W = (torch.randn(M, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    X = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.zeros(N, M, device="cuda", dtype=torch.float32)
    _run(b2a, W, X, out, M, N, K, 128, 0, 1, 0, ctas)
    ref = X.float() @ W.float().T
    assert _rel(out, ref) < 1e-5, _rel(out, ref)
