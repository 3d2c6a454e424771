"""tcgen05 building block: the hand-written UMMA descriptors / 128B swizzle / TMEM read-back of
csrc/g2048_tc.cuh against a plain PyTorch fp32 matmul of the same bf16-rounded operands."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(64, 64), (16, 208), (208, 208), (256, 256), (128, 16)])
def test_tc_gemm_selftest_matches_torch(K, N):
    from g2048 import _lib, env
    env.init(0)
    _lib.register("g2048_tc_gemm_selftest", [C.c_void_p] * 3 + [C.c_int32, C.c_int32, C.c_void_p])
    g = torch.Generator(device="cuda").manual_seed(K * 1000 + N)
    A = torch.randn((128, K), generator=g, device="cuda")
    W = torch.randn((N, K), generator=g, device="cuda")
    out = torch.full((128, N), float("nan"), device="cuda")
    _lib.call("g2048_tc_gemm_selftest", C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(out.data_ptr()),
              K, N, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ W.bfloat16().float().T
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("a_f16,w_f16", [(1, 1), (0, 1), (1, 0)])
def test_tc_gemm_operand_formats_are_independent(a_f16, w_f16):
    """kind::f16 with fp16 operands, and with one bf16 and one fp16 operand (the fused update multiplies bf16 gradient
    images with fp16 activation images): exact products of the rounded operands, fp32 accumulation."""
    from g2048 import _lib, env
    env.init(0)
    _lib.register("g2048_tc_gemm_selftest_fmt", [C.c_void_p] * 3 + [C.c_int32] * 4 + [C.c_void_p])
    K, N = 208, 208
    g = torch.Generator(device="cuda").manual_seed(7 + 2 * a_f16 + w_f16)
    A = torch.randn((128, K), generator=g, device="cuda")
    W = torch.randn((N, K), generator=g, device="cuda")
    out = torch.full((128, N), float("nan"), device="cuda")
    _lib.call("g2048_tc_gemm_selftest_fmt", C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(out.data_ptr()),
              K, N, a_f16, w_f16, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    r = lambda t, f16: t.half().float() if f16 else t.bfloat16().float()
    ref = r(A, a_f16) @ r(W, w_f16).T
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-3)
    assert float((out - A @ W.T).abs().max()) < (0.05 if a_f16 and w_f16 else 0.5)
