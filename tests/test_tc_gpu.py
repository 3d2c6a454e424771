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


def test_tc_gemm_fp16_operands_and_no_mixed_formats():
    """kind::f16 with fp16 operands (the term format of the split-fp16 kernels): exact products of the rounded operands,
    fp32 accumulation.  Mixing a bf16 and an fp16 operand in one MMA faults on B200 (measured in round 2), so the entry
    point refuses it instead of launching."""
    from g2048 import _lib, env
    env.init(0)
    _lib.register("g2048_tc_gemm_selftest_fmt", [C.c_void_p] * 3 + [C.c_int32] * 4 + [C.c_void_p])
    K, N = 208, 208
    g = torch.Generator(device="cuda").manual_seed(7)
    A = torch.randn((128, K), generator=g, device="cuda")
    W = torch.randn((N, K), generator=g, device="cuda")
    out = torch.full((128, N), float("nan"), device="cuda")
    args = lambda a, w: (C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(out.data_ptr()), K, N, a, w,
                         C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.call("g2048_tc_gemm_selftest_fmt", *args(1, 1))
    torch.cuda.synchronize()
    ref = A.half().float() @ W.half().float().T
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-3)
    assert float((out - A @ W.T).abs().max()) < 0.05          # bf16 operands: ~0.3 on this product
    for a, w in ((0, 1), (1, 0)):
        with pytest.raises(_lib.G2048Error):
            _lib.call("g2048_tc_gemm_selftest_fmt", *args(a, w))
