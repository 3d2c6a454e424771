"""Split-bf16 ("x3") tcgen05 GEMMs of the policy update (csrc/g2048_linear.cu) against float64 matmuls.

Tolerance: every product a*b is evaluated as alo*bhi + ahi*blo + ahi*bhi with bf16 terms, so the
per-product error is bounded by ~3 * 2^-17 |a||b|; the tests require the result within
2e-5 * (|A| |B|^T) elementwise of the float64 product (a bound, typical errors are a few 1e-6 relative)
and no worse than 4x the error of torch's own fp32 matmul plus that slack.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, generator=g, device="cuda") * scale


def _check(got, a64, b64, label):
    ref = a64 @ b64
    bound = 2e-5 * (a64.abs() @ b64.abs()) + 1e-30
    err = (got.double() - ref).abs()
    worst = float((err / bound).max())
    assert worst < 1.0, f"{label}: error {worst:.2f}x the 2e-5 |A||B| bound"
    return float(err.max() / ref.abs().max())


@pytest.mark.parametrize("M,N,K", [(128, 196, 196), (1000, 196, 48), (4096 + 77, 196, 196), (300, 64, 208), (129, 4, 196),
                                   (50000, 196, 196)])
def test_gemm_matches_float64(M, N, K):
    from g2048 import linear
    a, w = _rand((M, K), M + N), _rand((N, K), K + 7, 0.3)
    out = linear.gemm(a, linear.pack_weight(w), N)
    torch.cuda.synchronize()
    rel = _check(out, a.double(), w.double().T, "forward")
    assert rel < 1.5e-5
    # dgrad form: dy [M,N] @ W [N,K]
    dy = _rand((M, N), 3 * M + 1, 1e-3)
    dx = linear.gemm(dy, linear.pack_weight(w, transpose=True), K)
    torch.cuda.synchronize()
    _check(dx, dy.double(), w.double(), "dgrad")


@pytest.mark.parametrize("M,N,K", [(32, 196, 196), (1000, 196, 48), (4096 + 77, 196, 196), (70000, 196, 196), (333, 64, 208),
                                   (5000, 128, 16), (600000, 196, 196)])
def test_wgrad_matches_float64(M, N, K):
    from g2048 import linear
    dy, x = _rand((M, N), M + 11, 1e-2), _rand((M, K), M + 13)
    dw = linear.wgrad(dy, x)
    torch.cuda.synchronize()
    _check(dw, dy.double().T, x.double(), "wgrad")
    again = linear.wgrad(dy, x)
    assert torch.equal(dw, again), "wgrad must be deterministic"


def test_gemm_handles_special_rows():
    from g2048 import linear
    a = _rand((256, 196), 5)
    a[3] = 0.0
    a[7] = 1e-30          # lo terms underflow to zero: still exact enough
    a[9] = 3e4
    w = _rand((196, 196), 6, 0.1)
    out = linear.gemm(a, linear.pack_weight(w), 196)
    assert torch.isfinite(out).all()
    assert torch.equal(out[3], torch.zeros_like(out[3]))
    _check(out, a.double(), w.double().T, "special rows")


def test_linear_autograd_matches_torch_fp64():
    from g2048 import linear
    x = _rand((3000, 196), 21).requires_grad_(True)
    w = _rand((196, 196), 22, 0.2).requires_grad_(True)
    g = _rand((3000, 196), 23, 1e-2)
    y = linear.linear(x, w)
    y.backward(g)
    x64, w64 = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    (x64 @ w64.T).backward(g.double())
    # max-norm relative error of a 196-term x3 dot product is ~1e-5 of the output scale (fp32 SGEMM: ~1e-6)
    for got, ref in ((y.detach(), x64.detach() @ w64.detach().T), (x.grad, x64.grad), (w.grad, w64.grad)):
        scale = float(ref.abs().max())
        assert float((got.double() - ref).abs().max()) < 4e-5 * scale


def test_linear_rejects_unsupported_inputs():
    from g2048 import linear
    with pytest.raises(ValueError):
        linear.linear(torch.zeros(4, 210, device="cuda"), torch.zeros(8, 210, device="cuda"))
    with pytest.raises(ValueError):
        linear.gemm(torch.zeros(4, 16), torch.zeros(16, dtype=torch.uint8), 16)


def operand_image(x: torch.Tensor, hp: int, dtype=torch.bfloat16) -> torch.Tensor:
    """[m, f] fp32 -> the hi|lo operand image (bf16 or fp16 terms) the fused update kernel writes and x3_wgrad_kernel
    bulk-copies: per tile of 128 samples [hi | lo][16-feature block][sample][32 B, halves swapped on (sample >> 2) & 1]."""
    m, f = x.shape
    tiles = (m + 127) // 128
    xp = torch.zeros((tiles * 128, hp), dtype=torch.float32, device=x.device)
    xp[:m, :f] = x
    hi = xp.to(dtype)
    lo = (xp - hi.float()).to(dtype)
    parts = torch.stack([hi, lo], 0).view(2, tiles, 128, hp // 16, 2, 8)          # [part, tile, row, block, half, 8]
    swap = ((torch.arange(128, device=x.device) >> 2) & 1).view(1, 1, 128, 1, 1, 1).bool()
    parts = torch.where(swap, parts.flip(-2), parts)
    img = parts.permute(1, 0, 3, 2, 4, 5).contiguous()                            # [tile, part, block, row, half, 8]
    return img.view(torch.uint8).view(-1).view(torch.float32)


@pytest.mark.parametrize("fp16", [False, True])
@pytest.mark.parametrize("m,n,k", [(1000, 196, 196), (70000, 196, 48), (333, 8, 196), (4096, 64, 64)])
@pytest.mark.parametrize("dy_img,x_img", [(True, False), (False, True), (True, True)])
def test_wgrad_with_operand_images(m, n, k, dy_img, x_img, fp16):
    """dW = dY^T X with either operand as a hi|lo operand image (bulk-copied into the ring) instead of row-major fp32, in bf16
    terms (the x3 bound of the fp32 path) or in fp16 terms (the fused update's images; 22 mantissa bits: a bound 30 times
    tighter); the image decodes back (g2048.update.untile) to hi + lo."""
    from g2048 import linear, update
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    dy = torch.randn((m, n), generator=g, device="cuda") * 0.3
    x = torch.randn((m, k), generator=g, device="cuda")
    hp_n, hp_k = (n + 15) // 16 * 16, (k + 15) // 16 * 16
    dt = torch.float16 if fp16 else torch.bfloat16
    a = operand_image(dy, hp_n, dt) if dy_img else dy
    b = operand_image(x, hp_k, dt) if x_img else x
    if dy_img:
        back = update.untile(a, m, n, hp=hp_n, dtype=dt)
        torch.testing.assert_close(back, dy, rtol=1e-6 if fp16 else 2e-5, atol=1e-6)
    got = linear.wgrad_tiled(a, b, m, n, k, dy_hp=hp_n if dy_img else 0, x_hp=hp_k if x_img else 0, fp16=fp16)
    want = dy.double().T @ x.double()
    bound = (6e-7 if fp16 else 2e-5) * (dy.double().abs().T @ x.double().abs()) + 1e-9
    assert bool(((got.double() - want).abs() <= bound).all())


@pytest.mark.parametrize("m,dy_img", [(1000, True), (70001, True), (4097, False)])
def test_stem_wgrad_from_packed_boards(m, dy_img):
    """x_hp = -1: X is the packed boards; the kernel forms the 48 model inputs itself (game.py:92-101) and must give
    exactly what it gives on g2048_encode's output."""
    from g2048 import env, linear
    g = torch.Generator(device="cuda").manual_seed(m)
    e = torch.randint(0, 16, (m, 16), generator=g, device="cuda", dtype=torch.int64)
    boards = (e << (torch.arange(16, device="cuda") * 4)).sum(1)
    dy = torch.randn((m, 196), generator=g, device="cuda") * 0.1
    a = operand_image(dy, 208) if dy_img else dy
    hp = 208 if dy_img else 0
    want = linear.wgrad_tiled(a, env.encode(boards), m, 196, 48, dy_hp=hp)
    got = linear.wgrad_tiled(a, boards, m, 196, 48, dy_hp=hp, x_hp=-1)
    assert torch.equal(got, want)
