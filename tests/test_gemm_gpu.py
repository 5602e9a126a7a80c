"""GPU parity: tcgen05 GEMM (K1) vs a plain torch fp32 matmul of the same bf16-rounded operands.
Tolerance: fp32 accumulation of exact bf16 products -> differences are summation-order only
(rtol 1e-4 of the row scale); bf16 outputs add one rounding (2^-8 relative)."""
import pytest
import torch

from rcnn_ocr_b200 import ops

pytestmark = pytest.mark.gpu


def _ref(a, b, bias):
    r = a.float() @ b.float().t()
    return r + bias if bias is not None else r


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 512), (256, 384, 128), (16384, 4096, 512),
                                   (16384, 512, 1024), (16384, 195, 512), (100, 72, 40), (1, 8, 8),
                                   (257, 129, 72), (2048, 1024, 256), (300, 200, 1000),
                                   # few row blocks: tile widths 64 / 32 / 64 / 128 / 256 of the persistent kernel
                                   (256, 2048, 1024), (256, 736, 512), (256, 4100, 200), (384, 3200, 64), (384, 6400, 72)])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_gemm_matches_fp32_matmul(M, N, K, out_dtype):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    got = ops.gemm_bf16(a, b, bias, out_dtype)
    want = _ref(a, b, bias)
    scale = float(want.abs().max())
    tol = 1e-4 if out_dtype == torch.float32 else 2 ** -7
    torch.testing.assert_close(got.float(), want, rtol=tol, atol=tol * scale)
    got_nb = ops.gemm_bf16(a, b, None, out_dtype)
    torch.testing.assert_close(got_nb.float(), _ref(a, b, None), rtol=tol, atol=tol * scale)


def test_gemm_strided_operands_and_output():
    g = torch.Generator(device="cuda").manual_seed(1)
    big_a = torch.randn(200, 256, device="cuda", generator=g).bfloat16()
    big_b = torch.randn(96, 256, device="cuda", generator=g).bfloat16()
    a, b = big_a[:, 64:192], big_b[:, 64:192]                    # row pitch 256, K = 128
    out = torch.zeros(200, 128, device="cuda")
    ops.gemm_bf16(a, b, None, out=out[:, 16:112])
    torch.testing.assert_close(out[:, 16:112], _ref(a, b, None), rtol=1e-4, atol=1e-3)
    assert float(out[:, :16].abs().max()) == 0 and float(out[:, 112:].abs().max()) == 0


def test_gemm_exact_small_integers():
    """Small-integer operands make every product and partial sum exact: bit-exact result."""
    g = torch.Generator(device="cuda").manual_seed(2)
    a = torch.randint(-4, 5, (384, 320), device="cuda", generator=g).bfloat16()
    b = torch.randint(-4, 5, (256, 320), device="cuda", generator=g).bfloat16()
    got = ops.gemm_bf16(a, b)
    assert torch.equal(got, _ref(a, b, None))


@pytest.mark.parametrize("K,M,N", [(64, 128, 128), (128, 128, 128), (256, 64, 64), (16384, 4096, 512), (16384, 512, 1024),
                                   (16384, 195, 512), (1000, 200, 72), (99, 130, 40), (4096, 2048, 512)])
def test_gemm_atb_matches_fp32(K, M, N):
    """Weight-gradient shape D = A^T B with MN-major operands and split-K."""
    g = torch.Generator(device="cuda").manual_seed(K + M + N)
    a = torch.randn(K, M, device="cuda", generator=g).bfloat16()
    b = torch.randn(K, N, device="cuda", generator=g).bfloat16()
    got = ops.gemm_bf16_atb(a, b)
    want = a.float().t() @ b.float()
    scale = float(want.abs().max())
    torch.testing.assert_close(got, want, rtol=1e-4, atol=2e-4 * scale)
    got2 = ops.gemm_bf16_atb(a, b, out=got.clone(), accumulate=True)
    torch.testing.assert_close(got2, 2 * want, rtol=1e-4, atol=4e-4 * scale)


def test_gemm_atb_exact_small_integers_and_strided():
    g = torch.Generator(device="cuda").manual_seed(4)
    a_big = torch.randint(-3, 4, (320, 512), device="cuda", generator=g).bfloat16()
    b = torch.randint(-3, 4, (320, 192), device="cuda", generator=g).bfloat16()
    a = a_big[:, 128:384]                                     # column slice: lda = 512, M = 256
    got = ops.gemm_bf16_atb(a, b)
    assert torch.equal(got, a.float().t() @ b.float())


def test_gemm_atb_grouped_and_colsum():
    g = torch.Generator(device="cuda").manual_seed(6)
    a = torch.randint(-3, 4, (700, 2 * 256), device="cuda", generator=g).bfloat16()
    b = torch.randint(-3, 4, (700, 2 * 64), device="cuda", generator=g).bfloat16()
    got = ops.gemm_bf16_atb_grouped(a, b, 2, 256, 64)
    for d in range(2):
        want = a[:, d * 256:(d + 1) * 256].float().t() @ b[:, d * 64:(d + 1) * 64].float()
        assert torch.equal(got[d * 256:(d + 1) * 256], want)
    for rows, cols in ((700, 512), (1000, 200), (33, 195)):
        x = torch.randn(rows, cols, device="cuda", generator=g).bfloat16()
        torch.testing.assert_close(ops.colsum_bf16(x), x.float().sum(0), rtol=1e-4, atol=1e-3)
