"""GPU: the tcgen05 building blocks (shared-memory operand descriptors, UMMA issue, TMEM read-back) in the
four operand-major combinations the head kernels use, against a torch fp32 matmul of the bf16-rounded
operands (fp32 accumulation: only summation order differs)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("shape", [(176, 80), (64, 96), (112, 96), (256, 48), (96, 128), (16, 16)])
def test_umma_selftest(native_lib, mode, shape):
    from garbage_classification_rca_b200 import _native as N
    n, k = shape
    g = torch.Generator().manual_seed(mode * 100 + n + k)
    A = torch.randn(128, k, generator=g)
    B = torch.randn(n, k, generator=g)
    a_src = (A.t().contiguous() if mode & 2 else A).cuda()   # bit 2 (value 4): two M=64 MMAs
    b_src = (B.t().contiguous() if mode & 1 else B).cuda()
    out = torch.full((128, n), float("nan"), device="cuda")
    N.check(native_lib.mmrca_dev_umma_selftest(mode, a_src.data_ptr(), b_src.data_ptr(), out.data_ptr(), n, k,
                                               torch.cuda.current_stream().cuda_stream), "selftest")
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ B.bfloat16().float().t()
    err = (out.cpu() - ref).abs().max().item()
    assert err < 1e-3 * max(1.0, ref.abs().max().item()), f"mode {mode} n {n} k {k}: max err {err}"


@pytest.mark.parametrize("k", [16, 48, 128])
def test_umma_constant_ones_operand(native_lib, k):
    """mode 8: the all-ones [16 x K] operand of the column-sum MMAs as one re-read K = 16 slice (512 bytes)."""
    from garbage_classification_rca_b200 import _native as N
    n = 16
    g = torch.Generator().manual_seed(k)
    A = torch.randn(128, k, generator=g)
    a_dev, b_dev = A.cuda(), torch.zeros(n, k).cuda()      # (kept alive: the kernel reads them after this statement)
    for rep in range(3):
        out = torch.full((128, n), float("nan"), device="cuda")
        N.check(native_lib.mmrca_dev_umma_selftest(8, a_dev.data_ptr(), b_dev.data_ptr(), out.data_ptr(), n, k,
                                                   torch.cuda.current_stream().cuda_stream), "selftest")
        torch.cuda.synchronize()
        ref = A.bfloat16().float().sum(dim=1, keepdim=True).expand(128, n)
        err = (out.cpu() - ref).abs().max().item()
        assert err < 1e-3 * max(1.0, ref.abs().max().item()), f"k {k}: max err {err}"


@pytest.mark.parametrize("shape", [(64, 16), (128, 64), (192, 64), (160, 128), (176, 32), (256, 64)])
def test_umma_swizzled_mn_major_operands(native_lib, shape):
    """mode 16: both operands MN-major in the SWIZZLE_128B layout a tensor-map box lands from row-major [K][columns]
    memory (the token-level weight-gradient GEMM reads X and the gradients this way)."""
    from garbage_classification_rca_b200 import _native as N
    n, k = shape
    g = torch.Generator().manual_seed(n + k)
    A = torch.randn(128, k, generator=g)
    B = torch.randn(n, k, generator=g)
    a_dev, b_dev = A.t().contiguous().cuda(), B.t().contiguous().cuda()
    out = torch.full((128, n), float("nan"), device="cuda")
    N.check(native_lib.mmrca_dev_umma_selftest(16, a_dev.data_ptr(), b_dev.data_ptr(), out.data_ptr(), n, k,
                                               torch.cuda.current_stream().cuda_stream), "selftest")
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ B.bfloat16().float().t()
    err = (out.cpu() - ref).abs().max().item()
    assert err < 1e-3 * max(1.0, ref.abs().max().item()), f"n {n} k {k}: max err {err}"
