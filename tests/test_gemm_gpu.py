"""GPU unit tests of the hand-written tcgen05 GEMM kernels through the C-ABI self-test hooks.

Reference: torch fp32 matmul of the same bf16-rounded operands (the tensor cores multiply bf16 exactly and accumulate
in fp32, so only the summation order differs: tolerance 1e-3 relative to the row scale is generous)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from climsim_b200 import _lib
    return _lib.load(), _lib


def _bf16_operand(rows, cols, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    a = (scale * torch.randn(rows, cols, generator=g)).to(torch.bfloat16).cuda()
    return a


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 128, 64, 128),        # one tile, one k-block: the descriptor / swizzle smoke test
    (128, 128, 128, 128),       # two k-blocks (accumulate flag, smem ring advance)
    (128, 256, 64, 256),        # N = 256 instruction
    (256, 128, 512, 128),       # ring wraps (6 stages, 8 k-blocks)
    (1000, 640, 768, 256),      # ragged M, ragged last n-block (640 = 256 + 256 + 128), layer-1 shape
    (1000, 640, 768, 128),
    (4096, 768, 128, 256),      # layer-0 shape
    (333, 64, 192, 128),        # N = 64 (n_valid < BN)
    (70000, 128, 128, 128),     # many tiles per CTA: both TMEM accumulator buffers, phase flips
    (256, 256, 64, 512),        # bn code 512 = 256-wide tiles on CTA pairs (tcgen05 cta_group::2): smallest case
    (256, 256, 512, 512),       # pairs: ring wraps
    (1000, 640, 768, 512),      # pairs: odd number of m-blocks (idle half-pair at the tail), ragged last n-block
    (65536, 768, 128, 512),     # pairs: layer-0 shape at the benchmark batch
    (40000, 640, 640, 512),     # pairs: many tiles per pair
])
def test_gemm_tn(M, N, K, bn):
    lib, L = _lib()
    A = _bf16_operand(M, K, 1)
    Bt = _bf16_operand(N, K, 2)
    Cout = torch.full((M, N), float("nan"), device="cuda")
    L.check(lib.csb_test_gemm_tn(A.data_ptr(), Bt.data_ptr(), Cout.data_ptr(), M, N, K, bn, None), "csb_test_gemm_tn")
    torch.cuda.synchronize()
    ref = A.float() @ Bt.float().t()
    err = (Cout - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 1e-3 * scale, (err, scale)


@pytest.mark.parametrize("M,N,R,splits", [
    (128, 128, 64, 1),          # one 64-row block: MN-major descriptor smoke test
    (128, 128, 128, 1),
    (128, 256, 256, 1),
    (64, 64, 1000, 3),          # M = N = 64 (single chunk), ragged R, split over rows
    (768, 640, 4096, 4),        # dW of layer 1
    (128, 768, 5000, 7),        # dW of layer 0, ragged R, splits that do not divide
    (640, 128, 70000, 16),
])
def test_gemm_nt(M, N, R, splits):
    lib, L = _lib()
    A = _bf16_operand(R, M, 3)
    B = _bf16_operand(R, N, 4)
    Cout = torch.full((splits, M, N), float("nan"), device="cuda")
    colsum = torch.full((splits * ((M + 127) // 128), N), float("nan"), device="cuda")     # one partial per (split, m-block)
    L.check(lib.csb_test_gemm_nt(A.data_ptr(), B.data_ptr(), Cout.data_ptr(), colsum.data_ptr(), M, N, R, splits, None), "csb_test_gemm_nt")
    torch.cuda.synchronize()
    got = Cout.sum(dim=0)
    ref = A.float().t() @ B.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 1e-3 * scale, (err, scale)
    # fused bias gradient: column sums of B over the contraction rows
    cs_ref = B.float().sum(dim=0)
    cs_err = (colsum.sum(dim=0) - cs_ref).abs().max().item()
    assert cs_err <= 1e-3 * max(cs_ref.abs().max().item(), 1.0), cs_err


@pytest.mark.parametrize("M,N,R,splits", [
    (256, 128, 64, 1),          # smallest CTA pair: one 64-row block, 64 dZ columns per CTA
    (256, 256, 1024, 1),        # ring wraps (6 stages), no split
    (768, 640, 4096, 4),        # dW of layer 1: ragged last n-block (128 wide -> 64 columns per CTA)
    (640, 640, 5000, 7),        # odd m-block count (idle half pair), ragged R, splits that do not divide
    (640, 128, 70000, 16),      # 128-wide layer on pairs (BN = 128 instantiation)
    (512, 640, 100, 3),         # more splits than the rows need: trailing splits get no work and write zeros
    (192, 384, 3000, 5),        # M = 192: the last m-block is half a chunk short
])
def test_gemm_nt_cta_pairs(M, N, R, splits):
    """gemm_nt_kernel<.., CG = 2>: 256 x BN weight-gradient tiles on tcgen05 cta_group::2 pairs."""
    import ctypes
    lib, L = _lib()
    A = _bf16_operand(R, M, 5)
    B = _bf16_operand(R, N, 6)
    Cout = torch.full((splits, M, N), float("nan"), device="cuda")
    colsum = torch.full((splits * ((M + 127) // 128), N), float("nan"), device="cuda")
    m_tiles = ctypes.c_int(0)
    L.check(lib.csb_test_gemm_nt_cg(A.data_ptr(), B.data_ptr(), Cout.data_ptr(), colsum.data_ptr(), M, N, R, splits, 2,
                                    ctypes.byref(m_tiles), None), "csb_test_gemm_nt_cg")
    torch.cuda.synchronize()
    assert m_tiles.value == ((M + 127) // 128 + 1) // 2
    got = Cout.sum(dim=0)
    ref = A.float().t() @ B.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 1e-3 * scale, (err, scale)
    cs_ref = B.float().sum(dim=0)
    cs = colsum[: splits * m_tiles.value]
    cs_err = (cs.sum(dim=0) - cs_ref).abs().max().item()
    assert cs_err <= 1e-3 * max(cs_ref.abs().max().item(), 1.0), cs_err
    # bit-reproducible (no atomics, fixed split geometry)
    Cout2 = torch.empty_like(Cout)
    L.check(lib.csb_test_gemm_nt_cg(A.data_ptr(), B.data_ptr(), Cout2.data_ptr(), colsum.data_ptr(), M, N, R, splits, 2,
                                    ctypes.byref(m_tiles), None), "csb_test_gemm_nt_cg")
    torch.cuda.synchronize()
    assert torch.equal(Cout, Cout2)


@pytest.mark.parametrize("M,N,K,act", [
    (4096, 768, 128, 3),        # layer-0 shape: staged epilogue on CTA pairs (K <= 256), LeakyReLU
    (1000, 640, 128, 1),        # ragged M, ragged last n-block (128 of 256 columns), ReLU
    (333, 128, 128, 0),         # 128-wide tiles (32 columns per epilogue warp), no activation
    (70000, 128, 256, 3),       # many tiles per CTA: the staging tile is reused across tiles
    (513, 192, 64, 2),          # ELU instantiation, N not a multiple of 128
    (2048, 640, 768, 3),        # long contraction: register -> global epilogue (not staged), for comparison
])
def test_linear_fwd_engine_policy(M, N, K, act):
    """One forward layer through the engine's launch policy (csb_test_linear_fwd, pairs = 2) against torch fp32 on the same bf16
    operands: covers the staged (shared memory -> coalesced stores) epilogue of the short-contraction launches."""
    lib, L = _lib()
    A = _bf16_operand(M, K, 7)
    Wt = (_bf16_operand(N, K, 8).float() * 0.1).to(torch.bfloat16)
    bias = torch.linspace(-0.5, 0.5, N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(lib.csb_test_linear_fwd(A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, act, 0.15, 2, None), "csb_test_linear_fwd")
    torch.cuda.synchronize()
    z = A.float() @ Wt.float().t() + bias
    ref = {0: z, 1: torch.relu(z), 2: torch.nn.functional.elu(z), 3: torch.nn.functional.leaky_relu(z, 0.15)}[act]
    assert not torch.isnan(out.float()).any()
    err = (out.float() - ref).abs().max().item()
    assert err <= 1e-2 * ref.abs().max().item(), err          # one bf16 rounding of the result


@pytest.mark.parametrize("M,N,R,splits,cg", [
    (768, 640, 5000, 9, 2),     # dW of layer 1 with the engine's geometry: 9 slots, the 128-wide block takes 5 longer splits
    (640, 640, 4096, 8, 2),     # even slot count, odd m-block count
    (512, 640, 70000, 14, 2),   # layer 3 at a batch with ragged row blocks
    (256, 384, 300, 7, 1),      # single CTAs, 256 + 128 columns, more slots than the narrow block has row blocks for
    (768, 640, 64, 2, 2),       # one row block in total: most CTAs only write zeros
])
def test_gemm_nt_uneven_splits(M, N, R, splits, cg):
    """Uneven split counts for a half-width last n-block (NtParams.splits_narrow): every slot of every column is written (the NaN
    fill would survive otherwise), the slots sum to the product, the bias-gradient partials to the column sums."""
    import ctypes
    lib, L = _lib()
    A = _bf16_operand(R, M, 9)
    B = _bf16_operand(R, N, 10)
    Cout = torch.full((splits, M, N), float("nan"), device="cuda")
    m_blocks = (M + 127) // 128
    colsum = torch.full((splits * m_blocks, N), float("nan"), device="cuda")
    m_tiles = ctypes.c_int(0)
    L.check(lib.csb_test_gemm_nt_cg(A.data_ptr(), B.data_ptr(), Cout.data_ptr(), colsum.data_ptr(), M, N, R, -splits, cg,
                                    ctypes.byref(m_tiles), None), "csb_test_gemm_nt_cg")
    torch.cuda.synchronize()
    assert not torch.isnan(Cout).any()
    ref = A.float().t() @ B.float()
    assert (Cout.sum(dim=0) - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()
    narrow_slots = (splits + 1) // 2
    assert (Cout[narrow_slots:, :, N - 128:] == 0).all()                      # the slots the narrow block leaves unused hold zeros
    cs = colsum[: splits * m_tiles.value]
    assert not torch.isnan(cs).any()
    cs_ref = B.float().sum(dim=0)
    assert (cs.sum(dim=0) - cs_ref).abs().max().item() <= 1e-3 * max(cs_ref.abs().max().item(), 1.0)


def _tf32_trunc(t):
    """Keep sign, exponent and the top 10 mantissa bits: what kind::tf32 reads of an fp32 word."""
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 128, 32, 128),        # one tile, one k-block of 32 fp32 elements
    (1000, 640, 768, 128),
    (333, 64, 96, 128),         # N = 64, K = three k-blocks
    (70000, 128, 128, 128),     # both accumulator buffers, phase flips
    (256, 256, 64, 512),        # CTA pairs
    (1000, 640, 768, 512),      # pairs, ragged last n-block, odd m-block count
    (65536, 768, 128, 512),     # layer-0 shape at the benchmark batch
])
def test_gemm_tn_tf32(M, N, K, bn):
    """The same tcgen05 kernel on fp32 operands (kind::tf32): with operands that are exactly representable in TF32 the result is the
    fp32 product up to the summation order; with arbitrary fp32 operands it is the product of the TRUNCATED operands (the tensor core
    drops the low 13 mantissa bits), i.e. within 2^-10 relative of the fp32 product per term."""
    lib, L = _lib()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    Bt = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    Cout = torch.full((M, N), float("nan"), device="cuda")
    L.check(lib.csb_test_gemm_tn_tf32(A.data_ptr(), Bt.data_ptr(), Cout.data_ptr(), M, N, K, bn, None), "csb_test_gemm_tn_tf32")
    torch.cuda.synchronize()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref_trunc = (_tf32_trunc(A).double() @ _tf32_trunc(Bt).double().t()).float()
        ref_full = (A.double() @ Bt.double().t()).float()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    scale = ref_full.abs().max().item()
    err_t = (Cout - ref_trunc).abs().max().item()
    err_f = (Cout - ref_full).abs().max().item()
    print(f"tf32 gemm {M}x{N}x{K} bn {bn}: vs truncated operands {err_t / scale:.2e}, vs fp32 product {err_f / scale:.2e}")
    assert err_t <= 2e-6 * scale * max(1.0, (K / 128) ** 0.5), (err_t, scale)
    assert err_f <= 2e-3 * scale, (err_f, scale)
