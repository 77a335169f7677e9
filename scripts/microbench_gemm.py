"""Kernel micro-benchmark (GPU box): time the forward-layer GEMM (bias + LeakyReLU epilogue, bf16 TMA store) and the
weight-gradient GEMM as a function of the contraction length, to separate mainloop-bound from epilogue-bound time."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from climsim_b200 import _lib

lib = _lib.load()


def time_it(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3          # us


def fwd(M, N, K, pairs):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    Wt = (0.05 * torch.randn(N, K, device="cuda")).to(torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

    def run():
        _lib.check(lib.csb_test_linear_fwd(A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, 3, 0.15, pairs, None), "fwd")
    us = time_it(run)
    ref = torch.nn.functional.leaky_relu(A[:256].float() @ Wt.float().t(), 0.15)
    err = (out[:256].float() - ref).abs().max().item() / ref.abs().max().item()
    tiles_per_sm = (M / 128) * (N / 256) / 148
    print(f"fwd pairs={pairs} M={M} N={N} K={K}: {us:8.1f} us  {2 * M * N * K / us / 1e6:7.1f} TF/s  {us / tiles_per_sm:6.2f} us/tile/SM  err {err:.1e}", flush=True)


def wgrad(M, N, R, splits):
    A = torch.randn(R, M, device="cuda").to(torch.bfloat16)
    B = torch.randn(R, N, device="cuda").to(torch.bfloat16)
    C = torch.empty(splits, M, N, device="cuda")

    def run():
        _lib.check(lib.csb_test_gemm_nt(A.data_ptr(), B.data_ptr(), C.data_ptr(), None, M, N, R, splits, None), "nt")
    us = time_it(run)
    print(f"wgrad M={M} N={N} R={R} splits={splits}: {us:8.1f} us  {2 * M * N * R / us / 1e6:7.1f} TF/s", flush=True)


if __name__ == "__main__":
    for pairs in (0, 1):
        for K in (64, 128, 256, 512, 768, 1536, 3072):
            fwd(65536, 768, K, pairs)
    for pairs in (0, 1):
        fwd(65536, 640, 768, pairs)
        fwd(65536, 512, 640, pairs)
    torch_a = torch.randn(65536, 768, device="cuda").to(torch.bfloat16)
    torch_b = torch.randn(640, 768, device="cuda").to(torch.bfloat16)
    us = time_it(lambda: torch_a @ torch_b.t())
    print(f"cuBLAS bf16 65536x640x768: {us:8.1f} us  {2 * 65536 * 640 * 768 / us / 1e6:7.1f} TF/s")
    for splits in (1, 8):
        wgrad(768, 640, 65536, splits * 1)
    wgrad(768, 768, 65536, 8)
    wgrad(768, 768, 65536 * 4, 8)
    us = time_it(lambda: torch_a.t() @ torch_a[:, :640])
    print(f"cuBLAS bf16 wgrad 768x640x65536: {us:8.1f} us  {2 * 65536 * 640 * 768 / us / 1e6:7.1f} TF/s")
