"""Weight-gradient kernel micro-benchmark (GPU box): gemm_nt_kernel with single CTAs (cg = 1) against CTA pairs (cg = 2) on the
MLP_v1 layer shapes at the benchmark batch, operands larger than L2 rotated between launches.
    python scripts/microbench_nt.py [R]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from climsim_b200 import _lib

lib = _lib.load()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
SM = torch.cuda.get_device_properties(0).multi_processor_count


def time_it(fns, reps=24):
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3          # us


def wgrad(M, N, cg, splits=None):
    bn = 256 if N > 128 else 128
    m_tiles = ((M + 127) // 128 + cg - 1) // cg
    tiles = m_tiles * ((N + bn - 1) // bn)
    if splits is None:
        splits = max(1, min(64, (SM // cg) // tiles))
    ops = [(torch.randn(R, M, device="cuda").to(torch.bfloat16), torch.randn(R, N, device="cuda").to(torch.bfloat16)) for _ in range(3)]
    C = torch.empty(splits, M, N, device="cuda")
    cs = torch.empty(splits * ((M + 127) // 128), N, device="cuda")
    mt = ctypes.c_int(0)

    def mk(A, B):
        return lambda: _lib.check(lib.csb_test_gemm_nt_cg(A.data_ptr(), B.data_ptr(), C.data_ptr(), cs.data_ptr(), M, N, R, splits, cg,
                                                           ctypes.byref(mt), None), "nt")
    us = time_it([mk(a, b) for a, b in ops])
    print(f"wgrad M={M:4d} N={N:4d} R={R} cg={cg} splits={splits:2d} ctas={tiles * cg * splits:3d}: {us:8.1f} us  {2 * M * N * R / us / 1e6:7.1f} TF/s",
          flush=True)
    return us


if __name__ == "__main__":
    shapes = [(128, 768), (768, 640), (640, 512), (512, 640), (640, 640), (640, 128), (128, 128)]
    tot = {1: 0.0, 2: 0.0}
    for M, N in shapes:
        for cg in (1, 2):
            if cg == 2 and (N % 128 or M <= 128):
                tot[2] += wgrad(M, N, 1)
                continue
            tot[cg] += wgrad(M, N, cg)
    print(f"sum over the MLP_v1 layers: cg=1 {tot[1]:.1f} us, cg=2 where applicable {tot[2]:.1f} us")
    a = torch.randn(R, 768, device="cuda").to(torch.bfloat16)
    b = torch.randn(R, 640, device="cuda").to(torch.bfloat16)
    us = time_it([lambda: a.t() @ b])
    print(f"cuBLAS bf16 wgrad 768x640x{R}: {us:8.1f} us  {2 * R * 640 * 768 / us / 1e6:7.1f} TF/s")
