"""Epilogue ablation (GPU box): forward-layer kernel at K = 64 (epilogue-bound) and K = 768 with parts of the epilogue
switched off through csb_test_set_debug: 1 bias staging, 2 math + smem stores, 4 TMA store, 8 TMEM loads."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from climsim_b200 import _lib
from microbench_gemm import time_it
lib = _lib.load()
M, N = 65536, 768
for K in (64, 768):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    Wt = (0.05 * torch.randn(N, K, device="cuda")).to(torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for flags, name in [(0, "full"), (1, "-bias"), (4, "-tma_store"), (5, "-bias -store"), (2, "-math/sts"), (6, "-math -store"), (7, "-bias -math -store"),
                        (15, "-everything (only barriers)"), (8, "-ldtm")]:
        lib.csb_test_set_debug(flags)
        us = time_it(lambda: _lib.check(lib.csb_test_linear_fwd(A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, 3, 0.15, 1, None), "fwd"))
        print(f"K={K:4d} {name:30s} {us:7.1f} us  {us / (512 * 3 / 148):5.2f} us/tile/SM", flush=True)
lib.csb_test_set_debug(0)
