"""Mainloop probe (GPU box): forward-layer kernel with parts of the pipeline switched off (csb_test_set_debug) and the SM
clock measured inside the kernel (clock64 / globaltimer per CTA), to tell ingest-bound from MMA-bound from epilogue-bound.
flags: 1 bias, 2 math + stores, 4 stores, 8 TMEM loads, 16 no TMA loads, 32 no MMAs, 64 no B loads, 128 no A loads.
    python scripts/probe_mainloop.py [pairs]     pairs > 0: run on that many CTA pairs only (stays below the power cap)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from climsim_b200 import _lib
from microbench_gemm import time_it
lib = _lib.load()
M = 65536
stats = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
lib.csb_test_set_stats(stats.data_ptr())
CASES = [(0, "full"), (4, "no stores"), (2, "no math, no stores"), (1, "no bias"), (15, "epilogue off"), (15 + 16, "epi off, no loads"),
         (15 + 32, "epi off, no MMA (ingest only)"), (16, "full epilogue, no loads"), (32, "full epilogue, no MMA"), (16 + 32, "epilogue only")]
GRID = int(sys.argv[1]) if len(sys.argv) > 1 else 0      # CTA pairs to run on (0 = all): a few pairs stay far below the power cap
if GRID:
    M = 256 * GRID * 4
for (N, K) in ((768, 768),):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    Wt = (0.05 * torch.randn(N, K, device="cuda")).to(torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for pairs in (1,):
        for flags, name in CASES:
            lib.csb_test_set_debug(flags | (GRID << 16))
            us = time_it(lambda: _lib.check(lib.csb_test_linear_fwd(A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, 3, 0.15, pairs, None), "fwd"))
            if flags == 0 and False:
                ref = torch.nn.functional.leaky_relu(A[:256].float() @ Wt.float().t(), 0.15)
                print("   check err", ((out[:256].float() - ref).abs().max() / ref.abs().max()).item())
            s = stats.view(148, 8)[: (2 * GRID if GRID else 148)].cpu().double()
            mhz = ((s[:, 1] - s[:, 0]) / (s[:, 3] - s[:, 2]).clamp(min=1) * 1e3).median().item()
            cyc = (s[:, 1] - s[:, 0]).median().item()
            tiles = (M / 128) * ((N + 255) // 256) / (2 * GRID if GRID else 148)
            print(f"N={N} K={K} pairs={pairs} {name:32s} {us:7.1f} us  {2*M*N*K/us/1e6:7.1f} TF/s  SM {mhz:6.0f} MHz  {cyc/tiles:8.0f} cyc/tile/SM"
                  f"  (MMA floor {K * 8 * N / (((N + 255) // 256) * 256):6.0f})  issuer waits: acc {s[::2, 4].median().item()/tiles:6.0f} operands {s[::2, 5].median().item()/tiles:6.0f}"
                  f"  epi warp: wait {s[:, 6].median().item()/tiles:6.0f} busy {s[:, 7].median().item()/tiles:6.0f}", flush=True)
lib.csb_test_set_debug(0)
lib.csb_test_set_stats(None)
