"""One forward-layer launch per shape (for ncu captures): python scripts/one_fwd.py N K [N K ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from climsim_b200 import _lib
lib = _lib.load()
M = 65536
args = [int(a) for a in sys.argv[1:]] or [768, 128, 768, 768]
for N, K in zip(args[0::2], args[1::2]):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    Wt = (0.05 * torch.randn(N, K, device="cuda")).to(torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        _lib.check(lib.csb_test_linear_fwd(A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, 3, 0.15, 1, None), "fwd")
    torch.cuda.synchronize()
