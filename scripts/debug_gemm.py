"""Diagnostic driver for the tcgen05 GEMM kernels (run on the GPU box): prints per-case errors and, for a failing
case, enough structure (which rows / columns / k-slices are wrong) to diagnose descriptor or swizzle mistakes."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from climsim_b200 import _lib

lib = _lib.load()


def report(name, got, ref):
    err = (got - ref).abs()
    scale = ref.abs().max().item()
    nan = torch.isnan(got).sum().item()
    print(f"{name}: max_err {err.max().item():.4e} scale {scale:.3e} nan {nan} "
          f"bad_frac {(err > 1e-3 * scale).float().mean().item():.4f}", flush=True)
    if err.max().item() > 1e-3 * scale or nan:
        bad = (err > 1e-3 * scale) | torch.isnan(got)
        rows = bad.any(dim=1).nonzero().flatten()[:16].tolist()
        cols = bad.any(dim=0).nonzero().flatten()[:16].tolist()
        print("   bad rows (first 16):", rows, " bad cols (first 16):", cols)
        print("   got[0,:8]", got[0, :8].tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())
        return False
    return True


def tn(M, N, K, bn, structured=False):
    g = torch.Generator().manual_seed(1)
    if structured:
        A = torch.zeros(M, K); A[:, 0] = 1.0                       # picks column 0 of Bt^T: C[m, n] = Bt[n, 0]
        Bt = torch.arange(N * K, dtype=torch.float32).reshape(N, K) % 251
    else:
        A, Bt = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    A, Bt = A.to(torch.bfloat16).cuda(), Bt.to(torch.bfloat16).cuda()
    C = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.csb_test_gemm_tn(A.data_ptr(), Bt.data_ptr(), C.data_ptr(), M, N, K, bn, None)
    if rc:
        print("tn launch error", rc, lib.csb_last_error().decode()); return False
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print("tn sync error", e); return False
    return report(f"tn M{M} N{N} K{K} bn{bn} structured={structured}", C, A.float() @ Bt.float().t())


def nt(M, N, R, splits, structured=False):
    g = torch.Generator().manual_seed(2)
    if structured:
        A = torch.zeros(R, M); A[0, :] = 1.0                        # C[m, n] = B[0, n]
        B = torch.arange(R * N, dtype=torch.float32).reshape(R, N) % 251
    else:
        A, B = torch.randn(R, M, generator=g), torch.randn(R, N, generator=g)
    A, B = A.to(torch.bfloat16).cuda(), B.to(torch.bfloat16).cuda()
    C = torch.full((splits, M, N), float("nan"), device="cuda")
    rc = lib.csb_test_gemm_nt(A.data_ptr(), B.data_ptr(), C.data_ptr(), None, M, N, R, splits, None)
    if rc:
        print("nt launch error", rc, lib.csb_last_error().decode()); return False
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print("nt sync error", e); return False
    return report(f"nt M{M} N{N} R{R} s{splits} structured={structured}", C.sum(0), A.float().t() @ B.float())


TN_CASES = [(256, 256, 64, 512, True), (256, 256, 64, 512), (256, 256, 512, 512), (1000, 640, 768, 512), (65536, 640, 768, 512),
            (128, 128, 64, 128, True), (128, 128, 64, 128), (333, 64, 192, 128), (256, 128, 512, 128),
            (1000, 640, 768, 256), (70000, 128, 128, 128), (65536, 640, 768, 256)]
NT_CASES = [(128, 128, 64, 1, True), (128, 128, 64, 1), (128, 128, 128, 1), (128, 256, 256, 1), (64, 64, 1000, 3),
            (768, 640, 4096, 4), (640, 128, 70000, 16)]

if __name__ == "__main__":
    # every case in its own process: a trapped kernel poisons the CUDA context
    if len(sys.argv) >= 3 and sys.argv[1] == "one":
        kind, args = sys.argv[2], [int(a) for a in sys.argv[3:]]
        fn = tn if kind == "tn" else nt
        structured = bool(args[4]) if len(args) > 4 else False
        sys.exit(0 if fn(*args[:4], structured=structured) else 1)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    ok = True
    for kind, cases in (("tn", TN_CASES), ("nt", NT_CASES)):
        if which not in ("all", kind):
            continue
        for c in cases:
            r = subprocess.run([sys.executable, __file__, "one", kind] + [str(int(a)) for a in c], capture_output=True, text=True, timeout=120)
            print(r.stdout.strip() or r.stderr.strip()[-500:], flush=True)
            ok &= r.returncode == 0
    print("ALL OK" if ok else "FAILURES")
