"""Library yardsticks missing from MEASURED_PEAKS.json: FP64 DGEMM peak and batched cuSOLVER Cholesky (run on the B200)."""
import json
import sys
import time

import torch


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1e-3)
    return best


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    N = 8192
    a = torch.randn(N, N, dtype=torch.float64, device="cuda")
    b = torch.randn(N, N, dtype=torch.float64, device="cuda")
    t = timed(lambda: torch.matmul(a, b))
    out["dgemm_8192_tflops"] = 2 * N ** 3 / t / 1e12
    for (B, n) in [(90, 450), (90, 285), (150, 1050)]:
        A = torch.randn(B, n, n, dtype=torch.float64, device="cuda")
        A = A @ A.transpose(1, 2) + n * torch.eye(n, dtype=torch.float64, device="cuda")
        t = timed(lambda: torch.linalg.cholesky(A))
        out[f"potrf_batched_B{B}_n{n}_ms"] = t * 1e3
        out[f"potrf_batched_B{B}_n{n}_gflops"] = B * n ** 3 / 3 / t / 1e9
    # our batched GP fit (K build + Cholesky + L^-1 + alpha + transposition + packing: the tile pipeline of gp_chol.cu)
    # on the same shapes
    try:
        import os
        ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, os.path.join(ROOT, "para-ortho-pca-bo_b200"))
        from pcabo_b200 import ops
        for (B, n, k) in [(90, 450, 36), (30, 450, 36), (90, 285, 36), (150, 1050, 90)]:
            kmax = 40 if k <= 40 else 100
            Z = torch.rand(B, n, kmax, dtype=torch.float64, device="cuda")
            f = torch.randn(B, n, dtype=torch.float64, device="cuda")
            kk = torch.full((B,), k, dtype=torch.int32, device="cuda")
            buf = ops.GPBuffers(B, n, kmax, torch.device("cuda"))
            t = timed(lambda: ops.gp_fit(buf, Z, kk, f, n, mode=0))
            out[f"pcabo_gp_fit_B{B}_n{n}_ms"] = t * 1e3
            # n^3/3 Cholesky + n^3/3 triangular inverse + n^2 (3k + 14) / 2 kernel matrix
            out[f"pcabo_gp_fit_B{B}_n{n}_gflops"] = B * (2 * n ** 3 / 3 + n * n * (3 * k + 14) / 2) / t / 1e9
    except Exception as e:  # the yardsticks above do not need the library
        out["pcabo_gp_fit_error"] = repr(e)
    # HBM copy for reference
    x = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    t = timed(lambda: y.copy_(x))
    out["copy_gbs"] = 2 * x.numel() * 8 / t / 1e9
    print(json.dumps(out))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
