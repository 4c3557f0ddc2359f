import sys, os
sys.path.insert(0, "para-ortho-pca-bo_b200")
import torch
from pcabo_b200 import ops
B, n, kmax, k = 90, 450, 40, 36
Z = torch.rand(B, n, kmax, dtype=torch.float64, device="cuda")
f = torch.randn(B, n, dtype=torch.float64, device="cuda")
kk = torch.full((B,), k, dtype=torch.int32, device="cuda")
buf = ops.GPBuffers(B, n, kmax, torch.device("cuda"))
for _ in range(2):
    ops.gp_fit(buf, Z, kk, f, n, mode=0)
torch.cuda.synchronize()
print("ok", int(buf.info.abs().max()))
