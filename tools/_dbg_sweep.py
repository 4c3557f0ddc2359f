import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/para-ortho-pca-bo_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from test_gpu_sweep import make_gp, sobol_params
from pcabo_b200 import ops
for (n,r,kmax) in [(450,40,40),(450,8,40),(123,16,16)]:
    buf, acq, lo, hi = make_gp(n, r, kmax, seed=n + r)
    N=4096+5
    eng, shift, sst = sobol_params(r, kmax, 1234)
    pts = -1/3 + (5/3) * torch.quasirandom.SobolEngine(r, scramble=True, seed=1234).draw(N, dtype=torch.float64)
    ref = acq(pts).numpy()
    ws = ops.SweepWorkspace(torch.device("cuda"))
    for rep in range(3):
        vals = torch.empty(N, dtype=torch.float64, device="cuda")
        bv, bi = ops.ei_sweep(buf, ws, N, 0, sobol=(shift, sst, lo, hi), vals=vals)
        torch.cuda.synchronize()
        got = vals.cpu().numpy()
        err=np.abs(got-ref)/np.maximum(np.abs(ref),1e-6)
        bad=np.nonzero(err>1e-8)[0]
        print(n,r,kmax,'rep',rep,'nbad',len(bad),'tiles',sorted(set((bad//32).tolist()))[:20],'cols',sorted(set((bad%32).tolist())), 'maxerr',err.max())
