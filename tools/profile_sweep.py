"""Per-warp phase cycle counters of the EI sweep kernel (tuning aid). Needs a library built with -DSWEEP_PROF:
    PCABO_NVCC_EXTRA=-DSWEEP_PROF python para-ortho-pca-bo_b200/pcabo_b200/build.py --force
    python tools/profile_sweep.py            (on the B200)   then rebuild without the flag."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "para-ortho-pca-bo_b200"))
import numpy as np

NAMES = ["tile_top+loop", "produce", "issue_w", "wait_w", "wait_full", "mma", "epilogue", "-"]

if __name__ == "__main__":
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_sweep.py"), "--log2n", "22", "--dims", os.environ.get("SWEEP_R", "40"), "--reps", "1"])
    # same process is needed to read the symbol: rerun in-process
    import runpy
    sys.argv = ["bench_sweep.py", "--log2n", "20", "--dims", os.environ.get("SWEEP_R", "40"), "--reps", "1"]
    from pcabo_b200.lib import load_library
    lib = load_library()
    buf = (ctypes.c_longlong * 128)()
    runpy.run_path(os.path.join(ROOT, "tools", "bench_sweep.py"), run_name="__main__")
    lib.pcabo_sweep_prof_read.argtypes = [ctypes.c_void_p]
    assert lib.pcabo_sweep_prof_read(buf) == 0
    p = np.array(buf[:], dtype=np.float64).reshape(16, 8)
    tot = p.sum(1, keepdims=True)
    print("warp  " + "  ".join(f"{n:>13s}" for n in NAMES[:7]))
    for w in range(16):
        print(f"{w:4d}  " + "  ".join(f"{100 * v / tot[w, 0]:12.1f}%" for v in p[w, :7]))
    print("mean  " + "  ".join(f"{100 * v:12.1f}%" for v in (p / tot).mean(0)[:7]))
    print("Mcycles per warp (mean):", tot.mean() / 1e6)
