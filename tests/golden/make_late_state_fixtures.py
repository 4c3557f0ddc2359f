"""Real-trajectory GP states for the parity tests (inputs only: the oracle is evaluated by the test itself).

Made on the B200 box by `python tools/dump_late_state.py 330 0,39,30` (config C3 advanced to n = 330 by the engine:
clustered points, 1000-penalties, reduced-space scales from 1e2 to 1e12), then reduced to the fields the tests need:
    python tests/golden/make_late_state_fixtures.py gpurun_out/late_state_0.npz gpurun_out/late_state_39.npz ...
Nothing here comes from the reference repository; the files pin states that random synthetic histories do not reach."""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    for path in sys.argv[1:]:
        d = np.load(path)
        name = os.path.basename(path)
        np.savez_compressed(os.path.join(HERE, name), Z=d["Z"], f=d["f"], best=float(d["best"]), ics=d["ics"])
        print(name, d["Z"].shape)
