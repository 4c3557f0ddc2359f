"""Extract known-answer vectors from the reference's shipped IOHprofiler files into a small fixture.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/reference_kats.json:
  * doe[fid][instance]    : the 10 DoE rows (x printed to 1e-6, raw_y to 1e-10) of every run block of
                            {pca,vanilla}-experiment/data_f*/IOHprofiler_f*_DIM5.dat (lines 76i+2 .. 76i+11)
  * rows[fid]             : every 7th later row (x, raw_y) -- produced by an older, clipping code version,
                            usable as (x, f(x)) known answers at 1e-6 input precision only
  * best[fid]             : full-precision (x, y) of the `best` entry of each run in
                            pca-experiment/IOHprofiler_f*.json
  * ioh_format            : the layout of the shipped IOHprofiler_f15_RastriginRotated.json / ..._DIM5.dat of both
                            experiments (key order, attribute names, header line, the first rows of the first run
                            block verbatim) -- the known answer for the repo's IOHprofiler writer
`raw_y` is ioh's untransformed objective, i.e. f(x) - f_opt.
"""
import json
import os

REF = "/root/reference"
NAMES = {15: "RastriginRotated", 20: "Schwefel"}


def blocks(path):
    out, cur = [], None
    for line in open(path):
        if line.startswith("evaluations"):
            cur = []
            out.append(cur)
        else:
            cur.append([float(t) for t in line.split()])
    return out


def main():
    kat = {"doe": {}, "rows": {}, "best": {}, "source": "IvanBanny/para-ortho-pca-bo shipped IOHprofiler data (D=5)"}
    for fid, name in NAMES.items():
        bl = blocks(f"{REF}/pca-experiment/data_f{fid}_{name}/IOHprofiler_f{fid}_DIM5.dat")
        blv = blocks(f"{REF}/vanilla-experiment/data_f{fid}_{name}/IOHprofiler_f{fid}_DIM5.dat")
        kat["doe"][str(fid)] = [[{"x": r[3:], "raw_y": r[1]} for r in b[:10]] for b in bl]
        # the vanilla files repeat the same DoE (evaluations numbered 76-85)
        for inst, b in enumerate(blv):
            assert [r[3:] for r in b[:10]] == [r[3:] for r in bl[inst][:10]], (fid, inst)
        rows = []
        for inst, b in enumerate(bl):
            rows += [{"instance": inst, "x": r[3:], "raw_y": r[1]} for r in b[10::7]]
        kat["rows"][str(fid)] = rows
        js = json.load(open(f"{REF}/pca-experiment/IOHprofiler_f{fid}_{name}.json"))
        kat["best"][str(fid)] = [{"instance": r["instance"], "x": r["best"]["x"], "y": r["best"]["y"],
                                  "evals": r["best"]["evals"]}
                                 for r in js["scenarios"][0]["runs"]]
    fmt = {}
    for alg in ("pca", "vanilla"):
        js = json.load(open(f"{REF}/{alg}-experiment/IOHprofiler_f15_RastriginRotated.json"))
        lines = open(f"{REF}/{alg}-experiment/data_f15_RastriginRotated/IOHprofiler_f15_DIM5.dat").read().splitlines()
        run0 = js["scenarios"][0]["runs"][0]
        fmt[alg] = {
            "top_keys": list(js.keys()), "algorithm": js["algorithm"], "version": js["version"], "suite": js["suite"],
            "experiment_attribute_names": [list(a.keys())[0] for a in js["experiment_attributes"]],
            "run_attributes": js["run_attributes"], "attributes": js["attributes"],
            "scenario_keys": list(js["scenarios"][0].keys()), "path": js["scenarios"][0]["path"],
            "run_keys": list(run0.keys()), "best_keys": list(run0["best"].keys()),
            "run0_evals": run0["evals"], "dat_first_lines": lines[:4], "dat_block_lines": 76,
            "dat_blocks": sum(1 for l in lines if l.startswith("evaluations")),
        }
    kat["ioh_format"] = fmt
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")
    with open(out, "w") as fh:
        json.dump(kat, fh)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
