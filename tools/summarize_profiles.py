"""Turn the scratch outputs of tools/collect_round.sh (gpurun_out/) into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r01
Needs ncu on PATH (reads the .ncu-rep files with `ncu -i ... --page raw --csv`)."""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct")


def raw_summary(rep, dst, header):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(dst, "w") as f:
        f.write("# " + header + "\n")
        for h, u, v in sorted(zip(hdr, units, vals)):
            if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h):
                f.write(f"{h},{u},{v}\n")


def launch_summary(src, dst, header):
    agg = collections.OrderedDict()
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        name = r["Kernel Name"]
        if "at::" in name or "cub::" in name or name.startswith("void at") or "elementwise" in name:
            name = "(torch plumbing: fills / copies / index ops)"
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) * 1e-6
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# " + header + "\n")
        f.write("kernel,launches,total_ms,avg_us,share\n")
        for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{c},{ms:.3f},{1e3 * ms / c:.1f},{ms / tot:.4f}\n")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    os.makedirs(PROF, exist_ok=True)
    j = lambda n: os.path.join(OUT, n)
    for cand in ("launches_%s.csv" % tag, "launches_%s_final.csv" % tag):
        if os.path.exists(j(cand)):
            shutil.copy(j(cand), os.path.join(PROF, f"{tag}_launches_bench_steps2.csv"))
            launch_summary(j(cand), os.path.join(PROF, f"{tag}_launch_summary.csv"),
                           "ncu launch list summary (ncu --metrics gpu__time_duration.sum --clock-control none, "
                           "bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep --sub-batches 1, device optimizer, n = 120..125; "
                           "the cutlass d884gemm launches are the bench's own in-run DGEMM yardstick, not part of a step); per-launch "
                           "times are cold-cache and serialised: compare SHARES with bench.py's roofline.share_of_step")
            break
    for rep, dst, hd in [
        (f"prof_lbfgsb_{tag}.ncu-rep", f"{tag}_lbfgsb_ncu_raw_summary.csv",
         "ncu --set full, lbfgsb_logei_kernel<5, 32, 8, false> (resident W panels), tools/profile_lbfgsb.py 3 (n = 123, 90 runs x 2 chunks)"),
        (f"lb_v8_late.ncu-rep", f"{tag}_lbfgsb_n430_ncu_raw_summary.csv",
         "ncu --set full, lbfgsb_logei_kernel<5, 16, 8> streaming W panels at n ~ 430 (tools/profile_lbfgsb.py 310), mid-round build"),
        (f"prof_chol_{tag}.ncu-rep", f"{tag}_chol_update_ncu_raw_summary.csv",
         "ncu --set full, chol_update_kernel (one launch of the left-looking Cholesky pipeline), tools/measure_fp64.py"),
        (f"prof_triinv_{tag}.ncu-rep", f"{tag}_tri_inverse_ncu_raw_summary.csv",
         "ncu --set full, tri_inverse_kernel (block columns of L^-1), tools/measure_fp64.py"),
        (f"prof_lbfgsb_{tag}_final.ncu-rep", f"{tag}_lbfgsb_ncu_raw_summary.csv",
         "ncu --set full, lbfgsb_logei_kernel<5>, bench.py --steps 3 --warmup 1 --sub-batches 1, n = 122, 90 runs x 2 chunks"),
        (f"prof_sweep_{tag}_final.ncu-rep", f"{tag}_ei_sweep_ncu_raw_summary.csv",
         "ncu --set full, ei_sweep_kernel (v6: mbarrier K* ring, per-warp W staging, asynchronous epilogue), "
         "tools/bench_sweep.py --log2n 22 --dims 40 (n = 450)")]:
        if os.path.exists(j(rep)):
            raw_summary(j(rep), os.path.join(PROF, dst), hd)
    for src, dst in [("sweep_final.jsonl", f"{tag}_ei_sweep_bench.jsonl"), ("fp64_peaks2.json", f"{tag}_fp64_peaks.json"),
                     (f"{tag}_lbfgsb_phase_cycles_n123.json", f"{tag}_lbfgsb_phase_cycles_n123.json"),
                     (f"{tag}_lbfgsb_phase_cycles_n430.json", f"{tag}_lbfgsb_phase_cycles_n430.json"),
                     (f"bench_{tag}_final.json", f"{tag}_bench_c3_full_run.json"),
                     (f"bench_{tag}_reference.json", f"{tag}_bench_c3_full_run_reference_arm.json")]:
        if os.path.exists(j(src)):
            shutil.copy(j(src), os.path.join(PROF, dst))
    if os.path.exists(j("trace_ph3.json")):
        t = json.load(open(j("trace_ph3.json")))
        ph, s = t["step_phases_ms"], t["step_ms"]
        keys = sorted({k for p in ph for k in p})
        with open(os.path.join(PROF, f"{tag}_step_phases_by_n.csv"), "w") as f:
            f.write("# bench.py --phases --trace (adds a device sync per phase): mean ms per lock-step BO iteration of 90 runs, "
                    "buckets of 40 steps\n")
            f.write("n_first,step_ms," + ",".join(keys) + "\n")
            for i in range(0, len(s), 40):
                m = len(s[i:i + 40])
                f.write(f"{t['n_first'] + i},{sum(s[i:i + 40]) / m:.2f}," +
                        ",".join(f"{sum(p.get(k, 0) for p in ph[i:i + 40]) / m:.2f}" for k in keys) + "\n")


if __name__ == "__main__":
    main()
