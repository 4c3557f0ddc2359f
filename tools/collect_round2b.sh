#!/bin/bash
# Round-2 (second half) collection: full GPU tests, smoke, whole-run and stratified bench, reference arm, ncu launch list
# and one full capture of the raw-sample scoring launch of the sweep kernel.
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r02b_gputest.log 2>&1; tail -3 gpurun_out/r02b_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; tail -2 gpurun_out/r02b_smoke.log
python bench.py > gpurun_out/r02b_bench_c3_full_run.json 2> gpurun_out/r02b_bench_full.err; tail -c 300 gpurun_out/r02b_bench_full.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_c3_steps20.json 2> gpurun_out/r02b_bench_steps20.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02b_bench_reference_steps20.json 2> gpurun_out/r02b_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02b_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep --sub-batches 1 > gpurun_out/r02b_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ei_sweep --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02b_prof_raw_scoring python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep --sub-batches 1 > gpurun_out/r02b_ncu_raw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lbfgsb_logei --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02b_prof_lbfgsb python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep --sub-batches 1 > gpurun_out/r02b_ncu_lb.log 2>&1
python tools/profile_lbfgsb.py 3 > gpurun_out/r02b_lbfgsb_phase_cycles_n123.json 2> /dev/null
python tools/profile_lbfgsb.py 310 > gpurun_out/r02b_lbfgsb_phase_cycles_n430.json 2> /dev/null
python - <<EOF
import json
for f in ("r02b_bench_c3_full_run", "r02b_bench_c3_steps20", "r02b_bench_reference_steps20"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f)); print(f, d["value"], d.get("ms_per_step"), (d.get("ei_sweep") or {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(f, "failed", e)
EOF
