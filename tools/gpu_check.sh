#!/bin/bash
# Round-trip used while tuning the optimiser kernel: parity tests, short bench, phase-cycle profile.
tag=${1:-x}
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -m gpu -x -q -k "not large_configs" > gpurun_out/t_$tag.log 2>&1; tail -3 gpurun_out/t_$tag.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-sweep --no-cpu-baseline > gpurun_out/b_$tag.json 2> gpurun_out/b_$tag.err; echo "bench exit $?"
timeout 200 python tools/profile_lbfgsb.py 3 > gpurun_out/p_${tag}_n123.json 2> gpurun_out/p_$tag.err
python - <<EOF
import json
d=json.load(open("gpurun_out/b_$tag.json")); print(d["value"], d["ms_per_step"], d["run_info"]["window_seconds"], d["roofline"]["avg_launch_us"])
d=json.load(open("gpurun_out/p_${tag}_n123.json")); w=d["worst_chunk"]; print(d["kernel_ms"], w["nit"],w["nfev"],w["Mcycles"]); print({k:round(v) for k,v in w["cycles_per_iteration"].items() if v>0})
EOF
if [ -n "$2" ]; then
timeout 200 python tools/profile_lbfgsb.py $2 > gpurun_out/p_${tag}_late.json 2>/dev/null
python - <<EOF
import json
d=json.load(open("gpurun_out/p_${tag}_late.json")); w=d["worst_chunk"]; print(d["kernel_ms"], d["n"], w["nit"],w["nfev"],w["Mcycles"]); print({k:round(v) for k,v in w["cycles_per_iteration"].items() if v>0})
EOF
fi
