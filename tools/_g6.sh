for S in 8 9 10 12; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-sweep --no-cpu-baseline --sub-batches $S > gpurun_out/b3_S$S.json 2> gpurun_out/b3_S$S.err; echo "bench S=$S exit $?"; tail -3 gpurun_out/b3_S$S.err
done
python - <<EOF
import json
for S in (8,9,10,12):
    try:
        d=json.load(open("gpurun_out/b3_S%d.json"%S)); print(S, d["value"], d["ms_per_step"], d["run_info"]["window_seconds"], d["roofline"]["avg_launch_us"], d["run_info"]["host_blocked_ms_per_step"])
    except Exception as e: print(S, "failed", e)
EOF
