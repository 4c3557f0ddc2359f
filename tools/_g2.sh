for S in 3 9 18; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-sweep --no-cpu-baseline --sub-batches $S > gpurun_out/b_S$S.json 2> gpurun_out/b_S$S.err; echo "bench S=$S exit $?"
done
timeout 300 python -m cProfile -o gpurun_out/prof_S9.out bench.py --steps 20 --warmup 5 --no-sweep --no-cpu-baseline --sub-batches 9 > gpurun_out/b_S9p.json 2> gpurun_out/b_S9p.err
python - <<EOF
import json
for S in (3,9,18):
    d=json.load(open("gpurun_out/b_S%d.json"%S)); print(S, d["value"], d["ms_per_step"], d["run_info"]["window_seconds"], d["roofline"]["avg_launch_us"], d["run_info"]["host_blocked_ms_per_step"])
EOF
