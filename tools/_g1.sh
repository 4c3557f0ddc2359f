set -x
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "logei_raw" > gpurun_out/t_raw.log 2>&1; tail -5 gpurun_out/t_raw.log
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_teacher_forced.py -m gpu -x -q > gpurun_out/t_eng.log 2>&1; tail -5 gpurun_out/t_eng.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-sweep --no-cpu-baseline > gpurun_out/b_raw.json 2> gpurun_out/b_raw.err; echo "bench exit $?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-sweep --no-cpu-baseline --phases --sub-batches 1 > gpurun_out/b_raw_ph.json 2> gpurun_out/b_raw_ph.err; echo "bench exit $?"
tail -30 gpurun_out/b_raw_ph.err
timeout 200 python tools/profile_lbfgsb.py 3 > gpurun_out/p_raw_n123.json 2> gpurun_out/p_raw.err
python - <<EOF
import json
d=json.load(open("gpurun_out/b_raw.json")); print(d["value"], d["ms_per_step"], d["run_info"]["window_seconds"], d["roofline"]["avg_launch_us"])
d=json.load(open("gpurun_out/p_raw_n123.json")); print(d["kernel_ms"], d["nit_hist_25"])
EOF
