timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py tests/test_gpu_teacher_forced.py -m gpu -x -q -k "lbfgsb or device or engine or teacher or forced or pipelined" > gpurun_out/t_h2.log 2>&1; tail -4 gpurun_out/t_h2.log
timeout 200 python tools/profile_lbfgsb.py 3 > gpurun_out/p_h2_n123.json 2> gpurun_out/p_h2.err
python - <<EOF
import json
for f in ("gpurun_out/p_h2_n123.json",):
    d=json.load(open(f)); w=d["worst_chunk"]; print(d["kernel_ms"], d["n"], w["nit"],w["nfev"],w["Mcycles"]); print({k:round(v) for k,v in w["cycles_per_iteration"].items() if v>0})
EOF
timeout 300 python bench.py --steps 20 --warmup 5 --no-sweep --no-cpu-baseline > gpurun_out/b_h2.json 2> gpurun_out/b_h2.err; python -c "
import json; d=json.load(open('gpurun_out/b_h2.json')); print(d['value'], d['run_info']['window_seconds'], d['roofline']['avg_launch_us'])"
