set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/test_final.txt
python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err
python bench.py --impl reference > gpurun_out/bench_r01_reference.json 2> gpurun_out/bench_r01_reference.err
python tools/measure_fp64.py gpurun_out/fp64_peaks2.json > /dev/null 2> gpurun_out/fp64_2.err
python tools/bench_sweep.py --log2n 20 22 24 26 --dims 8 16 32 40 > gpurun_out/sweep_final.jsonl 2> gpurun_out/sweep_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --sub-batches 1 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lbfgsb_logei --launch-skip 2 --launch-count 1 -f -o gpurun_out/prof_lbfgsb_r01_final python bench.py --steps 3 --warmup 1 --no-cpu-baseline --sub-batches 1 > gpurun_out/ncu_lb_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ei_sweep --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_sweep_r01_final python tools/bench_sweep.py --log2n 22 --dims 40 --reps 1 > gpurun_out/ncu_sw_final.log 2>&1
cat gpurun_out/test_final.txt
tail -c 600 gpurun_out/bench_r01_final.json
tail -c 400 gpurun_out/bench_r01_reference.json
