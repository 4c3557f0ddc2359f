timeout 300 python -m cProfile -o gpurun_out/prof2_S9.out bench.py --steps 20 --warmup 5 --no-sweep --no-cpu-baseline --sub-batches 9 > gpurun_out/b_S9p2.json 2> gpurun_out/b_S9p2.err
nproc; python -c "import os;print(os.cpu_count())"
