set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2a_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
python bench.py --steps 20 --warmup 5 --no-graphs --no-train --no-cpu-baseline > gpurun_out/r2a_bench_eager.json 2> gpurun_out/r2a_bench_eager.err
python tools/step_cpu_time.py train > gpurun_out/r2a_cpu_train.txt 2>&1
python tools/step_cpu_time.py infer > gpurun_out/r2a_cpu_infer.txt 2>&1
tail -5 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_bench.json | head -c 3000; head -3 gpurun_out/r2a_cpu_train.txt; head -3 gpurun_out/r2a_cpu_infer.txt
