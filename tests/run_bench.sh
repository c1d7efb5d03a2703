#!/bin/bash
# On the GPU box: bench (both arms), ncu launch lists and full captures of the dominant kernels.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
python bench.py --mode stft > gpurun_out/bench_stft.json 2> gpurun_out/bench_stft.err
python bench.py --mode expand > gpurun_out/bench_expand.json 2> gpurun_out/bench_expand.err
python bench.py --mode aligner > gpurun_out/bench_aligner.json 2> gpurun_out/bench_aligner.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1200 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/ncu_launches_train.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 28 -c 6 -o gpurun_out/prof_gemm -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mha_tc_kernel -s 8 -c 2 -o gpurun_out/prof_mha -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mha.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bgemm_tc_kernel -s 60 -c 12 -o gpurun_out/prof_bgemm -f \
    python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/ncu_bgemm.log 2>&1
cat gpurun_out/bench_ours.json
