set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:bgemm_tc_kernel -s 28 -c 1 -o gpurun_out/r2f_ds python bench.py --mode train --steps 1 --warmup 3 --no-graphs > /dev/null 2> gpurun_out/r2f_ds.err
ncu --set full --clock-control none --import-source on -k regex:stft_mel_kernel -s 2 -c 1 -o gpurun_out/r2f_stft python bench.py --mode stft --steps 3 > /dev/null 2> gpurun_out/r2f_stft.err
ncu --set full --clock-control none --import-source on -k regex:"length_regulate_kernel|expand_indices_kernel|durations_to_int_kernel" -s 3 -c 3 -o gpurun_out/r2f_expand python bench.py --mode expand --steps 3 > /dev/null 2> gpurun_out/r2f_expand.err
ls -la gpurun_out/*.ncu-rep
