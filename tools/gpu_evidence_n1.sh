set -x
mkdir -p gpurun_out
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
python bench.py --mode stft --steps 30 > $O/r02_stft.json 2> $O/r02_stft.err
TTSB_STFT_V1=1 python bench.py --mode stft --steps 30 > $O/r02_stft_v1.json 2> $O/r02_stft_v1.err
python bench.py --mode expand --steps 30 > $O/r02_expand.json 2> $O/r02_expand.err
python bench.py --mode aligner --steps 20 --warmup 5 > $O/r02_aligner.json 2> $O/r02_aligner.err
python bench.py --steps 20 --warmup 5 --no-graphs --no-train --no-cpu-baseline > $O/r02_bench_n1_eager.json 2> $O/r02_bench_n1_eager.err
# launch lists (eager launches, serialised by ncu: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_infer.csv python bench.py --steps 1 --warmup 3 --no-graphs --no-train --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/r02_launches_train.csv python bench.py --mode train --steps 1 --warmup 3 --no-graphs > /dev/null 2>&1
# full captures
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 60 -c 12 -f -o $O/r02_gemm python bench.py --steps 1 --warmup 3 --no-graphs --no-train --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mha_tc_kernel -s 8 -c 1 -f -o $O/r02_mha python bench.py --steps 1 --warmup 3 --no-graphs --no-train --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft_mel_v2_kernel -s 2 -c 1 -f -o $O/r02_stft_v2 python bench.py --mode stft --steps 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"layernorm_bwd_vec_kernel" -s 20 -c 2 -f -o $O/r02_rowk python bench.py --mode train --steps 1 --warmup 3 --no-graphs > /dev/null 2>&1
# attention-probability kernels of the training step at the C3 decoder shape (tools/probs_bench.py): forward (mode 0) and dS (mode 1)
python tools/probs_bench.py > $O/r02_probs_bench.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_probs --launch-skip 2 -c 1 -f -o $O/r02_probs python tools/probs_bench.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_probs --launch-skip 26 -c 1 -f -o $O/r02_ds16 python tools/probs_bench.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"expand_ln_pe_kernel" -s 1 -c 1 -f -o $O/r02_expand_ln_pe python bench.py --steps 1 --warmup 3 --no-graphs --no-train --no-cpu-baseline > /dev/null 2>&1
ls -la $O/r02_*
