mkdir -p gpurun_out
python -m pytest tests/test_gpu_aligner.py -m gpu -q -x 2>&1 | tail -5
python bench.py --mode aligner --steps 20 --warmup 5 > gpurun_out/r2l_aligner.json 2> gpurun_out/r2l_aligner.err
python bench.py --mode aligner --steps 20 --warmup 5 --no-graphs --no-cpu-baseline > gpurun_out/r2l_aligner_eager.json 2> gpurun_out/r2l_aligner_eager.err
python - <<'PY'
import json
for f in ('gpurun_out/r2l_aligner.json','gpurun_out/r2l_aligner_eager.json'):
    try:
        t=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'fwd ms', t['ms_per_step'], 'steps/s', t['value'], 'e2e', t['e2e']['value'], 'train', t['train_step'])
    except Exception as e: print(f, e, open(f.replace('.json','.err')).read()[-800:])
PY
