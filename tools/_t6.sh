mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/t6_tests.log 2>&1; tail -6 gpurun_out/t6_tests.log | cut -c1-300
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/t6_bench.json 2> gpurun_out/t6_bench.err; tail -2 gpurun_out/t6_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t6_bench.json').read().strip().splitlines()[-1]); t=d['train']
print('infer', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], '| train', t['value'], t['ms_per_step'], t['e2e']['value'], 'launches', d['gpu_launches'])
PY
