set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2b_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python tools/step_cpu_time.py train 2>&1 | head -3 > gpurun_out/r2b_cpu_train.txt
tail -8 gpurun_out/r2b_tests.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
print('infer ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['avg_launch_ms'])
t=d['train']; print('train', t['value'], t['ms_per_step'], t['e2e']['value'], t['loss'])
PY
cat gpurun_out/r2b_cpu_train.txt
