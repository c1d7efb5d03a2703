set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_dp.py "tests/test_gpu_train.py::test_train_tts_from_disk_dataset_data_parallel" -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2e_dp_tests.log
tail -8 gpurun_out/r2e_dp_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode train --steps 20 --warmup 5 > gpurun_out/r2e_train_n2.json 2> gpurun_out/r2e_train_n2.err
TTSB_DP_BACKEND=torch python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode train --steps 20 --warmup 5 > gpurun_out/r2e_train_n2_torch.json 2> gpurun_out/r2e_train_n2_torch.err
python bench.py --mode train --steps 20 --warmup 5 > gpurun_out/r2e_train_n1.json 2> gpurun_out/r2e_train_n1.err
tail -3 gpurun_out/r2e_train_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/r2e_train_n1.json','gpurun_out/r2e_train_n2.json','gpurun_out/r2e_train_n2_torch.json'):
    try:
        t=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'steps/s', t['value'], 'ms', t['ms_per_step'], 'e2e', t['e2e']['value'], 'noar', t['ms_per_step_without_allreduce'], 'exposed', t['nccl_exposed_ms'], t['loss'])
    except Exception as e: print(f, e)
PY
