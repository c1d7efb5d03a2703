"""Debug aid: per-tile clock64 timeline of one attention CTA (softmax warp 0 and the TMA/MMA thread).

    python tools/att_trace.py build     # here: nvcc -DTTSB_ATT_TRACE -> transformertts_b200/libttsb_trace.so
    python tools/att_trace.py run       # on the GPU box: one launch, prints the per-tile event deltas in clocks

The product library never carries this instrumentation (the macro is off in transformertts_b200.build).
"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / 'transformertts_b200' / 'libttsb_trace.so'


def build():
    sys.path.insert(0, str(ROOT))
    from transformertts_b200 import build as b
    objs = []
    tmp = b.OBJ_DIR / 'trace'
    tmp.mkdir(parents=True, exist_ok=True)
    for src in b.SOURCES:
        obj = tmp / (src + '.o')
        cmd = [b._nvcc(), *b.NVCC_FLAGS, '-DTTSB_ATT_TRACE', '-DTTSB_GEMM_TRACE', '-I', str(b.INCLUDE), '-c', str(b.CSRC / src), '-o', str(obj)]
        subprocess.run(cmd, check=True)
        objs.append(str(obj))
    subprocess.run([b._nvcc(), '-shared', '-o', str(LIB), *objs, '-gencode', 'arch=compute_100a,code=sm_100a',
                    '-Xcompiler', '-fPIC'], check=True)
    print('built', LIB)


def run():
    import torch
    trace = torch.zeros(2 * 64 * 8, dtype=torch.int64, device='cuda:0')
    os.environ['TTSB_ATT_TRACE_PTR'] = hex(trace.data_ptr())
    os.environ['TTSB_LIB'] = str(LIB)
    sys.path.insert(0, str(ROOT))
    sys.argv = ['kbench', 'mha'] + sys.argv[2:]
    sys.path.insert(0, str(ROOT / 'tools'))
    import kbench
    kbench.main()
    torch.cuda.synchronize()
    t = trace.cpu().view(2, 64, 8)
    base = int(t[1, 0, 0])
    names = {0: ['loop_top', 's_ready', 's_loaded', 'exp_done', 'pv_prev_done', 'rescaled', 'p_arrived'],
             1: ['loop_top', 's_next_issued', 'k_loaded', 'v_ready', 'p_ready', 'pv_issued', 'v_next_loaded']}
    for role in (0, 1):
        print('role', 'softmax warp 0' if role == 0 else 'MMA thread', names[role])
        for j in range(64):
            if int(t[role, j, 0]) == 0:
                break
            row = [int(t[role, j, k]) - base for k in range(7)]
            print(f'  tile {j:2d}: ' + ' '.join(f'{v:7d}' for v in row))


if __name__ == '__main__':
    {'build': build, 'run': run}[sys.argv[1]]()
