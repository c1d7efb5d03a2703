"""Is a step host-launch bound?  Issues N steps back to back and reports the wall time until the python loop has ISSUED
them (queue depth permitting) next to the time until the GPU has finished them.

    python tools/step_cpu_time.py [train|infer|aligner]
"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import forward_oracle as fo  # noqa: E402  (input generators only)
from transformertts_b200.model.models import ForwardTransformer  # noqa: E402
from transformertts_b200.model.training import Adam  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else 'train'
    cfg = fo.CONFIGS['LJ256']
    p = fo.init_params(cfg, seed=7)
    dev = 'cuda:0'
    if mode == 'train':
        m = ForwardTransformer(**cfg, train_dropout=True)
        m.set_weights(p)
        m._compile(Adam(1e-4))
        tok, dur, pit = fo.make_inputs('full', 32, 128, 1000, seed=300)
        mel = fo.make_mel_targets(dur, 80, seed=400)
        a = [t.to(dev) for t in (tok, mel, dur, pit)]
        step = lambda: m.train_step(*a)  # noqa: E731
    else:
        m = ForwardTransformer(**cfg, cuda_graphs=(mode == 'infer-graph'))
        m.set_weights(p)
        tok, dur, pit = fo.make_inputs('full', 64, 128, 1000, seed=200)
        a = [tok.to(dev), dur.to(dev).float(), pit.to(dev)]
        step = lambda: m.call(a[0], target_durations=a[1], target_pitch=a[2])  # noqa: E731
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f'{mode}: issue {t_issue / n * 1e3:.2f} ms/step, complete {t_all / n * 1e3:.2f} ms/step '
          f'({"HOST-bound" if t_issue > 0.9 * t_all else "GPU-bound"})')
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        step()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(25)


if __name__ == '__main__':
    main()
