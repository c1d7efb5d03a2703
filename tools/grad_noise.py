"""Per-tensor gradient error of the bf16 training step vs fp32 autograd on the oracle, at several batch sizes (calibrates the
gates of tests/test_gpu_train.py)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import forward_oracle as fo  # noqa: E402
from transformertts_b200.model.models import ForwardTransformer  # noqa: E402
from transformertts_b200.model.training import Adam  # noqa: E402


EMU = '--emulate' in sys.argv


def main():
    torch.set_num_threads(16)
    print('oracle:', 'bf16-operand emulation' if EMU else 'fp32')
    for cfg_name, B, Tp, Tm in (('C1', 3, 24, 150), ('C1', 16, 48, 400), ('LJ256', 2, 32, 260), ('LJ256', 8, 48, 400)):
        cfg = fo.CONFIGS[cfg_name]
        p = fo.init_params(cfg, seed=7)
        tok, dur, pit = fo.make_inputs('ragged', B, Tp, Tm, seed=301)
        mel = fo.make_mel_targets(dur, 80, seed=302)
        ref_out, ref_g = fo.loss_and_grads(p, cfg, tok, mel, dur, pit, emulate_bf16=EMU)
        m = ForwardTransformer(**cfg, train_dropout=False)
        m.set_weights(p)
        m._compile(Adam(1e-4))
        eng = m._get_engine()
        out = eng.forward_backward(tok, mel, dur, pit, training=True)
        torch.cuda.synchronize()
        gscale = max(float(g.norm()) for g in ref_g.values())
        rows = []
        for name, gref in ref_g.items():
            got = eng.g[name].detach().double().cpu()
            gr = gref.double()
            if float(gr.norm()) < 1e-6 * gscale:
                rows.append((name, 'zero', float(got.norm()) / gscale, 1.0))
                continue
            rel = float((got - gr).norm() / gr.norm())
            cos = float((got * gr).sum() / (got.norm() * gr.norm())) if gr.dim() else (1.0 if float(got) * float(gr) > 0 else -1.0)
            rows.append((name, tuple(gr.shape), rel, cos))
        rows.sort(key=lambda r: -r[2] if r[1] != 'zero' else 0)
        print(f'== {cfg_name} B={B} Tp={Tp} Tm={Tm}: loss {float(out["loss"]):.5f} vs {float(ref_out["loss"]):.5f}')
        for r in rows[:14]:
            print('   %-28s %-18s rel %.4f cos %.5f' % (r[0], r[1], r[2], r[3]))
        small = [r for r in rows if r[1] != 'zero' and (r[0].endswith('.b') or 'ln' in r[0] or 'pos_scalar' in r[0] or r[0].endswith('beta') or r[0].endswith('gamma'))]
        small.sort(key=lambda r: -r[2])
        print('   worst small tensors:', [(r[0], round(r[2], 4)) for r in small[:8]])
        zeros = [r for r in rows if r[1] == 'zero']
        print('   analytically-zero tensors: max |g|/gscale', max([r[2] for r in zeros] + [0]))


if __name__ == '__main__':
    main()
