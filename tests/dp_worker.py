"""Worker of tests/test_gpu_dp.py (launched with torchrun, 2 ranks, NCCL): one data-parallel training step on two 16-row
shards vs. the same step on the whole 32-row batch in one process."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import forward_oracle as fo  # noqa: E402  (seeded weights / inputs only)
from transformertts_b200.model.models import ForwardTransformer  # noqa: E402
from transformertts_b200.model.training import Adam  # noqa: E402
from transformertts_b200.utils.data_parallel import global_loss  # noqa: E402


def main():
    out_path = sys.argv[1]
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    cfg = fo.CONFIGS['LJ256']
    p = fo.init_params(cfg, seed=7)
    B = 32
    tok, dur, pit = fo.make_inputs('ragged', B, 48, 320, seed=901)
    mel = fo.make_mel_targets(dur, 80, seed=902)

    def fresh():
        m = ForwardTransformer(**cfg, device=str(dev), train_dropout=False)
        m.set_weights(p)
        m._compile(Adam(1e-4))
        return m

    # ---- data parallel: rank r takes rows r, r+W, ... (what data/datasets.py hands out), padded to the GLOBAL lengths
    m = fresh()
    sl = slice(rank, None, world)
    o = m.train_step(tok[sl], mel[sl], dur[sl], pit[sl], data_parallel=True)
    eng = m._get_engine()
    g_dp = eng.flat_g.clone() / world                     # the summed buffer; 1/N is applied inside the Adam kernel
    w_dp = eng.flat_w.clone()
    loss_dp = global_loss(o['loss'], 1)                   # equal-shape shards: plain mean over ranks
    # both ranks must hold identical weights after the step
    w_all = [torch.empty_like(w_dp) for _ in range(world)]
    dist.all_gather(w_all, w_dp)
    same_across_ranks = all(torch.equal(w_all[0], w) for w in w_all)
    if rank == 0:
        s = fresh()
        o1 = s.train_step(tok, mel, dur, pit)
        e1 = s._get_engine()
        # the same two shards stepped one after the other in THIS process: what the all-reduce must reproduce exactly
        g_seq = torch.zeros_like(g_dp)
        for r in range(world):
            t = fresh()
            sr = slice(r, None, world)
            t.train_step(tok[sr], mel[sr], dur[sr], pit[sr])
            g_seq += t._get_engine().flat_g / world
        torch.save({'g_dp': g_dp.cpu(), 'w_dp': w_dp.cpu(), 'loss_dp': float(loss_dp), 'g_single': e1.flat_g.cpu(), 'g_seq': g_seq.cpu(),
                    'w_single': e1.flat_w.cpu(), 'loss_single': float(o1['loss']), 'same_across_ranks': same_across_ranks,
                    'offsets': {n: (eng.offsets[n][0], p[n].numel()) for n in eng.names}}, out_path)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
