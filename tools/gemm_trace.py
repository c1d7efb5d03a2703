"""Debug aid: per-tile clock64 timeline of CTA 0 of one ttsb_linear_fwd launch (MMA-issue warp and epilogue warp 2).

    python tools/att_trace.py build                      # builds transformertts_b200/libttsb_trace.so with both trace macros
    python tools/gemm_trace.py [--K 256 --N 768 --relu 0 --fp16 1]   # on the GPU box

Prints, per tile of CTA 0: MMA warp [loop top, accumulator free, first operands landed, all MMAs issued] and epilogue warp
[wait start, accumulator complete, epilogue done], in clocks relative to the first stamp.
"""
import argparse
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / 'transformertts_b200' / 'libttsb_trace.so'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=64)
    ap.add_argument('--T', type=int, default=1000)
    ap.add_argument('--K', type=int, default=256)
    ap.add_argument('--N', type=int, default=768)
    ap.add_argument('--segs', type=int, default=1)
    ap.add_argument('--relu', type=int, default=0)
    ap.add_argument('--fp16', type=int, default=1)
    a = ap.parse_args()
    import torch
    trace = torch.zeros(3 * 64 * 4, dtype=torch.int64, device='cuda:0')
    os.environ['TTSB_GEMM_TRACE_PTR'] = hex(trace.data_ptr())
    os.environ['TTSB_LIB'] = str(LIB)
    sys.path.insert(0, str(ROOT))
    from transformertts_b200 import lib
    from transformertts_b200.model.models import _PackedLinear
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(0)
    x = torch.randn(a.B, a.T, a.K, generator=g).to(dev)
    x_hi, x_lo = lib.split_bf16(x, True)
    w = torch.randn(a.segs * a.K, a.N, generator=g).to(dev) * 0.05
    pl = _PackedLinear(w, torch.zeros(a.N, device=dev), [a.K] * a.segs, True, block_n=256 if a.N % 256 == 0 else None)
    out_hi = torch.empty(a.B, a.T, pl.n_pad, dtype=torch.float16 if a.fp16 else torch.bfloat16, device=dev)
    out_lo = None if a.fp16 else torch.empty_like(out_hi)
    ga = lib.GemmArgs()
    ga.B, ga.T, ga.N, ga.block_n = a.B, a.T, pl.N, pl.block_n
    ga.num_segments = a.segs
    shifts = [0] if a.segs == 1 else [-1, 0, 1]
    for s in range(a.segs):
        ga.seg_src[s], ga.seg_shift[s], ga.seg_k[s] = 0, shifts[s], a.K
    ga.a_hi[0], ga.a_lo[0], ga.lda[0], ga.a_col0[0] = x_hi.data_ptr(), x_lo.data_ptr(), a.K, 0
    ga.w_hi, ga.w_lo, ga.bias = pl.w_hi.data_ptr(), pl.w_lo.data_ptr(), pl.bias.data_ptr()
    ga.relu = a.relu
    ga.out_hi = out_hi.data_ptr()
    ga.out_lo = out_lo.data_ptr() if out_lo is not None else None
    ga.ld_out = pl.n_pad
    ga.out_fp16 = a.fp16
    ga.precision, ga.impl = lib.PREC_BF16X3, lib.IMPL_TCGEN05
    for _ in range(3):
        lib.linear_fwd(ga)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.linear_fwd(ga)
    e1.record()
    torch.cuda.synchronize()
    print(f'K={a.segs}x{a.K} N={a.N} M={a.B * a.T}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch')
    t = trace.cpu().view(3, 64, 4)
    base = int(t[1, 0, 0])
    print('MMA warp      [loop_top, acc_free, operands_landed, mmas_issued]   | epilogue warp 2 [wait_start, acc_complete, epilogue_done]')
    for i in range(16):
        if int(t[1, i, 0]) == 0:
            break
        m = [int(t[1, i, k]) - base for k in range(4)]
        e = [int(t[0, i, k]) - base for k in range(3)]
        print(f'  tile {i:2d}: ' + ' '.join(f'{v:7d}' for v in m) + '   | ' + ' '.join(f'{v:7d}' for v in e))


    print('epilogue warp 2, last tile, per group of 4 chunks: [group start, TMEM loads landed, chunks stored] (relative to group 0 start)')
    b2 = int(t[2, 0, 0])
    for gidx in range(4):
        if int(t[2, gidx, 0]) == 0:
            break
        print('   group', gidx, [int(t[2, gidx, k]) - b2 for k in range(3)])


if __name__ == '__main__':
    main()
