"""Micro-benchmark of the attention-probability kernels of the training step at the C3 decoder shape: the fused kernel
(ttsb_attn_probs_fwd) against the materialised path (ttsb_bgemm fp32 logits + ttsb_softmax_fwd).  CUDA events, L2 flushed
between iterations by the 258 MB outputs themselves."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from transformertts_b200 import lib  # noqa: E402
from transformertts_b200.model.training import TrainEngine  # noqa: E402


def main():
    B, H, T, dh, rate = 32, 2, 1000, 128, 0.1
    if len(sys.argv) > 1:
        B, H, T, dh = [int(v) for v in sys.argv[1:5]]
    dev = torch.device('cuda:0')
    lib.load()
    d = H * dh
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(B, T, 3 * d, generator=g).bfloat16().to(dev)
    lens = torch.randint(int(0.6 * T), T + 1, (B,), generator=g, dtype=torch.int32)
    lens[0] = T
    lens = lens.to(dev)
    Z, ldp = B * H, (T + 15) // 16 * 16
    P, D = torch.empty(Z, T, ldp, dtype=torch.bfloat16, device=dev), torch.empty(Z, T, ldp, dtype=torch.bfloat16, device=dev)
    S = torch.empty(Z, T, ldp, device=dev)
    eng = TrainEngine.__new__(TrainEngine)
    eng.dev = dev

    def fused():
        lib.attn_probs_fwd(qkv, 3 * d, 0, d, B, H, T, dh, lens, 1.0 / math.sqrt(dh), rate, 7, 3, P, D, ldp)

    def two():
        eng._bgemm(B, H, T, T, dh, qkv, (3 * d, T, B), (3 * d, 3 * d * T), (dh, 0, 0, 0), qkv, (2 * d, T, B), (3 * d, 3 * d * T),
                   (dh, 0, 0, d), alpha=1.0 / math.sqrt(dh), out_f32=S, ld_out=ldp, out_batch_stride=T * ldp, out_cols=ldp)
        lib.softmax_fwd(S, B, H, T, T, ldp, lens, rate, 7, 3, P, D)

    dO = torch.randn(B, T, d, generator=g).bfloat16().to(dev)
    Dv = (torch.randn(Z * T, generator=g) * 0.1).to(dev)
    dS = torch.empty(Z, T, ldp, dtype=torch.bfloat16, device=dev)
    fused()
    scale = 1.0 / math.sqrt(dh)

    def ds_fused():
        lib.attn_ds_bwd(dO, d, 0, qkv, 3 * d, 2 * d, B, H, T, dh, lens, P, Dv, scale, rate, 7, 3, dS, ldp)

    def ds_bgemm():
        eng._bgemm(B, H, T, T, dh, dO, (d, T, B), (d, d * T), (dh, 0, 0, 0), qkv, (d, T, B), (3 * d, 3 * d * T), (dh, 0, 0, 2 * d),
                   out_bf16=dS, ld_out=ldp, out_batch_stride=T * ldp, out_cols=ldp, softmax_bwd=(P, Dv, scale, rate, 7, 3, 0, lens, None))

    for name, fn in (('fused', fused), ('two-kernel', two), ('dS fused', ds_fused), ('dS bgemm epilogue', ds_bgemm)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        print(f'{name}: {us:.1f} us  ({2 * Z * T * ldp * 2 / us / 1e3:.0f} GB/s of bf16 output)')


if __name__ == '__main__':
    main()
