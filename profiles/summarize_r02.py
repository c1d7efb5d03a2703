#!/usr/bin/env python
"""Round-2 evidence: turns the artifacts of `bash tools/_final_n1.sh` (run under gpurun; files in gpurun_out/) into the committed
summaries profiles/r02_summary.md and profiles/traffic.json.

    python profiles/summarize_r02.py
"""
import collections
import csv
import json
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / 'gpurun_out'

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def launches(fn):
    lines = [l for l in open(fn) if not l.startswith('==')]
    rows = []
    for r in csv.DictReader(lines):
        try:
            v = float(r['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        v = v / 1e3 if r['Metric Unit'] == 'ns' else v
        rows.append((re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('ttsb::', ''), v))
    return rows


def last_step(rows, marker, include_marker=True):
    idx = [i for i, (n, _) in enumerate(rows) if n.startswith(marker)]
    if len(idx) < 2:
        return rows
    return rows[idx[-2] + 1: idx[-1] + 1] if include_marker else rows[idx[-2]: idx[-1]]


def table(step, top=16):
    tot = sum(v for _, v in step)
    agg = collections.OrderedDict()
    for n, v in step:
        a = agg.setdefault(n[:64], [0, 0.0])
        a[0] += 1
        a[1] += v
    out = [f'{len(step)} launches, {tot:.0f} us summed kernel time (ncu serialises and cold-starts every launch: compare shares)\n',
           '| kernel | launches | total us | avg us | share |', '|---|---|---|---|---|']
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f'| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {t / tot * 100:.1f}% |')
    return '\n'.join(out)


def ncu_rows(rep):
    txt = subprocess.run(['ncu', '-i', str(rep), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        d = {'kernel': re.sub(r'\(.*', '', r[idx['Kernel Name']]).replace('void ', '').replace('ttsb::', '')[:48], 'grid': r[idx.get('Grid Size', 0)]}
        for k in KEYS:
            if k in idx:
                d[k] = (r[idx[k]], units[idx[k]])
        out.append(d)
    return out


def fmt(d, k, scale=1.0, nd=1):
    if k not in d:
        return '-'
    try:
        return f'{float(d[k][0].replace(",", "")) * scale:.{nd}f}'
    except ValueError:
        return d[k][0]


def bytes_of(d, k):
    v, u = d.get(k, ('0', 'byte'))
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)


def us_of(d):
    v, u = d.get('gpu__time_duration.sum', ('0', 'us'))
    v = float(v.replace(',', ''))
    return v * {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(u, 1)


def ncu_table(rep, title):
    rows = ncu_rows(rep)
    out = [f'### {title}  (`{Path(rep).name}`)', '',
           '| kernel | grid | us | DRAM rd MB | DRAM wr MB | DRAM GB/s | tensor % | fma-pipe % | issue % | warps % | regs | long-sb | short-sb |', '|' + '---|' * 13]
    for d in rows:
        t = us_of(d)
        rd, wr = bytes_of(d, 'dram__bytes_read.sum'), bytes_of(d, 'dram__bytes_write.sum')
        out.append(f"| `{d['kernel']}` | {d['grid']} | {t:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / t / 1e3 if t else 0:.0f} | "
                   f"{fmt(d, KEYS[3])} | {fmt(d, KEYS[4])} | {fmt(d, KEYS[5])} | {fmt(d, KEYS[6])} | {fmt(d, 'launch__registers_per_thread', nd=0)} | "
                   f"{fmt(d, KEYS[10], nd=2)} | {fmt(d, KEYS[11], nd=2)} |")
    return '\n'.join(out), rows


def jline(fn):
    try:
        return json.loads(open(fn).read().strip().splitlines()[-1])
    except Exception:
        return None


def main():
    md = ['# Round 2 profile summary', '',
          'Sources: `tools/gpu_evidence_n1.sh` under gpurun on one B200 (files named below live in gpurun_out/, which is scratch; this file and',
          '`traffic.json` are the committed digests).  Peaks: MEASURED_PEAKS.json (6585 GB/s HBM copy, 1415.6 TFLOP/s sustained bf16).', '']
    b = jline(OUT / 'r02_bench_n1.json')
    if b:
        t = b.get('train') or {}
        md += ['## bench.py lines (N = 1)', '',
               f"* inference C2 (CUDA graphs): **{b['ms_per_step']:.3f} ms/step, {b['value'] / 1e6:.2f} M frames/s**, e2e {b['e2e']['value'] / 1e6:.2f} M frames/s; "
               f"decoder conv GEMMs {b['roofline']['achieved']:.0f} TFLOP/s = {b['roofline']['frac']:.3f} of sustained bf16 peak (3 passes: cap 1/3); "
               f"whole model {b['model_tflops']:.0f} TFLOP/s; clocks {b['clocks']}",
               f"* training C3 (CUDA graphs): **{t.get('ms_per_step', 0):.3f} ms/step, {t.get('value', 0):.1f} steps/s**, e2e {t.get('e2e', {}).get('value', 0):.1f} steps/s, "
               f"frac {t.get('roofline', {}).get('frac', 0):.3f} of sustained bf16 peak (single pass)", '']
    e = jline(OUT / 'r02_bench_n1_eager.json')
    if e:
        md += [f"* inference C2, eager launches: {e['ms_per_step']:.3f} ms/step", '']
    for tag, fn in (('STFT->mel v2 (C4)', 'r02_stft.json'), ('STFT->mel v1 (round-1 kernel, TTSB_STFT_V1=1)', 'r02_stft_v1.json'), ('length regulator (C2-LR)', 'r02_expand.json')):
        j = jline(OUT / fn)
        if j:
            md += [f"* {tag}: {j['ms_per_step'] * 1e3:.1f} us, {j['roofline']['achieved']:.0f} GB/s = {j['roofline']['frac']:.3f} of HBM copy peak"]
    a = jline(OUT / 'r02_aligner.json')
    if a:
        md += [f"* Aligner C5: forward+losses {a['ms_per_step']:.2f} ms ({a['value']:.0f} steps/s), training step {a['train_step']['ms_per_step']:.2f} ms ({a['train_step']['value']:.0f} steps/s)"]
    md += ['']
    if (OUT / 'r02_launches_infer.csv').exists():
        md += ['## launch list, inference step (eager)', '', table(last_step(launches(OUT / 'r02_launches_infer.csv'), 'embed_ln_pe', False)), '']
    if (OUT / 'r02_launches_train.csv').exists():
        md += ['## launch list, training step (eager)', '', table(last_step(launches(OUT / 'r02_launches_train.csv'), 'adam_tf_kernel')), '']
    traffic = {}
    md += ['## ncu --set full captures', '']
    for rep, title in (('r02_gemm.ncu-rep', 'inference GEMMs of a decoder block (bf16x3)'), ('r02_mha.ncu-rep', 'fused attention (fp16, decoder)'),
                       ('r02_stft_v2.ncu-rep', 'STFT->mel v2'), ('r2f_stft.ncu-rep', 'STFT->mel v1 (round-1 kernel)'), ('r02_rowk.ncu-rep', 'training row kernels'),
                       ('r02_probs.ncu-rep', 'attention probabilities, training forward: QK^T + softmax + dropout in one kernel (C3 decoder shape, tools/probs_bench.py)'),
                       ('r02_ds16.ncu-rep', 'attention dS, training backward: dO V^T + softmax gradient in one kernel (same shape)'),
                       ('r02_ds.ncu-rep', 'the dS product as an epilogue of the generic batched GEMM (what the kernel above replaced)'), ('r2f_ds.ncu-rep', 'the same before its epilogue work'),
                       ('r02_expand_ln_pe.ncu-rep', 'length regulator gather + LayerNorm + PE (fused, real path)'),
                       ('r2f_expand.ncu-rep', 'length regulator kernels alone')):
        if (OUT / rep).exists():
            t, rows = ncu_table(OUT / rep, title)
            md += [t, '']
            if rep == 'r02_gemm.ncu-rep':
                big = [d for d in rows if us_of(d) > 120]
                if big:
                    tr = [bytes_of(d, 'dram__bytes_read.sum') + bytes_of(d, 'dram__bytes_write.sum') for d in big]
                    traffic = {'source': 'profiles/r02_summary.md (ncu --set full, decoder conv GEMMs)', 'conv_gemm_bytes': tr,
                               'conv_gemm_mean_bytes_per_launch': sum(tr) / len(tr)}
    (ROOT / 'profiles' / 'r02_summary.md').write_text('\n'.join(md) + '\n')
    # traffic.json is maintained by hand from the table above (conv2 = its single-CTA launch + its CTA-pair tail launch)
    print('wrote profiles/r02_summary.md; conv GEMM DRAM bytes of this capture:', traffic.get('conv_gemm_bytes'))


if __name__ == '__main__':
    main()
