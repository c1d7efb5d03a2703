#!/usr/bin/env python
"""Turn the artifacts a `bash tests/run_bench.sh` gpurun call leaves in gpurun_out/ into the committed summaries
profiles/<tag>_*.md (launch-list shares per step, ncu --set full key metrics per kernel, bench JSON lines).

    python profiles/summarize.py r01c
"""
import collections
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / 'gpurun_out'


def load_launches(fn):
    with open(fn) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = []
    for r in csv.DictReader(lines):
        v = float(r['Metric Value'])
        v = v / 1e3 if r['Metric Unit'] == 'ns' else v
        rows.append((re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('ttsb::', ''), v))
    return rows


def one_step(rows):
    idx = [i for i, (n, _) in enumerate(rows) if n.startswith('embed_ln_pe')]
    return rows[idx[0]:idx[1]] if len(idx) >= 2 else rows


def table(step):
    tot = sum(v for _, v in step)
    agg = collections.OrderedDict()
    for n, v in step:
        a = agg.setdefault(n[:60], [0, 0.0])
        a[0] += 1
        a[1] += v
    out = [f'{len(step)} launches, {tot:.0f} us summed kernel time (ncu serialises and cold-starts every launch: compare shares)\n',
           '| kernel | launches | total us | avg us | share |', '|---|---|---|---|---|']
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        out.append(f'| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {t / tot * 100:.1f}% |')
    return '\n'.join(out)


def ncu_raw(rep, wanted):
    csv_txt = subprocess.run(['ncu', '-i', str(rep), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(csv_txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        out.append({w: (r[idx[w]], units[idx[w]]) for w in wanted if w in idx})
    return out


METRICS = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
           'launch__shared_mem_per_block_dynamic', 'launch__grid_size']


def ncu_table(rep, labels):
    rows = ncu_raw(rep, METRICS)
    head = '| launch | ' + ' | '.join(m.split('.')[0].replace('__', ' ') for m in METRICS) + ' |'
    out = [head, '|' + '---|' * (len(METRICS) + 1)]
    for i, r in enumerate(rows):
        lab = labels[i] if i < len(labels) else f'#{i}'
        out.append(f'| {lab} | ' + ' | '.join(f'{r[m][0]} {r[m][1]}' if m in r else '-' for m in METRICS) + ' |')
    return '\n'.join(out)


def main():
    tag = sys.argv[1]
    md = [f'# {tag} -- bench lines, launch lists and ncu captures (B200, `bash tests/run_bench.sh`)\n']
    for f in ('bench_ours.json', 'bench_ref.json', 'bench_train.json', 'bench_stft.json', 'bench_expand.json', 'bench_aligner.json'):
        p = OUT / f
        if p.exists():
            md.append(f'## {f}\n```json\n{p.read_text().strip().splitlines()[-1]}\n```\n')
    if (OUT / 'launches.csv').exists():
        md.append('## inference step (C2, bf16x3 GEMMs + fp16 attention) -- `ncu --metrics gpu__time_duration.sum --clock-control none`\n')
        step = one_step(load_launches(OUT / 'launches.csv'))
        md.append(table(step) + '\n')
        dec = [(n, v) for n, v in step if 'gemm_tc' in n or 'mha_tc' in n]
        # one decoder block = the launches from the second-to-last attention kernel up to the last one; a LayerNorm GEMM
        # with a pair-mode tail is two launches (gemm_tc_kernel<1,0> then <1,1>)
        att = [i for i, (n, _) in enumerate(dec) if 'mha_tc' in n]
        if len(att) >= 3:
            blk = dec[att[-2] - 1:att[-1] - 1]
            md.append('One decoder block (M = 64 x 1000 rows), launch order (QKV GEMM, attention, concat-projection + LN '
                      '[single-CTA waves, pair-mode tail], conv1 + ReLU, conv2 + LN [single-CTA waves, pair-mode tail]):\n\n| launch | us |\n|---|---|')
            for n, v in blk:
                md.append(f'| `{n[:40]}` | {v:.1f} |')
            md.append(f'| block total | {sum(v for _, v in blk):.1f} |')
            md.append('')
    if (OUT / 'launches_train.csv').exists():
        md.append('## training step (C3, bf16, B=32, dropout 0.1) -- launch list\n')
        md.append(table(one_step(load_launches(OUT / 'launches_train.csv'))) + '\n')
    # DRAM traffic of the roofline kernels (decoder conv GEMMs) for bench.py's roofline.traffic: bytes per logical GEMM
    if (OUT / 'prof_gemm.ncu-rep').exists():
        rows = ncu_raw(OUT / 'prof_gemm.ncu-rep', ['dram__bytes_read.sum', 'dram__bytes_write.sum'])

        def to_bytes(cell):
            v, u = float(cell[0]), cell[1].lower()
            return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)

        if len(rows) >= 6:
            tot = [to_bytes(r['dram__bytes_read.sum']) + to_bytes(r['dram__bytes_write.sum']) for r in rows[:6]]
            traffic = {'source': f'profiles/{tag}_summary.md (ncu --set full, decoder block 0)', 'conv1_bytes': tot[3],
                       'conv2_bytes': tot[4] + tot[5], 'conv_gemm_mean_bytes_per_launch': (tot[3] + tot[4] + tot[5]) / 2}
            (ROOT / 'profiles' / 'traffic.json').write_text(json.dumps(traffic, indent=1))
    for rep, title, labels in (('prof_gemm.ncu-rep', 'gemm_tc_kernel<split=true>, decoder block 0', ['QKV', 'concat-proj + LN (single-CTA waves)', 'concat-proj + LN (pair tail)', 'conv1 + relu', 'conv2 + LN (single-CTA waves)', 'conv2 + LN (pair tail)']),
                               ('prof_mha.ncu-rep', 'mha_tc_kernel<128, fp16>', ['decoder layer', 'decoder layer']),
                               ('prof_bgemm.ncu-rep', 'bgemm_tc_kernel (training: S, PV of the first blocks)', [])):
        if (OUT / rep).exists():
            md.append(f'## `ncu --set full` -- {title}\n')
            md.append(ncu_table(OUT / rep, labels) + '\n')
    (ROOT / 'profiles' / f'{tag}_summary.md').write_text('\n'.join(md))
    print('\n'.join(md)[:6000])


if __name__ == '__main__':
    main()
