#!/usr/bin/env python
"""profiles/r2_kernels_ncu.md from an `ncu --set full` capture of profiles/evidence_run.py (every kernel of the library,
one launch each at benchmark size):

    ncu -i gpurun_out/r2_all_kernels.ncu-rep --page raw --csv > /tmp/r2_all.csv
    python profiles/summarize_ncu_r2.py /tmp/r2_all.csv > profiles/r2_kernels_ncu.md

Per kernel (mean over its launches): duration, DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), achieved DRAM
GB/s and its fraction of the measured HBM peak (MEASURED_PEAKS.json), issue-slot utilisation, registers, achieved occupancy.
ncu serialises the launches and flushes caches between its replay passes: durations are cold-cache, a few per cent above
the in-step times of bench.py."""
import collections
import csv
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def num(x):
    try:
        return float(x.replace(',', ''))
    except Exception:
        return float('nan')


def short(name):
    name = re.sub(r'\(.*', '', name)
    name = name.replace('void ', '').replace('fgs::', '').replace('<unnamed>::', '').replace('unnamed>::', '').replace('(anonymous namespace)::', '')
    return name.strip()


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    peak = 6552.0
    try:
        peak = float(json.loads((ROOT / 'MEASURED_PEAKS.json').read_text())['hbm_gbs'])
    except Exception:
        pass
    want = {'dur': 'gpu__time_duration.sum', 'rd': 'dram__bytes_read.sum', 'wr': 'dram__bytes_write.sum',
            'issue': 'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'regs': 'launch__registers_per_thread',
            'occ': 'sm__warps_active.avg.pct_of_peak_sustained_active', 'grid': 'launch__grid_size', 'block': 'launch__block_size',
            'inst': 'smsp__inst_executed.sum', 'lsu': 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
            'dramp': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'}
    unit_of = {k: units[col[v]] for k, v in want.items() if v in col}
    agg = collections.OrderedDict()
    for r in data:
        name = short(r[col['Kernel Name']])
        if 'at::native' in r[col['Kernel Name']] or 'vectorized_elementwise' in name or name.startswith('at::') or 'cub::' in name:
            continue  # torch's own kernels (fills, copies, index ops of the harness)
        d = agg.setdefault(name, collections.defaultdict(list))
        for k, v in want.items():
            if v in col:
                d[k].append(num(r[col[v]]))
    def scale(k, v):  # to bytes / microseconds
        u = unit_of.get(k, '')
        if k in ('rd', 'wr'):
            return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        if k == 'dur':
            return v * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1, 'msecond': 1e3}.get(u, 1)
        return v
    print('# Round 2 -- `ncu --set full` of the kernels round 1 left without evidence (B200, benchmark sizes)\n')
    print('`ncu --set full --clock-control none --profile-from-start off -k regex:... python profiles/evidence_run.py` '
          '(profiles/r2_call14_ncu.sh; profiler windows around the calls of interest; C2 scene: 1 M Gaussians, 1920x1080; 4D: 500 k; '
          'densification on a 1 M optimizer state).  The report stays on the box (hundreds of MB); its raw page was exported there with '
          '`ncu -i ... --page raw --csv` and is summarised here by profiles/summarize_ncu_r2.py.  The ten kernels of a plain training '
          'step are in profiles/r1_kernels_ncu.md (unchanged since); `blend_forward_kernel<1>`, `blend_backward_kernel<0>`, '
          '`preprocess_kernel<0,0,0>` appear again here because the windows contain whole steps.  '
          f'HBM peak = {peak:.0f} GB/s (MEASURED_PEAKS.json).  DRAM GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / duration; '
          'ncu serialises the launches and flushes caches between passes, so durations are cold-cache.  Kernels far below the HBM peak '
          'with a high issue-slot share are issue bound (blend, ssim), with both low are latency bound (plans, bbox) or tiny (the MCMC / '
          'filter operators run on 4 k test inputs here).  Template arguments: `preprocess_kernel<inference, cov3d, defer_colour>`, '
          '`blend_forward_kernel<training>`, `blend_backward_kernel<scores>`, `preprocess_backward_kernel<skip_invisible_rest, cov3d>`.\n')
    print('| kernel | launches | duration (us) | DRAM MB | DRAM GB/s | % of HBM peak | issue slots % | LSU pipe % | regs | warps active % | grid x block |')
    print('|---|---|---|---|---|---|---|---|---|---|---|')
    for name, d in agg.items():
        m = {k: sum(v) / len(v) for k, v in d.items() if v}
        dur = scale('dur', m.get('dur', float('nan')))
        by = scale('rd', m.get('rd', 0.0)) + scale('wr', m.get('wr', 0.0))
        gbs = by / (dur * 1e-6) / 1e9 if dur == dur and dur > 0 else float('nan')
        print(f"| `{name}` | {len(d['dur'])} | {dur:.1f} | {by / 1e6:.1f} | {gbs:.0f} | {100 * gbs / peak:.1f} | {m.get('issue', float('nan')):.1f} | "
              f"{m.get('lsu', float('nan')):.1f} | {m.get('regs', float('nan')):.0f} | {m.get('occ', float('nan')):.1f} | {m.get('grid', float('nan')):.0f} x {m.get('block', float('nan')):.0f} |")


    print('\nNot in this capture: the view-parallel kernels (`peer_adam_kernel`, `peer_adam_multicast_kernel`, `peer_adam_rows_kernel`, '
          '`peer_broadcast_multicast_kernel`) need several ranks and ncu profiles one process on one GPU (B200_PROFILING.md: never a '
          'multi-rank command); they are timed with CUDA events per phase and rank instead (`per_rank[].phases_ms` of the bench line, '
          'profiles/r2_results.md section 2: small-tensor kernel 0.16 ms = 98 MB each way over NVLink at 8 GPUs, sh_rest kernel 0.30 ms = '
          '157 MB each way).  `adam_rest_invisible_kernel` belongs to the rejected `--overlap-adam` experiment and is on no default path.')


if __name__ == '__main__':
    main()
