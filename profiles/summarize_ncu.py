"""Turn the artefacts a GPU call brought back into the tracked summaries under profiles/.

    ncu -i gpurun_out/f_step.ncu-rep --page raw --csv > /tmp/f_raw.csv
    python profiles/summarize_ncu.py /tmp/f_raw.csv gpurun_out/f_launches_mine.csv

Writes  profiles/r1_kernels_ncu.md   (per-kernel table from the `ncu --set full` capture),
        profiles/r1_traffic.json     (DRAM bytes per launch, read by bench.py for roofline.traffic),
        profiles/r1_launches_summary.md (launch lists: shares of GPU time, this repo and the reference).
"""
import collections
import csv
import json
import re
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent

WANT = [('gpu__time_duration.sum', 'duration'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__registers_per_thread', 'regs/thread'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active'),
        ('smsp__inst_executed.sum', 'warp instructions'),
        ('sm__inst_issued.avg.pct_of_peak_sustained_active', 'issue slots used'),
        ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma pipe'),
        ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram throughput'),
        ('lts__t_sectors_op_read.sum', 'L2 read sectors'), ('lts__t_sectors_op_red.sum', 'L2 red sectors'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_sb'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short_sb'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier'),
        ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall not_selected'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait')]

# kernel -> profiled group name of bench.py (kernels_ms keys)
GROUP = {'blend_backward_kernel': 'blend_backward', 'blend_forward_kernel': 'blend_forward',
         'preprocess_backward_kernel': 'preprocess_backward', 'preprocess_kernel': 'preprocess',
         'ranges_gather_kernel': 'ranges_and_gather', 'offsets_scan_kernel': 'offsets scan'}


def short(name):
    n = re.sub(r'\(.*', '', name).replace('void ', '')
    for junk in ('fgs::<unnamed>::', '<unnamed>::', 'unnamed>::', 'fgs::'):
        n = n.replace(junk, '')
    return n


def num(x):
    try:
        return float(x.replace(',', ''))
    except ValueError:
        return None


def kernels_table(raw_csv):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index('Kernel Name')
    groups = collections.OrderedDict()
    for r in data:
        groups.setdefault(short(r[ik]), []).append(r)
    order = sorted(groups, key=lambda g: -sum(num(r[hdr.index('gpu__time_duration.sum')]) for r in groups[g]))
    lines = ['| metric | ' + ' | '.join(f'`{g}` (x{len(groups[g])})' for g in order) + ' |', '|---|' + '---|' * len(order)]
    for key, label in WANT:
        if key not in hdr:
            continue
        i = hdr.index(key)
        cells = []
        for g in order:
            vals = [num(r[i]) for r in groups[g] if num(r[i]) is not None]
            m = sum(vals) / len(vals) if vals else float('nan')
            cells.append(f'{m:,.0f}' if abs(m) >= 1000 else f'{m:.2f}')
        lines.append(f'| {label} ({units[i] or "-"}) | ' + ' | '.join(cells) + ' |')
    ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    traffic = {}
    for g in order:
        base = g.split('<')[0]
        if base in GROUP:
            per = [num(r[ir]) * scale[units[ir]] + num(r[iw]) * scale[units[iw]] for r in groups[g]]
            traffic[GROUP[base]] = int(sum(per) / len(per))
    return lines, traffic


def launch_summary(path, title):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    d, total = collections.OrderedDict(), 0.0
    for r in rows[h + 1:]:
        if len(r) < 15:
            continue
        t = float(r[-1]) / 1000.0
        d.setdefault(short(r[4]), []).append(t)
        total += t
    out = [f'## {title}', '', '| kernel | launches | us / launch | share of GPU time |', '|---|---|---|---|']
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        if sum(v) / total < 0.003:
            continue
        out.append(f'| `{k[:80]}` | {len(v)} | {sum(v) / len(v):.1f} | {100 * sum(v) / total:.1f} % |')
    return out


def main():
    raw_csv, launches = sys.argv[1], sys.argv[2]
    lines, traffic = kernels_table(raw_csv)
    (HERE / 'r1_kernels_ncu.md').write_text(
        '# Round 1 -- `ncu --set full` of every kernel of one training step (B200, C2: 1M Gaussians, 1920x1080)\n\n'
        'Command: `ncu --set full --clock-control none --import-source on -k regex:"blend_|preprocess_backward|'
        'preprocess_kernel|onesweep|ranges_gather|create_instances|offsets_scan" -s 80 -c 16 python bench.py --steps 2 '
        '--warmup 3 --no-cpu-baseline` (report gpurun_out/f_step.ncu-rep, not tracked; extracted with '
        '`ncu -i ... --page raw --csv` and profiles/summarize_ncu.py; means over the launches captured).\n\n'
        + '\n'.join(lines) + '\n')
    (HERE / 'r1_traffic.json').write_text(json.dumps(
        {'source': 'ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch (profiles/r1_kernels_ncu.md)',
         'bytes_per_launch': traffic}, indent=1) + '\n')
    out = ['# Round 1 -- ncu launch lists (gpu__time_duration.sum, --clock-control none)', '',
           'Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv python bench.py '
           '--steps 2 --warmup 3 --no-cpu-baseline` (reference: `--impl reference`, captured earlier in the round, its '
           'code is unchanged).  Per-launch times are cold-cache and serialised: compare SHARES.  Raw lists: '
           '`r1_launches_mine.csv`, `r1_launches_ref.csv`.', '']
    out += launch_summary(launches, 'this repo (fused backward+Adam path)')
    ref = HERE / 'r1_launches_ref.csv'
    if ref.exists():
        out += [''] + launch_summary(ref, 'reference FasterGSCudaBackend (baseline/_ref)')
    (HERE / 'r1_launches_summary.md').write_text('\n'.join(out) + '\n')
    print(json.dumps(traffic))


if __name__ == '__main__':
    main()
