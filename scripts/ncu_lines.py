"""Summarise an ncu report per kernel: headline metrics + hottest source lines (needs -lineinfo, --import-source on).
Usage: python scripts/ncu_lines.py gpurun_out/x/prof.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 22
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__shared_mem_per_block_dynamic']
stall_keys = [h for h in hdr if 'issue_stalled' in h and h.endswith('.ratio')]
for r in rows[2:]:
    print('====', r[idx['Kernel Name']][:60])
    for w in want:
        if w in idx:
            print(f"  {w:78s} {r[idx[w]]} {units[idx[w]]}")
    vals = sorted(((float(r[idx[k]]), k) for k in stall_keys), reverse=True)
    print('  stalls/issue:', ', '.join(f"{k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')}={v:.2f}" for v, k in vals[:8]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kern = collections.OrderedDict()
cur_file = cur_fn = None
h = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    elif r[0] == 'Function Name':
        cur_fn = r[1][:46]
    elif r[0] == 'Line No':
        h = {n: i for i, n in enumerate(r)}
    elif h is not None and len(r) >= len(h):
        try:
            s = int(r[h['Warp Stall Sampling (All Samples)']])
            e = int(r[h['Instructions Executed']])
        except (ValueError, KeyError):
            continue
        if not r[0].strip():
            continue          # SASS-only row (the same samples are also attributed to a source line)
        d = kern.setdefault(cur_fn, {})
        key = (cur_file, r[0])
        ent = d.setdefault(key, [0, 0, r[1].strip()[:100]])
        ent[0] += s
        ent[1] += e
for fn, d in kern.items():
    ts = sum(v[0] for v in d.values()) or 1
    te = sum(v[1] for v in d.values()) or 1
    print(f"\n==== {fn}: {ts} samples, {te} warp-instructions")
    byfile = collections.defaultdict(lambda: [0, 0])
    for (f, _), v in d.items():
        byfile[f][0] += v[0]
        byfile[f][1] += v[1]
    print('  by file:', {f: (f"{100 * a / ts:.0f}% samples", f"{100 * b / te:.0f}% inst") for f, (a, b) in byfile.items()})
    for (f, ln), (s, e, text) in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {f[:14]:14s}:{ln:>4} {100 * s / ts:5.1f}% smp {100 * e / te:5.1f}% inst  {text}")
