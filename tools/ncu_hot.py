"""Per-opcode and per-block digest of one kernel from an .ncu-rep (source page, SASS).
usage: python tools/ncu_hot.py report.ncu-rep [block]"""
import csv, subprocess, sys
from collections import Counter

rep = sys.argv[1]
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, r = rows[0], rows[2]
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum']
for h, v in zip(hdr, r):
    if h in keys:
        print(h, v)
st = [(float(v), h) for h, v in zip(hdr, r)
      if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and 'not_issued' not in h]
for v, h in sorted(st, reverse=True)[:8]:
    print(f'{v:.3f}', h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
out, k = [], 0
for r in csv.reader(src.splitlines()):
    if r and r[0] == 'Kernel Name':
        k += 1
        if k == 2:
            break
        continue
    if len(r) > 6 and r[0].startswith('0x'):
        out.append((r[1].strip(), int(r[5]), int(r[2])))
tot = sum(x[1] for x in out); tots = sum(x[2] for x in out)
print('total warp-instructions', tot, 'stall samples', tots)
c = Counter(); s = Counter()
def opc(src):
    t = src.split()
    o = t[1] if t[0].startswith('@') else t[0]
    return o.split('.')[0]
for srcl, n, stl in out:
    c[opc(srcl)] += n; s[opc(srcl)] += stl
for op, n in c.most_common(16):
    print(f'{op:10s} {n:10d} {100*n/tot:5.1f}%  samples {s[op]}')
for b in range(0, len(out), blk):
    bl = out[b:b + blk]
    n = sum(x[1] for x in bl); stl = sum(x[2] for x in bl)
    if stl * 50 > tots or n * 50 > tot:
        print(b, 'samples', stl, 'inst', n, Counter(opc(x[0]) for x in bl).most_common(4))
