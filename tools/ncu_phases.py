"""Per-phase (between BAR.SYNC) and per-opcode dynamic instruction breakdown from an ncu source-page CSV.
Usage: ncu -i rep --page source --csv --kernel-name regex:K > src.csv; python tools/ncu_phases.py src.csv [ctas warps steps]"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
norm = 148 * 16 * 500
if len(sys.argv) > 4: norm = int(sys.argv[2]) * int(sys.argv[3]) * int(sys.argv[4])
idx = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
seg = 0; segs = collections.OrderedDict(); ops = collections.Counter(); tot = 0
for i, r in enumerate(data):
    src = r[idx['Source']].strip()
    n = f(r, 'Instructions Executed'); smp = f(r, '# Samples')
    segs.setdefault(seg, [0, 0, 0]); segs[seg][0] += n; segs[seg][1] += smp; segs[seg][2] += 1
    m = re.match(r'(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    if m:
        op = m.group(1); base = op.split('.')[0]; key = base
        if base == 'IMAD':
            key = 'IMAD.' + ('WIDE' if 'WIDE' in op else 'MOV' if 'MOV' in op else 'IADD' if 'IADD' in op else 'SHL' if 'SHL' in op else 'X' if '.X' in op else 'mul')
        ops[key] += n; tot += n
    if src.startswith('BAR'): seg += 1
ts = sum(v[1] for v in segs.values())
for k, v in segs.items():
    if v[0] / tot > 0.003:
        print('seg %2d static %5d  exec %5.1f%%  samples %5.1f%%  per-warp-per-step %.0f' % (k, v[2], 100 * v[0] / tot, 100 * v[1] / ts, v[0] / norm))
print('total per warp per step %.0f' % (tot / norm))
print(', '.join('%s %.1f%%' % (k, 100 * v / tot) for k, v in ops.most_common(16)))
