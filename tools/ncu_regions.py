"""Per-region stall summary of one kernel in an .ncu-rep (source page, SASS view): python tools/ncu_regions.py rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in data); ti = sum(f(r, 'Instructions Executed') for r in data)
keys = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print('samples', tot, 'warp-instructions', ti, 'sass lines', len(data))
print('stalls %:', ', '.join(f'{k[6:]} {v:.1f}' for v, k in sorted(((sum(f(r, k) for r in data) / tot * 100, k) for k in keys), reverse=True)[:8]))
# regions split at SYNCS / BAR instructions
bounds = [0] + [i for i, r in enumerate(data) if 'SYNCS' in r[idx['Source']] or 'BAR.' in r[idx['Source']]] + [len(data)]
for a, b in zip(bounds[:-1], bounds[1:]):
    seg = data[a:b]
    s = sum(f(r, '# Samples') for r in seg) / tot * 100; ins = sum(f(r, 'Instructions Executed') for r in seg) / ti * 100
    if s < 1.0 and ins < 1.0: continue
    st = sorted(((sum(f(r, k) for r in seg) / tot * 100, k[6:]) for k in keys), reverse=True)[:4]
    print(f'[{a:5d},{b:5d}) samples {s:5.1f}% inst {ins:5.1f}%  first: {seg[0][idx["Source"]][:50]:50s}', ' '.join(f'{k}:{v:.1f}' for v, k in st))
print('hottest:')
for i, r in sorted(enumerate(data), key=lambda ir: -f(ir[1], '# Samples'))[:top]:
    print(f'  {i:5d} {f(r, "# Samples") / tot * 100:5.2f}%  {r[idx["Source"]][:90]}')
