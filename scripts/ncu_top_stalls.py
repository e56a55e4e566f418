"""Reads `ncu -i X.ncu-rep --page source --csv` (SASS view) from stdin and prints the instructions with the most warp-stall
samples and their dominant stall reasons, per profiled kernel launch."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'data': []}
        blocks.append(cur)
    elif cur is not None and cur['hdr'] is None and r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] is not None and len(r) == len(cur['hdr']):
        cur['data'].append(r)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for b in blocks[:1]:
    hdr, data = b['hdr'], b['data']
    ix = {h: i for i, h in enumerate(hdr)}
    num = lambda r, k: int(float(r[ix[k]] or 0))      # noqa: E731
    tot = sum(num(r, '# Samples') for r in data)
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    print(b['name'][:90], 'samples', tot, 'instructions', len(data))
    agg = {s: sum(num(r, s) for r in data) for s in stalls}
    print('stall totals:', ' '.join(f'{k[6:]}:{v}' for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    for r in sorted(data, key=lambda r: -num(r, '# Samples'))[:N]:
        n = num(r, '# Samples')
        st = sorted(((num(r, s), s) for s in stalls), reverse=True)[:3]
        print(f"{n:6d} {100 * n / max(tot, 1):5.1f}%  {r[ix['Address']][-6:]} {r[ix['Source']][:64]:64s} {' '.join(f'{s[6:]}:{c}' for c, s in st if c)}")
