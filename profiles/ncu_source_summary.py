"""Summarise `ncu -i X.ncu-rep --page source --csv --kernel-name K` output:
opcode histogram + hottest sampled instructions."""
import csv, re, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1]))]
h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[h]; idx = {k: i for i, k in enumerate(hdr)}
data = []
for r in rows[h + 1:]:
    if not r or not r[0].startswith('0x'):
        break          # first section only
    data.append(r)
I, S = idx['Instructions Executed'], idx['# Samples']
ti = sum(int(r[I] or 0) for r in data); ts = sum(int(r[S] or 0) for r in data)
print('SASS lines', len(data), 'warp instr', ti, 'samples', ts)
op = collections.Counter(); ops = collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[idx['Source']])
    o = m.group(2).split('.')[0] if m else '?'
    op[o] += int(r[I] or 0); ops[o] += int(r[S] or 0)
for o, c in op.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 20):
    print('%-10s inst %9d (%4.1f%%) samples %7d (%4.1f%%)' % (o, c, 100 * c / ti, ops[o], 100 * ops[o] / max(ts, 1)))
print('--- hottest sampled instructions')
for r in sorted(data, key=lambda r: -int(r[S] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 20]:
    print('%7s %9s  %s' % (r[S], r[I], r[idx['Source']].strip()[:100]))
