"""Per-CUDA-source-line totals from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name K`:
warp instructions executed and stall samples attributed to each source line (file:line)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; hdr = None
agg = collections.OrderedDict()
srcs = {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; hdr = None; continue
    if r[0] == 'Line No':
        hdr = r; iI = hdr.index('Instructions Executed'); iS = hdr.index('# Samples'); continue
    if hdr is None or cur_file is None:
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    # a CUDA line row has an empty Address column; its totals aggregate its SASS rows
    if r[2] in ('', '-'):
        key = (cur_file, ln)
        try:
            agg[key] = (int(r[iI] or 0), int(r[iS] or 0)); srcs[key] = r[1].strip()
        except ValueError:
            pass
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print('total warp instr', ti, 'samples', ts)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%-16s %5d  inst %9d (%4.1f%%) samples %6d (%4.1f%%)  %s' % (k[0], k[1], v[0], 100 * v[0] / max(ti, 1), v[1], 100 * v[1] / max(ts, 1), srcs[k][:90]))
