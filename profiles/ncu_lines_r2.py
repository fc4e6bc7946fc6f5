"""Per-source-line warp-stall samples of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K
--print-source cuda,sass` (compiled with -lineinfo): which lines of OUR source the kernel's time sits on.
usage: python profiles/ncu_lines_r2.py <csv> <top-n>"""
import csv, os, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur, lines, per_file = None, [], collections.Counter()
col = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = os.path.basename(r[1]); continue
    if r[0] == 'Line No':
        col = {k: i for i, k in enumerate(r)}; continue
    if r[0] == 'Function Name' or col is None:
        continue
    if r[0].strip().isdigit():
        s = r[col['# Samples']]
        if s.isdigit() and int(s) > 0:
            lines.append((int(s), cur, int(r[0]), r[1].strip()))
            per_file[cur] += int(s)
tot = sum(s for s, *_ in lines) or 1
print('samples attributed to source lines: %d' % tot)
print('by file: ' + ', '.join('%s %.1f%%' % (f, 100. * n / tot) for f, n in per_file.most_common()))
print('| samples | share | file:line | source |\n|---:|---:|---|---|')
for s, f, ln, src in sorted(lines, reverse=True)[:top]:
    print('| %d | %.1f%% | %s:%d | `%s` |' % (s, 100. * s / tot, f, ln, src[:110].replace('|', '/')))
