"""Turn ncu outputs (gpurun_out/) into the tracked summaries under profiles/.
usage: python profiles/summarise.py launches <csv> <out.md> "<title>"
       python profiles/summarise.py full <ncu-rep> <out.md> "<title>"   (also writes roofline_traffic.json)"""
import collections, csv, json, os, subprocess, sys

def launches(src, out, title):
    lines = [l for l in open(src) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = row['Kernel Name'].split('(')[0]
        v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, 'w') as f:
        f.write('# %s\n\n(cold-cache, serialised per-launch times from `ncu --metrics gpu__time_duration.sum --clock-control none`: compare SHARES)\n\n' % title)
        f.write('| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n')
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write('| %s | %d | %.1f | %.2f | %.1f%% |\n' % (k, n, t, t / n, 100 * t / tot))

def full(rep, out, title):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
    want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
            'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg',
            'sm__inst_executed_pipe_tensor.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
            'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size']
    stalls = [h for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')]
    traffic = {}
    with open(out, 'w') as f:
        f.write('# %s\n\n`ncu --set full --clock-control none --import-source on` (one launch per kernel, warm; the default of ncu, `--cache-control all`: caches are flushed before every replay pass, so dram__bytes are cold-cache figures)\n\n' % title)
        for r in rows[2:]:
            name = r[idx['Kernel Name']].split('(')[0].replace('void ', '')
            f.write('## %s\n\n| metric | value |\n|---|---|\n' % name)
            for w in want:
                if w in idx:
                    f.write('| %s | %s %s |\n' % (w, r[idx[w]], units[idx[w]]))
            tot = sum(float(r[idx[n]] or 0) for n in stalls) or 1.0
            top = sorted([(float(r[idx[n]] or 0), n) for n in stalls], reverse=True)[:6]
            f.write('| warp stall samples | %s |\n\n' % ', '.join('%s %.0f%%' % (n.replace('smsp__pcsamp_warps_issue_stalled_', ''), 100 * v / tot) for v, n in top))
            def mb(key):
                v = float(r[idx[key]]); u = units[idx[key]]
                return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
            traffic[name.split('<')[0].replace('bs::', '')] = mb('dram__bytes_read.sum') + mb('dram__bytes_write.sum')
    json.dump(traffic, open(os.path.join(os.path.dirname(out), 'roofline_traffic.json'), 'w'), indent=1)

if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](*sys.argv[2:5])
