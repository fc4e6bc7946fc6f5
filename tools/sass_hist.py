"""Opcode histograms of the shipped kernels: `cuobjdump -sass pyslam_b200/libbslam.so` -> profiles/sass_<kernel>.txt
(evidence of which Blackwell/legacy paths each kernel uses: DMMA = fp64 tensor core, LDGSTS = cp.async,
RED/ATOMG = global reductions, LDGMC = multimem.ld_reduce over NVLS, ...)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'pyslam_b200', 'libbslam.so')
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
kern, hist = None, {}
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        kern = m.group(1); hist[kern] = collections.Counter(); continue
    m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
    if m and kern:
        hist[kern][m.group(1)] += 1
want = sys.argv[1:] or ['fused_panel_kernelILi3', 'panel_finish_kernelILi3', 'chol_solve_kernel', 'prepare_kernel', 'peer_pack_signal_kernel',
                        'peer_scalar_exchange_kernel', 'motion_only_kernelILb0', 'photometric_kernelILb0', 'ransac_count_kernel',
                        'reproj_block_kernelILi3', 'schur_block_kernel']
for w in want:
    for k, h in hist.items():
        if w in k:
            demangled = subprocess.run(['c++filt', k], capture_output=True, text=True).stdout.strip()
            name = re.sub(r'[^A-Za-z0-9_]', '_', demangled.split('(')[0].replace('void ', '').replace('bs::', ''))
            tot = sum(h.values())
            fam = collections.Counter()
            for op, n in h.items():
                fam[op.split('.')[0]] += n
            with open(os.path.join(ROOT, 'profiles', 'sass_%s.txt' % name), 'w') as f:
                f.write('# %s\n# %d SASS instructions (sm_100a), by opcode family then by full opcode\n' % (demangled, tot))
                for op, n in fam.most_common():
                    f.write('%-12s %6d\n' % (op, n))
                f.write('\n')
                for op, n in h.most_common(60):
                    f.write('%-40s %6d\n' % (op, n))
            print(name, tot, {k2: fam[k2] for k2 in ('DMMA', 'DFMA', 'LDGSTS', 'RED', 'ATOMG', 'LDGMC', 'SHFL', 'LDS', 'STS', 'BAR', 'MUFU', 'UBLKCP', 'SYNCS', 'UTMALDG') if fam[k2]})
            break
