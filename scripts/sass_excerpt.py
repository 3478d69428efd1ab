#!/usr/bin/env python3
"""SASS evidence for the hot kernels (runs in the build container): per kernel the opcode histogram of the sm_100a code
in the built object files, the memory / asynchronous-copy instructions that matter for the data-movement claims, and the
first lines of the listing.  Writes profiles/<tag>_sass.txt.

    python scripts/sass_excerpt.py r2"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = os.path.join(ROOT, 'pylabfea_b200', 'csrc', 'build')
KERNELS = [('response_kernels.o', '_ZN3plf21response_passA_kernelILb1ELi4ELb0EEE', 'response_passA_kernel<1, 4, 0> (production pass A: cp.async ring, staged stores)'),
           ('response_kernels.o', '_ZN3plf26response_passA_bulk_kernelILb1ELi4EEE', 'response_passA_bulk_kernel<1, 4> (option passa_bulk: cp.async.bulk + mbarrier)'),
           ('response_kernels.o', '_ZN3plf21response_passB_kernelILi3ELb1ELi8EEE', 'response_passB_kernel<3, 1, 8> (planar isotropic sub-step loop, recurrence form)'),
           ('assembly_kernels.o', '_ZN3plf27assemble_rows_staged_kernelILb0ELb1ELi5ELb1EEE', 'assemble_rows_staged_kernel<0, 1, 5, 1> (fused BtCB + CSR assembly)'),
           ('mg_kernels.o', '_ZN3plf14mg_coop_kernelENS_8CoopArgsEPj', 'mg_coop_kernel (cooperative middle of the V-cycle + fused tail)')]
WATCH = ('LDGSTS', 'UBLKCP', 'UTMA', 'SYNCS', 'LDG', 'STG', 'LDS', 'STS', 'DFMA', 'DMUL', 'DADD', 'MUFU', 'BAR', 'ATOM', 'RED', 'SHFL', 'LDC', 'ULDC')


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r2'
    out = []
    names = {}
    for obj in sorted(set(k[0] for k in KERNELS)):
        r = subprocess.run(['cuobjdump', '-sass', os.path.join(B, obj)], capture_output=True, text=True)
        names[obj] = re.findall(r'Function : (\S+)', r.stdout)
    for obj, prefix, title in KERNELS:
        fun = [n for n in names[obj] if n.startswith(prefix)][0]   # the template kernels carry their parameter types in the name
        r = subprocess.run(['cuobjdump', '-sass', '-fun', fun, os.path.join(B, obj)], capture_output=True, text=True)
        lines = [ln for ln in r.stdout.splitlines() if re.search(r'/\*[0-9a-f]{4}\*/', ln)]
        ops = collections.Counter()
        for ln in lines:
            m = re.search(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
            if m:
                ops[m.group(1)] += 1
        out.append('== %s\n   %s : %s\n   %d SASS instructions' % (title, obj, fun, len(lines)))
        fam = collections.Counter()
        for op, n in ops.items():
            for w in WATCH:
                if op.startswith(w):
                    fam[w] += n
                    break
        out.append('   families: ' + ', '.join('%s %d' % (w, fam[w]) for w in WATCH if fam[w]))
        wide = {op: n for op, n in ops.items() if ('128' in op or '256' in op or op.startswith(('UBLKCP', 'SYNCS', 'LDGSTS')))}
        out.append('   wide / asynchronous memory ops: ' + ', '.join('%s x%d' % kv for kv in sorted(wide.items())))
        out.append('   first instructions:')
        out += ['      ' + re.sub(r'\s+/\* 0x[0-9a-f]+ \*/', '', ln).strip() for ln in lines[:14]]
        out.append('')
    path = os.path.join(ROOT, 'profiles', tag + '_sass.txt')
    with open(path, 'w') as f:
        f.write('# cuobjdump -sass of the built sm_100a objects (pylabfea_b200/csrc/build), scripts/sass_excerpt.py\n\n' + '\n'.join(out))
    print(open(path).read()[:3000])


if __name__ == '__main__':
    main()
