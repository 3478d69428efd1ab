"""Register / spill report of the CUDA kernels: python scripts/ptxas_report.py [filter]  (rebuilds with -Xptxas -v)."""
import re, subprocess, sys
from pylabfea_b200 import build as b
import os
flt = sys.argv[1] if len(sys.argv) > 1 else ''
for src, extra in b.UNITS.items():
    cmd = [b.NVCC] + b.COMMON + extra + ['-Xptxas', '-v', '-c', os.path.join(b.CSRC, src), '-o', '/dev/null']
    err = subprocess.run(cmd, capture_output=True, text=True).stderr
    cur = None
    for line in err.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
            spill = ''
        elif 'spill' in line:
            spill = line.strip()
        elif 'Used' in line and cur:
            if flt in cur:
                r = re.search(r'Used (\d+) registers', line).group(1)
                print('%-70s regs %s | %s' % (cur.replace('void plf::', ''), r, spill))
            cur = None
