"""Instruction histogram per kernel of the built library (cuobjdump -sass):
evidence for what the kernels are made of (packed FP32, bulk/tensor copies,
mbarriers).  Usage: python tools/sass_digest.py [regex] > profiles/..."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'baseband-tasks_b200', 'csrc', 'libbbt_b200.so')
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else
                 r'dd_row2_kernel.*14|dd_row2_tile.*14|dd_col_tma_kernel.*10.*5.*512'
                 r'|chanpow_tma_kernel.*10.*5.*256|dd_row_tma_tile.*14|fold_kernel')
out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True,
                     text=True, check=True).stdout
mangled_names = re.findall(r'Function : (\S+)', out)
demangle = subprocess.run(['cu++filt'], input='\n'.join(mangled_names),
                          capture_output=True, text=True)
names = dict(zip(mangled_names, demangle.stdout.splitlines()))
arch = re.search(r'arch = (\S+)', out)
print(f'# SASS digest of {os.path.relpath(LIB, ROOT)} ({arch.group(1)})\n')
print('Made by `python tools/sass_digest.py` (cuobjdump -sass, CUDA 12.9); '
      'counts are static instructions.\n')
interesting = ('FFMA2', 'FADD2', 'FMUL2', 'FFMA', 'FADD', 'FMUL', 'LDS', 'STS',
               'LDG', 'STG', 'LDGSTS', 'UBLKCP', 'UTMALDG', 'UTMASTG',
               'UTMAPF', 'UBLKPF', 'SYNCS', 'BAR', 'SHFL', 'ATOMG', 'REDG',
               'RED', 'ATOMS', 'DFMA', 'DMUL', 'DADD', 'MUFU', 'CCTL',
               'WARPSYNC', 'NANOSLEEP', 'LDL', 'STL')
for block in out.split('Function : ')[1:]:
    mangled = block.split()[0]
    name = names.get(mangled, mangled)
    if not pat.search(name):
        continue
    ops = collections.Counter()
    for line in block.splitlines():
        m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)',
                     line)
        if m:
            ops[m.group(1)] += 1
    total = sum(ops.values())
    print(f'## {name[:150]}\n')
    print(f'{total} instructions; ' + ', '.join(
        f'{k} {ops[k]}' for k in interesting if ops.get(k)))
    rest = [(k, v) for k, v in ops.most_common(14) if k not in interesting]
    print('other frequent: ' + ', '.join(f'{k} {v}' for k, v in rest) + '\n')
