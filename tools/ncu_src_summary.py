"""Summarise an `ncu --page source --csv` export: stall reasons and opcode mix
per kernel section (development aid).  Usage: ncu_src_summary.py file.csv [regex]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
sections = []
for i, r in enumerate(rows):
    if r and r[0] == 'Kernel Name':
        sections.append(i)
sections.append(len(rows))
seen = set()
for a, b in zip(sections[:-1], sections[1:]):
    name = rows[a][1]
    if pat and not pat.search(name):
        continue
    key = name
    if key in seen:
        continue
    seen.add(key)
    hdr = rows[a + 1]
    data = rows[a + 2:b]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = collections.Counter()
    opc = collections.Counter()
    smp = collections.Counter()
    n_inst = 0
    for r in data:
        if len(r) < len(hdr):
            continue
        src = r[ix['Source']].strip()
        parts = src.split()
        op = parts[1] if parts[0].startswith('@') else parts[0]
        op = op.split('.')[0]
        ex = int(r[ix['Instructions Executed']] or 0)
        n_inst += ex
        opc[op] += ex
        smp[op] += int(r[ix['# Samples']] or 0)
        for s in stalls:
            tot[s] += int(r[ix[s]] or 0)
    print('=====', name[:120])
    print('static instructions', len(data), 'executed warp-inst', n_inst)
    ts = sum(tot.values()) or 1
    print('--- stalls:', ', '.join(f'{k[6:]} {100 * v / ts:.1f}%'
                                   for k, v in tot.most_common(9)))
    print('--- opcodes:')
    for k, v in opc.most_common(22):
        print(f'   {k:10s}{v:12d} {100 * v / n_inst:5.1f}%  samples {smp[k]}')
