#!/usr/bin/env python
"""Counts, per kernel of the shipped libhsc_b200.so, the SASS mnemonics that show which hardware paths it uses
(`cuobjdump -sass`): UTCHMMA / UTCQMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk,
the non-tensor TMA form), SYNCS (mbarrier), REDUX (warp reduce), ATOMS / ATOMG / RED (atomics), LDL / STL (spills), and the
registers per thread.  Runs on the CPU box:   python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'hierarchical_sparse_coding_b200', 'libhsc_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', so], stdout=subprocess.PIPE, text=True).stdout
res = subprocess.run(['cuobjdump', '-res-usage', so], stdout=subprocess.PIPE, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.match(r'\s*Function (\S+):', line)
    if m:
        cur = m.group(1)
    m = re.search(r'REG:(\d+)', line)
    if m and cur:
        regs[cur] = int(m.group(1))
KEYS = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'UTCBAR', 'UBLKCP', 'UBLKPF', 'SYNCS', 'REDUX', 'ATOMS', 'ATOMG', 'RED.', 'BAR.', 'LDL', 'STL', 'DFMA', 'FFMA']
demangle = lambda n: subprocess.run(['c++filt', n], stdout=subprocess.PIPE, text=True).stdout.strip()      # noqa: E731
print('# %s (%d bytes), sm_100a SASS; counts of static instructions per kernel' % (os.path.basename(so), os.path.getsize(so)))
print('%-100s %5s %6s  %s' % ('kernel', 'regs', 'instr', '  '.join(KEYS)))
for blk in re.split(r'\n\s*Function : ', sass)[1:]:
    name, body = blk.split('\n', 1)
    ops = re.findall(r'\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)', body)
    counts = [sum(1 for o in ops if o.startswith(k)) for k in KEYS]
    short = re.sub(r'hsc::', '', demangle(name.strip()))
    short = re.sub(r'\(.*$', '', short)
    print('%-100s %5s %6d  %s' % (short[:100], regs.get(name.strip(), '?'), len(ops), '  '.join('%*d' % (len(k), c) for k, c in zip(KEYS, counts))))
