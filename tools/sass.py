#!/usr/bin/env python
"""Summarise SASS of one kernel in libraja_b200.so: tools_sass.py <regex> [maxlines]"""
import re, subprocess, sys
import os
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'raja_b200/lib/libraja_b200.so')
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
funcs = re.split(r'\n\s*Function : ', txt)
pat = sys.argv[1]; maxlines = int(sys.argv[2]) if len(sys.argv) > 2 else 60
filt = sys.argv[3] if len(sys.argv) > 3 else r'LDG|STG|BRA|RED|ATOM|BAR|LDS|STS|MATCH|SHFL|UBLKCP|SYNCS'
for f in funcs[1:]:
    name = f.split('\n', 1)[0].strip()
    dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    if re.search(pat, dem):
        print('==', dem[:200])
        ops = [m.group(1) for m in re.finditer(r'/\*[0-9a-f]{4,6}\*/\s+(.*?);', f)]
        print(len(ops), 'instrs')
        for o in [o for o in ops if re.search(filt, o)][:maxlines]:
            print('   ', o)
        break
