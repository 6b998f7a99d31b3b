#!/usr/bin/env python
"""Opcode histogram of one kernel: tools/sass_ops.py <binary-or-.so> <regex on demangled name>"""
import re, subprocess, sys, collections
so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
funcs = re.split(r'\n\s*Function : ', txt)
for f in funcs[1:]:
    name = f.split('\n', 1)[0].strip()
    dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    if re.search(pat, dem):
        ops = [m.group(1) for m in re.finditer(r'/\*[0-9a-f]{4,6}\*/\s+(.*?);', f)]
        print(dem[:170], len(ops), 'instrs')
        c = collections.Counter()
        for o in ops:
            o = re.sub(r'^@!?U?P\d+\s+', '', o)
            c[o.split()[0].split('.')[0]] += 1
        print(' '.join(f'{k}:{v}' for k, v in c.most_common(40)))
        if len(sys.argv) > 3: print('\n'.join(ops))
        break
