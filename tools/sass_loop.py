#!/usr/bin/env python3
"""Count SASS instructions of the main loop of a kernel (between the loop head and its back edge)."""
import re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
cur = None; funcs = {}
for l in out.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name: continue
    back = [(a, int(re.search(r"BRA (?:[!P0-9a-zA-Z, ]*?)?0x([0-9a-f]+)", t).group(1), 16)) for a, t in ins if re.match(r"(@!?U?P\d+ )?BRA ", t) and re.search(r"0x([0-9a-f]+)", t)]
    loops = [(a, tg) for a, tg in back if tg < a]
    if not loops: continue
    a, tg = max(loops, key=lambda x: x[0] - x[1])
    body = [t for ad, t in ins if tg <= ad <= a]
    print(name[:70], "total", len(ins), "loop", len(body), "[0x%x..0x%x]" % (tg, a))
