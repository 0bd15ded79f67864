#!/usr/bin/env python3
"""DEVELOPMENT AID: opcode histogram of ddc_front_bt_kernel's hottest loop (the largest backward-branch body)."""
import re, subprocess, sys, collections
for lib in sys.argv[1:]:
    out = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN3ua319ddc_front_bt_kernelEPKsPKijPKjS5_S5_jPmj", lib], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    # backward branches
    best = None
    for idx, (addr, txt) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`\(\.L_x_\d+\)|BRA\S*.*0x([0-9a-f]+)", txt)
        if "BRA" in txt:
            t = re.search(r"0x([0-9a-f]+)", txt)
            if t:
                tgt = int(t.group(1), 16)
                if tgt < addr:
                    n = sum(1 for a, _ in ins if tgt <= a <= addr)
                    if best is None or n > best[0]:
                        best = (n, tgt, addr)
    n, tgt, addr = best
    body = [t for a, t in ins if tgt <= a <= addr]
    hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in body)
    print(lib.split("/")[-1], "loop instrs:", n, dict(hist.most_common()))
