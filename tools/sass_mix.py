#!/usr/bin/env python
"""Instruction mix of a SASS address range: python tools/sass_mix.py file.sass 0xa160 0xce40"""
import re, sys, collections
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
mix = collections.Counter()
n = 0
for line in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if not m: continue
    a = int(m.group(1), 16)
    if not (lo <= a <= hi): continue
    ins = re.sub(r"^@!?U?P\d+\s+", "", m.group(2)).split()[0]
    mix[ins.split(".")[0]] += 1; n += 1
print(n, "instructions")
print("  ".join(f"{k}:{v}" for k, v in mix.most_common()))
