#!/usr/bin/env python
"""Compact register / spill / shared-memory table from `nvcc -Xptxas -v` output (stdin or files).
  nvcc ... -Xptxas -v -c x.cu -o /dev/null 2>&1 | python tools/ptxas_report.py
"""
import re
import subprocess
import sys


def demangle(name):
    try:
        out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except OSError:
        return name
    out = re.sub(r"\(anonymous namespace\)::", "", out)
    return re.sub(r"\(.*", "", out)


def main():
    text = sys.stdin.read()
    cur = None
    rows = []
    for line in text.splitlines():
        m = re.search(r"Compiling entry function '([^']+)'", line)
        if m:
            cur = {"name": demangle(m.group(1)), "spill": "0/0"}
            rows.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            cur["spill"] = f"{m.group(2)}/{m.group(3)}"
        m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", line)
        if m:
            cur["regs"] = m.group(1)
            cur["smem"] = m.group(3) or "0"
    for r in rows:
        print(f"{r.get('regs', '?'):>4} regs  spill st/ld {r['spill']:>9}  smem {r.get('smem', '0'):>6}  {r['name']}")


if __name__ == "__main__":
    main()
