#!/usr/bin/env python
"""Static view of a kernel's SASS: instructions between consecutive CTA barriers / calls, with an opcode histogram.
Development aid (no GPU needed): python tools/sass_regions.py <lib.so> <substring of the mangled kernel name>"""
import collections
import re
import subprocess
import sys


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, funcs = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return funcs


def opcode(text):
    parts = text.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    return op.split(".")[0]


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    show = len(sys.argv) > 3
    for name, ins in functions(lib).items():
        if pat not in name:
            continue
        print(f"== {name}: {len(ins)} instructions ({16 * len(ins) / 1024:.1f} KB)")
        start, hist = 0, collections.Counter()
        for i, (addr, text) in enumerate(ins):
            hist[opcode(text)] += 1
            if text.startswith(("BAR.SYNC", "EXIT")) or "RET." in text or i == len(ins) - 1:
                top = " ".join(f"{k}:{v}" for k, v in hist.most_common(9))
                print(f"  [{ins[start][0]:05x}..{addr:05x}] {i - start + 1:4d}  {text.split()[0]:10s} {top}")
                start, hist = i + 1, collections.Counter()
        if show:
            for addr, text in ins:
                print(f"    {addr:05x}  {text}")


if __name__ == "__main__":
    main()
