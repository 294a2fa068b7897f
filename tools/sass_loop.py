#!/usr/bin/env python
"""Opcode histogram of the innermost MUFU-carrying loop of a kernel in libb200mc.so (runs here, no GPU).
Usage: tools/sass_loop.py <mangled-name-regex> [--list] [--all]   e.g.  tools/sass_loop.py 'european_kernelILi1ELb1ELi6ELb0'"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "optionslab_b200", "libb200mc.so")
if "--so" in sys.argv:  # inspect another binary (scratch experiments)
    SO = sys.argv[sys.argv.index("--so") + 1]


def functions():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        else:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m and name:
                body.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        yield name, body


def loops(body):
    """Every backward-branch loop that carries MUFU work, innermost (shortest) first."""
    found = []
    for addr, text in body:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= addr:
            continue
        loop = [t for a, t in body if tgt <= a <= addr]
        if any("MUFU" in t for t in loop):
            found.append((tgt, addr, loop))
    return sorted(found, key=lambda x: len(x[2]))


def hot_loop(body):
    best = None
    for addr, text in body:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= addr:
            continue
        loop = [t for a, t in body if tgt <= a <= addr]
        mufu = sum("MUFU" in t for t in loop)
        if mufu and (best is None or len(loop) < len(best[2])):
            best = (tgt, addr, loop)
    return best


def main():
    pat = re.compile(sys.argv[1])
    for name, body in functions():
        if not pat.search(name):
            continue
        hl = hot_loop(body)
        if hl is None:
            continue
        for tgt, addr, loop in (loops(body) if "--all" in sys.argv else [hl]):
            report(name, tgt, addr, loop)


def report(name, tgt, addr, loop):
    if True:
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for t in loop)
        mufu = sum(v for k, v in ops.items() if k.startswith("MUFU"))
        print(f"== {name}\n   loop 0x{tgt:04x}..0x{addr:04x}: {len(loop)} instructions, {mufu} MUFU")
        print("   " + ", ".join(f"{v} {k}" for k, v in ops.most_common()))
        if "--list" in sys.argv:
            for t in loop:
                print("      " + t)


if __name__ == "__main__":
    main()
