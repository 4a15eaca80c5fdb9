#!/usr/bin/env python
"""Static view of a kernel's SASS: basic blocks with their opcode mix and a dispatch-model cycle estimate
(2 x FP64 instructions + 1 x every other instruction, the model that fits the ncu captures of k_advance:
profiles/README.md).  No GPU needed.

    python profiles/sass_blocks.py [lib.so] [mangled-kernel-substring] [--top N] [--dump 0xADDR]
"""
import collections
import re
import subprocess
import sys

FP64 = ("DFMA", "DMUL", "DADD", "DSETP")


def kernel_sass(lib, key):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    out, on = [], False
    for l in txt.split("\n"):
        if "Function :" in l:
            on = key in l
            continue
        if on:
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                out.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opcode(t):
    return re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]


def blocks(ins):
    tg = set()
    for _, t in ins:
        if "BRA" in t:
            m = re.search(r"0x([0-9a-f]+)", t)
            if m:
                tg.add(int(m.group(1), 16))
    bl, cur = [], []
    for a, t in ins:
        if a in tg and cur:
            bl.append(cur)
            cur = []
        cur.append((a, t))
        if opcode(t).startswith(("BRA", "EXIT", "RET", "BRX", "CALL")):
            bl.append(cur)
            cur = []
    if cur:
        bl.append(cur)
    return bl


def main():
    argv, a, opts = sys.argv[1:], [], {}
    while argv:
        x = argv.pop(0)
        if x.startswith("--"):
            opts[x] = argv.pop(0)
        else:
            a.append(x)
    lib = a[0] if a else "picles_b200/libpicles_b200.so"
    key = a[1] if len(a) > 1 else "k_advanceILb0ELb0ELi1E"
    top = int(opts.get("--top", 12))
    ins = kernel_sass(lib, key)
    if "--dump" in opts:
        at = int(opts["--dump"], 16)
        for b in blocks(ins):
            if b[0][0] <= at <= b[-1][0]:
                for x, t in b:
                    print("%05x  %s" % (x, t))
        return
    c = collections.Counter(opcode(t).split(".")[0] for _, t in ins)
    print(f"{key}: {len(ins)} instructions; " + ", ".join(f"{k} {v}" for k, v in c.most_common(14)))
    for b in sorted(blocks(ins), key=len, reverse=True)[:top]:
        c = collections.Counter(opcode(t).split(".")[0] for _, t in b)
        fp = sum(c[k] for k in FP64)
        print("%05x len %4d fp64 %3d other %3d model-cycles %4d MUFU %2d  %s" % (
            b[0][0], len(b), fp, len(b) - fp, 2 * fp + len(b) - fp, c["MUFU"],
            " ".join(f"{k}:{v}" for k, v in c.most_common(12))))


if __name__ == "__main__":
    main()
