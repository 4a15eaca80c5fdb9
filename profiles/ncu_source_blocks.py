#!/usr/bin/env python
"""Dynamic view of a kernel from an ncu capture with source counters (--import-source on): executed warp
instructions and stall samples per basic block and per opcode.  Read here, no GPU needed.

    python profiles/ncu_source_blocks.py capture.ncu-rep [--kernel NAME] [--top N] [--dump BLOCKSTART_HEX]
"""
import collections
import csv
import io
import re
import subprocess
import sys

FP64 = ("DFMA", "DMUL", "DADD", "DSETP")


def opcode(t):
    return re.sub(r"^@!?U?P\d+\s+", "", t.strip()).split()[0]


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 14
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = txt.split("\n")
    # one section per profiled launch ("Kernel Name" line, header line, rows): take the first whose name holds --kernel
    want = sys.argv[sys.argv.index("--kernel") + 1] if "--kernel" in sys.argv else "k_advance"
    heads = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
    sec = next(i for i in heads if want in lines[i])
    end = next((j for j in heads if j > sec), len(lines))
    start = next(i for i in range(sec, end) if lines[i].startswith('"Address"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
    base = int(rows[0]["Address"], 16)
    ins = []
    for r in rows:
        try:
            a = int(r["Address"], 16) - base
        except (ValueError, TypeError):
            continue
        ins.append(dict(a=a, t=r["Source"].strip(), n=int(r["Instructions Executed"] or 0), s=int(r["# Samples"] or 0), r=r))
    tot_n = sum(i["n"] for i in ins)
    tot_s = sum(i["s"] for i in ins)
    # basic blocks: split at branch targets and after branches
    tg = set()
    for i in ins:
        if "BRA" in i["t"]:
            m = re.search(r"0x([0-9a-f]+)", i["t"])
            if m:
                tg.add(int(m.group(1), 16))
    blocks, cur = [], []
    for i in ins:
        if i["a"] in tg and cur:
            blocks.append(cur)
            cur = []
        cur.append(i)
        if opcode(i["t"]).startswith(("BRA", "EXIT", "RET", "BRX", "CALL")):
            blocks.append(cur)
            cur = []
    if cur:
        blocks.append(cur)
    if "--dump" in sys.argv:
        at = int(sys.argv[sys.argv.index("--dump") + 1], 16)
        for b in blocks:
            if b[0]["a"] <= at <= b[-1]["a"]:
                for i in b:
                    print("%05x %9d %6d  %s" % (i["a"], i["n"], i["s"], i["t"]))
        return
    print(f"{rep}: {len(ins)} SASS instructions, {tot_n:.4e} executed warp instructions, {tot_s} samples")
    byop = collections.Counter()
    sop = collections.Counter()
    for i in ins:
        byop[opcode(i["t"]).split(".")[0]] += i["n"]
        sop[opcode(i["t"]).split(".")[0]] += i["s"]
    fp = sum(byop[k] for k in FP64)
    print("executed by opcode: FP64 %.1f %%; " % (100 * fp / tot_n) + ", ".join(f"{k} {100 * v / tot_n:.1f}" for k, v in byop.most_common(22)))
    print("blocks by stall samples (start, static length, executions of the block, share of executed instructions, share of samples, samples per executed instruction relative to the kernel mean):")
    mean = tot_s / tot_n
    for b in sorted(blocks, key=lambda b: -sum(i["s"] for i in b))[:top]:
        n = sum(i["n"] for i in b)
        s = sum(i["s"] for i in b)
        c = collections.Counter()
        for i in b:
            c[opcode(i["t"]).split(".")[0]] += 1
        f = sum(c[k] for k in FP64)
        print("%05x len %4d exec %10d  inst %5.1f %%  samples %5.1f %%  rel-cost %.2f  fp64 %3d other %3d" % (
            b[0]["a"], len(b), max(i["n"] for i in b), 100 * n / tot_n, 100 * s / tot_s, (s / max(n, 1)) / mean, f, len(b) - f))


if __name__ == "__main__":
    main()
