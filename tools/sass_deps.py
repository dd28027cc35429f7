#!/usr/bin/env python3
"""Static dependency distances in a kernel's hottest loop: for every instruction, how many instructions earlier its most
recent FP64 (DFMA / DADD) or shuffle producer was issued.  Short distances behind a long-latency producer are the "wait"
stalls ncu reports; used to judge a row-loop schedule before spending GPU time.

    python tools/sass_deps.py <lib.so|.o> <kernel-name-substring> [loop-index]"""
import collections
import re
import subprocess
import sys


def regs_of(tok):
    m = re.match(r"-?\|?(R)(\d+)", tok)
    return int(m.group(2)) if m else None


def main():
    path, pat = sys.argv[1], sys.argv[2]
    want_loop = int(sys.argv[3]) if len(sys.argv) > 3 else None
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
    for name, ins in funcs.items():
        if pat not in name:
            continue
        addr_index = {a: i for i, (a, _) in enumerate(ins)}
        loops = []
        for i, (a, text) in enumerate(ins):
            m = re.search(r"\bBRA(?:\.U)?\b.*?(0x[0-9a-f]+)", text)
            if m:
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addr_index:
                    loops.append((addr_index[tgt], i))
        # hottest = the innermost loop with the most DFMA
        best = None
        for n, (s, e) in enumerate(loops):
            dfma = sum(1 for _, t in ins[s:e + 1] if "DFMA" in t)
            if want_loop is not None:
                if n == want_loop:
                    best = (s, e)
            elif best is None or (dfma >= 400 and (e - s) < (best[1] - best[0])) or sum(1 for _, t in ins[best[0]:best[1] + 1] if "DFMA" in t) < 400:
                if dfma >= 400:
                    if best is None or (e - s) < (best[1] - best[0]):
                        best = (s, e)
        s, e = best
        body = [t for _, t in ins[s:e + 1]]
        last_writer = {}
        hist = collections.Counter()
        detail = collections.defaultdict(list)
        for pos, text in enumerate(body):
            toks = text.replace(",", " ").split()
            if toks[0].startswith("@"):
                toks = toks[1:]
            op = toks[0]
            ops = toks[1:]
            wide = op.startswith(("DFMA", "DADD", "DMUL")) or ".64" in op
            dst = regs_of(ops[0]) if ops else None
            srcs = []
            for t in ops[1:]:
                r = regs_of(t)
                if r is not None:
                    srcs.append(r)
                    if wide or op.startswith(("IADD3", "IMAD.X", "LOP3", "SHF")) and False:
                        srcs.append(r + 1)
            if op.startswith(("DFMA", "DADD")):
                srcs = [x for r in srcs for x in (r, r + 1)]
            dist = None
            for r in srcs:
                if r in last_writer:
                    wpos, wop = last_writer[r]
                    if wop.startswith(("DFMA", "DADD", "SHFL", "LDS")):
                        d = pos - wpos
                        dist = d if dist is None else min(dist, d)
                        prod = wop.split(".")[0]
            if dist is not None:
                hist[min(dist, 12)] += 1
                detail[(op.split(".")[0], min(dist, 12))].append(pos)
            if dst is not None and not op.startswith(("ST", "BRA", "ISETP", "UISETP")):
                last_writer[dst] = (pos, op)
                if op.startswith(("DFMA", "DADD")) or op.startswith("LDS.64"):
                    last_writer[dst + 1] = (pos, op)
        total = sum(hist.values())
        print(f"== {name[:60]} loop [{s}..{e}] {len(body)} instr; consumers of FP64/SHFL/LDS results: {total}")
        for d in sorted(hist):
            print(f"   distance {d:>2}{'+' if d == 12 else ' '}: {hist[d]:4d}  " + " ".join(f"{k[0]}:{len(v)}" for k, v in sorted(detail.items()) if k[1] == d))


if __name__ == "__main__":
    main()
