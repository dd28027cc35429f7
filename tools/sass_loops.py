#!/usr/bin/env python3
"""Summarise the loops of a kernel's SASS: for every backward branch, the instruction mix of its body.

    python tools/sass_loops.py <lib.so|cubin> <kernel-name-substring> [--dump N]

Used to check that the Montgomery row loop is IMAD.WIDE-dominated (no MOV / carry-chain overhead)
before spending GPU time."""
import collections
import re
import subprocess
import sys


def main():
    path, pat = sys.argv[1], sys.argv[2]
    dump = int(sys.argv[sys.argv.index("--dump") + 1]) if "--dump" in sys.argv else -1
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
        print(f"== {name}: {len(ins)} instructions")
        addr_index = {a: i for i, (a, _) in enumerate(ins)}
        loops = []
        for i, (a, text) in enumerate(ins):
            m = re.search(r"\bBRA(?:\.U)?\b.*?(0x[0-9a-f]+)", text)
            if m:
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addr_index:
                    loops.append((addr_index[tgt], i))
        for n, (s, e) in enumerate(loops):
            body = ins[s:e + 1]
            mix = collections.Counter()
            for _, text in body:
                op = text.split()[0]
                if op.startswith("@"):
                    op = text.split()[1]
                mix[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith("IMAD") and "." in op else "")] += 1
            total = len(body)
            wide = sum(v for k, v in mix.items() if k.startswith("IMAD.WIDE"))
            print(f"  loop {n}: [{ins[s][0]:#x}..{ins[e][0]:#x}] {total} instr, IMAD.WIDE {wide} ({100.0 * wide / total:.0f}%)  "
                  + " ".join(f"{k}:{v}" for k, v in mix.most_common(14)))
            if n == dump:
                for a, text in body:
                    print(f"      {a:#06x}  {text}")


if __name__ == "__main__":
    main()
