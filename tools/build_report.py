#!/usr/bin/env python3
"""Compile the library with -Xptxas -v and print registers / spills per kernel (one line each)."""
import re, subprocess, sys
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
       "-Xptxas", "-v", "--threads", "2", "-o", "protocols/secure_comparison_b200/libscb200.so", "protocols/secure_comparison_b200/csrc/scb200.cu", "protocols/secure_comparison_b200/csrc/scb200_protocol.cu"] + sys.argv[1:]
out = subprocess.run(cmd, capture_output=True, text=True)
text = out.stderr + out.stdout
cur = None
rows = []
for line in text.splitlines():
    if "error" in line or "warning" in line:
        print(line)
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1)
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", line)
    if m and cur:
        spill = m.group(2)
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        name = re.sub(r"^_ZN3scb\d+", "", cur)
        name = re.sub(r"EEvP.*", "", name).replace("ILi", "<").replace("ELi", ",")
        rows.append((name, int(m.group(1)), spill))
        cur = None
for name, regs, spill in sorted(rows):
    print(f"{name:32s} regs={regs:3d} spill={spill}")
print("exit", out.returncode)
