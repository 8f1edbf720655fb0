#!/usr/bin/env python
"""Development aid: registers / stack / spills per kernel from `nvcc -Xptxas -v` (compiles lf_kernels.cu for sm_100a).
usage: python tools/ptxas_usage.py [extra nvcc flags, e.g. -DLF_SHADE_MINBLOCKS=6]"""
import re
import subprocess
import sys

cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-Xcompiler", "-fPIC", "-std=c++17",
       "-Iinclude", "-Ilavaframe_b200/csrc", "-Xptxas", "-v", "-c", "lavaframe_b200/csrc/lf_kernels.cu", "-o", "/tmp/ptxas_usage.o"]
if not any(a.startswith("-fmad") for a in sys.argv[1:]):
    cmd.append("-fmad=false")
out = subprocess.run(cmd + sys.argv[1:], capture_output=True, text=True).stderr
name = None
rows = {}
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void lf::", "")
        rows[name] = {}
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and name:
        rows[name].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        rows[name]["regs"] = int(m.group(1))
        m2 = re.search(r"(\d+) bytes smem", line)
        rows[name]["smem"] = int(m2.group(1)) if m2 else 0
if not rows:
    print(out)
for k, v in sorted(rows.items()):
    print(f"{k:58s} regs {v.get('regs', '?'):>3}  stack {v.get('stack', 0):>4}  spill st/ld {v.get('spill_st', 0):>4}/{v.get('spill_ld', 0):<4}  smem {v.get('smem', 0)}")
