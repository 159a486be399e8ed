#!/usr/bin/env python
"""Summarise nvcc -Xptxas -v logs: kernel, registers, spills, smem."""
import re, sys, subprocess
for path in sys.argv[1:]:
    txt = open(path).read()
    for m in re.finditer(r"Compiling entry function '([^']+)' for 'sm_100a'\n(?:ptxas info\s+: Function properties for [^\n]+\n)?\s*(?:ptxas info\s+:\s*)?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", txt):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        print(f"{name:70s} regs={m.group(5):>3s} stack={m.group(2):>4s} spill_st={m.group(3):>4s} spill_ld={m.group(4):>4s} smem={m.group(6) or 0}")
