#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` logs: registers, spills, static shared memory per kernel."""
import re
import subprocess
import sys


def main(paths):
    for path in paths:
        txt = open(path).read()
        blocks = re.split(r"ptxas info\s+: Compiling entry function '", txt)[1:]
        for b in blocks:
            name = b.split("'")[0]
            try:
                name = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            except OSError:
                pass
            name = re.sub(r"sdsp_b200::", "", name)
            name = re.sub(r"\(.*", "", name)
            name = re.sub(r"^void ", "", name)
            spill = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
            regs = re.search(r"Used (\d+) registers", b)
            smem = re.search(r"(\d+) bytes smem", b)
            print(f"{name[:110]:110s} regs={regs.group(1) if regs else '?':>3s} stack={spill.group(1) if spill else '?':>4s} "
                  f"spill={spill.group(2) if spill else '?'}/{spill.group(3) if spill else '?'} smem={smem.group(1) if smem else 0}")


if __name__ == "__main__":
    main(sys.argv[1:])
