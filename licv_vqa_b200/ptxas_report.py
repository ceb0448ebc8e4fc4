"""Summarise lib/ptxas.log: registers / spills / static smem per kernel (python -m licv_vqa_b200.ptxas_report)."""
import os
import re
import subprocess


def report(path=None):
    path = path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "ptxas.log")
    log = open(path).read()
    rows = []
    for b in log.split("Compiling entry function '")[1:]:
        name = b.split("'")[0]
        regs = re.search(r"Used (\d+) registers", b)
        sp = re.search(r"(\d+) bytes spill stores", b)
        sm = re.search(r"(\d+) bytes smem", b)
        rows.append((name, int(regs.group(1)) if regs else -1, int(sp.group(1)) if sp else 0,
                     int(sm.group(1)) if sm else 0))
    dem = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True,
                         text=True).stdout.splitlines()
    out = []
    for d, r in zip(dem, rows):
        d = d.replace("licv::(anonymous namespace)::", "").replace("void ", "")
        d = re.sub(r"\(.*", "", d)
        out.append((d, r[1], r[2], r[3]))
    return out


if __name__ == "__main__":
    for name, regs, spill, smem in report():
        print(f"{regs:>4} regs {spill:>4} B spill {smem:>6} B smem  {name}")
