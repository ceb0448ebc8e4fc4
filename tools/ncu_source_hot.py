"""Aggregate `ncu --page source --csv --print-source sass` output: executed warp instructions and
stall samples per code region (regions split where the execution count changes level), top opcodes.

    python tools/ncu_source_hot.py src.csv [kernel-index]"""
import csv
import sys
from collections import Counter


def main(path, which=0):
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    starts.append(len(rows))
    seg = rows[starts[which]:starts[which + 1]]
    print(seg[0][1][:120])
    hdr = seg[1]
    ia, isrc, iex, ismp = (hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"),
                           hdr.index("# Samples"))
    stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = seg[2:]
    base = int(data[0][ia], 16)
    tot_ex = sum(int(r[iex]) for r in data)
    tot_smp = sum(int(r[ismp]) for r in data)
    print(f"instructions {len(data)}  executed {tot_ex}  samples {tot_smp}")
    regions = []
    cur = None
    for r in data:
        ex = int(r[iex])
        lvl = 0 if ex == 0 else len(bin(ex))
        if cur is None or abs(lvl - cur["lvl"]) > 1:
            cur = dict(lvl=lvl, a0=int(r[ia], 16) - base, n=0, ex=0, smp=0, ops=Counter(), stalls=Counter())
            regions.append(cur)
        cur["n"] += 1
        cur["ex"] += ex
        cur["smp"] += int(r[ismp])
        cur["a1"] = int(r[ia], 16) - base
        op = r[isrc].split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
        cur["ops"][op.split(".")[0]] += ex
        for h, i in stall_cols:
            cur["stalls"][h] += int(r[i])
    for g in regions:
        if g["ex"] < tot_ex * 0.01 and g["smp"] < tot_smp * 0.01:
            continue
        st = ", ".join(f"{k[6:]} {v * 100 // max(tot_smp, 1)}%" for k, v in g["stalls"].most_common(6)
                       if v * 100 // max(tot_smp, 1) > 0)
        ops = ", ".join(f"{k} {v * 100 // max(g['ex'], 1)}%" for k, v in g["ops"].most_common(10))
        print(f"[{g['a0']:#07x}-{g['a1']:#07x}] {g['n']:4d} instrs  exec {g['ex'] * 100 / tot_ex:5.1f}%  "
              f"samples {g['smp'] * 100 / max(tot_smp, 1):5.1f}%  | {st}")
        print(f"        ops: {ops}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
