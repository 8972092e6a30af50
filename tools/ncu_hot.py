#!/usr/bin/env python
"""Top stall-sample SASS instructions of an ncu report (source page, needs -lineinfo builds).

    python tools/ncu_hot.py gpurun_out/x.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    # several kernels may follow each other: split on "Kernel Name" rows
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1], rows=[])
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for blk in blocks:
        hdr = blk["rows"][0]
        ci = {h: i for i, h in enumerate(hdr)}
        body = [r for r in blk["rows"][1:] if len(r) == len(hdr)]
        s_i, x_i = ci["# Samples"], ci["Instructions Executed"]
        total = sum(int(r[s_i] or 0) for r in body)
        print(f"== {blk['name'][:110]}\n   total samples {total}, instructions {len(body)}")
        order = sorted(range(len(body)), key=lambda i: -int(body[i][s_i] or 0))[:top]
        for i in sorted(order):
            r = body[i]
            print(f"  [{i:5d}] {int(r[s_i] or 0):7d} ({100.0 * int(r[s_i] or 0) / max(total, 1):5.1f}%) exec {r[x_i]:>9}  {r[ci['Source']].strip()[:90]}")


if __name__ == "__main__":
    main()
