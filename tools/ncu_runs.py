#!/usr/bin/env python
"""Where a kernel's warp instructions go: contiguous SASS runs with the same execution count, largest first,
with the CUDA source line of the run's first instruction (needs --import-source on and -lineinfo).

    python tools/ncu_runs.py rep.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, body = None, []
    for r in rows:
        if hdr is None and "Instructions Executed" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            body.append(r)
    ci = {h: i for i, h in enumerate(hdr)}
    x, src, smp = ci["Instructions Executed"], ci["Source"], ci["# Samples"]
    tot = sum(int(r[x] or 0) for r in body)
    tots = sum(int(r[smp] or 0) for r in body)
    print(f"total warp instructions {tot}, SASS rows {len(body)}, stall samples {tots}")
    runs = []
    for i, r in enumerate(body):
        e = int(r[x] or 0)
        if runs and runs[-1][2] == e:
            runs[-1][1] = i
            runs[-1][3] += int(r[smp] or 0)
        else:
            runs.append([i, i, e, int(r[smp] or 0)])
    big = sorted(runs, key=lambda t: -(t[1] - t[0] + 1) * t[2])[:top]
    for a, b, e, s in sorted(big):
        n = b - a + 1
        print(f"[{a:5d}-{b:5d}] n={n:4d} exec={e:8d} inst={n * e:9d} ({100 * n * e / tot:4.1f}%) samples {100 * s / max(tots, 1):4.1f}%  {body[a][src].strip()[:70]}")


if __name__ == "__main__":
    main()
