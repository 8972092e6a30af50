#!/usr/bin/env python
"""Per-source-line totals (instructions executed, stall samples) of one launch in an .ncu-rep captured with
--import-source on.   python tools/ncu_lines.py rep.ncu-rep [launch_index] [top_n]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    # sections: a header row starting with "Line No"; the first section of each launch is the main .cu file
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Line No":
            cur = dict(hdr=r, rows=[])
            secs.append(cur)
        elif r and r[0] in ("File Name", "File Path"):
            fname = r[1]
            cur = None
            secs.append(dict(file=fname))
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    # group sections per launch: a launch restarts when the same file name re-appears
    launches, seen = [[]], set()
    fname = None
    for s in secs:
        if "file" in s:
            fname = s["file"]
            if fname in seen:
                launches.append([])
                seen = set()
            seen.add(fname)
        else:
            s["file"] = fname
            launches[-1].append(s)
    L = launches[launch]
    inst, samp, src = defaultdict(float), defaultdict(float), {}
    tot_i = tot_s = 0.0
    for s in L:
        h = s["hdr"]
        ci, cs = h.index("Instructions Executed"), h.index("# Samples")
        for r in s["rows"]:
            key = (s["file"].split("/")[-1], r[0])
            try:
                i, sm = float(r[ci] or 0), float(r[cs] or 0)
            except ValueError:
                continue
            inst[key] += i
            samp[key] += sm
            src[key] = r[1].strip()[:100]
            tot_i += i
            tot_s += sm
    print(f"launch {launch}: {tot_i:.0f} warp instructions, {tot_s:.0f} stall samples")
    order = (lambda k: -inst[k]) if len(sys.argv) > 4 and sys.argv[4] == "inst" else (lambda k: -samp[k])
    for key in sorted(inst, key=order)[:top]:
        print(f"{key[0]}:{key[1]:>4}  inst {inst[key] / tot_i:6.1%}  samples {samp[key] / max(tot_s, 1):6.1%}  {src[key]}")


if __name__ == "__main__":
    main()
