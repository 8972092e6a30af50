#!/usr/bin/env python
"""Summarise an ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum)
into per-kernel-class DRAM bytes per launch -> profiles/traffic.json (read by bench.py's roofline).

    python tools/ncu_traffic.py gpurun_out/traffic_step.csv [gpurun_out/traffic_decode.csv] > profiles/traffic.json
"""
import csv
import json
import sys


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ci = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[start + 1:]:
        if len(r) != len(hdr):
            continue
        key = (r[ci["ID"]], r[ci["Kernel Name"]])
        val = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ns": 1e-3, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1)
        per.setdefault(key, {})[r[ci["Metric Name"]]] = val * mul
    return per


def main():
    out = {}
    step = load(sys.argv[1])
    gem = [v for (i, k), v in step.items() if "gemm_tc_kernel" in k]
    n = len(gem)
    if n:
        out["gemm_launches_captured"] = n
        out["gemm_dram_bytes_per_launch"] = sum(v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0) for v in gem) / n
        out["gemm_us_per_launch_ncu"] = sum(v.get("gpu__time_duration.sum", 0) for v in gem) / n
    att = [v for (i, k), v in step.items() if "attention" in k]
    if att:
        out["attention_dram_bytes_per_launch"] = sum(v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0) for v in att) / len(att)
    if len(sys.argv) > 2:
        dec = load(sys.argv[2])
        d = [v for (i, k), v in sorted(dec.items(), key=lambda kv: int(kv[0][0])) if "decode_kernel" in k]
        # tools/decode_probe.py model256: launches alternate plain / tta
        plain, tta = d[0::2], d[1::2]
        for name, grp in (("decode_b256_dram_bytes", plain), ("decode_b256_tta_dram_bytes", tta)):
            if grp:
                out[name] = sum(v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0) for v in grp) / len(grp)
                out[name.replace("dram_bytes", "us_ncu")] = sum(v.get("gpu__time_duration.sum", 0) for v in grp) / len(grp)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
