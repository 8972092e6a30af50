#!/usr/bin/env python
"""Blackwell-native evidence that lives in the tree: per object file, how many tcgen05 / TMA / TMEM / mbarrier /
packed-fp32 instructions the shipped SASS contains (cuobjdump -sass of probpose_code_b200/build/*.o).

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = [("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA(?!\.2CTA)"), ("UTCBAR", r"\bUTCBAR"), ("UTMALDG", r"\bUTMALDG"),
             ("UBLKCP", r"\bUBLKCP"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("SYNCS", r"\bSYNCS"), ("FFMA2", r"\bFFMA2"),
             ("FADD2/FMUL2", r"\bF(ADD|MUL)2"), ("REDUX", r"\b(C?REDUX)"), ("HMMA (mma.sync)", r"\bHMMA"), ("STG.E.256 / ENL2", r"\bSTG\.E\.(ENL2\.)?256")]


def main():
    print("# SASS instruction summary of the shipped objects (sm_100a)\n")
    print("`cuobjdump -sass probpose_code_b200/build/<file>.o`, instruction counts over every kernel in the object. "
          "UTCHMMA = tcgen05.mma, UTCHMMA.2CTA = cta_group::2, UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor (TMA), "
          "UBLKCP = cp.async.bulk, LDTM / STTM = tcgen05.ld / st (TMEM), SYNCS = mbarrier, FFMA2 = packed fp32 FMA.\n")
    print("| object | SASS instr. | " + " | ".join(m for m, _ in MNEMONICS) + " |")
    print("|---|---:|" + "---:|" * len(MNEMONICS))
    for obj in sorted(glob.glob(os.path.join(ROOT, "probpose_code_b200", "build", "*.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        lines = [l for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s", l)]
        counts = [sum(1 for l in lines if re.search(rx, l)) for _, rx in MNEMONICS]
        print(f"| `{os.path.basename(obj)}` | {len(lines)} | " + " | ".join(str(c) for c in counts) + " |")
    ver = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
    print(f"\nToolchain: {ver}; flags `-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo`.")


if __name__ == "__main__":
    main()
