#!/usr/bin/env python
"""Turn the outputs of `bash tools/profile_round.sh <tag>` (in gpurun_out/) into the tracked files under profiles/:
the bench lines, the ncu launch list, the DRAM-traffic summary (profiles/traffic.json, read by bench.py) and
profiles/<tag>_summary.md (bench digest, one steady-state step, the `ncu --set full` captures).

    python tools/profile_summary.py r02e [--launches-per-step 111] [--notes notes.md]
"""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def run(*cmd):
    return subprocess.run([sys.executable, *cmd], capture_output=True, text=True, cwd=ROOT).stdout


def main():
    tag = sys.argv[1]
    per_step = int(sys.argv[sys.argv.index("--launches-per-step") + 1]) if "--launches-per-step" in sys.argv else 0
    notes = open(sys.argv[sys.argv.index("--notes") + 1]).read() if "--notes" in sys.argv else ""
    for src, dst in ((f"{tag}_bench_fp16x3.json", None), (f"{tag}_bench_batch1.json", None), (f"{tag}_launches.csv", f"{tag}_launches_bench_fp16x3.csv"),
                     (f"{tag}_traffic_decode.csv", None), (f"{tag}_kernel_bench.jsonl", None), (f"{tag}_vitb_bench.jsonl", None),
                     (f"{tag}_udp_revert.csv", None)):
        if os.path.exists(os.path.join(OUT, src)):
            shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst or src))
    step_csv = os.path.join(OUT, f"{tag}_traffic_step.csv")
    if os.path.exists(step_csv):
        with open(step_csv, errors="replace") as f, open(os.path.join(PROF, f"{tag}_traffic_step_head.csv"), "w") as g:
            g.writelines(f.readlines()[:300])
        traffic = run("tools/ncu_traffic.py", step_csv, os.path.join(OUT, f"{tag}_traffic_decode.csv"))
        if traffic.strip().startswith("{"):
            open(os.path.join(PROF, "traffic.json"), "w").write(traffic)
    d = json.load(open(os.path.join(PROF, f"{tag}_bench_fp16x3.json")))
    if not per_step:
        per_step = int(round(d["gpu_launches"] / d["steps"]))
    launch = run("tools/summarize_launches.py", os.path.join(OUT, f"{tag}_launches.csv"), "--last", str(per_step))
    reps = [os.path.join(OUT, f"{tag}_{k}.ncu-rep") for k in ("attention", "gemm_qkv", "gemm_fc1", "gemm_fc2", "decode")]
    ncu = run("tools/ncu_summary.py", *[r for r in reps if os.path.exists(r)])
    dr, lb, rf = d.get("decode_roofline") or {}, d.get("library_baseline") or {}, d["roofline"]
    lines = [f"# {tag} - ncu summaries and measurements (one B200; `bash tools/profile_round.sh {tag}`, `python tools/profile_summary.py {tag}`; "
             "the `.ncu-rep` files stay in `gpurun_out/`)", "",
             f"Files under `profiles/`: `{tag}_bench_fp16x3.json` (default `python bench.py`, every leg), `{tag}_bench_batch1.json`, "
             f"`{tag}_launches_bench_fp16x3.csv` (ncu launch list, `--metrics gpu__time_duration.sum --clock-control none`), "
             f"`{tag}_traffic_step_head.csv` / `{tag}_traffic_decode.csv` / `traffic.json` (DRAM bytes per launch), `{tag}_kernel_bench.jsonl`, "
             f"`{tag}_vitb_bench.jsonl`, `{tag}_udp_revert.csv`.", "", f"## Bench line (`{tag}_bench_fp16x3.json`)", "",
             f"* value **{d['value']:.0f} persons/s** resident ({d['ms_per_step']:.2f} ms per step, batch 64, flip-TTA, fp16x3), e2e "
             f"**{d['e2e']['value']:.0f}** (pinned host crops -> host records; {d['e2e']['value'] / d['value']:.3f} of the resident number)"
             + (f", e2e_list {d['e2e_list']['value']:.0f} (list of 64 pageable per-person tensors)" if d.get("e2e_list") else "")
             + (f", config4 (256 crops per GPU) {d['config4']['value']:.0f}" if d.get("config4") else "")
             + (f", config5 (ViT-B backbone, batch 128) {d['config5']['crops_per_s']:.0f} crops/s = {d['config5']['mma_frac_of_burst']:.2f} of the "
                "burst bf16 peak on the MMA pipe" if d.get("config5") else "")
             + (f", CPU oracle {d['cpu_baseline']['value']:.1f} persons/s on {d['cpu_baseline']['cores']} cores" if d.get("cpu_baseline") else "")
             + f"; {per_step} launches per step.",
             f"* roofline (GEMM class): {rf['achieved']:.0f} algorithmic TFLOP/s = {rf['frac']:.3f} of the peak ({rf['mma_frac']:.2f} on the MMA pipe: "
             f"3 MMAs per product); `{rf['peak_source']}`."]
    if dr:
        lines.append(f"* decode kernel at batch 256 (launches replayed from one CUDA graph): model logits {dr['us_per_launch']:.1f} us "
                     f"({dr['frac']:.2f}) / TTA {dr['tta']['us_per_launch']:.1f} us ({dr['tta']['frac']:.2f}); planted peaks "
                     f"{dr['planted']['us_per_launch']:.1f} us ({dr['planted']['frac']:.2f}) / TTA {dr['planted']['tta']['us_per_launch']:.1f} us "
                     f"({dr['planted']['tta']['frac']:.2f}).  ncu: DRAM bytes = algorithmic bytes (`traffic.json`).")
    if lb:
        lines.append(f"* library baseline on the same GPU (torch eager, model only): fp32 {lb['fp32']['persons_per_s']:.0f} persons/s "
                     f"(err {lb['fp32']['max_keypoint_err_px']:.1e} px), TF32 {lb['tf32']['persons_per_s']:.0f} ({lb['tf32']['max_keypoint_err_px']:.1e} px), "
                     f"bf16 autocast {lb['bf16_autocast']['persons_per_s']:.0f} ({lb['bf16_autocast']['max_keypoint_err_px']:.0f} px).")
    lines += [f"* kernel_ms_per_step (event pair around every launch, serialised): {json.dumps(d['kernel_ms_per_step'])}.", ""]
    if notes:
        lines += [notes.rstrip(), ""]
    lines += ["## One steady-state step (ncu launch list)", "", launch.rstrip(), "",
              "## Captures (`ncu --set full --clock-control none --import-source on`)", "", ncu.rstrip(), ""]
    open(os.path.join(PROF, f"{tag}_summary.md"), "w").write("\n".join(lines))
    print("wrote", os.path.join(PROF, f"{tag}_summary.md"))


if __name__ == "__main__":
    main()
