#!/bin/bash
# One profiling pass on the GPU box (run through gpurun): launch list, DRAM traffic, ncu --set full captures, bench.
#   bash tools/profile_round.sh r01c
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench_fp16x3.json 2> $out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode-leg --no-extra-legs > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"gemm_tc|attention" --csv --log-file $out/${tag}_traffic_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decode-leg --no-extra-legs > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:decode_kernel --csv --log-file $out/${tag}_traffic_decode.csv python tools/decode_probe.py pair_alt 256 0 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 14 -c 1 -o $out/${tag}_attention python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-decode-leg --no-extra-legs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 2 -c 2 -o $out/${tag}_decode python tools/decode_probe.py pair_alt 256 0 3 > /dev/null 2>&1
for shape in qkv fc1 fc2; do
  ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -o $out/${tag}_gemm_$shape python tools/gemm_probe.py $shape fp16x3 2 > /dev/null 2>&1
done
python tools/vitb_bench.py > $out/${tag}_vitb_bench.jsonl 2>/dev/null
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"udp_decode|revert_merge" --csv --log-file $out/${tag}_udp_revert.csv python tools/kernel_bench.py udp > /dev/null 2>&1
python tools/kernel_bench.py decode udp 2>/dev/null | grep -v Warn > $out/${tag}_kernel_bench.jsonl
python bench.py --batch 1 --steps 50 --warmup 10 --no-cpu-baseline --no-decode-leg --no-extra-legs > $out/${tag}_bench_batch1.json 2>/dev/null
ls -la $out | grep $tag
