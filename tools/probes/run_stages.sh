for lib in libprobpose_b200.so libpp_var_2_8.so libpp_var_3_5.so; do
export PROBPOSE_B200_LIB=$PWD/probpose_code_b200/$lib
for fam in pair_alt noise1_alt; do
  for b in 256 64; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:decode_kernel --csv --log-file gpurun_out/st.csv python tools/decode_probe.py $fam $b 0 3 > /dev/null 2>&1
  echo "$lib $fam $b: $(grep decode_kernel gpurun_out/st.csv | awk -F'","' '{printf "%s ", $NF}' | tr -d '"')"
  done
done
done
