timeout 200 python -m pytest tests/test_attention_gpu.py -x -q 2>&1 | tail -1
for d in 0; do
  PP_ATT_DBG=$d ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:attention_tc --csv --log-file gpurun_out/st.csv python tools/probes/att_probe.py > /dev/null 2>&1
  echo "dbg $d: $(grep attention_tc gpurun_out/st.csv | grep duration | awk -F'","' '{printf "%s ", $NF}' | tr -d '"')"
  echo "inst: $(grep attention_tc gpurun_out/st.csv | grep inst_exec | awk -F'","' '{printf "%s ", $NF}' | tr -d '"')"
done
