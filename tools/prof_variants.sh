python -c "from tools import scene_fixture as sf; sf.unpack(sf.fixture('cornell-box'),'/tmp/cb')"
for v in nb00 nb14; do
  cp cudaraytracing_b200/variants/libcrt_$v.so cudaraytracing_b200/libcrt.so
  ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum --clock-control none -k regex:"k_extend|k_shadow" -s 4 -c 2 --csv cudaraytracing_b200/crt --config /tmp/cb/config.json --spp 16 --out /tmp/o.png 2>/dev/null | grep -E "k_extend|k_shadow" | awk -F'","' '{print $5, $(NF-2), $NF}' | sed "s/^/$v /"
done
