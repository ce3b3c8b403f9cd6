python scripts/coarse_modes.py && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/coarse_modes.csv python scripts/coarse_modes.py > /dev/null 2>&1; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/coarse_modes.csv")) if len(r)>10]
h=rows[0]; i_n=h.index("Kernel Name"); i_v=h.index("Metric Value"); i_id=h.index("ID")
for r in rows[1:]:
    if "tc_pre" in r[i_n] or "corr_tc" in r[i_n]: print(r[i_id], r[i_n][:50], r[i_v])
PY
