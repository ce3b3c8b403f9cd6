python scripts/former_profile.py coarse ${NP:-3} > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/former_list_np${NP:-3}.csv python scripts/former_profile.py coarse ${NP:-3} > /dev/null 2>&1
python - <<'PY'
import csv, collections
import os
rows=[r for r in csv.reader(open("gpurun_out/former_list_np" + os.environ.get("NP", "3") + ".csv")) if len(r)>10]
h=rows[0]; i_n=h.index("Kernel Name"); i_v=h.index("Metric Value"); i_g=h.index("Grid Size")
data=[(r[i_n][:48], r[i_g], float(r[i_v].replace(',',''))) for r in rows[1:]]
# last forward = last quarter of launches
n=len(data)//4
last=data[-n:]
d=collections.OrderedDict()
for name,grid,t in last:
    k=(name,grid); d.setdefault(k,[]).append(t)
tot=sum(t for _,_,t in last)
print("launches",len(last),"total us",round(tot/1000,1))
for (name,grid),v in sorted(d.items(), key=lambda kv:-sum(kv[1])):
    print(name.ljust(50), grid.ljust(16), len(v), round(sum(v)/len(v)/1000,1), "us avg", round(sum(v)/1000,1), "us total")
PY
