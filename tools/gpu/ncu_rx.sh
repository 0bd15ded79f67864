set -e
python tools/gpu/rx_kernels_once.py 1024 3 > /dev/null
ncu --metrics gpu__time_duration.sum,sm__inst_issued.avg.per_cycle_active,smsp__inst_executed.sum --clock-control none -k regex:"rx_|duc_|tx_" --csv --log-file gpurun_out/r2_rx_launches.csv python tools/gpu/rx_kernels_once.py 1024 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_rx_launches.csv')) if len(r)>6]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value')
for r in rows[1:]:
    print(r[ki].split('(')[0], r[mi], r[vi])
PY
