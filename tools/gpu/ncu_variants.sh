V=ua3reo-ddc-transceiver_b200/lib/variants
for lib in "" $V/minb1.so $V/kpre4.so $V/kpre4minb1.so; do
  for split in 0 1; do
  UA3REO_RX_SPLIT=$split UA3REO_LIB=$lib ncu --metrics gpu__time_duration.sum,sm__inst_issued.avg.per_cycle_active,smsp__inst_executed.sum --clock-control none -k regex:"rx_audio|rx_filter|rx_post" --csv --log-file gpurun_out/v.csv python tools/gpu/rx_kernels_once.py ${NCH:-1024} 2 > /dev/null 2>&1
  python - "$lib" $split <<'PY'
import csv,sys
rows=[r for r in csv.reader(open('gpurun_out/v.csv')) if len(r)>6]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value')
out={}
for r in rows[1:]:
    out.setdefault(r[ki].split('(')[0],{}).setdefault(r[mi],[]).append(r[vi])
print(sys.argv[1] or 'default', 'split', sys.argv[2], {k:{m:v[-1] for m,v in d.items()} for k,d in out.items()})
PY
  done
done
