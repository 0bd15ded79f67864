# side measurements quoted in DESIGN.md: the DDC alone at the full chain's channel count, the full chain at 1024 channels, one channel
R=${1:-r02i}
python bench.py --workload ddc --channels-per-gpu 4096 --no-cpu-baseline --no-sustained > gpurun_out/${R}_bench_ddc_4096.json 2>/dev/null
python bench.py --workload full_chain --channels-per-gpu 1024 --no-cpu-baseline --no-sustained > gpurun_out/${R}_bench_full_1024.json 2>/dev/null
python bench.py --workload ddc --channels-per-gpu 1 --no-cpu-baseline --no-sustained > gpurun_out/${R}_bench_ddc_1ch.json 2>/dev/null
for f in ddc_4096 full_1024 ddc_1ch; do python -c "
import json; d=json.load(open('gpurun_out/${R}_bench_$f.json')); print('$f', 'ms %.4f value %.4e e2e %.4e'%(d['ms_per_step'], d['value'], d['e2e']['value']), d['roofline']['all_kernels_ms_per_step'], d['parity']['ddc_ranks_ok'])"; done
