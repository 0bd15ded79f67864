# A/B of programmatic dependent launch between the DDC kernels (UA3REO_PDL=0|1): tests first, then the step time
R=${1:-r02g}
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_gputest.log 2>&1; tail -3 gpurun_out/${R}_gputest.log
for rep in 1 2; do for v in 0 1; do UA3REO_PDL=$v python bench.py --workload ddc --no-cpu-baseline --no-sustained --steps 64 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('pdl $v ddc', 'step %.4f ms front %.4f ms e2e %.4e'%(d['ms_per_step'], r['kernel_ms'], d['e2e']['value']), d['parity']['ddc_ranks_ok'])"; done; done
for v in 0 1; do UA3REO_PDL=$v python bench.py --workload full_chain --no-cpu-baseline --no-sustained --steps 32 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('pdl $v full', 'step %.4f ms e2e %.4e'%(d['ms_per_step'], d['e2e']['value']), d['parity']['ddc_ranks_ok'], d['parity']['stm32_ranks_ok'])"; done
UA3REO_PDL=1 python bench.py --workload full_chain --channels-per-gpu 1024 --no-cpu-baseline --no-sustained --steps 32 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('pdl 1 full 1024ch', 'step %.4f ms'%(d['ms_per_step']), d['parity']['ddc_ranks_ok'], d['parity']['stm32_ranks_ok'])"
