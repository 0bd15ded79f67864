# A/B of the tensor-core front kernel's ADC delivery: staged in shared memory (1) against broadcast global loads (0)
python -m pytest tests/test_ddc_gpu.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do for v in 0 1; do UA3REO_TC_ADC_STAGE=$v python bench.py --workload ddc --no-cpu-baseline --no-sustained --steps 40 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('adc stage $v', 'step %.4f ms front %.4f ms'%(d['ms_per_step'], r['kernel_ms']), d['parity']['ddc_ranks_ok'])"; done; done
