# after a change of the STM32-stage kernels: their parity tests, then the full chain at 4096 and 1024 channels
R=${1:-r02j}
python -m pytest tests/test_rx_gpu.py tests/test_rx_scale_gpu.py tests/test_fw_shim_gpu.py -m gpu -x -q > gpurun_out/${R}_rxtests.log 2>&1; tail -3 gpurun_out/${R}_rxtests.log
python bench.py --workload full_chain --no-cpu-baseline --no-sustained > gpurun_out/${R}_bench_full_4096.json 2>/dev/null
python bench.py --workload full_chain --channels-per-gpu 1024 --no-cpu-baseline --no-sustained > gpurun_out/${R}_bench_full_1024.json 2>/dev/null
for f in full_4096 full_1024; do python -c "
import json; d=json.load(open('gpurun_out/${R}_bench_$f.json')); print('$f', 'ms %.4f value %.4e e2e %.4e'%(d['ms_per_step'], d['value'], d['e2e']['value']), d['roofline']['all_kernels_ms_per_step'], d['parity']['ddc_ranks_ok'], d['parity']['stm32_ranks_ok'])"; done
