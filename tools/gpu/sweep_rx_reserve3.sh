# the 1024-channel full chain (192-thread stage CTAs) with fewer SMs set aside, and the new defaults at 4096 / 2048
for r in 0 2 5; do UA3REO_RX_RESERVE_SMS=$r python bench.py --workload full_chain --channels-per-gpu 1024 --no-cpu-baseline --no-sustained --steps 32 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('1024 ch reserve $r: step %.4f ms'%(d['ms_per_step']), d['parity']['stm32_ranks_ok'])"; done
for n in 4096 2048; do python bench.py --workload full_chain --channels-per-gpu $n --no-cpu-baseline --no-sustained --steps 32 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$n ch default: step %.4f ms e2e %.4e'%(d['ms_per_step'], d['e2e']['value']), d['parity']['stm32_ranks_ok'])"; done
python -m pytest tests/test_rx_gpu.py tests/test_fw_shim_gpu.py tests/test_rx_scale_gpu.py -m gpu -x -q 2>&1 | tail -2
