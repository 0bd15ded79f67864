# SMs the front kernel sets aside for the STM32 stage, with the packed CTAs: full chain at 4096 and 2048 channels
for r in 0 1 2 4 6; do UA3REO_RX_RESERVE_SMS=$r python bench.py --workload full_chain --no-cpu-baseline --no-sustained --steps 32 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('4096 ch reserve $r: step %.4f ms e2e %.4e'%(d['ms_per_step'], d['e2e']['value']), d['parity']['stm32_ranks_ok'])"; done
for r in 0 2 4 6; do UA3REO_RX_RESERVE_SMS=$r python bench.py --workload full_chain --channels-per-gpu 2048 --no-cpu-baseline --no-sustained --steps 32 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('2048 ch reserve $r: step %.4f ms'%(d['ms_per_step']), d['parity']['stm32_ranks_ok'])"; done
