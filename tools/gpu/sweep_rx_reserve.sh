# full-chain step time against the number of SMs the front kernel leaves to the STM32 stage: sweep_rx_reserve.sh "1024:6 1024:8 4096:4 ..."
for spec in ${1:-1024:4 1024:8 4096:4 4096:8}; do
ch=${spec%%:*}; r=${spec##*:}
UA3REO_RX_RESERVE_SMS=$r python bench.py --workload full_chain --channels-per-gpu $ch --steps 40 --warmup 4 --no-cpu-baseline --no-sustained 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('ch',$ch,'reserve',$r,'ms/step %.4f e2e %.4f'%(d['ms_per_step'],d['e2e']['ms_per_step']), {k:round(v,3) for k,v in d['roofline']['all_kernels_ms_per_step'].items()}, d['parity']['stm32_ranks_ok'])"
done
