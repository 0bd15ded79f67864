# N>1: DDC step time against the SMs left to the NCCL broadcast: sweep_comm_sms.sh N "0 1 2 4"
N=$1
for r in $2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$r bench.py --gpus $N --steps 40 --warmup 4 --workload ddc --comm-sms $r --no-cpu-baseline --no-sustained 2>/dev/null | grep "^{" | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N',$N,'comm_sms',$r,'ms/step %.4f e2e %.4f enqueue %.3f'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['e2e']['host_enqueue_ms_per_step']), 'front %.4f share %.3f'%(d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step']), d['parity']['ddc_ranks_ok'])"
done
