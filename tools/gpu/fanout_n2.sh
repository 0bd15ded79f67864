# ua3reo_fanout_* on two GPUs: the two-process test, then the DDC bench with either transport and the full chain
R=${1:-r02e}
timeout 300 python -m pytest tests/test_fanout_gpu.py -m gpu -x -q 2>&1 | tail -15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for t in ipc nccl ipc nccl; do
  timeout 240 $TR bench.py --gpus 2 --workload ddc --no-cpu-baseline --no-sustained --adc-transport $t > gpurun_out/${R}_n2_ddc_$t.json 2> gpurun_out/${R}_n2_ddc_$t.err || { echo "bench $t failed"; tail -20 gpurun_out/${R}_n2_ddc_$t.err; }
  python -c "
import sys,json; d=json.loads(open('gpurun_out/${R}_n2_ddc_$t.json').read()); print('$t', d['config']['adc_transport'], 'comm_sms', d['config']['comm_sms'], 'ms %.4f'%d['ms_per_step'], 'value %.4e'%d['value'], 'e2e %.4e'%d['e2e']['value'], 'front %.4f'%d['roofline']['kernel_ms'], d['parity'])"
done
timeout 300 $TR bench.py --gpus 2 --workload full_chain --no-cpu-baseline --no-sustained --adc-transport ipc > gpurun_out/${R}_n2_full_ipc.json 2> gpurun_out/${R}_n2_full_ipc.err || { echo "bench full failed"; tail -20 gpurun_out/${R}_n2_full_ipc.err; }
python -c "
import sys,json; d=json.loads(open('gpurun_out/${R}_n2_full_ipc.json').read()); print('full ipc', d['config']['adc_transport'], 'comm_sms', d['config']['comm_sms'], 'ms %.4f'%d['ms_per_step'], 'value %.4e'%d['value'], 'e2e %.4e'%d['e2e']['value'], d['parity'])"
