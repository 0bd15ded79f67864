# ua3reo_gather_* on two GPUs: the two-process tests, then the full chain with the copy-engine gather against the NCCL one
R=${1:-r02f}
timeout 300 python -m pytest tests/test_fanout_gpu.py -m gpu -x -q 2>&1 | tail -15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
for t in ipc nccl; do
  timeout 300 $TR bench.py --gpus 2 --workload full_chain --no-cpu-baseline --no-sustained --adc-transport $t > gpurun_out/${R}_n2_full_$t.json 2> gpurun_out/${R}_n2_full_$t.err || { echo "bench full $t failed"; tail -20 gpurun_out/${R}_n2_full_$t.err; }
  python -c "
import sys,json; d=json.loads(open('gpurun_out/${R}_n2_full_$t.json').read()); print('full $t', d['config']['adc_transport'], '|', d['config']['spectra_gather'], 'comm_sms', d['config']['comm_sms'], 'ms %.4f'%d['ms_per_step'], 'value %.4e'%d['value'], 'e2e %.4e'%d['e2e']['value'], d['parity'])"
done
timeout 400 $TR bench.py --gpus 2 > gpurun_out/${R}_bench_n2.json 2> gpurun_out/${R}_bench_n2.err || { echo "default bench failed"; tail -20 gpurun_out/${R}_bench_n2.err; }
python -c "
import sys,json; d=json.loads(open('gpurun_out/${R}_bench_n2.json').read()); f=d['full_chain']; print('default n2: ddc ms %.4f value %.4e e2e %.4e | full ms %.4f value %.4e e2e %.4e'%(d['ms_per_step'], d['value'], d['e2e']['value'], f['ms_per_step'], f['value'], f['e2e']['value']), d['config']['adc_transport'], f['config']['spectra_gather'], f['parity'])"
