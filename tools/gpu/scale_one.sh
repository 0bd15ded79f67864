# one point of the scaling run, as the driver launches it.  usage: scale_one.sh <N> <tag>
N=$1; R=${2:-r02h}
if [ "$N" = 1 ]; then python bench.py --gpus 1 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N > gpurun_out/${R}_bench_n$N.json 2> gpurun_out/${R}_bench_n$N.err; fi
python -c "
import sys,json; d=json.loads(open('gpurun_out/${R}_bench_n$N.json').read()); f=d['full_chain']; print('N=$N ddc ms %.4f value %.4e e2e %.4e | full ms %.4f value %.4e e2e %.4e'%(d['ms_per_step'], d['value'], d['e2e']['value'], f['ms_per_step'], f['value'], f['e2e']['value']), d['config']['adc_transport'], '|', f['config'].get('spectra_gather'), d['parity'], f['parity'], d['clocks'])" || tail -30 gpurun_out/${R}_bench_n$N.err
