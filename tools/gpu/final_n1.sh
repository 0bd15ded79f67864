# Final single-GPU evidence of a round (no --set full captures: profile_round.sh does those).  usage: final_n1.sh r02c
R=${1:-r02c}
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_gputest.log 2>&1; tail -3 gpurun_out/${R}_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -2 gpurun_out/${R}_smoke.log
python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; tail -c 200 gpurun_out/${R}_bench_n1.json; echo
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sustained > /dev/null 2>&1
python tools/summarize_ncu.py launches gpurun_out/${R}_launches_bench.csv > gpurun_out/${R}_launches_bench_summary.md; head -12 gpurun_out/${R}_launches_bench_summary.md
python bench.py --workload ddc --channels-per-gpu 1 --no-cpu-baseline --no-sustained > gpurun_out/${R}_bench_ddc_1ch.json 2>/dev/null; head -c 260 gpurun_out/${R}_bench_ddc_1ch.json; echo
python bench.py --workload full_chain --channels-per-gpu 1024 --no-cpu-baseline > gpurun_out/${R}_bench_full_1024.json 2>/dev/null; head -c 260 gpurun_out/${R}_bench_full_1024.json; echo
python bench.py --workload ddc --scaling strong --total-channels 8192 --no-cpu-baseline > gpurun_out/${R}_bench_strong_n1.json 2>/dev/null; head -c 260 gpurun_out/${R}_bench_strong_n1.json; echo
for v in 1 2 3; do UA3REO_FRONT_VARIANT=$v python bench.py --workload ddc --no-cpu-baseline --no-sustained --steps 20 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('front variant $v', r['kernel'], 'step %.4f ms front %.4f ms'%(d['ms_per_step'], r['kernel_ms']), 'frac', r['frac'], 'issue', r.get('issue_frac'))"; done
