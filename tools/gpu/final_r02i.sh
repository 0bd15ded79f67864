# Final single-GPU evidence of the round: every GPU test, smoke, fuzz seeds on the final kernels, PDL A/B, the default bench line,
# the reference arm, the launch list and a --set full capture of the DDC kernels.
R=${1:-r02i}
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_gputest.log 2>&1; tail -3 gpurun_out/${R}_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -2 gpurun_out/${R}_smoke.log
for seed in 201 202 203; do python tools/fuzz_ddc_duc_vs_golden.py $seed 2>&1 | tail -1; done > gpurun_out/${R}_fuzz_ddc_duc.log; cat gpurun_out/${R}_fuzz_ddc_duc.log
for v in 0 1; do UA3REO_PDL=$v python bench.py --workload ddc --no-cpu-baseline --no-sustained --steps 64 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('pdl $v ddc', 'step %.4f ms front %.4f ms e2e %.4e (%.4f ms)'%(d['ms_per_step'], r['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step']), d['parity']['ddc_ranks_ok'])"; done | tee gpurun_out/${R}_ab_pdl.log
python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; tail -c 200 gpurun_out/${R}_bench_n1.json; echo
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sustained > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ddc_front_tc|ddc_ciccomp|ddc_hilb|ddc_rotate" -s 16 -c 4 -f -o gpurun_out/${R}_ddc_kernels \
    python bench.py --workload ddc --steps 4 --warmup 3 --no-cpu-baseline --no-sustained > /dev/null 2>&1
python tools/summarize_ncu.py full gpurun_out/${R}_ddc_kernels.ncu-rep gpurun_out/${R}_ddc_kernels_ncu_full_selected.csv > /dev/null
python tools/summarize_ncu.py launches gpurun_out/${R}_launches_bench.csv > gpurun_out/${R}_launches_bench_summary.md; head -12 gpurun_out/${R}_launches_bench_summary.md
