# Evidence for a change of the DDC kernels only: tests, smoke, default bench line, launch list, --set full of the DDC kernels.
# usage: profile_ddc.sh r02d
R=${1:-r02d}
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_gputest.log 2>&1; tail -3 gpurun_out/${R}_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -2 gpurun_out/${R}_smoke.log
python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; tail -c 200 gpurun_out/${R}_bench_n1.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sustained > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ddc_front_tc|ddc_ciccomp|ddc_hilb" -s 12 -c 3 -f -o gpurun_out/${R}_ddc_kernels \
    python bench.py --workload ddc --steps 4 --warmup 3 --no-cpu-baseline --no-sustained > /dev/null 2>&1
python tools/summarize_ncu.py full gpurun_out/${R}_ddc_kernels.ncu-rep gpurun_out/${R}_ddc_kernels_ncu_full_selected.csv > /dev/null
python tools/summarize_ncu.py launches gpurun_out/${R}_launches_bench.csv > gpurun_out/${R}_launches_bench_summary.md; head -12 gpurun_out/${R}_launches_bench_summary.md
