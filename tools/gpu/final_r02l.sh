# Final single-GPU evidence after the STM32-stage CTA packing: every GPU test, smoke, the default bench line, the launch list,
# --set full captures of the DDC kernels and of the STM32 kernels at 4096 channels (the packed form).
R=${1:-r02l}
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_gputest.log 2>&1; tail -3 gpurun_out/${R}_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -2 gpurun_out/${R}_smoke.log
python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; tail -c 200 gpurun_out/${R}_bench_n1.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sustained > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ddc_front_tc|ddc_ciccomp|ddc_hilb|ddc_rotate" -s 16 -c 4 -f -o gpurun_out/${R}_ddc_kernels \
    python bench.py --workload ddc --steps 4 --warmup 3 --no-cpu-baseline --no-sustained > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rx_audio_kernel|rx_fft_pre|rx_fft_kernel" -s 3 -c 3 -f -o gpurun_out/${R}_stm32_kernels \
    python tools/gpu/rx_kernels_once.py 4096 3 > /dev/null 2>&1
for k in ddc stm32; do python tools/summarize_ncu.py full gpurun_out/${R}_${k}_kernels.ncu-rep gpurun_out/${R}_${k}_kernels_ncu_full_selected.csv > /dev/null; done
python tools/summarize_ncu.py launches gpurun_out/${R}_launches_bench.csv > gpurun_out/${R}_launches_bench_summary.md; head -12 gpurun_out/${R}_launches_bench_summary.md
