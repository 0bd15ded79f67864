# Round evidence on one GPU: tests, smoke, the default bench line, the ncu launch list of the same command and --set full
# captures of the dominant kernels, all into gpurun_out/ (copied to profiles/ by hand).  usage: profile_round.sh r02
R=${1:-r02}
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_gputest.log 2>&1; tail -3 gpurun_out/${R}_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -2 gpurun_out/${R}_smoke.log
python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; tail -c 200 gpurun_out/${R}_bench_n1.json; echo
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2>/dev/null
# launch list of the bench command (short run; shares must agree with the live CUDA-event times)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sustained > /dev/null 2>&1
# --set full of one launch of each kernel of the DDC workload (after warm-up) ...
ncu --set full --clock-control none --import-source on -k regex:"ddc_front_tc|ddc_ciccomp|ddc_hilb" -s 12 -c 3 -f -o gpurun_out/${R}_ddc_kernels \
    python bench.py --workload ddc --steps 4 --warmup 3 --no-cpu-baseline --no-sustained > /dev/null 2>&1
# ... of the STM32 stage (1024 channels, mode mix) ...
ncu --set full --clock-control none --import-source on -k regex:"rx_audio_kernel|rx_fft_pre|rx_fft_kernel" -s 3 -c 3 -f -o gpurun_out/${R}_stm32_kernels \
    python tools/gpu/rx_kernels_once.py 1024 3 > /dev/null 2>&1
# ... and of the transmit mirror (4096 channels)
ncu --set full --clock-control none --import-source on -k regex:"duc_kernel|tx_audio_kernel" -s 2 -c 2 -f -o gpurun_out/${R}_tx_kernels \
    python tools/gpu/tx_kernels_once.py 4096 > /dev/null 2>&1
for k in ddc stm32 tx; do python tools/summarize_ncu.py full gpurun_out/${R}_${k}_kernels.ncu-rep gpurun_out/${R}_${k}_kernels_ncu_full_selected.csv > /dev/null; done
python tools/summarize_ncu.py launches gpurun_out/${R}_launches_bench.csv > gpurun_out/${R}_launches_bench_summary.md; head -12 gpurun_out/${R}_launches_bench_summary.md
python tools/bench_duc.py > gpurun_out/${R}_bench_duc.json 2>/dev/null; cat gpurun_out/${R}_bench_duc.json
