# the shipped state, one GPU: every GPU test, smoke, the default bench line
R=${1:-r02m}
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_gputest.log 2>&1; tail -3 gpurun_out/${R}_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -2 gpurun_out/${R}_smoke.log
python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; tail -c 200 gpurun_out/${R}_bench_n1.json; echo
