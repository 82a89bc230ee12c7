# round-end pass on one GPU: full GPU suite, smoke, bench line, launch list of the bench command
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --months 120 --month-steps 1 --no-kriging-1deg --no-stochastic --no-ellipse-fit --no-notebook --no-fp64-peaks > gpurun_out/r02_launches.log 2>&1
echo ncu rc=$?
