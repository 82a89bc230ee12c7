# short check of the late round-2 additions on one GPU: the reference-style round-trip test, the configs[0] bench leg,
# a fresh ncu capture of the diagonal-block kernel
timeout 300 python -m pytest tests/test_ellipse_fit_gpu.py -q -x -s -k const_ellipse 2>&1 | grep -v "Numerical error" | tail -8
timeout 300 python bench.py --steps 3 --warmup 3 --no-kriging --no-kriging-1deg --no-stochastic --no-ellipse-fit --no-cpu-baseline --no-fp64-peaks > gpurun_out/r02b_notebook.json 2> gpurun_out/r02b_notebook.err; echo bench rc=$?
tail -3 gpurun_out/r02b_notebook.err
python -c "
import json; d=json.loads(open('gpurun_out/r02b_notebook.json').read().strip().splitlines()[-1]); print(json.dumps(d['notebook'], indent=1)[:2500]); print(len(open('gpurun_out/r02b_notebook.json').read().strip().splitlines()), 'stdout lines')"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:potrf_diag -c 1 -s 3 -f \
    -o gpurun_out/r02b_potrf_diag python tools/time_kriging.py --months 60 --reps 1 > gpurun_out/r02b_potrf_diag.log 2>&1; echo ncu rc=$?
