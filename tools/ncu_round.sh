#!/bin/bash
# Round profile pass (run under gpurun, after the plain bench exited 0):
#   1. launch list of the bench command (gpu__time_duration per launch)
#   2. ncu --set full of the ellipse tile kernel on the full 1-degree launch (nu = 2.2 and 1.5)
#   3. ncu --set full of the DMMA GEMM inside the batched triangular solve
TAG=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --months 120 --month-steps 1 --no-kriging-1deg --no-stochastic --no-ellipse-fit --no-notebook --no-fp64-peaks > gpurun_out/${TAG}_launches.log 2>&1
for NU in 2.2 1.5; do
  ncu --set full --clock-control none --import-source on -k regex:ellipse_cov_tile -c 1 -s 1 -f \
      -o gpurun_out/${TAG}_ellipse_full_nu${NU} python tools/profile_ellipse.py --nu $NU --rows 64800 --reps 2 \
      > gpurun_out/${TAG}_ellipse_full_nu${NU}.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:trsm_step_fwd -c 1 -s 10 -f \
    -o gpurun_out/${TAG}_trsm_step python tools/time_kriging.py --months 60 --reps 1 > gpurun_out/${TAG}_trsm_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:potrf_diag -c 1 -s 3 -f \
    -o gpurun_out/${TAG}_potrf_diag python tools/time_kriging.py --months 60 --reps 1 > gpurun_out/${TAG}_potrf_diag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sytrd_kernel -c 1 -s 1 -f \
    -o gpurun_out/${TAG}_sytrd python tools/time_clip.py > gpurun_out/${TAG}_sytrd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:potrf_colupdate -c 1 -s 8 -f \
    -o gpurun_out/${TAG}_potrf_colupdate python tools/time_kriging.py --months 60 --reps 1 > gpurun_out/${TAG}_potrf_colupdate.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ellipse_fit_kernel -c 1 -s 1 -f \
    -o gpurun_out/${TAG}_ellipse_fit python tools/fit_stats.py > gpurun_out/${TAG}_ellipse_fit.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lu_panel_kernel -c 1 -s 10 -f \
    -o gpurun_out/${TAG}_lu_panel python tools/time_lu.py 8000 1 0 > gpurun_out/${TAG}_lu_panel.log 2>&1
