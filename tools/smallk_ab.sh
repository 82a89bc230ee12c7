timeout 600 python -m pytest tests -m gpu -x -q -k "potrf or factor or gemm or solve or kriging_gpu or batched or multigpu" 2>&1 | tail -3
for K in 0 512 1024 2048 100000; do GMG_SMALL_K=$K python tools/time_kriging.py --months 480 --reps 2 2>&1 | grep -v setup | tail -3; done
GMG_SMALL_K=0 python tools/time_kriging.py --nobs 20000 --res 1 --reps 1 2>&1 | tail -2
GMG_SMALL_K=1024 python tools/time_kriging.py --nobs 20000 --res 1 --reps 1 2>&1 | tail -2
