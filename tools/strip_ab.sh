timeout 600 python -m pytest tests/test_multigpu_paths_gpu.py -q -k "strip or overlapped" 2>&1 | tail -3
for W in 128 64 56 auto; do if [ $W = auto ]; then unset GMG_STRIP; else export GMG_STRIP=$W; fi; python tools/time_rank_solve.py; done
unset GMG_STRIP
for W in 64 56; do GMG_STRIP=$W python tools/time_rank_solve.py --cols 32402; done
