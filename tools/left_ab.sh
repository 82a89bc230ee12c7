timeout 900 python -m pytest tests -m gpu -x -q -k "potrf or factor or kriging_gpu or multigpu or stochastic" 2>&1 | tail -3
python tools/time_kriging.py --months 480 --reps 2 2>&1 | grep -v setup | tail -3
python tools/time_kriging.py --months 60 --reps 2 2>&1 | grep -v setup | tail -3
