#!/bin/bash
# ncu captures of the ellipse tile kernel (one 1-degree row block): launch list + full set with source.
# usage (under gpurun): tools/ncu_ellipse.sh <tag> <nu> [rows]
set -e
TAG=$1; NU=$2; ROWS=${3:-16384}
mkdir -p gpurun_out
python tools/profile_ellipse.py --nu $NU --rows $ROWS --reps 2 > gpurun_out/${TAG}_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ellipse_cov_tile -c 1 -s 1 \
    -o gpurun_out/${TAG} -f python tools/profile_ellipse.py --nu $NU --rows $ROWS --reps 2 > gpurun_out/${TAG}_ncu.log 2>&1
