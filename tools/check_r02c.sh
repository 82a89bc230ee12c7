# one call: month-batch A/B of a kernel variant (variants/old = the previous build) and, with the new build, the whole GPU
# suite and a bench line
GLOMAR_B200_LIB=variants/old/libglomar_b200.so python tools/months_ab.py old && python tools/months_ab.py new
timeout 120 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 150 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo bench rc=$?
python tools/show_bench.py gpurun_out/r02c_bench_n1.json
