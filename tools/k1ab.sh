for rep in 1 2; do
for v in default k1_e k1_f k1_g; do
  if [ $v = default ]; then unset GLOMAR_B200_LIB; else export GLOMAR_B200_LIB=$PWD/variants/$v/libglomar_b200.so; fi
  python tools/time_ellipse.py --nus 2.2 --reps 7
done; done > gpurun_out/r02_k1_ab2.log 2>&1
unset GLOMAR_B200_LIB
GLOMAR_B200_LIB=$PWD/variants/k1_e/libglomar_b200.so python tests/manual/ellipse_accuracy.py >> gpurun_out/r02_k1_ab2.log 2>&1
cat gpurun_out/r02_k1_ab2.log
