# 8-GPU A/B of the forward-solve strip width inside the distributed 1-degree kriging (run under gpurun --gpus 8)
for s in 0 64 0 64; do
  echo "GMG_STRIP=$s"
  GMG_STRIP=$s timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 8 --steps 5 --warmup 3 --no-kriging --no-stochastic --no-ellipse-fit --no-cpu-baseline --no-fp64-peaks 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kriging_1deg']; print('  seconds', k['seconds'], 'e2e', (k.get('e2e') or {}).get('seconds'))"
done
