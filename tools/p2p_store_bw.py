"""torchrun (2 ranks): bandwidth of SM stores into a peer's memory through an IPC mapping."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from glomargridding_b200.ellipse.p2p import PeerRowBlocks, _CudaBlock
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 32768
pb = PeerRowBlocks(N)
peer = (rank + 1) % world
rows = int(pb.blocks[peer + 1] - pb.blocks[peer])
remote = torch.as_tensor(_CudaBlock(pb.ptrs[peer], (rows, N), "<f8"), device=torch.device("cuda", local))
src = torch.randn((rows, N), dtype=torch.float64, device="cuda")
def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n
gb = rows * N * 8 / 1e9
res = {
  "fill_remote (SM stores, 16 B/lane)": gb / timeit(lambda: remote.fill_(1.5)),
  "fill_local": gb / timeit(lambda: pb.local.fill_(1.5)),
  "copy_ local->remote (copy engine)": gb / timeit(lambda: remote.copy_(src)),
  "add_ into remote (SM load+store remote)": gb / timeit(lambda: remote.add_(1.0)),
  "transpose-store remote (t().contiguous kernel)": (N * min(rows, N) * 8 / 1e9) / timeit(lambda: remote[:, :rows].copy_(src[:, :rows].t())),
}
for w in (64, 256, 1024, 4096):
    nb = N // w
    def f(w=w, nb=nb):
        for b in range(0, nb, max(1, nb // 16)):
            remote[:, b * w:(b + 1) * w].fill_(2.5)
    calls = len(range(0, nb, max(1, nb // 16)))
    res[f"strided fill remote, {w * 8} B per row per call"] = (rows * w * 8 * calls / 1e9) / timeit(f, 3)
    def g(w=w, nb=nb):
        for b in range(0, nb, max(1, nb // 16)):
            pb.local[:, b * w:(b + 1) * w].fill_(2.5)
    res[f"strided fill local,  {w * 8} B per row per call"] = (rows * w * 8 * calls / 1e9) / timeit(g, 3)
pb.barrier()
if rank == 0:
    for k, v in res.items(): print(f"{k}: {v:.1f} GB/s")
pb.close()
dist.destroy_process_group()
