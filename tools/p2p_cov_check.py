"""torchrun script: peer-to-peer symmetric covariance build vs the plain row-block build (bit-equal), and timing.

    torchrun --nproc-per-node G tools/p2p_cov_check.py [--res 1]
"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from glomargridding_b200 import synthetic as synth
from glomargridding_b200._lib import check, lib
from glomargridding_b200.ellipse import EllipseCovarianceBuilder
from glomargridding_b200.ellipse.p2p import PeerRowBlocks

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=float, default=1.0)
ap.add_argument("--nus", default="1.5,2.2")
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

out = {"world": world, "res": a.res}
for res_deg, check_equal in ((10.0, True), (a.res, True)):
    lats, lons = synth.global_grid(res_deg)
    Lx, Ly, th, sd = synth.ellipse_fields(lats, lons, land_mask=res_deg == 10.0)
    N = int((~np.ma.getmaskarray(Lx)).sum())
    pb = PeerRowBlocks(N)
    if os.environ.get("P2P_FAKE"):       # experiment: mirrored tiles go to a LOCAL scratch block instead of the peer
        import ctypes
        scratch = [torch.empty((int(pb.blocks[g + 1] - pb.blocks[g]), N), dtype=torch.float64, device="cuda") for g in range(world)]
        for g in range(world):
            if g != rank:
                pb.ptrs[g] = scratch[g].data_ptr()
        check_equal = False
    for nu in [float(x) for x in a.nus.split(",")]:
        b = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=nu, precision=np.float64, peer_blocks=pb)
        plain = EllipseCovarianceBuilder(Lx, Ly, th, sd, lats, lons, v=nu, precision=np.float64,
                                         rows=(pb.row0, pb.nrows))
        same = bool(torch.equal(b.cov_ns_device, plain.cov_ns_device))
        flag = torch.tensor([int(same)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        assert int(flag.item()) == 1 or os.environ.get("P2P_FAKE"), f"p2p build differs from the row-block build (res {res_deg}, nu {nu}, rank {rank})"
        if res_deg == a.res:
            L, s = lib(), torch.cuda.current_stream().cuda_stream
            tp, tn = [], []
            for rep in range(a.reps):
                pb.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                check(L.gmg_ellipse_cov_p2p_f64(b._ws.data_ptr(), N, nu, -1.0, 0, pb.world, pb.rank, pb.block_row0,
                                                pb.ptrs, N, 0, s), "p2p")
                e1.record(); e1.synchronize()
                t = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
                tp.append(float(t.item()))
                pb.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                check(L.gmg_ellipse_cov_f64(b._ws.data_ptr(), N, pb.row0, pb.nrows, nu, -1.0, 0,
                                            plain.cov_ns_device.data_ptr(), N, 0, s), "plain")
                e1.record(); e1.synchronize()
                t = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
                tn.append(float(t.item()))
            out[f"nu{nu}"] = {"p2p_ms": float(np.median(tp)), "row_block_ms": float(np.median(tn)),
                              "p2p_entries_per_s": N * N / np.median(tp) * 1e3}
        del b, plain
    pb.close()
    del pb
    torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
