"""Row-sharded FITC M = 20 evaluation under torchrun: ms per evaluation with the in-kernel peer-memory exchange and
with ncclAllReduce between the kernels, at N = 1e4 and N = 1e6 (max over ranks, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from gpscore_b200 import api, synth
from gpscore_b200 import dist as gd

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
theta = synth.hyper_point("P1"); U = synth.inducing_init(20)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n in (10000, 1000000):
    X, y = synth.kin40k_like(n, seed=7)
    lo, hi = gd.row_block(n, rank, world)
    ctx = api.Context(local); ctx.set_stream(stream); ctx.comm_init()
    ctx.set_data(torch.from_numpy(X[lo:hi]).cuda(), torch.from_numpy(y[lo:hi]).cuda())
    for name, p2p in (("peer-memory in-kernel", True), ("ncclAllReduce between kernels", False)):
        act = ctx.comm_transport(p2p)
        for _ in range(10): ctx.fitc_eval_sharded(theta, U, "crps", n)
        dist.barrier(); torch.cuda.synchronize()
        reps = 200
        e0.record(stream)
        for _ in range(reps): v = ctx.fitc_eval_sharded(theta, U, "crps", n)
        e1.record(stream); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier(); torch.cuda.synchronize()
        iters = 300
        ctx.fitc_descend_sharded(theta, U, "crps", n, 1e-4, 1e-4, 20)
        dist.barrier(); torch.cuda.synchronize()
        e0.record(stream)
        ctx.fitc_descend_sharded(theta, U, "crps", n, 1e-4, 1e-4, iters)
        e1.record(stream); torch.cuda.synchronize()
        td = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda"); dist.all_reduce(td, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("N=%d world=%d %-32s descent loop: %.1f us/iteration" % (n, world, name, float(td[0]) * 1e3), flush=True)
        if rank == 0:
            print("N=%d world=%d %-32s (peer memory active: %s): %.1f us/eval  obj %.12g" % (n, world, name, act, float(t[0]) * 1e3, v[0]), flush=True)
    ctx.close()
dist.destroy_process_group()
