"""Multi-GPU parity check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py

Every rank holds a contiguous block of rows; the row-sharded FITC evaluation (library NCCL all-reduces between the
fused passes, and the matrix form for M > 32) must equal the single-GPU evaluation of the whole problem that rank 0
runs beside it, and the CPU oracle.  Also: dist.ShardedFitc over api.Context with torch.distributed doing the
all-reduces (the staged protocol — the advisor's stream-ordering case), sharded prediction + metrics, the sharded
grid.  Exits non-zero on any mismatch; rank 0 prints one line per check."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gpscore_b200 import api, synth
    from gpscore_b200 import dist as gd
    from oracle import gp_oracle as O
    from oracle import woodbury as W

    fails = []

    def check(name, err, tol):
        ok = err <= tol
        if rank == 0:
            print("%-58s err %.3e tol %.0e %s" % (name, err, tol, "ok" if ok else "FAIL"), flush=True)
        if not ok:
            fails.append(name)

    n = 20011                                  # ragged: blocks differ by one row
    X, y = synth.kin40k_like(n, seed=3)
    theta = synth.hyper_point("P1")
    lo, hi = gd.row_block(n, rank, world)
    part = api.Context(local)
    part.comm_init()
    part.set_data(torch.from_numpy(X[lo:hi]).cuda(), torch.from_numpy(y[lo:hi]).cuda())
    whole = api.Context(local)
    whole.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
    rng = np.random.default_rng(4)
    p2p = part.comm_transport(True)
    if rank == 0:
        print("peer-memory transport available: %s" % p2p, flush=True)
    for m_ind in (20, 31, 64):
        U = X[rng.choice(n, m_ind, replace=False)] + 0.05 * rng.standard_normal((m_ind, 8))
        for score in ("crps", "logs", "nlml"):
            part.comm_transport(False)
            sv, sg, sgu = part.fitc_eval_sharded(theta, U, score, n)
            v1, g1, gu1 = whole.fitc_eval(theta, U, score)
            check("M=%d %s library-NCCL sharded vs single GPU" % (m_ind, score),
                  max(abs(sv - v1) / abs(v1), rel(sg, g1), rel(sgu, gu1)), 1e-10)
            if p2p and m_ind <= 31:
                part.comm_transport(True)
                for rep in range(3):      # consecutive exchanges alternate parity
                    pv, pg, pgu = part.fitc_eval_sharded(theta, U, score, n)
                check("M=%d %s in-kernel peer-memory exchange vs single GPU" % (m_ind, score),
                      max(abs(pv - v1) / abs(v1), rel(pg, g1), rel(pgu, gu1)), 1e-10)
                ov_only = part.fitc_eval_sharded(theta, U, score, n)[0]
                check("M=%d %s peer-memory exchange repeatable" % (m_ind, score), abs(ov_only - pv), 0.0)
            if m_ind == 20:
                ov, og, ogu = W.fitc_obj_grad(X, y, U, theta, O.SCORES[score])[:3]
                check("M=%d %s library-NCCL sharded vs oracle (objective)" % (m_ind, score), abs(sv - ov) / abs(ov), 1e-8)
                check("M=%d %s library-NCCL sharded vs oracle (gradient)" % (m_ind, score), max(rel(sg, og), rel(sgu, ogu)), 1e-6)
    # block objectives (4-fold DSS, kc): folds are ranges of the GLOBAL row order and straddle the ranks' uneven blocks
    n4 = 20012
    X4, y4 = synth.kin40k_like(n4, seed=5)
    cuts = [0] + [gd.row_block(n4, r, world)[1] + 37 for r in range(world - 1)] + [n4]
    part.set_data(torch.from_numpy(X4[cuts[rank]:cuts[rank + 1]]).cuda(), torch.from_numpy(y4[cuts[rank]:cuts[rank + 1]]).cuda())
    whole.set_data(torch.from_numpy(X4).cuda(), torch.from_numpy(y4).cuda())
    for m_ind in (20, 64):
        U = X4[rng.choice(n4, m_ind, replace=False)] + 0.05 * rng.standard_normal((m_ind, 8))
        for kind in ("dss", "kc"):
            sv, sg, sgu = part.fitc_eval_sharded(theta, U, kind, n4)
            v1, g1, gu1 = whole.fitc_eval(theta, U, kind)
            check("M=%d %s block objective row-sharded vs single GPU" % (m_ind, kind),
                  max(abs(sv - v1) / abs(v1), rel(sg, g1), rel(sgu, gu1)), 1e-9)
            if m_ind == 20:
                ov, og, ogu = W.fitc_block_obj_grad(X4, y4, U, theta, kind)
                check("M=%d %s block objective row-sharded vs oracle" % (m_ind, kind),
                      max(abs(sv - ov) / abs(ov) * 100, rel(sg, og), rel(sgu, ogu)), 1e-6)
    part.set_data(torch.from_numpy(X[lo:hi]).cuda(), torch.from_numpy(y[lo:hi]).cuda())
    whole.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
    part.comm_transport(True)
    # every rank got the same numbers
    U = synth.inducing_init(20)
    sv, sg, sgu = part.fitc_eval_sharded(theta, U, "crps", n)
    t = torch.tensor(np.concatenate([[sv], sg, sgu.ravel()]), device="cuda")
    tmin, tmax = t.clone(), t.clone()
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    check("all ranks return identical results", float((tmax - tmin).abs().max()), 0.0)
    # row-sharded device-resident descent loop == single-GPU loop
    for p2pflag in ((True, False) if p2p else (False,)):
        part.comm_transport(p2pflag)
        th_s, U_s, tr_s = part.fitc_descend_sharded(theta, U, "crps", n, 0.3, 0.3, 15)
        th_1, U_1, tr_1 = whole.fitc_descend(theta, U, "crps", 0.3, 0.3, 15)
        check("sharded descent loop (peer memory: %s) vs single GPU" % p2pflag,
              max(rel(th_s, th_1), rel(U_s, U_1), rel(tr_s, tr_1)), 1e-10)
    part.comm_transport(True)
    # staged protocol with torch.distributed all-reduces on torch's stream (no synchronize in the callback)
    sf = gd.ShardedFitc(part, n)
    for score in ("crps", "nlml"):
        v2, g2, gu2 = sf.eval(theta, U, score)
        v1, g1, gu1 = whole.fitc_eval(theta, U, score)
        check("%s dist.ShardedFitc(api.Context) staged + torch all-reduce" % score,
              max(abs(v2 - v1) / abs(v1), rel(g2, g1), rel(gu2, gu1)), 1e-10)
    sfl = gd.ShardedFitc(part, n, library_comm=True)
    v3, g3, gu3 = sfl.eval(theta, U, "crps")
    check("dist.ShardedFitc(library_comm=True)", max(abs(v3 - sv) / abs(sv), rel(g3, sg)), 0.0)
    # prediction + scoring by test rows
    Xt, yt, Xs, ys = synth.kin40k_like(1500, 4001, seed=9)
    whole.set_data(torch.from_numpy(Xt).cuda(), torch.from_numpy(yt).cuda())
    Xsd, ysd = torch.from_numpy(Xs).cuda(), torch.from_numpy(ys).cuda()
    for label, uarg in (("full", None), ("fitc", U)):
        met = gd.sharded_predict_metrics(whole, theta, Xsd, ysd, inducing_x=uarg)
        if uarg is None:
            m1, v1 = whole.full_predict(theta, Xsd)
        else:
            m1, v1 = whole.fitc_predict(theta, uarg, Xsd)
        single = whole.test_metrics(m1, v1, ysd)
        check("%s prediction + metrics split by test rows vs single GPU" % label,
              max(abs(met[k] - single[k]) / max(abs(single[k]), 1e-300) for k in single), 1e-10)
    # grid round-robin
    xg = np.linspace(-6, 6, 20)
    yg = np.sin(xg)
    ls = np.repeat(np.linspace(0.05, 2.0, 16), 16)
    sd = np.tile(np.linspace(0.05, 1.0, 16), 16)
    xgd, ygd = torch.from_numpy(xg).cuda(), torch.from_numpy(yg).cuda()
    surf = gd.sharded_grid(lambda l_, s_: whole.grid_eval(xgd, ygd, l_, s_, "crps"), ls, sd, device="cuda")
    ref = whole.grid_eval(xgd, ygd, ls, sd, "crps")
    check("grid dealt round-robin vs single GPU", rel(surf, ref), 0.0)
    part.close()
    whole.close()
    dist.barrier()
    dist.destroy_process_group()
    if fails:
        print("rank %d FAILED: %s" % (rank, fails), flush=True)
        return 1
    if rank == 0:
        print("mgpu_check: all checks passed on %d GPUs" % world, flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
