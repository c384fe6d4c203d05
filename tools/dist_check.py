"""Partitioned solve vs single-GPU solve of the same graph (run under torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_check.py [laps] [poses_per_lap] [lm_iters] [preconditioner 0|1|2] [sphere|manhattan]
Prints "DIST_CHECK PASS" on rank 0 when chi2 histories agree to 1e-9 relative and the estimates to 1e-8.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import sim3opt_b200 as s3
from sim3opt_b200 import synth


def main():
    laps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    precond = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    graph = sys.argv[5] if len(sys.argv) > 5 else "sphere"
    vtol = float(sys.argv[6]) if len(sys.argv) > 6 else 1e-8
    g = synth.sphere(laps, per, seed=11) if graph == "sphere" else synth.manhattan3d(laps * per, seed=11)
    box = [s3.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)

    def make(comm):
        p = s3.Problem(s3.KIND_SIM3, device=local)
        p.set_math_mode(s3.MATH_CORRECTED)
        p.set_pcg(1e-12, 20000)
        p.set_preconditioner(precond)
        if comm:
            p.set_comm(rank, world, box[0])
        p.set_vertices(g["est"], g["fixed"])
        p.set_edges(g["v0"], g["v1"], g["meas"], g["info"])
        return p

    pd = make(True)
    chi0 = pd.chi2()
    n, chi2, lam, hist = pd.optimize(iters)
    vd = pd.vertices()
    ok = True
    if rank == 0:
        ps = make(False)
        chi0s = ps.chi2()
        ns, chi2s, lams, hists = ps.optimize(iters)
        vs = ps.vertices()
        rel = np.abs(hist[:, 0] - hists[:, 0]) / hists[:, 0]
        dv = np.abs(vd - vs).max()
        ok = (abs(chi0 - chi0s) <= 1e-12 * chi0s and n == ns and rel.max() <= 1e-9 and dv <= vtol
              and np.array_equal(hist[:, 2], hists[:, 2]))
        print("chi2_0 dist %.12g single %.12g" % (chi0, chi0s))
        print("chi2 history rel diff", rel)
        print("pcg iters dist", hist[:, 4], "single", hists[:, 4])
        print("max |vertex diff| %.3e" % dv)
        print("DIST_CHECK", "PASS" if ok else "FAIL", "world", world)
        print("P2P_HALO", pd.stats()["p2p_halo"], "MULTILEVEL_LEVELS", pd.stats()["multilevel_levels"])
    # s3o_update in the partitioned solve: every rank passes its OWNED rows of one global step, the step is
    # all-gathered inside, and all ranks must end with the estimates a single-GPU s3o_update produces
    nf = int((np.asarray(g["fixed"]) == 0).sum())
    seg = -(-nf // world)
    rng = np.random.default_rng(5)
    step = 1e-3 * rng.standard_normal((nf, 7))
    lo, hi = min(nf, rank * seg), min(nf, (rank + 1) * seg)
    pd.update(step[lo:hi])
    vu = pd.vertices()
    upd_ok = True
    if rank == 0:
        dv = np.abs(vd - vs).max()
        ps.update(step)
        du = np.abs(vu - ps.vertices()).max()
        upd_ok = du <= dv + 1e-10          # the update adds nothing to the difference the solves left
        print("UPDATE_PARTITIONED max diff %.3e %s" % (du, "PASS" if upd_ok else "FAIL"))
        ok = ok and upd_ok
    vd = vu
    # sharded host round trip: every rank writes only ITS slice of a new estimate vector (the rest of its host copy is
    # garbage on purpose), after the collective every rank must hold all of it, and read its own slice back unchanged
    first, count = pd.estimate_slice()
    target = g["est"] + 0.0
    target[:, 4:7] += 0.01 * np.arange(len(target))[:, None]
    mine = np.ascontiguousarray(target[first:first + count])
    pd.set_estimates_slice(mine)
    back = np.zeros_like(mine)
    pd.vertices_slice(back)
    full = pd.vertices()
    slice_ok = bool(np.array_equal(back, mine) and np.array_equal(full, target))
    sl = torch.tensor([1.0 if slice_ok else 0.0], device="cuda")
    dist.all_reduce(sl, op=dist.ReduceOp.MIN)
    if rank == 0:
        covered = first == 0 and count == -(-len(target) // world)
        print("ESTIMATE_SLICES", "PASS" if (sl.item() == 1.0 and covered) else "FAIL")
        ok = ok and sl.item() == 1.0 and covered
    vd = full
    # every rank holds the same estimates after the solve, the update and the sliced upload
    t = torch.from_numpy(vd.copy()).cuda()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    same = bool((t == ref).all().item())
    flags = torch.tensor([1.0 if same else 0.0], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("ESTIMATES_IDENTICAL_ACROSS_RANKS", bool(flags.item() == 1.0))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
