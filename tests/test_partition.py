"""CPU tests of the multi-GPU host logic: the vertex-range partition plan.

The N>1 path is exercised with a real world_size-2 gloo group: each rank computes its own plan through
the C ABI, the ranks exchange their halo lists and check that they agree.
"""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _plans(g, world):
    from sim3opt_b200 import api
    n = len(g["est"])
    return [api.host_partition(n, g["fixed"], g["v0"], g["v1"], r, world) for r in range(world)]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_covers_graph(sphere_small, world):
    g = sphere_small
    plans = _plans(g, world)
    nf = int((g["fixed"] == 0).sum())
    assert sum(p["n_own"] for p in plans) == nf
    # every active edge has exactly one primary rank
    assert sum(p["n_primary"] for p in plans) == len(g["v0"])
    seg = -(-nf // world)
    for r, p in enumerate(plans):
        assert p["n_own"] == max(0, min(nf, (r + 1) * seg) - min(nf, r * seg))
        assert np.all(np.diff(p["ghosts"]) > 0)
        assert np.all((p["ghosts"] // seg) != r)                      # a ghost is never owned
        assert p["recv_count"].sum() == p["n_ghost"] and p["recv_count"][r] == 0
    # what q receives from r is exactly what r sends to q, in the same order
    for r in range(world):
        off = np.concatenate([[0], np.cumsum(plans[r]["send_count"])])
        for q in range(world):
            sent = plans[r]["send_idx"][off[q]:off[q + 1]]
            recv = plans[q]["ghosts"][(plans[q]["ghosts"] // seg) == r]
            assert np.array_equal(sent, recv)
    if world == 1:
        assert plans[0]["n_ghost"] == 0 and plans[0]["n_local_edges"] == len(g["v0"])


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from sim3opt_b200 import api, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = synth.sphere(n_laps=10, poses_per_lap=30, seed=3)
        n = len(g["est"])
        plan = api.host_partition(n, g["fixed"], g["v0"], g["v1"], rank, world)
        # exchange ghost lists and send lists
        objs = [None] * world
        dist.all_gather_object(objs, (plan["ghosts"].tolist(), plan["send_idx"].tolist(), plan["send_count"].tolist(),
                                      plan["n_own"], plan["n_primary"]))
        nf = int((g["fixed"] == 0).sum())
        seg = -(-nf // world)
        ok = sum(o[3] for o in objs) == nf and sum(o[4] for o in objs) == len(g["v0"])
        for r in range(world):
            off = np.concatenate([[0], np.cumsum(objs[r][2])])
            sent_to_me = objs[r][1][off[rank]:off[rank + 1]]
            mine = [x for x in plan["ghosts"].tolist() if x // seg == r]
            ok = ok and sent_to_me == mine
        # a partitioned product y = (H + lambda I) x over the plan's halo lists: every rank holds its own rows
        # of H (from the oracle's linearisation) and its own segment of x, receives exactly its ghosts, and the
        # gathered result equals the global product -- i.e. the ghosts cover every column the owned rows touch
        from oracle import oracle as orc
        p = orc.Problem(orc.KIND_SIM3)
        p.set_vertices(g["est"], g["fixed"])
        p.set_edges(g["v0"], g["v1"], g["meas"], g["info"])
        p.set_jacobian_mode(orc.JAC_ANALYTIC)
        colptr, rowidx = p.build_structure()
        H, _b = p.linearize()
        d = 7
        A = np.zeros((nf * d, nf * d))
        for c in range(nf):
            for k in range(colptr[c], colptr[c + 1]):
                r = rowidx[k]
                A[r * d:(r + 1) * d, c * d:(c + 1) * d] = H[k]
                A[c * d:(c + 1) * d, r * d:(r + 1) * d] = H[k].T
        A += 0.3 * np.eye(nf * d)
        x = np.random.default_rng(1).normal(size=(nf, d))
        lo, hi = min(nf, rank * seg), min(nf, (rank + 1) * seg)
        known = {i: x[i] for i in range(lo, hi)}                       # what this rank may read
        off = np.concatenate([[0], np.cumsum(plan["send_count"])])
        reqs, bufs = [], {}
        for peer in range(world):
            if peer == rank:
                continue
            send = plan["send_idx"][off[peer]:off[peer + 1]]
            if len(send):
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(x[send])), peer))
            gh = [gi for gi in plan["ghosts"].tolist() if gi // seg == peer]
            if gh:
                bufs[peer] = (gh, torch.zeros((len(gh), d), dtype=torch.float64))
                reqs.append(dist.irecv(bufs[peer][1], peer))
        for rq in reqs:
            rq.wait()
        for peer, (gh, buf) in bufs.items():
            for gi, row in zip(gh, buf.numpy()):
                known[gi] = row
        y_own = np.zeros((hi - lo, d))
        for i in range(lo, hi):
            cols = np.nonzero(np.abs(A[i * d:(i + 1) * d]).reshape(d, nf, d).max(axis=(0, 2)))[0]
            for j in cols:
                y_own[i - lo] += A[i * d:(i + 1) * d, j * d:(j + 1) * d] @ known[int(j)]     # KeyError = a missing ghost
        ys = [None] * world
        dist.all_gather_object(ys, y_own)
        ok = ok and np.abs(np.concatenate(ys).reshape(-1) - A @ x.reshape(-1)).max() <= 1e-9 * np.abs(A @ x.reshape(-1)).max()
        # the allreduce every PCG iteration relies on: a sum of per-rank partials
        t = torch.tensor([float(plan["n_primary"])], dtype=torch.float64)
        dist.all_reduce(t)
        ok = ok and int(t.item()) == len(g["v0"])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_partition_world2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
