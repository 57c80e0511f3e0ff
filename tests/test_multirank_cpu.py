"""world_size-2 `gloo` test of the multi-rank path's host logic (no GPU):
(1) every rank regenerates exactly its own contiguous slice of the global particle set
    from the counter-based IC stream;
(2) coefficients are additive over shards, so summing per-rank vectors with one
    all-reduce reproduces the single-rank result (the reference's MPI_Allreduce,
    src/SphericalBasis.cc:864-903, replaced on the GPU by one NCCL all-reduce);
(3) forces computed per shard from the reduced coefficients equal the global ones;
(4) the centring pass (Component::fix_positions): per-level {mass, m*pos, m*vel, m*acc} rows are
    additive over shards, reduced the way the peer-memory kernel does it -- every rank gathers
    all ranks' vectors and adds them in rank order, so the result is bitwise the same on
    every rank -- and give the single-rank centre of mass.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_per, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import get_table
    from exp_b200 import ics
    from oracle import oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tab = get_table("hernquist", 2, 8, numr=400)
    pos, vel, mass = ics.hernquist_halo(n_per * world, lo=rank * n_per, hi=(rank + 1) * n_per, seed=5)
    pos = pos + 0.01
    g = O.Grid(tab)
    P = O.sph_params(rmin=tab.rmin, rmax=tab.rmax)
    S = O.ParticleSet(pos, vel, mass)
    coef, used = O.sph_coefficients(g, P, S)
    t = torch.from_numpy(coef.copy())
    dist.all_reduce(t)                       # SUM
    u = torch.tensor([used])
    dist.all_reduce(u)
    O.sph_accel(g, P, S, t.numpy())
    # (4) centring sums with levels; rank-ordered reduction (what peer_allreduce_kernel does)
    multistep = 2
    S.level[:] = (np.arange(rank * n_per, (rank + 1) * n_per) % (multistep + 1)).astype(np.uint32)
    lev, _ = O.fix_positions(S, 0, multistep)
    parts = [torch.zeros_like(torch.from_numpy(lev)) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(lev.copy()))
    tot = np.zeros_like(lev)
    for q in range(world):
        tot += parts[q].numpy()
    np.savez(os.path.join(out, f"rank{rank}.npz"), pos=pos, coef=t.numpy(), used=u.numpy(),
             acc=S.acc, pot=S.pot, com_lev=tot)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import get_table, rel_rms
    from exp_b200 import ics
    from oracle import oracle as O

    world, n_per = 2, 3000
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_per, str(tmp_path)), nprocs=world, join=True)

    tab = get_table("hernquist", 2, 8, numr=400)
    pos, vel, mass = ics.hernquist_halo(n_per * world, seed=5)
    pos = pos + 0.01
    g = O.Grid(tab)
    P = O.sph_params(rmin=tab.rmin, rmax=tab.rmax)
    S = O.ParticleSet(pos, vel, mass)
    coef, used = O.sph_coefficients(g, P, S)
    O.sph_accel(g, P, S, coef)
    r = [np.load(os.path.join(tmp_path, f"rank{k}.npz")) for k in range(world)]
    assert np.array_equal(np.concatenate([r[0]["pos"], r[1]["pos"]]), pos)      # (1)
    assert r[0]["coef"].tobytes() == r[1]["coef"].tobytes()                      # same on all ranks
    assert rel_rms(r[0]["coef"], coef) < 1e-13 and int(r[0]["used"][0]) == used  # (2)
    acc = np.concatenate([r[0]["acc"], r[1]["acc"]])
    an = np.linalg.norm(S.acc, axis=1)
    assert (np.linalg.norm(acc - S.acc, axis=1) / an).max() < 1e-11              # (3)
    # (4)
    multistep = 2
    S.level[:] = (np.arange(n_per * world) % (multistep + 1)).astype(np.uint32)
    lev, out = O.fix_positions(S, 0, multistep)
    assert r[0]["com_lev"].tobytes() == r[1]["com_lev"].tobytes()                # rank-ordered sum
    scale = np.abs(lev).max() + 1e-300
    assert np.abs(r[0]["com_lev"] - lev).max() <= 1e-13 * scale
    tot = r[0]["com_lev"].sum(axis=0)
    assert np.abs(tot[1:] / tot[0] - out[1:]).max() <= 1e-12 * max(1.0, np.abs(out[1:]).max())
