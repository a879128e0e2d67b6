"""CPU, world_size 2 over gloo: the host-side plumbing of the multi-GPU path — block placement,
the ncclUniqueId hand-off, and that summing per-rank partial consensus sums with an all-reduce
reproduces the oracle's single-process x-bar (framework/LOne.java:395-417)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as O
    from sequnwinder_b200 import parallel
    from helpers import make_problem

    info = parallel.rank_info_from_env()
    assert (info.rank, info.world) == (rank, world)
    # 1. fixed-size byte broadcast (the route the ncclUniqueId takes)
    payload = bytes(range(128)) if rank == 0 else None
    got = parallel.broadcast_bytes(payload, 128)
    assert got == bytes(range(128))

    # 2. placement: every block exactly once, rows as LOne.java:337-357
    T = 5
    mine = parallel.local_blocks(T, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    assert sorted(sum(gathered, [])) == list(range(T))
    assert all(t % world == rank for t in mine)

    # 3. sharded first ADMM iteration: each rank solves its own blocks with the restated solver,
    #    partial sums are all-reduced, x-bar must equal the oracle's single-process result
    p = make_problem("cfg1", 240, max_features=16)
    n, dim = p.data.shape
    C, NN = p.num_classes, p.st.num_nodes
    nb = n // T
    z0 = np.zeros(NN * NN * dim)
    part = np.zeros(C * dim)
    for t in mine:
        rows = np.array(parallel.block_rows(n, T, t))
        ut = np.zeros(NN * NN * dim)
        for e in range(p.st.num_nodes):
            if p.st.is_leaf[e]:
                leaf = int(p.st.node_index[e])
                for par in p.st.parent_idx[p.st.parent_ptr[e]:p.st.parent_ptr[e + 1]]:
                    zo = (leaf * NN + int(par)) * dim
                    ut[zo:zo + dim] = 1.9 * p.x0[leaf * dim:(leaf + 1) * dim] - \
                        1.9 * p.smx0[int(par) * dim:(int(par) + 1) * dim]
        fg = lambda x: O.eval_fg(p.data[rows], p.cls[rows], p.weights[rows], x, p.smx0, z0, ut, 1.7,
                                 p.st, C)
        xt, _, _, _ = O.lbfgs_minimize(fg, p.x0.copy(), eps=0.1)
        part += xt
    tot = torch.from_numpy(part)
    dist.all_reduce(tot)
    xbar = tot.numpy() / T
    xo, _, _ = O.lone_execute(p.x0, p.smx0, p.data, p.cls, p.weights, p.st, C, num_threads=T,
                              admm_max_itr=1, nodes_max_itr=1)
    err = float(np.abs(xbar - xo).max())
    out[rank] = err
    dist.destroy_process_group()


def test_world_size_two_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    assert all(v < 1e-13 for v in out.values()), dict(out)


def test_sweep_values_are_dealt_out_as_replicas():
    from sequnwinder_b200 import parallel
    ridges = [0.1 * 10 ** (i / 5) for i in range(16)]
    seen = []
    for world in (1, 2, 8):
        parts = [parallel.local_ridges(ridges, r, world) for r in range(world)]
        flat = sorted(p for part in parts for p in part)
        assert flat == list(enumerate(ridges))                  # every value exactly once
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
        seen.append(parts)
    with pytest.raises(ValueError):
        parallel.local_ridges(ridges, 2, 2)


def test_local_blocks_errors():
    sys.path.insert(0, ROOT)
    from sequnwinder_b200 import parallel
    assert sorted(parallel.local_blocks(8, 3, 4)) == [3, 7]
    with pytest.raises(ValueError):
        parallel.local_blocks(2, 3, 4)      # rank 3 of 4 owns nothing when T == 2
    assert parallel.block_rows(11, 3, 2) == [2, 5, 8]   # rows 9, 10 are dropped


def test_local_row_indices_match_the_library_convention():
    """rows_local form of sqw_admm_set_problem_counts_dev (include/sequnwinder_b200.h): a rank holds
    the rows i < floor(N/T) T with (i mod T) mod world == rank in ascending order, so row s of block
    t sits at local position s * T_local + t // world - the index the device gather uses."""
    sys.path.insert(0, ROOT)
    from sequnwinder_b200 import parallel
    for (n, T, world) in [(103, 8, 2), (64, 8, 8), (1999999, 8, 4), (50, 5, 3), (40, 8, 1)]:
        nb = n // T
        seen = []
        for rank in range(world):
            try:
                blocks = sorted(parallel.local_blocks(T, rank, world))
            except ValueError:
                assert not any(t % world == rank for t in range(T))
                continue
            idx = parallel.local_row_indices(n, T, rank, world)
            assert np.all(np.diff(idx) > 0) and idx.size == nb * len(blocks)
            TL = len(blocks)
            for t in blocks[:3]:
                for s_ in (0, nb // 2, nb - 1):
                    assert idx[s_ * TL + t // world] == t + T * s_
            seen.append(idx)
        allrows = np.sort(np.concatenate(seen))
        assert np.array_equal(allrows, np.arange(nb * T))          # every kept row exactly once


def test_sharded_column_statistics_combine_to_the_global_plan():
    """The sharded column plan (sqw_prepare_stats_dev per rank -> all-reduce min / max / sums ->
    sqw_prepare_finish_dev): combining per-shard min / max / sum w x / sum w x^2 reproduces the
    statistics Classifier.java:337-372 computes over all rows (host arithmetic of the same
    combination; the device kernels are checked on the GPU)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_problem
    from sequnwinder_b200 import parallel
    p = make_problem("cfg1", 403, max_features=None)
    counts, w = p.counts.astype(np.float64), p.weights
    T, world = 8, 4
    parts = [parallel.local_row_indices(counts.shape[0], T, r, world) for r in range(world)]
    dropped = np.arange((counts.shape[0] // T) * T, counts.shape[0])
    parts[0] = np.concatenate([parts[0], dropped])                 # rank 0 adds the rows the blocking drops
    mn = np.min([counts[i].min(0) for i in parts], axis=0)
    mx = np.max([counts[i].max(0) for i in parts], axis=0)
    swx = sum((w[i, None] * counts[i]).sum(0) for i in parts)
    swxx = sum((w[i, None] * counts[i] ** 2).sum(0) for i in parts)
    tot = sum(w[i].sum() for i in parts)
    keep = np.nonzero(mn != mx)[0]
    mean = swx / tot
    sd = np.sqrt(np.abs(swxx - tot * mean * mean) / (tot - 1))
    assert np.array_equal(keep, p.keep)
    assert np.allclose(mean[keep], p.mean[1:], rtol=1e-12) and np.allclose(sd[keep], p.sd[1:], rtol=1e-11)
