"""N > 1 host logic on CPU: world_size-2 gloo run of the sharding + all-gather + un-permute driver
(tlc_b200.multi), with the oracle standing in for the per-rank CUDA compute.  Every rank must end up with
the same [E, 25] table as a single-process run, in the caller's target order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as orc
from tlc_b200 import graphgen as gg
from tlc_b200 import multi


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    c = gg.make_config("pubmed", scale=0.05)
    labels, ne = gg.relabel_first_appearance(c["edges"])
    csr = gg.build_csr(len(labels), ne, c["kappa"])
    rng = np.random.default_rng(3)
    tg = ne[rng.choice(len(ne), 101, replace=False)].astype(np.int32)   # odd count: shards are ragged
    tg = np.concatenate([tg, np.array([[-1, 2], [5, 5]], np.int32)])
    return csr, tg


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    csr, tg = _problem()
    og = orc.OracleGraph(*csr)

    def local_fn(tshard):
        r = og.run_batch(tshard, hop=2, flags=orc.F_NORM)
        return torch.from_numpy(r["pi"].astype(np.float32)), torch.from_numpy(r["status"])

    sv = multi.ShardedVicinity(csr[0], local_fn, torch.device("cpu"))
    pi, st = sv.compute(tg)
    np.save(out_path % rank, np.concatenate([pi.numpy(), st.numpy()[:, None].astype(np.float32)], 1))
    dist.barrier()
    dist.destroy_process_group()


def test_plan_shards_partition():
    csr, tg = _problem()
    for world in (1, 2, 3, 8):
        order, L = multi.plan_shards(tg, csr[0], world)
        assert sorted(order.tolist()) == list(range(len(tg)))
        parts = [multi.shard_of(order, r, world) for r in range(world)]
        assert sum(len(p) for p in parts) == len(tg) and max(len(p) for p in parts) <= L
        g = np.zeros((world, L, 3), np.float32)
        for r, p in enumerate(parts):
            g[r, : len(p)] = tg[p].sum(axis=1)[:, None]
        assert np.array_equal(multi.unshard(g, order, world, len(tg))[:, 0], tg.sum(axis=1).astype(np.float32))


@pytest.mark.timeout(300)
def test_world2_gloo_allgather(tmp_path):
    out = str(tmp_path / "rank%d.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    csr, tg = _problem()
    ref = orc.OracleGraph(*csr).run_batch(tg, hop=2, flags=orc.F_NORM)
    for r in range(2):
        a = np.load(out % r)
        assert np.array_equal(a[:, :25], ref["pi"].astype(np.float32))
        assert np.array_equal(a[:, 25].astype(np.uint8), ref["status"])
