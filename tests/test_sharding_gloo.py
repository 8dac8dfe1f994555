"""CPU, world_size 2 over gloo: the N>1 host logic of the point-sharded MSM (shard ranges, the
all-gather of per-rank result records, the final sum).  The per-shard MSM and the final addition
are done by the ORACLE here (no GPU in this container); on the GPU box bench.py runs the same
helpers with the CUDA path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from blaze_b200.sharding import gather_records, shard_range
    from oracle import capi
    from oracle.py import curves, ec
    from util import chain_points, random_scalars
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = curves.BLS12_381
    pts, p0, qq = chain_points(c, n, seed=5)
    sc = random_scalars(c, n, seed=6)
    first, cnt = shard_range(n, rank, world)
    part = capi.msm_pippenger("BLS12_381", pts[first * 96:(first + cnt) * 96], sc[first * 32:(first + cnt) * 32], cnt)
    recs = gather_records(dist, part)
    acc = None
    for r in recs:
        acc = ec.add(c, acc, ec.decode_result(c, r))
    full = capi.msm_pippenger("BLS12_381", pts, sc, n)
    ok = ec.encode_result(c, acc) == full and len(recs) == world
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok, first, cnt))


def test_point_sharded_msm_world2():
    world, n = 2, 1001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res)
    assert [(f, c) for _, _, f, c in res] == [(0, 501), (501, 500)]


def test_shard_range_partitions():
    from blaze_b200.sharding import shard_range
    for n in (0, 1, 7, 64, 1001):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1


def _ntt_worker(rank, world, port, log_n, q):
    """Host logic of bench.py's distributed-NTT verification on 2 gloo ranks: the per-index element generator is the
    same stream on torch and numpy, the ranks' column slabs tile the input exactly once, and the (k2, t) -> global
    output index map of the rank blocks X[(h T + t) + N1 k2] tiles the output exactly once."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1 << log_n
    l1 = log_n // 3                      # any split with both factors >= world works for the index maps
    N1, N2 = 1 << l1, 1 << (log_n - l1)
    C, T = N2 // world, N1 // world
    seed = 31
    r = torch.arange(N1, dtype=torch.int64)
    idx = (r[:, None] * N2 + rank * C + torch.arange(C, dtype=torch.int64)[None, :]).reshape(-1)
    slab = bench.elems_torch(torch, idx, seed)
    full = torch.from_numpy(bench.elems_numpy(n, seed)).view(n, 32)
    ok = bool(torch.equal(slab, full[idx]))                      # torch stream == numpy stream, at the slab's indices
    seen = torch.zeros(n, dtype=torch.int32)
    seen[idx] += 1
    k2 = torch.arange(N2, dtype=torch.int64)
    out_idx = (rank * T + torch.arange(T, dtype=torch.int64)[None, :] + N1 * k2[:, None]).reshape(-1)
    seen_out = torch.zeros(n, dtype=torch.int32)
    seen_out[out_idx] += 1
    dist.all_reduce(seen)
    dist.all_reduce(seen_out)
    ok = ok and bool((seen == 1).all()) and bool((seen_out == 1).all())
    ok = ok and bool((full[:, 31] < 0x40).all())                 # canonical: < 2^254 < r
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok))


def test_distributed_ntt_index_maps_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ntt_worker, args=(r, world, port, 12, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res)
