"""world_size-2 gloo tests of the N>1 host path (CPU): the sharding rule of the C ABI and the single integer tally
reduce.  The tallies fed to the reduce come from the CPU oracle run on each rank's shard (tests may use the oracle);
the sum over ranks must equal the unsharded run bit for bit once quantised to the device's fixed-point format."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_voxel_slabs_tile_the_grid():
    """the slab rule of the fused exchange, the sharded upload and the sharded read-out: disjoint, ordered, complete."""
    from opendxmc_b200 import distributed as D
    for n, size in [(78_643_200, 8), (78_643_200, 2), (7_161_276, 4), (17, 8), (5, 8), (1, 1)]:
        edges = [D.slab(n, r, size) for r in range(size)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        for (b0, e0), (b1, e1) in zip(edges[:-1], edges[1:]):
            assert b0 <= e0 == b1 <= e1
        assert max(e - b for b, e in edges) - min(e - b for b, e in edges) <= 1


def test_shard_rule_partitions_all_histories():
    from opendxmc_b200 import distributed as D
    for n_total, world in [(1, 1), (65536, 2), (65537, 2), (1_000_003, 3), (10 * 65536 + 17, 8), (200_000, 4)]:
        seen = np.zeros(n_total, dtype=np.uint8)
        for rank in range(world):
            cnt = D.shard_local_count(n_total, rank, world)
            assert cnt % 65536 == 0
            loc = np.arange(0, cnt, 4099, dtype=np.uint64)
            loc = np.unique(np.concatenate([loc, np.arange(min(cnt, 70000), dtype=np.uint64), np.arange(max(0, cnt - 70000), cnt, dtype=np.uint64)]))
            ids = np.array([D.shard_history_id(int(v), rank, world) for v in loc[:2000]], dtype=np.uint64)
            # vectorised restatement for the full check
            allloc = np.arange(cnt, dtype=np.uint64)
            allids = (allloc // 65536 * world + rank) * 65536 + allloc % 65536
            assert np.array_equal(ids, allids[loc[:2000].astype(np.int64)])
            valid = allids[allids < n_total]
            assert not seen[valid.astype(np.int64)].any()
            seen[valid.astype(np.int64)] = 1
        assert seen.all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import opendxmc_b200 as dx
    from opendxmc_b200 import distributed as D
    from oracle import oracle_py as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wl = dx.workloads.ctdi_body_phantom(n=16, histories=200_000, step_deg=10.0)
    ow = orc.OracleWorld.from_workload(wl)
    e, e2, cnt, st = ow.run(wl.beam, 1, seed=77, threads=1, rank=rank, world=world)
    # the device's tally layout: 4 x int64 per voxel {E * 2^24, E^2 * 2^16, events, pad}
    tally = np.zeros((e.size, 4), dtype=np.int64)
    tally[:, 2] = cnt.astype(np.int64)
    t = torch.from_numpy(tally.reshape(-1))
    hist = torch.tensor([st["histories"]], dtype=torch.int64)
    D.reduce_tallies(t, 0)
    dist.reduce(hist, 0)
    if rank == 0:
        full_e, full_e2, full_cnt, full_st = ow.run(wl.beam, 1, seed=77, threads=1)
        ok = bool(np.array_equal(t.numpy().reshape(-1, 4)[:, 2], full_cnt.astype(np.int64))) and int(hist[0]) == full_st["histories"]
        out.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_reduce_of_sharded_tallies():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
