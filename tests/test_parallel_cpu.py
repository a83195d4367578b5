"""CPU, world_size 2 over gloo: the sharding / gather plumbing of the multi-GPU path (pylinac_b200/parallel.py)."""
import os
import socket
import sys

import numpy as np
import pytest

from pylinac_b200 import _native as nat
from pylinac_b200 import parallel


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 512, 2048, 4097):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
            assert sizes == parallel.shard_sizes(n, world)
    with pytest.raises(ValueError):
        parallel.shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    try:
        lo, hi = parallel.shard_range(n_total, world, rank)
        rows = np.zeros(hi - lo, nat.STAR_RESULT_DTYPE)
        rows["start_x"] = np.arange(lo, hi)                 # stands for "the result of frame i"
        rows["wobble_x"] = np.arange(lo, hi) * 0.5
        rows["peak_idx"][:, 3] = np.arange(lo, hi) + 7
        allrows = parallel.gather_rows(rows, n_total, dist=dist)
        ok = (len(allrows) == n_total and np.array_equal(allrows["start_x"], np.arange(n_total))
              and np.array_equal(allrows["wobble_x"], np.arange(n_total) * 0.5)
              and np.array_equal(allrows["peak_idx"][:, 3], np.arange(n_total) + 7))
        # analyze_sharded: each rank analyses only its shard (a stand-in analysis: no GPU here), everyone gets all rows
        touched = []

        class Lazy:
            def __len__(self):
                return n_total

            def __getitem__(self, sl):
                touched.append((sl.start, sl.stop))
                return np.arange(sl.start, sl.stop)

        def fake_analyze(shard):
            r = np.zeros(len(shard), nat.WL_RESULT_DTYPE)
            r["bb_x"] = shard * 2.0
            r["n_bbs"] = shard
            return r

        allwl = parallel.analyze_sharded(fake_analyze, Lazy(), dist=dist)
        ok = ok and touched == [(lo, hi)] and np.array_equal(allwl["bb_x"], np.arange(n_total) * 2.0) and \
            np.array_equal(allwl["n_bbs"], np.arange(n_total))
        # max-over-ranks of a timing, as bench.py does
        import torch

        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, bool(ok), float(t[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [5, 64])
def test_gather_rows_world_size_2_gloo(n_total):
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, 2.0), (1, True, 2.0)]


def test_bind_host_to_gpu_is_a_noop_without_a_device(tmp_path):
    """No GPU here: the PCI bus id query fails and the helper must leave the affinity untouched and say so."""
    import os

    from pylinac_b200 import parallel as par

    before = os.sched_getaffinity(0)
    info = par.bind_host_to_gpu(0, sysfs=str(tmp_path))
    assert info == {"numa_node": None, "cpus": None}
    assert os.sched_getaffinity(0) == before
    assert par._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
