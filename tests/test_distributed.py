"""N > 1 path on CPU: world_size-2 gloo run of the view-sharding logic (deepdrr_b200/parallel.py).

The per-rank projector is replaced by the CPU oracle on a tiny scene (the oracle is allowed in tests);
what is under test is the sharding, ordering and gather -- there is no data-path collective to test.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _OracleProjector:
    def __init__(self):
        from deepdrr_b200 import phantoms
        from deepdrr_b200.scene import SceneTables

        self.v = phantoms.c1_volume(16)
        self.st = SceneTables([self.v], "60KV_AL35")

    def project(self, *poses):
        from deepdrr_b200 import geo
        from oracle import cpu_oracle

        out = []
        for p in poses:
            w2i, src, ijk = geo.pose_arrays(p, [self.v])
            r = cpu_oracle.project([self.v.data], self.st.labels, self.st.M, 12, 10, 0.5, w2i, src, ijk, 1100.0, self.st.energies,
                                   self.st.pdf, self.st.mu, want_area=False, nthreads=1)
            out.append(r.intensity)
        return np.stack(out)


def _poses(n):
    from deepdrr_b200 import geo, phantoms

    k = geo.CameraIntrinsicTransform.from_sizes((12, 10), 20.0, 1000.0)
    return [phantoms.look_at_projection((-500.0 * np.cos(0.3 * i), -500.0 * np.sin(0.3 * i), 10.0 * i),
                                        (np.cos(0.3 * i), np.sin(0.3 * i), -0.02 * i), (0, 0, 1), k) for i in range(n)]


def _worker(rank, world, port, n_views, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepdrr_b200.parallel import ViewShardedProjector, shard_range

    sp = ViewShardedProjector(_OracleProjector())
    poses = _poses(n_views)
    out = sp.project(poses, gather_to=0)
    a, b = shard_range(n_views, rank, world)
    assert sp.project_local(poses).shape[0] == b - a
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_views", [4, 5])
def test_two_rank_view_sharding_matches_single_rank(n_views):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_views
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_views, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=180)
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    single = _OracleProjector().project(*_poses(n_views))
    assert gathered.shape == single.shape
    assert np.array_equal(gathered, single)


def _tally_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepdrr_b200.scatter import reduce_over_ranks

    rng = np.random.default_rng(100 + rank)
    mine = rng.integers(0, 2**40, size=(6, 5), dtype=np.uint64)
    total = reduce_over_ranks(mine)
    q.put((rank, mine, total))
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_tally_reduce_is_an_exact_integer_sum():
    """The scatter tallies of all ranks are summed with one all-reduce (NCCL on GPUs, gloo here)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 17
    procs = [ctx.Process(target=_tally_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    expect = got[0][1] + got[1][1]
    assert np.array_equal(got[0][2], expect) and np.array_equal(got[1][2], expect)
