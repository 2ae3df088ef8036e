"""The N > 1 path on the CPU: world_size-2 and -3 process groups over gloo.  Each rank reduces its
contiguous shard to one partial point (the pipeline is run by the CPU emulation of the kernels,
tests/host_emu), the partials travel through the same all-gather the NCCL path uses
(zk-toolkit_b200/sharding.py), every rank combines them, and the result must equal the
single-rank result and the oracle -- independent of the partition."""
import ctypes
import importlib
import os
import random
import subprocess

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import zkt_oracle as O
from tests import util as U

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "host_emu")


def build_emu():
    so = os.path.join(EMU, "libemu_msm.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(EMU, "emu_msm.cpp")])
    return so


def _worker(rank, world, port, so, xy, sc, n, ret, split="points"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharding = importlib.import_module("zk-toolkit_b200.sharding")
    lib = ctypes.CDLL(so)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    part = np.zeros(48, dtype=np.uint32)
    if split == "range":   # every rank: all terms, 1 / world of the bucket range (window 6, subgroup fold)
        assert lib.emu_g1_msm_partial_range(P(xy), P(sc), n, 6, 1, rank, world, P(part)) == 0
    else:
        lo, hi = sharding.shard_range(n, rank, world)
        xs, ss = np.ascontiguousarray(xy[lo:hi]), np.ascontiguousarray(sc[lo:hi])
        assert lib.emu_g1_msm_partial(P(xs), P(ss), hi - lo, P(part)) == 0
    gathered = sharding.gather_partials(torch.from_numpy(part.view(np.int32)))
    assert gathered.shape == (world, 48)
    g = np.ascontiguousarray(gathered.numpy().view(np.uint32))
    out = np.zeros(24, dtype=np.uint32)
    inf = ctypes.c_uint32(0)
    lib.emu_g1_combine(P(g), world, P(out), ctypes.byref(inf))
    ret[rank] = (out.tolist(), int(inf.value))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,split", [(2, "points"), (3, "points"), (2, "range"), (4, "range")])
def test_sharded_msm_over_gloo(world, split):
    so = build_emu()
    rnd = random.Random(100 + world)
    n = 37
    dlogs = [rnd.randrange(1, O.R) for _ in range(n)]
    pts = [O.scalar_mul(O.G1_GEN, k) for k in dlogs]
    scalars = U.rand_scalars(rnd, n)
    xy, _ = U.g1_points_to_array(pts)
    sc = U.scalars_to_array(scalars)
    exp = U.expected_from_dlogs(O.G1_GEN, dlogs, scalars)
    port = 29500 + random.randrange(2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, so, xy, sc, n, ret, split), nprocs=world, join=True)
        results = [ret[r] for r in range(world)]
    for out, inf in results:
        assert U.g1_from_array(np.array(out, dtype=np.uint32), inf) == exp


def test_shard_ranges_partition():
    sharding = importlib.import_module("zk-toolkit_b200.sharding")
    for n in (0, 1, 7, 1 << 20):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
