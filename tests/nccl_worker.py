"""Worker of tests/test_gpu_nccl.py: one rank of a torchrun launch (one process per GPU, NCCL).  Runs the
multi-GPU paths of the product -- point split, bucket-range split, one Groth16 proof over all ranks -- and
compares every result, on every rank, with the oracle (closed forms / the oracle's own prover)."""
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from importlib import import_module  # noqa: E402

import zk_toolkit_b200 as z  # noqa: E402
from oracle import zkt_oracle as O  # noqa: E402
from tests import util as U  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sharding = import_module("zk-toolkit_b200.sharding")
    G = import_module("zk-toolkit_b200.groth16")
    S = import_module("zk-toolkit_b200.synthetic")
    ctx = z.default_context()
    assert ctx.device == local
    # the library's kernels and the NCCL exchange must be ordered on ONE stream (as bench.py does): the context's own
    # stream is non-blocking with respect to torch's
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    torch.cuda.set_stream(stream)
    n = 1 << 14
    rnd = random.Random(1234)                      # the same instance on every rank
    dlogs = [rnd.randrange(1, O.R) for _ in range(n)]
    sc = U.rand_scalars(rnd, n)
    want = U.expected_from_dlogs(O.G1_GEN, dlogs, sc)
    arr = z.scalars_to_array(sc)
    # ---- point split: contiguous shards, device-resident partials, one all-gather, combine on every rank
    lo, hi = sharding.shard_range(n, rank, world)
    shard = z.G1Points.generator_multiples(dlogs[lo:hi], precompute=True)
    d_sc = torch.from_numpy(arr[lo:hi].view(np.int32)).cuda()
    d_part = torch.zeros(48, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.msm_partial_device(shard.set, d_sc.data_ptr(), hi - lo, d_part.data_ptr())
    torch.cuda.synchronize()
    g = sharding.gather_partials(d_part)
    assert U.g1_from_array(*ctx.combine_device(g.data_ptr(), world)) == want, "point split"
    # ---- bucket-range split: whole set and all scalars on every rank
    full = z.G1Points.generator_multiples(dlogs, precompute=True)
    d_all = torch.from_numpy(arr.view(np.int32)).cuda()
    torch.cuda.synchronize()
    ctx.msm_partial_range_device(full.set, d_all.data_ptr(), n, rank, world, d_part.data_ptr())
    torch.cuda.synchronize()
    g = sharding.gather_partials(d_part)
    ctx.combine_enqueue(g.data_ptr(), world)
    assert U.g1_from_array(*ctx.msm_result(1)) == want, "range split"
    # ---- a bad scalar on ONE rank is reported by every rank's combine
    bad = arr.copy()
    if rank == world - 1:
        bad[3, 7] |= 0x80000000
    d_bad = torch.from_numpy(bad.view(np.int32)).cuda()
    torch.cuda.synchronize()
    ctx.msm_partial_range_device(full.set, d_bad.data_ptr(), n, rank, world, d_part.data_ptr())
    torch.cuda.synchronize()
    g = sharding.gather_partials(d_part)
    try:
        ctx.combine_device(g.data_ptr(), world)
        raise AssertionError("out-of-range scalar was not reported")
    except z.ZkmsmError as e:
        assert e.code == -3
    # ---- one Groth16 proof over all ranks == the proof of one GPU == its closed form by the oracle
    inst = S.build(1 << 10, 1 << 10, seed=99)
    r, s = 0xabcdef123, 0x987654321
    one = inst["prover"].prove(inst["crs"], r, s)
    many = G.prove_distributed(inst["prover"], inst["crs"], r, s)
    assert (many.A, many.B, many.C) == (one.A, one.B, one.C), "distributed proof differs"
    a, b, c = S.expected_dlogs(inst, r, s)
    assert (many.A.x, many.A.y) == tuple(v.e for v in O.scalar_mul(O.G1_GEN, a))
    assert (many.C.x, many.C.y) == tuple(v.e for v in O.scalar_mul(O.G1_GEN, c))
    dist.barrier()
    if rank == 0:
        print(f"NCCL_WORKER_OK world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
