"""Multi-rank host logic on CPU: world_size 2 over gloo (no GPU).  Each rank takes its slab of one seeded
batch, "computes" it with the CPU oracle (standing in for the per-rank GPU call), and the ranks agree on the
job time (max) and on a checksum of checksums that must equal the single-process value."""
import hashlib
import os
import socket

import numpy as np
import pytest

import _oracle as O

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from libsafecrypto_b200 import sharding  # noqa: E402


def test_shard_range_partitions():
    for total in (0, 1, 7, 1 << 20, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q_out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q, n = 12289, 512
    w, r = O.tables(q, n, 16)
    rng = np.random.default_rng(20261017)
    a = rng.integers(0, q, size=(total, n)).astype(np.int32)
    b = rng.integers(0, q, size=(total, n)).astype(np.int32)
    lo, hi = sharding.shard_range(total, world, rank)
    out = O.port().ntt_batch(O.REFERENCE, O.OP_POLYMUL, n, q, 16, a[lo:hi], b[lo:hi], w, r, threads=1)
    job = sharding.gather_digests(sharding.digest_rows(out))
    slowest = sharding.max_over_ranks(1.0 + rank)
    rows = sharding.sum_over_ranks(hi - lo)
    dist.barrier()
    if rank == 0:
        q_out.put((job, slowest, rows))
    dist.destroy_process_group()


def test_two_ranks_agree_with_single_process():
    total, world = 37, 2
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    job, slowest, rows = q_out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert slowest == 2.0 and rows == total
    # single-process checksum of checksums over the same slabs
    q, n = 12289, 512
    w, r = O.tables(q, n, 16)
    rng = np.random.default_rng(20261017)
    a = rng.integers(0, q, size=(total, n)).astype(np.int32)
    b = rng.integers(0, q, size=(total, n)).astype(np.int32)
    full = O.port().ntt_batch(O.REFERENCE, O.OP_POLYMUL, n, q, 16, a, b, w, r)
    parts = b"".join(sharding.digest_rows(full[slice(*sharding.shard_range(total, world, g))]) for g in range(world))
    assert job == hashlib.sha256(parts).digest()
