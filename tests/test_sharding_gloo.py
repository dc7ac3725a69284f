"""The N>1 path on CPU: two gloo ranks shard a batch by image, run the (oracle) pipeline on
their shards, and the gathered per-image checksums equal a single-process run -- i.e. the
sharding is a partition with no cross-image state, which is all the multi-GPU path relies
on (there is no data-path collective to test).  Also: the barrier / max / sum plumbing."""
import os
import socket
import zlib

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from imageprocessor_b200 import sharding as S
from tests.util import rgba_random

N_IMAGES = 7      # ragged on purpose: 4 + 3


def image(i):
    return rgba_random(160 + 8 * i, 120 + 4 * i, 500 + i)


def checksum(O, a):
    h, w = a.shape[:2]
    R = O.Raster.rgba(a)
    nw, nh = O.keep_aspect_dims(w, h, 64, 48)
    c = zlib.crc32(O.resize_image(R, nw, nh).tobytes())
    return zlib.crc32(O.crop_and_resize(R, 20).tobytes(), c)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    mine = S.shard_indices(N_IMAGES, rank, world)
    local = [checksum(O, image(i)) for i in mine]
    S.barrier()
    allsums = S.gather_checksums(local, N_IMAGES, rank, world)
    slow = S.reduce_max(1.0 + rank)            # "seconds" of the slowest rank
    total = S.reduce_sum(float(len(mine)))     # units over all ranks
    thr = S.whole_job_throughput(float(len(mine)), 1.0 + rank)
    q.put((rank, mine, allsums, slow, total, thr))
    dist.destroy_process_group()


def test_two_gloo_ranks_shard_by_image(oracle):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = [checksum(oracle, image(i)) for i in range(N_IMAGES)]
    owned = sorted(i for _, mine, *_ in got for i in mine)
    assert owned == list(range(N_IMAGES)), "every image belongs to exactly one rank"
    for rank, mine, allsums, slow, total, thr in got:
        assert allsums == want, "gathered checksums differ from the single-process run"
        assert slow == 2.0 and total == float(N_IMAGES) and abs(thr - N_IMAGES / 2.0) < 1e-9
    assert zlib.crc32(np.array(got[0][2], np.int64).tobytes()) == zlib.crc32(np.array(want, np.int64).tobytes())


def test_shard_and_routing_rules():
    assert S.shard_indices(10, 0, 4) == [0, 4, 8] and S.shard_indices(10, 3, 4) == [3, 7]
    assert S.shard_indices(2, 3, 4) == [] and S.shard_indices(0, 0, 1) == []
    with pytest.raises(ValueError):
        S.shard_indices(4, 4, 4)
    assert S.least_loaded([5, 3, 3, 9]) == 1 and S.least_loaded([0]) == 0
    assert S.reduce_max(3.5) == 3.5 and S.reduce_sum(2.0) == 2.0       # no process group: identity
