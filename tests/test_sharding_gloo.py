"""CPU (-m "not gpu"): the N>1 host logic with world_size 2 over gloo -- batch sharding covers every image
exactly once, the single gradient all-reduce reproduces the global-batch mean, timing takes the max."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yolo_tf_b200 import parallel


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 32, 256, 257):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                b, e = parallel.shard_range(total, r, world)
                assert 0 <= b <= e <= total
                seen.extend(range(b, e))
            assert seen == list(range(total))
    b, e = parallel.shard_range(256, 3, 8)
    assert (b, e) == (96, 128)                      # BASELINE config 4: 256 images over 8 GPUs = 32 each


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        images = torch.arange(10 * 4, dtype=torch.float32).reshape(10, 4)        # 10 "images"
        mine = parallel.shard_batch(images, rank, world)
        # per-replica "gradient" of a mean loss over the local shard
        local_grad = mine.mean(0).clone()
        weight = torch.tensor([float(len(mine))])
        # the library's rule: equal shards -> plain mean of replicas == global-batch mean
        g = parallel.allreduce_mean_(local_grad.clone())
        t = parallel.max_over_ranks(1.0 + rank)
        ids = parallel.gather_detections([int(v) for v in mine[:, 0]])
        out.put((rank, g.numpy().tolist(), t, ids, float(weight)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_allreduce_and_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    images = np.arange(40, dtype=np.float32).reshape(10, 4)
    global_mean = images.mean(0)                     # 10 images split 5/5 -> mean of replica means == global mean
    for rank, g, t, ids, _ in res:
        np.testing.assert_allclose(g, global_mean, rtol=1e-6)
        assert t == 2.0                              # max over ranks
        assert [x for part in ids for x in part] == images[:, 0].astype(int).tolist()
