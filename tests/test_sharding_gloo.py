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


def _loss_worker(rank, world, port, out):
    """Each replica: loss gradient of ITS shard (cnt = local B * cells * A, model/yolo2/__init__.py:89), then the one
    all-reduce + 1/G of parallel.allreduce_mean_."""
    from oracle import head_oracle as ho
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        classes, cells_w, cells_h, batch = 20, 5, 4, 6
        rs = np.random.RandomState(11)
        net = rs.normal(0, 1, size=(batch, cells_h, cells_w, 5 * (5 + classes)))
        labels = ho.synthetic_labels(batch, classes, cells_w, cells_h, seed=3)
        b, e = parallel.shard_range(batch, rank, world)
        obj, g = ho.loss_grad_oracle(net[b:e], classes, ho.ANCHORS_VOC, tuple(t[b:e] for t in labels), dtype=np.float64)
        # stand-in for the backbone's backward: a linear map of dL/dnet summed over the shard's images (a weight gradient)
        wgrad = torch.from_numpy(g.reshape(e - b, -1).sum(0).copy())
        loss = torch.tensor([ho.total_loss_oracle(obj)], dtype=torch.float64)
        parallel.allreduce_mean_(wgrad)
        parallel.allreduce_mean_(loss)
        out.put((rank, wgrad.numpy(), float(loss)))
    finally:
        dist.destroy_process_group()


def test_world2_replica_mean_of_loss_gradients_equals_global_batch_gradient():
    """the loss normalises by the LOCAL batch, so averaging equal-sized replicas reproduces the gradient (and
    the loss) of the whole batch on one GPU -- checked with the real loss (oracle, float64) on a batch of 6 split 3 / 3."""
    from oracle import head_oracle as ho
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_loss_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((r, g, l) for r, g, l in (q.get(timeout=120) for _ in range(world)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    classes, cells_w, cells_h, batch = 20, 5, 4, 6
    rs = np.random.RandomState(11)
    net = rs.normal(0, 1, size=(batch, cells_h, cells_w, 5 * (5 + classes)))
    labels = ho.synthetic_labels(batch, classes, cells_w, cells_h, seed=3)
    obj, g = ho.loss_grad_oracle(net, classes, ho.ANCHORS_VOC, labels, dtype=np.float64)
    want = g.reshape(batch, -1).sum(0)
    for rank, got, loss in res:
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(loss, ho.total_loss_oracle(obj), rtol=1e-12)
