"""-m gpu, needs >= 2 GPUs (skipped otherwise): the data-parallel path over NCCL.  Two ranks each run the training
step on their own shard; the ONE all-reduce on the flat gradient bucket must give the mean of the two per-rank buckets
(bit-for-bit equal on both ranks), and inference on batch shards needs no collective at all."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from oracle import head_oracle as ho
    from oracle.darknet_oracle import init_params
    from yolo_tf_b200 import parallel, variables
    from yolo_tf_b200.model.yolo2 import Builder
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    try:
        classes, size, batch = 20, 64, 4
        params = init_params(classes, 5, seed=1)
        store = variables.reset_default_store()
        store.assign({"yolo2_darknet/" + k: v for k, v in params.items()})
        rs = np.random.RandomState(7)
        x_all = rs.normal(0, 1, size=(batch, size, size, 3)).astype(np.float32)
        labels_all = ho.synthetic_labels(batch, classes, 2, 2, seed=1)
        b0, b1 = parallel.shard_range(batch, rank, world)
        builder = Builder.from_values([str(i) for i in range(classes)], size, size, ho.ANCHORS_VOC)
        builder(torch.from_numpy(x_all[b0:b1]).to(dev), training=True)
        builder.create_objectives([t[b0:b1] for t in labels_all])
        local, _ = builder.backward(allreduce=False)
        local = local.clone()
        builder(torch.from_numpy(x_all[b0:b1]).to(dev), training=True)
        builder.create_objectives([t[b0:b1] for t in labels_all])
        reduced, _ = builder.backward(allreduce=True)
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, "local%d.npy" % rank), local.cpu().numpy())
        np.save(os.path.join(out_dir, "reduced%d.npy" % rank), reduced.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    l0, l1 = np.load(tmp_path / "local0.npy"), np.load(tmp_path / "local1.npy")
    r0, r1 = np.load(tmp_path / "reduced0.npy"), np.load(tmp_path / "reduced1.npy")
    assert np.array_equal(r0.view(np.uint32), r1.view(np.uint32))               # both replicas hold the same bucket
    mean = (l0.astype(np.float64) + l1.astype(np.float64)) / 2
    assert np.abs(r0 - mean).max() <= 1e-6 * max(np.abs(mean).max(), 1e-30)
    assert np.abs(l0 - l1).max() > 0                                             # the shards really differ
