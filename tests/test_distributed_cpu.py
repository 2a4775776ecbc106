"""Host-side logic of the image-sharded (multi-GPU) path on CPU: world_size = 2, gloo backend.
The per-shard numbers come from the CPU oracle standing in for the kernels (tests may use oracle/);
what is under test is the sharding arithmetic and the one collective (12-byte loss exchange)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuralnetworklibrary_b200 import distributed as nd
from tests import synth as syn


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 5, 16, 17, 255, 256):
        for w in (1, 2, 3, 4, 8):
            blocks = [nd.shard_bounds(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        nd.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, H, W, C, M, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from neuralnetworklibrary_b200.vision import reduce_loss_scalars
        from oracle import oracle as orc
        an = orc.anchors(H, W)
        gb, gc = syn.make_targets(B, M, H, W, C, seed=91, min_side=10.0, max_frac=0.7)
        clas, reg = syn.make_train_activations(B, an.shape[0], C, seed=91)
        activ, target = nd.shard_batch([torch.from_numpy(an), reg, clas], [gb, gc], world, rank)
        o = orc.loss(an, activ[2].numpy(), activ[1].numpy(), target[0].numpy(), target[1].numpy(), B_global=B)
        total = reduce_loss_scalars(torch.from_numpy(o["out3"]))       # the one collective of the path
        lists = ([["r%d" % rank]], [[rank]], [[float(rank)]])
        merged = nd.gather_detections(lists)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), total=total.numpy(), dclas=o["dclas"], dreg=o["dreg"],
                 merged=np.array([m[0] for m in merged[1]]))
    finally:
        dist.destroy_process_group()


def test_sharded_loss_world2_gloo(tmp_path):
    from oracle import oracle as orc
    B, H, W, C, M, world = 5, 64, 96, 20, 4, 2
    mp.spawn(_worker, args=(world, _free_port(), B, H, W, C, M, str(tmp_path)), nprocs=world, join=True)
    an = orc.anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=91, min_side=10.0, max_frac=0.7)
    clas, reg = syn.make_train_activations(B, an.shape[0], C, seed=91)
    full = orc.loss(an, clas.numpy(), reg.numpy(), gb.numpy(), gc.numpy())
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    # every rank holds the same, bit-identical total (all_gather + fixed rank-order sum)
    assert np.array_equal(parts[0]["total"], parts[1]["total"])
    np.testing.assert_allclose(parts[0]["total"], full["out3"], rtol=1e-6)
    # per-image gradients need no exchange: the shards' gradients are the full batch's, bit for bit
    assert np.array_equal(np.concatenate([p["dclas"] for p in parts]), full["dclas"])
    assert np.array_equal(np.concatenate([p["dreg"] for p in parts]), full["dreg"])
    assert parts[0]["merged"].tolist() == [0, 1]
