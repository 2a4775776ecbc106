"""Multi-GPU functional check (launched by tests/test_gpu_distributed.py, or by hand with torchrun on a box with >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tests/dist_gpu_check.py

Every rank computes the SSD_loss of its image shard (NCCL exchange of the three loss scalars) and, in addition, the full
batch on its own GPU; the checker is the CPU oracle on the full batch: the sharded loss equals the
single-GPU loss within rtol 1e-6 and is bit-identical on all ranks, each rank's gradients are bit-identical to the
corresponding slices of the full-batch gradients, and the sharded detections concatenate to the full-batch detections."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuralnetworklibrary_b200 import distributed as nd  # noqa: E402
from tests import synth as syn  # noqa: E402
from neuralnetworklibrary_b200.retinanet import AnchorGenerator, BBoxPredictor  # noqa: E402
from neuralnetworklibrary_b200.vision import SSD_loss  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    H, W, C, B, M = 256, 320, 20, 7, 6            # 7 images: uneven shards
    an = orc.anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=77, min_side=12.0, max_frac=0.6)
    clas, reg = syn.make_train_activations(B, an.shape[0], C, seed=77, edge_cases=64)
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev))

    # full batch on this GPU
    cf, rf = clas.to(dev).requires_grad_(True), reg.to(dev).requires_grad_(True)
    full = SSD_loss()
    lf = full([anchors, rf, cf], [gb.to(dev), gc.to(dev)])
    lf.backward()
    # this rank's shard of the same batch
    lo, hi = nd.shard_bounds(B, world, rank)
    activ, target = nd.shard_batch([anchors, reg.to(dev), clas.to(dev)], [gb.to(dev), gc.to(dev)], world, rank)
    rs, cs = activ[1].clone().requires_grad_(True), activ[2].clone().requires_grad_(True)
    shard = nd.sharded_ssd_loss(SSD_loss, B)
    ls = shard([anchors, rs, cs], target)
    ls.backward()

    o = orc.loss(an, clas.numpy(), reg.numpy(), gb.numpy(), gc.numpy())
    got = np.array([ls.item(), shard.reg_loss.item(), shard.clas_loss.item()], np.float32)
    np.testing.assert_allclose(got, o["out3"], rtol=1e-5, atol=0)
    np.testing.assert_allclose(ls.item(), lf.item(), rtol=1e-6)
    assert torch.equal(cs.grad, cf.grad[lo:hi]) and torch.equal(rs.grad, rf.grad[lo:hi]), "shard gradients differ from full-batch slices"
    every = [torch.zeros(1, device=dev) for _ in range(world)]
    dist.all_gather(every, ls.detach().reshape(1))
    assert all(torch.equal(every[0], e) for e in every), "loss differs between ranks"

    # the same shard through the peer-memory exchange (rn_peer_exchange): bit-identical to the NCCL all-gather + rank-order sum,
    # eagerly and as CUDA-graph replays with the exchange inside the graph
    peer_ok = "skipped"
    try:
        from neuralnetworklibrary_b200.vision import PeerExchange
        px = PeerExchange()
    except Exception as exc:   # no symmetric memory on this box: the NCCL path stands
        px, peer_ok = None, "unavailable (%s)" % type(exc).__name__
    if px is not None:
        peer = SSD_loss(distributed=True, global_batch=B, peer_exchange=px)
        rs2, cs2 = activ[1].clone().requires_grad_(True), activ[2].clone().requires_grad_(True)
        lp = peer([anchors, rs2, cs2], target)
        lp.backward()
        assert lp.item() == ls.item() and peer.reg_loss.item() == shard.reg_loss.item() and peer.clas_loss.item() == shard.clas_loss.item()
        assert torch.equal(cs2.grad, cs.grad) and torch.equal(rs2.grad, rs.grad)
        cap = peer.capture([anchors, activ[1].contiguous(), activ[2].contiguous()], [target[0].contiguous(), target[1].contiguous()])
        for _ in range(5):
            cap.replay()
        torch.cuda.synchronize()
        assert cap.exchange_in_graph and cap.loss.item() == ls.item(), (cap.loss.item(), ls.item())
        assert torch.equal(cap.dclas, cs.grad)
        # pipelined form: the exchange of a step runs on a parallel branch at the start of the NEXT replay's graph
        # (rn_peer_exchange_to); the static regression input changes between replays, so a stale or mixed-up sum would show.
        # Everything is bit-identical to the all-gather path and the gradients are untouched.
        t0, t1 = target[0].contiguous(), target[1].contiguous()
        rs3, cs3 = (activ[1] * 0.5).requires_grad_(True), activ[2].clone().requires_grad_(True)
        lb = shard([anchors, rs3, cs3], target)
        want = {"A": ls.item(), "B": lb.item()}
        assert want["A"] != want["B"]
        reg_in = activ[1].contiguous().clone()
        capp = peer.capture([anchors, reg_in, activ[2].contiguous()], [t0, t1], pipelined_exchange=True)
        assert capp.pipelined
        prev = None
        for name in "ABBABA":
            reg_in.copy_(activ[1] if name == "A" else activ[1] * 0.5)
            capp.replay()
            tot = capp.total().clone()
            torch.cuda.synchronize()
            assert tot[0].item() == want[name], (name, tot[0].item(), want[name])
            if prev is not None:
                assert capp.previous_total[0].item() == want[prev], (name, prev, capp.previous_total[0].item())
            prev = name
        assert torch.equal(capp.dclas, cs.grad)
        peer_ok = "ok (%d kernels per replay; pipelined exchange ok)" % cap.kernels_per_replay

    # detections: shard + gather == full batch
    ci, ri = syn.make_infer_activations(B, an.shape[0], C, seed=78, anchors=an, mu=-5.0, clusters=5)
    bp = BBoxPredictor()
    img = torch.zeros(B, 3, H, W, device=dev)
    full_det = bp(img, ri.to(dev), ci.to(dev), anchors)
    part = bp(img[lo:hi], ri[lo:hi].to(dev), ci[lo:hi].to(dev), anchors)
    merged = nd.gather_detections(part)
    for a, b in zip(full_det, merged):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert len(x) == len(y) and all(np.array_equal(u, v) for u, v in zip(x, y))
    dist.barrier()
    if rank == 0:
        print("dist_gpu_check ok: world=%d loss=%.6f (single GPU %.6f), shards %s, peer-memory exchange %s" %
              (world, ls.item(), lf.item(), [nd.shard_bounds(B, world, r) for r in range(world)], peer_ok))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
