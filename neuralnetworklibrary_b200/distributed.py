"""Image sharding for the data-parallel hot path (one process per GPU).

The reference is single-GPU by design (README.md:11-12); this is the B200-native extension named in
BASELINE.json: the batch shards by image, every per-image quantity (assignment, both losses, all reg/clas
gradients, detections) depends only on that image (reference Vision.py:1636-1641, retinanet.py:756), and
the only exchange is the sum of the three loss scalars (`vision.reduce_loss_scalars`, 12 bytes).
"""
import torch


def shard_bounds(n_images, world_size, rank):
    """Contiguous block [lo, hi) of images for `rank`; the remainder goes to the low ranks."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank %d / world size %d" % (rank, world_size))
    base, rem = divmod(int(n_images), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(activ, target, world_size, rank):
    """Slices a full batch ([anchors, reg, clas], [BBoxes, Cats]) to this rank's images.  Anchors are
    shared by all images and pass through untouched (they keep their geometry tag)."""
    anchors, reg, clas = activ
    lo, hi = shard_bounds(clas.shape[0], world_size, rank)
    return [anchors, reg[lo:hi], clas[lo:hi]], [target[0][lo:hi], target[1][lo:hi]]


def sharded_ssd_loss(loss_cls, full_batch, process_group=None, **kw):
    """An SSD_loss configured for a shard of a `full_batch`-image global batch."""
    return loss_cls(distributed=True, process_group=process_group, global_batch=int(full_batch), **kw)


def gather_detections(local_lists, process_group=None):
    """Concatenates per-image detection lists from all ranks in rank order (rank 0's images first), the
    only exchange post-processing ever needs and only when one process wants the whole result."""
    import torch.distributed as dist
    world = dist.get_world_size(process_group)
    out = [None] * world
    dist.all_gather_object(out, local_lists, group=process_group)
    merged = tuple([] for _ in local_lists)
    for part in out:
        for dst, src in zip(merged, part):
            dst.extend(src)
    return merged
