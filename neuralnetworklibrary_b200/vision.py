"""Drop-in replacements for the bounding-box loss objects of the reference's Applications/Vision.py,
backed by libretina_sm100.so:

    SSD_loss                  Vision.py:1607-1644  -> rn_assign + rn_loss (+ rn_scale_grads in backward)
    SSD_RegLoss / SSD_ClasLoss Vision.py:1646-1663 -> read the attributes SSD_loss stores
    match_anchors_objects     Vision.py:1474-1511  -> rn_assign on one image
    ComputeMaxOverlaps        Vision.py:1666-1694  -> rn_max_overlaps

`ImageLearner(..., loss_func=SSD_loss())` works unchanged: __call__(activ=[anchors, reg, clas],
target=[BBoxes, Cats]) returns a 0-dim float32 tensor that is differentiable w.r.t. reg and clas and
sets .reg_loss / .clas_loss on every call (Vision.py:1643).
"""
import numpy as np
import torch

from . import _lib
from .core import ARR, TEN
from .retinanet import anchor_args

_ws = _lib.Workspace()
_step_state = _lib.Workspace(zeroed=True)   # rn_loss_step's state buffer: zero-initialised once, left zeroed by every call


def _targets(target, device):
    BBoxes, Cats = target[0], target[1]
    _lib.require_cuda(BBoxes, "BBoxes")
    _lib.require_cuda(Cats, "Cats")
    # (the common case -- float32 / int64, contiguous, no grad -- costs no torch call at all: this runs once per step)
    gt_boxes = BBoxes if (BBoxes.dtype == torch.float32 and BBoxes.is_contiguous() and not BBoxes.requires_grad) \
        else BBoxes.detach().to(dtype=torch.float32).contiguous()
    gt_cats = Cats if (Cats.dtype == torch.int64 and Cats.is_contiguous()) else Cats.detach().to(dtype=torch.int64).contiguous()
    if gt_boxes.dim() != 3 or gt_boxes.shape[2] != 4 or gt_cats.shape != gt_boxes.shape[:2]:
        raise ValueError("target must be [BBoxes (bs x M x 4), Cats (bs x M)]")
    return gt_boxes, gt_cats


def assign_batch(anchors, gt_boxes, gt_cats, pos_thresh=0.5, neg_thresh=0.4, want_iou=False):
    """rn_assign on a batch: returns (matches [B,A] int32, npos [B] int32, max_iou [B,A] or None).
    matches: >= 0 matched object (index among the image's non-padding rows), -1 background, -2 ignored."""
    lib = _lib.load()
    B, M = int(gt_cats.shape[0]), int(gt_cats.shape[1])
    H, W, base, K, table, A = anchor_args(anchors)
    dev = gt_boxes.device
    matches = torch.empty((B, A), dtype=torch.int32, device=dev)
    npos = torch.empty((B,), dtype=torch.int32, device=dev)
    miou = torch.empty((B, A), dtype=torch.float32, device=dev) if want_iou else None
    with torch.cuda.device(dev):
        _lib.check(lib.rn_assign(_lib.ptr(gt_boxes), _lib.ptr(gt_cats), B, M, H, W, base, K, table, A,
                                 float(pos_thresh), float(neg_thresh), _lib.ptr(matches), _lib.ptr(npos),
                                 _lib.ptr(miou), _lib.stream_ptr(dev)))
    return matches, npos, miou


def _launch_loss(anchors, reg, clas, gt_boxes, gt_cats, cfg, need_grad, bufs=None):
    """rn_loss_step on the current stream: assignment + loss forward/backward + final reduction in one library call (the
    kernels of rn_assign + rn_loss by default; a byte-map chain or ONE persistent kernel behind rn_set_option).
    Returns (out3, dreg, dclas, matches, npos); `matches` is None unless cfg["keep_matches"]; `bufs` lets a caller
    (CUDA-graph capture) supply persistent output tensors and the zero-initialised workspace."""
    lib = _lib.load()
    B, A, Cn = (int(v) for v in clas.shape)
    M = int(gt_cats.shape[1])
    dev = clas.device
    H, W, base, K, table, _ = anchor_args(anchors)
    B_global = cfg["global_batch"] if cfg["global_batch"] else B * cfg["world_size"]
    if bufs is None:
        bufs = {}
    matches = bufs.get("matches")
    if matches is None and cfg.get("keep_matches"):
        matches = torch.empty((B, A), dtype=torch.int32, device=dev)
    npos = bufs.get("npos")
    if npos is None:
        npos = torch.empty((B,), dtype=torch.int32, device=dev)
    dclas = dreg = None
    if need_grad:
        dclas = bufs.get("dclas")
        if dclas is None:
            dclas = torch.empty_like(clas)
        dreg = bufs.get("dreg")
        if dreg is None:
            dreg = torch.empty_like(reg)
    out3 = bufs.get("out3")
    if out3 is None:
        out3 = torch.empty(3, dtype=torch.float32, device=dev)
    if B == 0:   # an empty image shard: nothing to launch but the zeroing of the three scalars
        npos = torch.empty((0,), dtype=torch.int32, device=dev)
        matches = torch.empty((0, A), dtype=torch.int32, device=dev) if cfg.get("keep_matches") else None
    ws = bufs.get("ws")
    if ws is None:
        ws = _ws.get(lib.rn_loss_step_workspace_bytes(B, A, Cn), dev)
    state = bufs.get("state")
    if state is None:
        state = _step_state.get(lib.rn_loss_step_state_bytes(B, A), dev)
    probs = None
    from_logits = bool(cfg.get("from_logits"))
    if from_logits:
        probs = torch.empty_like(clas) if cfg.get("want_probs") else None
        cfg["last_probs"] = probs
    _lib.check(lib.rn_loss_step(_lib.ptr(clas), _lib.ptr(reg), _lib.ptr(gt_boxes), _lib.ptr(gt_cats), B, A, Cn, M, H, W,
                                base, K, table, float(cfg["pos_thresh"]), float(cfg["neg_thresh"]), float(cfg["alpha"]),
                                float(cfg["gamma"]), float(cfg["beta"]), int(B_global), int(from_logits), _lib.ptr(dclas),
                                _lib.ptr(dreg), _lib.ptr(probs), _lib.ptr(out3), _lib.ptr(npos), _lib.ptr(matches),
                                _lib.ptr(state), state.numel(), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
    return out3, dreg, dclas, matches, npos


def _step_is_fused(anchors, gt_cats, cfg):
    """True when rn_loss_step runs as the single persistent kernel (opt-in: rn_set_option("step_fused", 1); see
    include/retina_b200.h)."""
    return _lib.load().rn_get_option(b"step_fused") > 0 and anchor_args(anchors)[4] is None \
        and 1 <= int(gt_cats.shape[1]) <= 128 and cfg["neg_thresh"] >= 0.2 and cfg["pos_thresh"] >= cfg["neg_thresh"]


class _SSDLossFunction(torch.autograd.Function):
    """Forward computes the loss AND both gradients in one streaming pass; backward hands the stored
    gradients back, scaled on the device by the upstream gradient (a no-op launch when it is 1)."""

    @staticmethod
    def forward(ctx, reg, clas, anchors, gt_boxes, gt_cats, cfg):
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        dev = clas.device
        if dev.index != torch.cuda.current_device():
            with torch.cuda.device(dev):
                out3, dreg, dclas, matches, npos = _launch_loss(anchors, reg, clas, gt_boxes, gt_cats, cfg, need_grad)
        else:
            out3, dreg, dclas, matches, npos = _launch_loss(anchors, reg, clas, gt_boxes, gt_cats, cfg, need_grad)
        if cfg["world_size"] > 1:
            out3 = cfg["exchange"](out3) if cfg.get("exchange") is not None else reduce_loss_scalars(out3, cfg["group"])
        ctx.grads = (dreg, dclas)
        ctx.used = False
        cfg["last_matches"], cfg["last_npos"] = matches, npos
        cfg["assign_inputs"] = (anchors, gt_boxes, gt_cats)   # for a lazy last_assignment
        loss, reg_loss, clas_loss = out3.unbind(0)
        ctx.mark_non_differentiable(reg_loss, clas_loss)
        return loss, reg_loss, clas_loss

    @staticmethod
    def backward(ctx, g_loss, _g_reg, _g_clas):
        dreg, dclas = ctx.grads
        if dreg is None:
            return None, None, None, None, None, None
        if ctx.used:
            raise RuntimeError("SSD_loss: backward through the same loss value twice is not supported; "
                               "the gradients were produced (and scaled in place) by the first backward")
        ctx.used = True
        ctx.grads = None  # hand over the only reference: autograd then adopts the buffers instead of cloning 1 GB
        lib = _lib.load()
        g = g_loss if (g_loss.dtype == torch.float32 and g_loss.is_contiguous()) else g_loss.detach().to(dtype=torch.float32).contiguous()
        if dclas.device.index != torch.cuda.current_device():
            with torch.cuda.device(dclas.device):
                _lib.check(lib.rn_scale_grads(_lib.ptr(dclas), dclas.numel(), _lib.ptr(dreg), dreg.numel(), _lib.ptr(g),
                                              _lib.stream_ptr(dclas.device)))
        else:
            _lib.check(lib.rn_scale_grads(_lib.ptr(dclas), dclas.numel(), _lib.ptr(dreg), dreg.numel(), _lib.ptr(g),
                                          _lib.stream_ptr(dclas.device)))
        return dreg, dclas, None, None, None, None


def level_shapes(H, W, K, n):
    """Shapes [(K*n, gh_l, gw_l)] of the five NCHW head outputs for an H x W image (retinanet.py:488)."""
    return [(K * n, -(-H // (8 << l)), -(-W // (8 << l))) for l in range(_lib.NUM_LEVELS)]


def _launch_loss_levels(anchors, reg_levels, clas_levels, gt_boxes, gt_cats, cfg, need_grad):
    """rn_assign + rn_loss_levels on the current stream: the loss on the heads' NCHW level tensors (no permute /
    view / cat pass, retinanet.py:215-217, :289-295, Vision.py:1467-1468).  Returns (out3, dreg_levels,
    dclas_levels, matches, npos)."""
    lib = _lib.load()
    H, W, base, K, table, A = anchor_args(anchors)
    if table is not None:
        raise ValueError("the level-tensor loss needs anchors from this package's AnchorGenerator (geometry tag)")
    B, M = int(gt_cats.shape[0]), int(gt_cats.shape[1])
    if len(reg_levels) != _lib.NUM_LEVELS or len(clas_levels) != _lib.NUM_LEVELS:
        raise ValueError("expected %d level tensors (P3..P7)" % _lib.NUM_LEVELS)
    if int(clas_levels[0].shape[1]) % K:
        raise ValueError("class head channels must be a multiple of the %d anchors per cell" % K)
    Cn = int(clas_levels[0].shape[1]) // K
    for t, shp in zip(list(clas_levels) + list(reg_levels), level_shapes(H, W, K, Cn) + level_shapes(H, W, K, 4)):
        _lib.require_cuda(t, "level tensor", torch.float32)
        if tuple(t.shape) != (B,) + shp or not t.is_contiguous():
            raise ValueError("level tensor must be contiguous NCHW of shape %s, got %s" % ((B,) + shp, tuple(t.shape)))
    dev = clas_levels[0].device
    B_global = cfg["global_batch"] if cfg["global_batch"] else B * cfg["world_size"]
    matches = torch.empty((B, A), dtype=torch.int32, device=dev)
    npos = torch.empty((B,), dtype=torch.int32, device=dev)
    dclas = dreg = probs = None
    if need_grad:
        dclas = [torch.empty_like(t) for t in clas_levels]
        dreg = [torch.empty_like(t) for t in reg_levels]
    if cfg.get("from_logits") and cfg.get("want_probs"):
        probs = [torch.empty_like(t) for t in clas_levels]
    cfg["last_probs"] = probs
    out3 = torch.empty(3, dtype=torch.float32, device=dev)
    ws = _ws.get(lib.rn_loss_levels_workspace_bytes(B, H, W, K, Cn), dev)
    stream = _lib.stream_ptr(dev)
    _lib.check(lib.rn_assign(_lib.ptr(gt_boxes), _lib.ptr(gt_cats), B, M, H, W, base, K, None, A,
                             float(cfg["pos_thresh"]), float(cfg["neg_thresh"]), _lib.ptr(matches), _lib.ptr(npos),
                             None, stream))
    _lib.check(lib.rn_loss_levels(_lib.ptr_array(clas_levels), _lib.ptr_array(reg_levels), int(bool(cfg.get("from_logits"))),
                                  _lib.ptr(gt_boxes), _lib.ptr(gt_cats), _lib.ptr(matches), _lib.ptr(npos), B, Cn, M, H, W,
                                  base, K, float(cfg["alpha"]), float(cfg["gamma"]), float(cfg["beta"]), int(B_global),
                                  _lib.ptr_array(dclas), _lib.ptr_array(dreg), _lib.ptr_array(probs), _lib.ptr(out3),
                                  _lib.ptr(ws), ws.numel(), stream))
    return out3, dreg, dclas, matches, npos


class _SSDLossLevelsFunction(torch.autograd.Function):
    """_SSDLossFunction for per-level NCHW head outputs: inputs are (anchors, gt_boxes, gt_cats, cfg, reg P3..P7,
    clas P3..P7); backward returns one gradient per level tensor, in the layout of its input."""

    @staticmethod
    def forward(ctx, anchors, gt_boxes, gt_cats, cfg, *levels):
        n = _lib.NUM_LEVELS
        reg_levels, clas_levels = levels[:n], levels[n:]
        need_grad = any(ctx.needs_input_grad[4:])
        with torch.cuda.device(clas_levels[0].device):
            out3, dreg, dclas, matches, npos = _launch_loss_levels(anchors, reg_levels, clas_levels, gt_boxes, gt_cats,
                                                                  cfg, need_grad)
        if cfg["world_size"] > 1:
            out3 = reduce_loss_scalars(out3, cfg["group"])
        ctx.grads = (dreg, dclas)
        ctx.used = False
        cfg["last_matches"], cfg["last_npos"] = matches, npos
        loss, reg_loss, clas_loss = out3.unbind(0)
        ctx.mark_non_differentiable(reg_loss, clas_loss)
        return loss, reg_loss, clas_loss

    @staticmethod
    def backward(ctx, g_loss, _g_reg, _g_clas):
        dreg, dclas = ctx.grads
        n = _lib.NUM_LEVELS
        if dreg is None:
            return (None,) * (4 + 2 * n)
        if ctx.used:
            raise RuntimeError("SSD_loss: backward through the same loss value twice is not supported; "
                               "the gradients were produced (and scaled in place) by the first backward")
        ctx.used = True
        ctx.grads = None
        lib = _lib.load()
        g = g_loss.detach().to(dtype=torch.float32).contiguous()
        dev = dclas[0].device
        with torch.cuda.device(dev):
            for dc, dr in zip(dclas, dreg):
                _lib.check(lib.rn_scale_grads(_lib.ptr(dc), dc.numel(), _lib.ptr(dr), dr.numel(), _lib.ptr(g),
                                              _lib.stream_ptr(dev)))
        return (None, None, None, None) + tuple(dreg) + tuple(dclas)


class CapturedLossStep(object):
    """SSD_loss forward+backward for fixed shapes, captured once into a CUDA graph and replayed with a
    single launch (assignment kernel + fused loss/gradient/reduction kernel).  For launch-bound loops: the
    whole step is ~0.1-0.5 ms of GPU time, less than the Python / launch overhead of issuing its pieces.

    The tensors passed to SSD_loss.capture() are the static inputs: copy new data into them (`.copy_`)
    before each replay().  After replay(): `.loss`, `.reg_loss`, `.clas_loss` (0-dim views, upstream
    gradient 1), `.dreg`, `.dclas` (d loss / d reg, d loss / d clas) and `.matches`, `.npos`."""

    def __init__(self, cfg, anchors, reg, clas, gt_boxes, gt_cats):
        dev = clas.device
        self.inputs = (anchors, reg, clas, gt_boxes, gt_cats)
        self.cfg = cfg
        B, A, Cn = clas.shape
        lib = _lib.load()
        self.bufs = dict(npos=torch.empty((B,), dtype=torch.int32, device=dev),
                         dclas=torch.empty_like(clas), dreg=torch.empty_like(reg),
                         out3=torch.empty(3, dtype=torch.float32, device=dev),
                         ws=torch.empty(int(lib.rn_loss_step_workspace_bytes(int(B), int(A), int(Cn))), dtype=torch.uint8, device=dev),
                         state=torch.zeros(int(lib.rn_loss_step_state_bytes(int(B), int(A))), dtype=torch.uint8, device=dev))
        if cfg.get("keep_matches"):
            self.bufs["matches"] = torch.empty((B, A), dtype=torch.int32, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        exchange = cfg.get("exchange")
        self.pipelined = bool(cfg.get("exchange_pipelined")) and exchange is not None
        self.exchange = exchange
        with torch.cuda.stream(side):   # warm-up outside capture (lazy module load); every rank runs the same sequence
            _launch_loss(anchors, reg, clas, gt_boxes, gt_cats, cfg, True, self.bufs)
            if exchange is not None:
                exchange(self.bufs["out3"])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if self.pipelined:
            # Two graphs that differ only in the buffer of the three scalars.  Graph p computes step k into share[p] and, on a
            # parallel branch that starts with the graph, sums the PREVIOUS step over the ranks (share[1-p] -> totals[1-p],
            # rn_peer_exchange_to): the NVLink round trip and the wait for the peers run beside the assignment and streaming
            # kernels instead of behind the final reduction, where they cost 5-10 us per step.
            self._share = [self.bufs["out3"], torch.zeros_like(self.bufs["out3"])]
            self._totals = [torch.zeros_like(self.bufs["out3"]) for _ in range(2)]
            self._graphs, self._parity = [], 1
            for p in range(2):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    main = torch.cuda.current_stream(dev)
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        exchange.exchange_to(self._share[1 - p], self._totals[1 - p])
                    _launch_loss(anchors, reg, clas, gt_boxes, gt_cats, cfg, True, dict(self.bufs, out3=self._share[p]))
                    main.wait_stream(side)
                self._graphs.append(g)
            self.graph = self._graphs[0]
        else:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                _launch_loss(anchors, reg, clas, gt_boxes, gt_cats, cfg, True, self.bufs)
                if exchange is not None:
                    exchange(self.bufs["out3"])
        self.exchange_in_graph = exchange is not None
        self.out3 = self.bufs["out3"]
        self.loss, self.reg_loss, self.clas_loss = self.out3.unbind(0)
        self.dreg, self.dclas = self.bufs["dreg"], self.bufs["dclas"]
        self.npos = self.bufs["npos"]
        # default: rn_assign (fill + sparse = 2 kernels, or a memset node + the dense kernel) + loss + final reduction.
        # Options: byte-map chain (3 kernels), fused step (rn_step_kernel only).
        sparse = anchor_args(anchors)[4] is None and 1 <= int(gt_cats.shape[1]) <= 128 and cfg["neg_thresh"] >= 0.2 \
            and cfg["pos_thresh"] >= cfg["neg_thresh"]
        if _step_is_fused(anchors, gt_cats, cfg):
            self.kernels_per_replay = 1
        elif sparse and int(gt_cats.shape[1]) < 128 and _lib.load().rn_get_option(b"step_bytemap") > 0:
            self.kernels_per_replay = 3
        else:
            self.kernels_per_replay = 4 if sparse else 3
        self.kernels_per_replay += 1 if self.exchange_in_graph else 0

    @property
    def matches(self):
        """[B,A] int32 assignment of the static inputs: the step's own output when the loss was built with
        keep_matches=True, else computed on demand by rn_assign."""
        if "matches" in self.bufs:
            return self.bufs["matches"]
        anchors, _, _, gt_boxes, gt_cats = self.inputs
        return assign_batch(anchors, gt_boxes, gt_cats, self.cfg["pos_thresh"], self.cfg["neg_thresh"])[0]

    def replay(self):
        if self.pipelined:
            self._parity ^= 1
            self._graphs[self._parity].replay()
            self.out3 = self._share[self._parity]   # this rank's share of the newest step
            self.loss, self.reg_loss, self.clas_loss = self.out3.unbind(0)
            return self.loss
        self.graph.replay()
        return self.loss

    def total(self):
        """[loss, reg_loss, clas_loss] of the newest replay summed over the ranks (device tensor).  Pipelined exchange: one
        rn_peer_exchange_to launch now (a collective: every rank calls it at the same point); otherwise the step's own out3,
        which already holds the sums (exchange inside the graph) or this rank's share (no exchange configured)."""
        if not self.pipelined:
            return self.out3
        p = self._parity
        return self.exchange.exchange_to(self._share[p], self._totals[p])

    @property
    def previous_total(self):
        """Pipelined exchange: the sums of the replay BEFORE the newest one, which the newest replay computed on its parallel
        branch (device tensor [3]; None without the pipelined exchange)."""
        return self._totals[1 - self._parity] if self.pipelined else None


class CapturedLevelLossStep(object):
    """CapturedLossStep for the level-tensor loss: the static inputs are the per-level NCHW lists handed to
    SSD_loss.capture(); after replay() `.loss/.reg_loss/.clas_loss`, `.dreg_levels`, `.dclas_levels` (lists in the layout of
    their inputs) and `.matches`, `.npos` hold the results (their buffers live in the graph's memory pool)."""

    def __init__(self, cfg, anchors, reg_levels, clas_levels, gt_boxes, gt_cats):
        dev = clas_levels[0].device
        self.inputs = (anchors, reg_levels, clas_levels, gt_boxes, gt_cats)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):   # warm-up outside capture (workspace allocation, lazy module load)
            _launch_loss_levels(anchors, reg_levels, clas_levels, gt_boxes, gt_cats, cfg, True)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out3, self.dreg_levels, self.dclas_levels, self.matches, self.npos = _launch_loss_levels(
                anchors, reg_levels, clas_levels, gt_boxes, gt_cats, cfg, True)
        self.loss, self.reg_loss, self.clas_loss = self.out3.unbind(0)
        self.kernels_per_replay = 4   # rn_assign_fill_kernel, rn_assign_sparse_kernel, rn_loss_levels_kernel, rn_loss_final_kernel

    def replay(self):
        self.graph.replay()
        return self.loss


def reduce_loss_scalars(out3, group=None):
    """The one collective of the path: sums the three per-rank loss scalars over the image shards.
    all_gather + a sequential sum in rank order (not all_reduce, and not torch.sum, whose reduction tree is an
    implementation detail), so the result is bit-identical on every rank, independent of the reduction algorithm NCCL
    picks, and bit-identical to rn_peer_exchange.  12 bytes per rank: pure latency."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out3 = out3.contiguous()
    if out3.is_cuda:
        gathered = torch.empty((world, out3.numel()), dtype=out3.dtype, device=out3.device)
        dist.all_gather_into_tensor(gathered, out3, group=group)   # NCCL
    else:                                                           # gloo (CPU tests of the host logic)
        parts = [torch.empty_like(out3) for _ in range(world)]
        dist.all_gather(parts, out3, group=group)
        gathered = torch.stack(parts)
    total = gathered[0]
    for r in range(1, world):
        total = total + gathered[r]
    return total


class PeerExchange(object):
    """The loss exchange of the image-sharded path without NCCL on the step's critical path: rn_peer_exchange, one tiny
    kernel on the step's own stream (capturable in its CUDA graph) that stores this rank's three scalars into every peer's
    buffer through peer-mapped memory over NVLink and sums all ranks' scalars in rank order (bit-identical on every rank).
    The buffers are torch symmetric memory (torch.distributed._symmetric_memory: allocation + rendezvous are plumbing, the
    kernel is this library's).  Construction is a collective; afterwards every rank must call the exchange the same number
    of times."""

    def __init__(self, group=None, device=None):
        import ctypes as C

        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        lib = _lib.load()
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        nbytes = int(lib.rn_peer_exchange_bytes(self.world))
        self.buf = symm_mem.empty((nbytes + 3) // 4, dtype=torch.float32, device=dev)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)   # every rank's buffer is zeroed before any rank stores into it
        self.ptrs = (C.c_void_p * self.world)(*[int(p) for p in self.handle.buffer_ptrs])
        self.seq = torch.zeros(1, dtype=torch.int32, device=dev)

    @classmethod
    def create(cls, group=None, device=None):
        """A PeerExchange if EVERY rank of the group could set one up, else None on every rank (a collective: the ranks
        agree through one NCCL all-reduce, so that no rank waits in rn_peer_exchange for a peer that fell back to NCCL)."""
        import torch.distributed as dist
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        try:
            px, ok = cls(group, dev), 1
        except Exception:
            px, ok = None, 0
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        return px if int(flag.item()) == 1 else None

    def __call__(self, out3):
        """Replaces out3 (this rank's share) by the sum over the ranks, in place, on the current stream."""
        _lib.check(_lib.load().rn_peer_exchange(_lib.ptr(out3), self.ptrs, self.rank, self.world, _lib.ptr(self.seq),
                                                _lib.stream_ptr(out3.device)))
        return out3

    def exchange_to(self, in3, total3):
        """total3 <- sum over the ranks of in3 (rn_peer_exchange_to); in3 keeps this rank's share, so exchanging the same
        step twice is harmless."""
        _lib.check(_lib.load().rn_peer_exchange_to(_lib.ptr(in3), _lib.ptr(total3), self.ptrs, self.rank, self.world,
                                                   _lib.ptr(self.seq), _lib.stream_ptr(in3.device)))
        return total3


class SSD_loss(object):
    """SSD / RetinaNet loss: (1-beta) * smooth-L1 + beta * focal (reference Vision.py:1607-1644).

    Extra, optional arguments (the reference is single-GPU): `distributed=True` treats the batch given to
    __call__ as this rank's image shard of a global batch of `global_batch` images (default: local batch x
    world size): the kernels scale by 1/global_batch and the three scalars are summed over ranks with one
    12-byte NCCL exchange; per-rank reg/clas gradients need no communication."""

    def __init__(self, beta=0.5, alpha=0.25, gamma=2.0, distributed=False, process_group=None, global_batch=None,
                 from_logits=False, keep_probs=False, keep_matches=False, peer_exchange=False, validate_targets=False):
        self.beta, self.alpha, self.gamma = beta, alpha, gamma
        # validate_targets: check on the host (one device->host sync per call) that every category is < C and raise the
        # IndexError the reference raises at Vision.py:1593; off by default -- the kernels then treat such an anchor as
        # positive without a class target
        self.validate_targets = bool(validate_targets)
        # peer_exchange (with distributed=True): sum the loss scalars with rn_peer_exchange (a kernel over peer-mapped
        # memory on the step's own stream, part of the captured graph) instead of a NCCL all-gather; True creates the
        # PeerExchange on first use (a collective), or pass one.
        self.peer_exchange = peer_exchange
        self.keep_probs = bool(keep_probs)   # from_logits only: also store sigmoid(logits) in .last_probs
        # keep_matches: the step also writes its dense [bs,A] int32 assignment (4*A*bs bytes of extra stores; for
        # inspection and tests).  Without it last_assignment computes the matches on demand with rn_assign.
        self.keep_matches = bool(keep_matches)
        # from_logits=True: `clas` holds logits and the head's nn.Sigmoid (reference retinanet.py:258,286) is
        # fused into the loss kernel (needs a head that returns logits)
        self.from_logits = bool(from_logits)
        self.distributed, self.process_group, self.global_batch = distributed, process_group, global_batch
        self.pos_thresh, self.neg_thresh = 0.5, 0.4   # defaults of match_anchors_objects, Vision.py:1474
        self.reg_loss = self.clas_loss = None
        self._cfg = {}

    def __call__(self, activ, target):
        anchors, reg, clas = activ[0], activ[1], activ[2]
        if isinstance(clas, (list, tuple)):
            return self._call_levels(anchors, reg, clas, target)
        _lib.require_cuda(reg, "reg", torch.float32)
        _lib.require_cuda(clas, "clas", torch.float32)
        gt_boxes, gt_cats = _targets(target, clas.device)
        if clas.dim() != 3 or reg.shape != clas.shape[:2] + (4,) or anchors.shape != (clas.shape[1], 4) \
                or gt_cats.shape[0] != clas.shape[0]:
            raise ValueError("expected anchors [A,4], reg [bs,A,4], clas [bs,A,C], targets of the same batch size")
        if self.validate_targets and gt_cats.numel() and int(gt_cats.max()) >= int(clas.shape[2]):
            raise IndexError("category %d is out of bounds for %d classes" % (int(gt_cats.max()), int(clas.shape[2])))
        world = 1
        if self.distributed:
            import torch.distributed as dist
            world = dist.get_world_size(self.process_group)
        cfg = dict(alpha=self.alpha, gamma=self.gamma, beta=self.beta, pos_thresh=self.pos_thresh,
                   neg_thresh=self.neg_thresh, world_size=world, group=self.process_group,
                   global_batch=self.global_batch, from_logits=self.from_logits, want_probs=self.keep_probs,
                   keep_matches=self.keep_matches, exchange=self._exchange() if world > 1 else None)
        loss, reg_loss, clas_loss = _SSDLossFunction.apply(reg if reg.is_contiguous() else reg.contiguous(),
                                                           clas if clas.is_contiguous() else clas.contiguous(), anchors,
                                                           gt_boxes, gt_cats, cfg)
        self._cfg = cfg
        self.reg_loss, self.clas_loss = reg_loss, clas_loss   # Vision.py:1643
        return loss

    def _exchange(self):
        """The PeerExchange of this loss (created on first use), or None for the NCCL all-gather."""
        if not self.peer_exchange or not self.distributed:
            return None
        if self.peer_exchange is True:
            self.peer_exchange = PeerExchange.create(self.process_group) or False   # False: every rank falls back to NCCL
        return self.peer_exchange or None

    def _cfg_for_call(self):
        world = 1
        if self.distributed:
            import torch.distributed as dist
            world = dist.get_world_size(self.process_group)
        return dict(alpha=self.alpha, gamma=self.gamma, beta=self.beta, pos_thresh=self.pos_thresh,
                    neg_thresh=self.neg_thresh, world_size=world, group=self.process_group,
                    global_batch=self.global_batch, from_logits=self.from_logits, want_probs=self.keep_probs)

    def _call_levels(self, anchors, reg_levels, clas_levels, target):
        """activ = [anchors, [reg P3..P7], [clas P3..P7]] with the heads' NCHW conv outputs (reg_l [bs, 9*4, gh, gw],
        clas_l [bs, 9*C, gh, gw]) instead of the permuted + concatenated [bs, A, 4|C] tensors: same loss and metrics,
        gradients arrive in the level tensors' own layout (needs an
        ObjectDetectionNet.forward that skips retinanet.py:215-217, :289-295 and Vision.py:1467-1468)."""
        gt_boxes, gt_cats = _targets(target, clas_levels[0].device)
        cfg = self._cfg_for_call()
        loss, reg_loss, clas_loss = _SSDLossLevelsFunction.apply(anchors, gt_boxes, gt_cats, cfg, *reg_levels, *clas_levels)
        self._cfg = cfg
        self.reg_loss, self.clas_loss = reg_loss, clas_loss   # Vision.py:1643
        return loss

    def capture(self, activ, target, pipelined_exchange=False):
        """Captures forward+backward for the given static tensors into a CUDA graph.  Returns a CapturedLossStep.  With
        distributed=True and peer_exchange the sum of the loss scalars over the ranks (rn_peer_exchange) is part of the graph
        and .loss / .reg_loss / .clas_loss hold the global values; with the NCCL exchange the graph holds this rank's share
        and the caller reduces it (reduce_loss_scalars).  pipelined_exchange=True (with peer_exchange): the graph ends with
        the exchange of a step's scalars runs at the START of the next replay, on a parallel branch of its graph beside the
        assignment and streaming kernels -- .loss etc. stay this rank's share, .previous_total holds the sums of the step
        before the newest one and .total() sums the newest step over the ranks on demand (a collective)."""
        anchors, reg, clas = activ[0], activ[1], activ[2]
        BBoxes, Cats = target[0], target[1]
        _lib.require_cuda(BBoxes, "BBoxes", torch.float32)
        _lib.require_cuda(Cats, "Cats", torch.int64)
        if isinstance(clas, (list, tuple)):   # per-level NCHW tensors (see _call_levels)
            if not (BBoxes.is_contiguous() and Cats.is_contiguous()):
                raise ValueError("capture() needs contiguous static tensors")
            cfg = self._cfg_for_call()
            cfg.update(world_size=1, group=None, want_probs=False)
            return CapturedLevelLossStep(cfg, anchors, [t.detach() for t in reg], [t.detach() for t in clas], BBoxes, Cats)
        _lib.require_cuda(reg, "reg", torch.float32)
        _lib.require_cuda(clas, "clas", torch.float32)
        for t in (reg, clas, BBoxes, Cats):
            if not t.is_contiguous():
                raise ValueError("capture() needs contiguous static tensors")
        cfg = dict(alpha=self.alpha, gamma=self.gamma, beta=self.beta, pos_thresh=self.pos_thresh,
                   neg_thresh=self.neg_thresh, world_size=1, group=None, global_batch=self.global_batch,
                   from_logits=self.from_logits, keep_matches=self.keep_matches, exchange=self._exchange(),
                   exchange_pipelined=bool(pipelined_exchange))
        return CapturedLossStep(cfg, anchors, reg.detach(), clas.detach(), BBoxes, Cats)

    @property
    def last_probs(self):
        """sigmoid(logits) exactly as the kernel used it (from_logits=True, keep_probs=True), else None."""
        return self._cfg.get("last_probs")

    @property
    def last_assignment(self):
        """(matches [bs,A] int32, npos [bs] int32) of the most recent call (device tensors).  npos is always the step's
        own count; matches is the step's own output with keep_matches=True (or on the level-tensor path) and is otherwise
        computed here, on demand, by rn_assign on the same targets."""
        matches, npos = self._cfg.get("last_matches"), self._cfg.get("last_npos")
        if matches is None and self._cfg.get("assign_inputs") is not None:
            anchors, gt_boxes, gt_cats = self._cfg["assign_inputs"]
            matches = assign_batch(anchors, gt_boxes, gt_cats, self._cfg["pos_thresh"], self._cfg["neg_thresh"])[0]
            self._cfg["last_matches"] = matches
        return matches, npos


class SSD_RegLoss(object):
    """Metric that reads SSD_loss.reg_loss (reference Vision.py:1646-1654)."""

    def __init__(self, loss_func):
        self.loss_func = loss_func

    def __call__(self, pred, target):
        return self.loss_func.reg_loss


class SSD_ClasLoss(object):
    """Metric that reads SSD_loss.clas_loss (reference Vision.py:1656-1663)."""

    def __init__(self, loss_func):
        self.loss_func = loss_func

    def __call__(self, pred, target):
        return self.loss_func.clas_loss


def match_anchors_objects(objects, anchors, pos_thresh=0.5, neg_thresh=0.4):
    """Single-image assignment with the reference's return convention (Vision.py:1474-1511):
    pos_idxs, neg_idxs (int64, ascending) and matches [N] int64 with -1 for every non-positive anchor.
    `objects` is an (m x 4) tensor of real (non-padding) boxes."""
    _lib.require_cuda(anchors, "anchors", torch.float32)
    objects = torch.as_tensor(objects).to(device=anchors.device, dtype=torch.float32).reshape(-1, 4).contiguous()
    m = int(objects.shape[0])
    gt_boxes = objects.unsqueeze(0) if m else torch.zeros((1, 1, 4), dtype=torch.float32, device=anchors.device)
    gt_cats = torch.zeros((1, max(m, 1)), dtype=torch.int64, device=anchors.device)
    if m == 0:
        gt_cats -= 1
    mt, _, _ = assign_batch(anchors, gt_boxes.contiguous(), gt_cats, pos_thresh, neg_thresh)
    mt = mt[0].long()
    pos_idxs = (mt >= 0).nonzero().view(-1)
    neg_idxs = (mt == _lib.MATCH_NEG).nonzero().view(-1)
    return pos_idxs, neg_idxs, torch.where(mt >= 0, mt, torch.full_like(mt, -1))


class ComputeMaxOverlaps(object):
    """Mean over images of the mean over objects of each object's best IoU with any anchor (reference
    Vision.py:1666-1694); all per-object values are appended to self.max_overlaps."""

    def __init__(self):
        self.max_overlaps = []

    def __call__(self, activ, target):
        lib = _lib.load()
        anchors = activ[0]
        gt_boxes, gt_cats = _targets(target, anchors.device)
        B, M = int(gt_cats.shape[0]), int(gt_cats.shape[1])
        H, W, base, K, table, A = anchor_args(anchors)
        out = torch.empty((B, M), dtype=torch.float32, device=gt_boxes.device)
        with torch.cuda.device(out.device):
            _lib.check(lib.rn_max_overlaps(_lib.ptr(gt_boxes), _lib.ptr(gt_cats), B, M, H, W, base, K, table, A,
                                           _lib.ptr(out), _lib.stream_ptr(out.device)))
        host, valid = ARR(out), ARR(gt_cats) >= 0
        means = []
        for i in range(B):
            v = host[i][valid[i]]
            if len(v) == 0:
                continue
            self.max_overlaps += list(v)
            means.append(v.mean())
        return TEN(float(np.array(means).mean()) if means else 0.0)


def stage_targets(bboxes, cats, scales, rand_scale=1.0, row_jit=0, col_jit=0, device=None):
    """Device-side version of the target half of AspectRatioCollater (reference Vision.py:770-785, :798-812):
    per-image box arrays are rescaled (box * scale_i * rand_scale), shifted by the jitter, rounded to float32
    and padded with -1 to the longest image -- from ONE pinned, ragged host->device upload and one small kernel
    instead of a padded host array per batch.  Returns [BBoxes (bs x M x 4) float32, Cats (bs x M) int64] on
    the device, ready for SSD_loss.  Arithmetic is float64 like NumPy's for integer / float64 box arrays."""
    lib = _lib.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    B = len(bboxes)
    counts = [0 if b is None else len(b) for b in bboxes]
    M = max(1, max(counts) if counts else 1)
    N = sum(counts)
    offsets = np.zeros(B + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(counts)
    # one pinned staging buffer: boxes f64 [N,4] | cats i64 [N] | scales f64 [B] | offsets i32 [B+1]
    nb, nc, ns, no = 32 * N, 8 * N, 8 * B, 4 * (B + 1)
    host = torch.empty(nb + nc + ns + no + 8, dtype=torch.uint8, pin_memory=True)
    hv = host.numpy()
    hb, hc = hv[:nb].view(np.float64).reshape(N, 4), hv[nb:nb + nc].view(np.int64)
    for i, (b, c) in enumerate(zip(bboxes, cats)):
        if counts[i]:
            hb[offsets[i]:offsets[i + 1]] = np.asarray(b, dtype=np.float64).reshape(-1, 4)
            hc[offsets[i]:offsets[i + 1]] = np.asarray(c, dtype=np.int64).reshape(-1)
    hv[nb + nc:nb + nc + ns].view(np.float64)[:] = np.asarray(scales, dtype=np.float64)
    hv[nb + nc + ns:nb + nc + ns + no].view(np.int32)[:] = offsets
    with torch.cuda.device(device):
        dev_buf = host.to(device, non_blocking=True)
        out_boxes = torch.empty((B, M, 4), dtype=torch.float32, device=device)
        out_cats = torch.empty((B, M), dtype=torch.int64, device=device)
        p = dev_buf.data_ptr()
        import ctypes as C
        _lib.check(lib.rn_stage_targets(C.c_void_p(p), C.c_void_p(p + nb), C.c_void_p(p + nb + nc + ns), C.c_void_p(p + nb + nc),
                                        float(rand_scale), int(row_jit), int(col_jit), B, M, _lib.ptr(out_boxes),
                                        _lib.ptr(out_cats), _lib.stream_ptr(device)))
        dev_buf.record_stream(torch.cuda.current_stream(device))
    return [out_boxes, out_cats]


def stage_images(images, row_jit=0, col_jit=0, device=None, mean=None, std=None):
    """Device-side version of the pixel half of AspectRatioCollater after its cv2.resize (reference Vision.py:775-777,
    :786, :790-796): every (already resized) H x W x C image is placed at (row_jit, col_jit), transposed to C x H x W and
    zero-padded to the batch's common size, height and width rounded up to multiples of 32 -- from ONE pinned, ragged
    host->device upload and one kernel instead of two padded host arrays and a transpose per batch.  Returns a float32
    tensor [bs, C, H, W] on the device (what the reference hands to to_cuda).

    Extension: uint8 images (0..255) are uploaded as bytes (4x less PCIe traffic) and converted on the device; with
    `mean` / `std` (per channel) the result is (x / 255 - mean) / std in float32, the normalisation the reference's
    transforms apply on the host before the collater."""
    lib = _lib.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    B = len(images)
    if B == 0:
        raise ValueError("stage_images needs at least one image")
    shapes = [tuple(int(d) for d in np.shape(im)) for im in images]
    if any(len(s) != 3 or s[2] != shapes[0][2] for s in shapes):
        raise ValueError("images must be H x W x C arrays with the same number of channels")
    Cn = shapes[0][2]
    as_bytes = all(np.asarray(im).dtype == np.uint8 for im in images)
    if (mean is None) != (std is None):
        raise ValueError("mean and std go together")
    if mean is not None and not as_bytes:
        raise ValueError("mean / std apply to uint8 images only (float images arrive normalised, as in the reference)")
    row_jit, col_jit = int(row_jit), int(col_jit)
    Hp = int(32 * np.ceil(max(s[0] + row_jit for s in shapes) / 32))     # Vision.py:790-792
    Wp = int(32 * np.ceil(max(s[1] + col_jit for s in shapes) / 32))
    sizes = [s[0] * s[1] * Cn for s in shapes]
    offsets = np.zeros(B, dtype=np.int64)
    offsets[1:] = np.cumsum(sizes)[:-1]
    total = int(sum(sizes))
    # one pinned staging buffer: pixels (f32 or u8) [total] | offsets i64 [B] | dims i32 [B,2]
    esz = 1 if as_bytes else 4
    npx, nof, ndm = esz * total, 8 * B, 8 * B
    pad = (-npx) % 8
    host = torch.empty(npx + pad + nof + ndm, dtype=torch.uint8, pin_memory=True)
    hv = host.numpy()
    px = hv[:npx] if as_bytes else hv[:npx].view(np.float32)
    for i, im in enumerate(images):
        px[offsets[i]:offsets[i] + sizes[i]] = np.asarray(im).astype(np.uint8 if as_bytes else np.float32, copy=False).reshape(-1)   # Vision.py:776
    hv[npx + pad:npx + pad + nof].view(np.int64)[:] = offsets
    hv[npx + pad + nof:].view(np.int32)[:] = np.array([[s[0], s[1]] for s in shapes], dtype=np.int32).reshape(-1)
    with torch.cuda.device(device):
        dev_buf = host.to(device, non_blocking=True)
        out = torch.empty((B, Cn, Hp, Wp), dtype=torch.float32, device=device)
        p = dev_buf.data_ptr()
        import ctypes as C
        if as_bytes:
            m32 = s32 = None
            if mean is not None:
                m32 = np.ascontiguousarray(mean, dtype=np.float32).reshape(Cn)
                s32 = np.ascontiguousarray(std, dtype=np.float32).reshape(Cn)
            _lib.check(lib.rn_stage_images_u8(C.c_void_p(p), C.c_void_p(p + npx + pad), C.c_void_p(p + npx + pad + nof), B, Cn, Hp, Wp,
                                              row_jit, col_jit, None if m32 is None else m32.ctypes.data_as(_lib._hf32p),
                                              None if s32 is None else s32.ctypes.data_as(_lib._hf32p), _lib.ptr(out),
                                              _lib.stream_ptr(device)))
        else:
            _lib.check(lib.rn_stage_images(C.c_void_p(p), C.c_void_p(p + npx + pad), C.c_void_p(p + npx + pad + nof), B, Cn, Hp, Wp,
                                           row_jit, col_jit, _lib.ptr(out), _lib.stream_ptr(device)))
        dev_buf.record_stream(torch.cuda.current_stream(device))
    return out


def merge_tta_predictions(passes, max_overlap=0.5, rel_thresh=None, top_k=1000, max_boxes=20, dup=None, inc=None,
                          transforms=None, device=None):
    """The merge step of ImageLearner.TTA_bbox (reference Vision.py:2104-2119) for ALL images in one launch: `passes` is a
    list (one entry per augmentation pass) of per-image [boxes, classes, scores] lists; the passes' predictions of each
    image are concatenated and nms is applied to the union -- rn_nms_batch: one ragged pinned upload, one launch
    sequence, one device->host copy, instead of one nms() call (and several copies) per image.

    transforms: None when the boxes are already mapped back to the original image (what TTA_bbox's loop does on the host,
    Vision.py:2091-2097); else transforms[i][l] = dict(row_jit, col_jit, rand_scale, scale, flip, cols) for pass i, image l
    and the un-transform runs on the device, fused in front of the NMS (float64 arithmetic rounded to float32 once, like
    NumPy >= 2 evaluates those lines for int64 jitter values)."""
    from .retinanet import _host_stages, decode_nms_buffer, nms_batch_device
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    L = len(passes[0])
    counts = np.zeros(L, np.int64)
    seg_counts, seg_par = [], []
    for l in range(L):
        for i, p in enumerate(passes):
            k = len(p[l][0])
            counts[l] += k
            if transforms is not None:
                t = transforms[i][l]
                seg_counts.append(k)
                seg_par.append([float(t["col_jit"]), float(t["row_jit"]), 1 / (float(t["rand_scale"]) * float(t["scale"])),
                                1.0 if (i > 0 and int(t["flip"]) == 1) else 0.0, float(t["cols"])])   # Vision.py:2095: i>0 and flip==1
    n = int(counts.sum())
    if n == 0:
        return [[[], [], []] for _ in range(L)]
    top_k = int(top_k)
    if top_k < 1:
        return [[[], [], []] for _ in range(L)]
    extra = bool(rel_thresh) or bool(inc) or bool(dup)
    max_keep = top_k if extra else max(1, min(int(max_boxes), top_k))
    S = len(seg_counts)
    # one pinned staging buffer: boxes f32 [n,4] | classes i64 [n] | scores f32 [n] | offsets i32 [L+1] | seg_off i32 [S+1] | seg_par f64 [S,5]
    nb, nc, ns, no = 16 * n, 8 * n, 4 * n, 4 * (L + 1)
    nso, nsp = 4 * (S + 1), 40 * S
    pad1 = (-(nb + nc + ns + no + nso)) % 8
    host = torch.empty(nb + nc + ns + no + nso + pad1 + nsp + 16, dtype=torch.uint8, pin_memory=True)
    hv = host.numpy()
    hb = hv[:nb].view(np.float32).reshape(n, 4)
    hc = hv[nb:nb + nc].view(np.int64)
    hs = hv[nb + nc:nb + nc + ns].view(np.float32)
    pos = 0
    for l in range(L):
        for p in passes:
            k = len(p[l][0])
            if k:
                hb[pos:pos + k] = np.asarray(p[l][0], dtype=np.float32).reshape(k, 4)
                hc[pos:pos + k] = np.asarray([int(c) for c in p[l][1]], dtype=np.int64)
                hs[pos:pos + k] = np.asarray(p[l][2], dtype=np.float32)
                pos += k
    offs = np.zeros(L + 1, np.int32)
    offs[1:] = np.cumsum(counts)
    hv[nb + nc + ns:nb + nc + ns + no].view(np.int32)[:] = offs
    o_so = nb + nc + ns + no
    o_sp = o_so + nso + pad1
    if S:
        so = np.zeros(S + 1, np.int32)
        so[1:] = np.cumsum(seg_counts)
        hv[o_so:o_so + nso].view(np.int32)[:] = so
        hv[o_sp:o_sp + nsp].view(np.float64)[:] = np.asarray(seg_par, np.float64).reshape(-1)
    with torch.cuda.device(device):
        d = host.to(device, non_blocking=True)
        boxes = d[:nb].view(torch.float32).view(n, 4)
        classes = d[nb:nb + nc].view(torch.int64)
        scores = d[nb + nc:nb + nc + ns].view(torch.float32)
        offsets = d[nb + nc + ns:nb + nc + ns + no].view(torch.int32)
        seg_off = d[o_so:o_so + nso].view(torch.int32) if S else None
        seg_pr = d[o_sp:o_sp + nsp].view(torch.float64) if S else None
        buf, K = nms_batch_device(boxes, classes, scores, offsets, max_overlap, top_k, max_keep, seg_off, seg_pr)
        kb, kc, ks, _, cnt = decode_nms_buffer(buf.cpu().numpy(), L, K)   # the one device->host copy
    merged = []
    m = max(int(max_boxes), 0)
    for l in range(L):
        k = int(cnt[l])
        b, c, s = kb[l, :k], kc[l, :k], ks[l, :k]
        if k and extra:
            sel = _host_stages(b, c, s, rel_thresh, inc, dup)
            b, c, s = b[sel], c[sel], s[sel]
        merged.append([list(b[:m]), list(c[:m]), list(s[:m])])
    return merged
