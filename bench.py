#!/usr/bin/env python
"""Benchmark of the RetinaNet loss / post-processing hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches N ranks)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the UNMODIFIED reference (oracle/_ref) on the host cores

A "step" is one pass of the hot path over one synthetic COCO-shaped batch per GPU: anchor assignment +
focal / smooth-L1 loss forward AND backward (SSD_loss(...) then loss.backward()), B=16 images of
800x1344 (A = 201 600 anchors, 80 classes, M = 20 ground-truth slots) per GPU -- BASELINE.json
configs[2].  Scaling is weak (per-GPU batch fixed); images shard over ranks with one 12-byte exchange of
the loss scalars per step (--exchange: by default a kernel over peer-mapped memory that runs beside the NEXT
step's kernels, see time_loss_graph).  One JSON line is printed by rank 0.

  value     images/s, inputs resident in HBM: CUDA-graph replays of the step captured from the public drop-in API
            (SSD_loss.capture); `eager` is the same step call by call (SSD_loss(...) + loss.backward())
  e2e       same API, but every step first copies that step's inputs from pinned host memory and ends
            with a device->host read of the loss
  roofline  the streaming loss kernel: algorithmic bytes 8*A*(C+4) per image / CUDA-event time of the
            rn_loss call measured inside the timed region, against MEASURED_PEAKS.json
  cpu_baseline  the unmodified reference on this box's host cores (kind "reference"; the C port of its algorithm as
            cpu_baseline_port), reference_cuda the reference on the B200 itself
Extra keys `postproc`, `pascal`, `coco_b256_sharded` report BASELINE.json configs[3] / configs[1] / configs[4] the same
way; `logits`, `levels`, `levels_logits`, `postproc_levels*` the variants that read the heads' logits / NCHW level tensors;
`aux_kernels` the kernels around the two hot paths; `parity` ties the first timed batch to the CPU oracle.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

COCO = dict(H=800, W=1344, C=80, M=20)
PASCAL = dict(H=512, W=512, C=20, M=10)
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only when MEASURED_PEAKS.json is absent


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------------------
def device_activations(B, A, C, seed, device, mu, logits=False):
    """clas = sigmoid(N(mu,1)) (or the logits themselves) and reg ~ N(0,0.25) generated on the device
    (SURVEY.md section 8d)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    clas = torch.empty((B, A, C), dtype=torch.float32, device=device)
    for i in range(B):  # per image to bound the temporaries
        z = torch.randn((A, C), generator=g, device=device) + mu
        clas[i] = z if logits else torch.sigmoid(z)
    reg = torch.randn((B, A, 4), generator=g, device=device) * 0.5
    return clas, reg


def measured_traffic(kernel, B_per_launch, ref_B):
    """DRAM bytes per launch from the committed ncu capture (profiles/r02_traffic.json), scaled to the batch of
    this run if it differs from the captured one (traffic is linear in the image count); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f)[kernel]["traffic_bytes"]
        return int(round(t * B_per_launch / ref_B))
    except Exception:
        return None


def loss_bytes(B, A, C, grad=True):
    return (8 if grad else 4) * A * (C + 4) * B


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def make_loss_sets(cfg, B, device, nbuf, seed, logits=False):
    import torch

    from tests import synth as syn
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator

    H, W, C, M = cfg["H"], cfg["W"], cfg["C"], cfg["M"]
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=device))
    A = anchors.shape[0]
    sets = []
    for k in range(nbuf):
        clas, reg = device_activations(B, A, C, seed + 17 * k, device, mu=-4.6, logits=logits)
        gb, gc = syn.make_targets(B, M, H, W, C, seed=seed + 17 * k)
        sets.append((clas, reg, gb.to(device), gc.to(device)))
    return anchors, sets


MODE = {"exchange": "pipelined"}
EXCHANGE = {"kind": None}   # how the loss scalars were summed over the ranks in the last time_loss_graph call
PEER = {}                   # the PeerExchange shared by the benchmark's loss objects (its construction is a collective)


RANK_MS = {}   # time_loss_graph: per-rank ms/step of its last multi-rank call


def time_loss_graph(anchors, sets, steps, warmup, device, world, from_logits=False):
    """`steps` replays of the captured step (assign kernel + fused loss fwd/bwd/reduction kernel),
    rotating over the input sets; with several ranks each step ends with the 12-byte loss exchange."""
    import torch
    import torch.distributed as dist

    from neuralnetworklibrary_b200.vision import SSD_loss, reduce_loss_scalars

    # Several ranks: the 12-byte loss exchange.  MODE["exchange"]:
    #   "stream" rn_peer_exchange -- one kernel over peer-mapped memory -- on a side stream: step k+1's kernels do not
    #            wait for step k's exchange (nothing on the GPU consumes the summed loss), so ranks may drift by a step
    #            instead of synchronising every 0.36 ms;
    #   "graph"  the same kernel INSIDE the step's CUDA graph (lowest latency to the global loss; every step then costs the
    #            slowest rank's time);
    #   "nccl"   all_gather_into_tensor on a side stream (also the fallback when symmetric memory is unavailable);
    #   "pipelined" (default) the exchange of step k runs on a parallel branch at the START of step k+1's CUDA graph
    #            (rn_peer_exchange_to beside the assignment and streaming kernels), so neither the NVLink round trip nor the wait
    #            for the peers sits behind the final reduction; the newest step's sums are exchanged once more, explicitly,
    #            at the end of the timed region.  2 GPUs: 0.3458 ms against 0.3474 (stream) and 0.3428 (one GPU alone).
    from neuralnetworklibrary_b200.vision import PeerExchange
    EXCHANGE["kind"] = "none (one rank)"
    mode = MODE["exchange"] if world > 1 else "none"
    px = None
    if mode in ("stream", "graph", "pipelined"):
        if "obj" not in PEER:
            PEER["obj"] = PeerExchange.create()          # a collective; None on every rank if any rank failed
        px = PEER["obj"]
        if px is None:
            mode = "nccl"
    B_glob = sets[0][0].shape[0] * world
    if mode == "graph":
        loss_fn = SSD_loss(global_batch=B_glob, from_logits=from_logits, distributed=True, peer_exchange=px)
        EXCHANGE["kind"] = "rn_peer_exchange: one kernel over peer-mapped (symmetric) memory inside the step's CUDA graph"
    elif mode == "pipelined":
        loss_fn = SSD_loss(global_batch=B_glob, from_logits=from_logits, distributed=True, peer_exchange=px)
        EXCHANGE["kind"] = ("rn_peer_exchange_to (one kernel over peer-mapped memory) on a parallel branch at the start of the NEXT "
                            "step's CUDA graph, beside its assignment and streaming kernels; the last step's sums are exchanged "
                            "explicitly at the end of the timed region")
    else:
        loss_fn = SSD_loss(global_batch=B_glob, from_logits=from_logits)
        if mode == "stream":
            EXCHANGE["kind"] = "rn_peer_exchange: one kernel over peer-mapped (symmetric) memory on a side stream (pipelined with the next step)"
        elif mode == "nccl":
            EXCHANGE["kind"] = "nccl all_gather_into_tensor on a side stream"
    in_graph = mode in ("graph", "pipelined")
    caps = [loss_fn.capture([anchors, reg, clas], [gb, gc], pipelined_exchange=(mode == "pipelined")) for clas, reg, gb, gc in sets]

    # Side-stream exchange: each captured step owns its out3 buffer, and a step waits for the exchange that last read that
    # buffer before overwriting it.
    comm = torch.cuda.Stream(device=device) if world > 1 else None
    ready = [None] * len(caps)   # event: the exchange that read caps[i].out3 has finished
    totals = [torch.empty(3, dtype=torch.float32, device=device) for _ in caps]

    def step(k):
        i = k % len(caps)
        cap = caps[i]
        if world == 1 or in_graph:
            cap.replay()
            return cap.out3
        main = torch.cuda.current_stream(device)
        if ready[i] is not None:
            main.wait_event(ready[i])
        cap.replay()
        done = torch.cuda.Event()
        done.record(main)
        with torch.cuda.stream(comm):
            comm.wait_event(done)
            if mode == "stream":
                totals[i].copy_(cap.out3)
                px(totals[i])
            else:
                totals[i] = reduce_loss_scalars(cap.out3)
            ready[i] = torch.cuda.Event()
            ready[i].record(comm)
        return totals[i]

    for k in range(warmup):
        step(k)
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(device)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(steps):
        out3 = step(warmup + k)
    if mode == "pipelined":
        out3 = caps[(warmup + steps - 1) % len(caps)].total()   # the newest step's sums: the one wait for the peers
    if world > 1 and not in_graph:
        torch.cuda.current_stream(device).wait_stream(comm)   # the timed region ends when the last exchange has
    t1.record()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    total_ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([total_ms], device=device)
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        RANK_MS["last"] = [round(float(v.item()) / steps, 4) for v in every]   # each rank's own device time per step
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    launches = caps[0].kernels_per_replay * steps + (steps if (world > 1 and mode == "stream") else 0) \
        + (1 if mode == "pipelined" else 0)   # + rn_peer_exchange_kernel per step / once more for the last step
    return total_ms, launches, float(out3[0].item())


def time_loss_levels(cfg, B, steps, warmup, device, from_logits):
    """The loss step on the heads' NCHW level tensors (rn_assign + rn_loss_levels + final reduction) as CUDA-graph
    replays; 2 rotating input sets (each larger than L2)."""
    import torch

    from tests import synth as syn
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    from neuralnetworklibrary_b200.vision import SSD_loss, level_shapes

    H, W, C, M = cfg["H"], cfg["W"], cfg["C"], cfg["M"]
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=device))
    g = torch.Generator(device=device).manual_seed(1007)
    sets = []
    for k in range(2):
        clas = []
        for shp in level_shapes(H, W, 9, C):
            z = torch.randn((B,) + shp, generator=g, device=device) - 4.6
            clas.append(z if from_logits else torch.sigmoid_(z))
        reg = [torch.randn((B,) + shp, generator=g, device=device) * 0.5 for shp in level_shapes(H, W, 9, 4)]
        gb, gc = syn.make_targets(B, M, H, W, C, seed=1007 + k)
        sets.append((reg, clas, gb.to(device), gc.to(device)))
    loss_fn = SSD_loss(global_batch=B, from_logits=from_logits)
    graphs = [loss_fn.capture([anchors, reg, clas], [gb, gc]) for reg, clas, gb, gc in sets]   # one CUDA graph per input set
    outs = [(gr.out3,) for gr in graphs]

    def step(k):
        graphs[k % 2].replay()
        return outs[k % 2][0]

    for k in range(warmup):
        step(k)
    torch.cuda.synchronize(device)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(steps):
        out3 = step(warmup + k)
    t1.record()
    torch.cuda.synchronize(device)
    # what this path removes from the model: the heads' sigmoid + permute(0,2,3,1).contiguous().view() and the cat over
    # levels (retinanet.py:215-217, :258, :286-295; Vision.py:1467-1468), forward only, with torch's own kernels
    reg, clas, _, _ = sets[0]
    K = 9

    def layout_ops():
        c = torch.cat([(torch.sigmoid(x) if from_logits else x).permute(0, 2, 3, 1).contiguous().view(x.shape[0], -1, C) for x in clas], dim=1)
        r = torch.cat([x.permute(0, 2, 3, 1).contiguous().view(x.shape[0], -1, 4) for x in reg], dim=1)
        return c, r

    layout_ms = None
    try:
        for _ in range(2):
            layout_ops()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            layout_ops()
        e1.record()
        torch.cuda.synchronize(device)
        layout_ms = e0.elapsed_time(e1) / 5
    except Exception:
        pass
    return t0.elapsed_time(t1), float(out3[0].item()), layout_ms


def time_loss_eager(anchors, sets, steps, warmup, device):
    """The same step through SSD_loss(...) + loss.backward() call by call.  For the roofline of the dominant kernel the one
    library call of the step (rn_loss_step) is replaced, for this pass only, by the two calls it makes itself -- rn_assign, then
    rn_loss -- with CUDA events around rn_loss (the streaming kernel + its small final reduction), same buffers, same stream."""
    import torch

    from neuralnetworklibrary_b200 import _lib
    from neuralnetworklibrary_b200.vision import SSD_loss

    loss_fn = SSD_loss()
    lib = _lib.load()
    ev = []
    orig = lib.rn_loss_step
    scratch = {}

    def timed_rn_loss(clas, reg, gtb, gtc, B, A, C, M, H, W, base, K, table, pos, neg, alpha, gamma, beta, Bg, logits, dclas, dreg,
                      probs, out3, npos, matches, state, state_n, ws, ws_n, stream):
        if logits or lib.rn_get_option(b"step_fused") > 0 or lib.rn_get_option(b"step_bytemap") > 0:
            return orig(clas, reg, gtb, gtc, B, A, C, M, H, W, base, K, table, pos, neg, alpha, gamma, beta, Bg, logits, dclas,
                        dreg, probs, out3, npos, matches, state, state_n, ws, ws_n, stream)
        key = (B, A, C)
        if key not in scratch:
            scratch[key] = (torch.empty((B, A), dtype=torch.int32, device=device),
                            torch.empty(int(lib.rn_loss_workspace_bytes(B, A, C)), dtype=torch.uint8, device=device))
        m32, lws = scratch[key]
        mp = matches if matches is not None and matches.value else _lib.ptr(m32)
        rc = lib.rn_assign(gtb, gtc, B, M, H, W, base, K, table, A, pos, neg, mp, npos, None, stream)
        if rc:
            return rc
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.rn_loss(clas, reg, gtb, gtc, mp, npos, B, A, C, M, H, W, base, K, table, alpha, gamma, beta, Bg, dclas, dreg, out3,
                         _lib.ptr(lws), lws.numel(), stream)
        e1.record()
        ev.append((e0, e1))
        return rc

    leaves = [(clas.detach().requires_grad_(True), reg.detach().requires_grad_(True), gb, gc) for clas, reg, gb, gc in sets]

    def step(k):
        clas, reg, gb, gc = leaves[k % len(leaves)]
        clas.grad = None
        reg.grad = None
        loss = loss_fn([anchors, reg, clas], [gb, gc])
        loss.backward()
        return loss

    # (1) the public API as it is, uninstrumented: this is the eager number
    for k in range(warmup):
        step(k)
    torch.cuda.synchronize(device)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(steps):
        last = step(warmup + k)
    t1.record()
    torch.cuda.synchronize(device)
    # (2) the same pass with the step split and events around rn_loss: only the kernel time is taken from it (the events
    # between the assignment and the loss kernel break their programmatic dependent launch, so this pass is a little slower)
    lib.rn_loss_step = timed_rn_loss
    try:
        for k in range(min(warmup, 3)):
            step(k)
        del ev[:]
        for k in range(steps):
            step(warmup + k)
        torch.cuda.synchronize(device)
    finally:
        lib.rn_loss_step = orig
    kern_ms = statistics.mean(a.elapsed_time(b) for a, b in ev)
    return t0.elapsed_time(t1), kern_ms, float(last.item())


def time_loss_e2e(cfg, B, steps, warmup, device, world, seed=1002):
    """Same step, but inputs start in pinned host memory every step and the loss is read back."""
    import torch
    import torch.distributed as dist

    from tests import synth as syn
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    from neuralnetworklibrary_b200.vision import SSD_loss

    H, W, C, M = cfg["H"], cfg["W"], cfg["C"], cfg["M"]
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=device))
    A = anchors.shape[0]
    clas_d, reg_d = device_activations(B, A, C, seed, device, mu=-4.6)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=seed)
    h_clas = torch.empty(clas_d.shape, dtype=torch.float32, pin_memory=True).copy_(clas_d)
    h_reg = torch.empty(reg_d.shape, dtype=torch.float32, pin_memory=True).copy_(reg_d)
    h_gb, h_gc = gb.pin_memory(), gc.pin_memory()
    loss_fn = SSD_loss(distributed=world > 1, peer_exchange=PEER.get("obj") or False)
    h2d = h_clas.numel() * 4 + h_reg.numel() * 4 + h_gb.numel() * 4 + h_gc.numel() * 8
    out = []

    def step():
        clas_d.requires_grad_(False).copy_(h_clas, non_blocking=True)
        reg_d.requires_grad_(False).copy_(h_reg, non_blocking=True)
        g1, g2 = h_gb.to(device, non_blocking=True), h_gc.to(device, non_blocking=True)
        clas_d.grad = None
        reg_d.grad = None
        loss = loss_fn([anchors, reg_d.requires_grad_(True), clas_d.requires_grad_(True)], [g1, g2])
        loss.backward()
        out.append(loss.item())  # device -> host read of the step's result

    for _ in range(max(1, min(warmup, 2))):
        step()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize(device)
    ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, h2d, 4


def time_postproc(cfg, B, steps, warmup, device, seed=1004):
    """BASELINE.json configs[3]: decode + clip + threshold + top-k + NMS over a COCO-shaped batch."""
    import torch

    from neuralnetworklibrary_b200.retinanet import AnchorGenerator, BBoxPredictor

    H, W, C = cfg["H"], cfg["W"], cfg["C"]
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=device))
    A = anchors.shape[0]
    clas, reg = device_activations(B, A, C, seed, device, mu=-6.0)
    bp = BBoxPredictor()
    # (a) device time of the library call alone (no host synchronisation between calls)
    for _ in range(warmup):
        bp.predict_device(H, W, reg, clas, anchors)
    torch.cuda.synchronize(device)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        bp.predict_device(H, W, reg, clas, anchors)
    t1.record()
    torch.cuda.synchronize(device)
    dev_ms = t0.elapsed_time(t1)
    # (b) the public call including the copy of the result lists to the host (wall clock)
    w0 = time.perf_counter()
    for _ in range(steps):
        out = bp.predict_arrays(H, W, reg, clas, anchors)
    wall_ms = (time.perf_counter() - w0) * 1e3
    return dev_ms, wall_ms, int(out["n_candidates"].mean()), int(out["counts"].sum()), A


def time_postproc_levels(cfg, B, steps, warmup, device, from_logits, seed=1011):
    """configs[3] on the heads' NCHW level tensors (rn_postproc_levels), plus the torch layout ops it removes
    (sigmoid if logits, permute/contiguous/view, cat) on the same tensors."""
    import torch

    from neuralnetworklibrary_b200.retinanet import AnchorGenerator, BBoxPredictor
    from neuralnetworklibrary_b200.vision import level_shapes

    H, W, C = cfg["H"], cfg["W"], cfg["C"]
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=device))
    g = torch.Generator(device=device).manual_seed(seed)
    clas = []
    for shp in level_shapes(H, W, 9, C):
        z = torch.randn((B,) + shp, generator=g, device=device) - 6.0
        clas.append(z if from_logits else torch.sigmoid_(z))
    reg = [torch.randn((B,) + shp, generator=g, device=device) * 0.5 for shp in level_shapes(H, W, 9, 4)]
    bp = BBoxPredictor()
    bp.from_logits = from_logits
    for _ in range(warmup):
        bp.predict_device(H, W, reg, clas, anchors)
    torch.cuda.synchronize(device)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        bp.predict_device(H, W, reg, clas, anchors)
    t1.record()
    torch.cuda.synchronize(device)
    dev_ms = t0.elapsed_time(t1) / steps
    out = bp.predict_arrays(H, W, reg, clas, anchors)
    lay_ms = None
    try:
        bp.flatten_levels(reg, clas)
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            bp.flatten_levels(reg, clas)
        e1.record()
        torch.cuda.synchronize(device)
        lay_ms = e0.elapsed_time(e1) / 3
    except Exception:
        pass
    return dev_ms, lay_ms, int(out["n_candidates"].mean()), int(out["counts"].sum())


def time_aux_kernels(device, peak):
    """Device time, algorithmic bytes and HBM fraction of the kernels around the two hot paths (SURVEY.md section 8f rows
    2-4): rn_max_overlaps, rn_stage_targets, rn_stage_images (fp32 and uint8 upload), rn_nms_batch (TTA merge) and the
    mAP pair rn_map_match + rn_map_ap.  Most are latency-bound (a few microseconds); the fraction says so."""
    import numpy as np
    import torch

    from neuralnetworklibrary_b200 import metrics
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    from neuralnetworklibrary_b200.vision import ComputeMaxOverlaps, merge_tta_predictions, stage_images, stage_targets
    from tests import synth as syn

    out = {}

    def timed(fn, reps=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(device)
        return e0.elapsed_time(e1) / reps

    def entry(ms, nbytes, note):
        gbs = nbytes / (ms * 1e-3) / 1e9
        return {"ms": round(ms, 4), "algorithmic_bytes": int(nbytes), "gb_per_s": round(gbs, 1), "hbm_frac": round(gbs / peak, 4), "note": note}

    H, W, C, M, B = COCO["H"], COCO["W"], COCO["C"], COCO["M"], 16
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=device))
    gb, gc = syn.make_targets(B, M, H, W, C, seed=21)
    gbd, gcd = gb.to(device), gc.to(device)
    lib = __import__("neuralnetworklibrary_b200._lib", fromlist=["x"])
    L = lib.load()
    mo = torch.empty((B, M), dtype=torch.float32, device=device)
    from neuralnetworklibrary_b200.retinanet import anchor_args
    Hh, Ww, base, K, table, A = anchor_args(anchors)
    ms = timed(lambda: lib.check(L.rn_max_overlaps(lib.ptr(gbd), lib.ptr(gcd), B, M, Hh, Ww, base, K, table, A, lib.ptr(mo), lib.stream_ptr(device))))
    out["rn_max_overlaps"] = entry(ms, B * M * 24, "compute bound: B*A*M IoU evaluations (%.0f M) on generated anchors, 24 B per box in/out" % (B * A * M / 1e6))

    rng = np.random.RandomState(4)
    boxes = [rng.uniform(0, 500, size=(int(rng.randint(1, M + 1)), 4)) for _ in range(B)]
    cats = [rng.randint(0, C, size=len(b)) for b in boxes]
    scales = rng.uniform(0.5, 1.5, size=B)
    ms = timed(lambda: stage_targets(boxes, cats, scales, 1.05, 3, 5, device=device))
    out["stage_targets (host packing + H2D + rn_stage_targets)"] = entry(ms, B * M * 24 + sum(len(b) for b in boxes) * 40, "wall-clock dominated by the host packing; the kernel is one launch of B*M threads")

    imgs = [rng.rand(int(rng.randint(700, 801)), int(rng.randint(1100, 1334)), 3).astype(np.float32) for _ in range(8)]
    px = sum(im.size for im in imgs)
    res = stage_images(imgs, 4, 6, device=device)
    obytes = res.numel() * 4

    def kernel_only(u8):
        # the kernel alone on device-resident pixels (the public call also packs and uploads)
        import ctypes as Cc
        flat = torch.from_numpy(np.concatenate([(im * 255).astype(np.uint8).reshape(-1) if u8 else im.reshape(-1) for im in imgs])).to(device)
        offs = np.zeros(len(imgs), np.int64)
        offs[1:] = np.cumsum([im.size for im in imgs])[:-1]
        d_off = torch.from_numpy(offs).to(device)
        d_dim = torch.from_numpy(np.array([[im.shape[0], im.shape[1]] for im in imgs], np.int32)).to(device)
        o = torch.empty_like(res)
        if u8:
            f = lambda: lib.check(L.rn_stage_images_u8(lib.ptr(flat), lib.ptr(d_off), lib.ptr(d_dim), len(imgs), 3, res.shape[2], res.shape[3], 4, 6, None, None, lib.ptr(o), lib.stream_ptr(device)))
        else:
            f = lambda: lib.check(L.rn_stage_images(lib.ptr(flat), lib.ptr(d_off), lib.ptr(d_dim), len(imgs), 3, res.shape[2], res.shape[3], 4, 6, lib.ptr(o), lib.stream_ptr(device)))
        return timed(f)

    out["rn_stage_images fp32 (kernel, 8 COCO-sized images)"] = entry(kernel_only(False), px * 4 + obytes, "read HWC fp32 + write padded CHW fp32")
    out["rn_stage_images uint8 (kernel, 8 COCO-sized images)"] = entry(kernel_only(True), px + obytes, "read HWC uint8 (4x fewer upload bytes) + write padded CHW fp32")
    ms_pub32 = timed(lambda: stage_images(imgs, 4, 6, device=device), reps=3, warm=1)
    imgs8 = [(im * 255).astype(np.uint8) for im in imgs]
    ms_pub8 = timed(lambda: stage_images(imgs8, 4, 6, device=device), reps=3, warm=1)
    out["stage_images public call (pack + H2D + kernel)"] = {"fp32_ms": round(ms_pub32, 3), "uint8_ms": round(ms_pub8, 3), "h2d_bytes_fp32": px * 4, "h2d_bytes_uint8": px}

    # TTA merge: 64 images x 5 passes x ~20 boxes
    passes = []
    for p in range(5):
        per = []
        for l in range(64):
            n = int(rng.randint(5, 21))
            xy = rng.uniform(0, 600, (n, 2)); wh = rng.uniform(20, 200, (n, 2))
            per.append([list(np.concatenate([xy, xy + wh], 1).astype(np.float32)), list(rng.randint(0, 80, n).astype(np.int64)),
                        list(rng.uniform(0.05, 1, n).astype(np.float32))])
        passes.append(per)
    ms = timed(lambda: merge_tta_predictions(passes, device=device), reps=5, warm=2)
    out["merge_tta_predictions (64 images x 5 passes: pack + H2D + rn_nms_batch + one D2H)"] = {"ms": round(ms, 3), "launches": 3, "d2h_copies": 1}

    # mAP: 2000 images x 80 classes x 10 thresholds
    N = 2000
    predictions, targets = [], []
    for i in range(N):
        nt = int(rng.randint(0, 8))
        xy = rng.uniform(0, 600, size=(nt, 2)); wh = rng.uniform(20, 300, size=(nt, 2))
        tb = np.concatenate([xy, xy + wh], 1); tcs = rng.randint(0, C, size=nt)
        targets.append([(tb[j], int(tcs[j])) for j in range(nt)])
        pb, pc, ps = [], [], []
        for j in range(nt):
            for _ in range(int(rng.randint(0, 4))):
                pb.append((tb[j] + rng.normal(0, 10, size=4)).astype(np.float32)); pc.append(np.int64(tcs[j])); ps.append(np.float32(rng.uniform(0.05, 1)))
        predictions.append([pb, pc, ps])
    t0 = time.perf_counter()
    tab = metrics.mAP_table(predictions, targets, C, metrics.COCO_thresholds, device=device)
    wall = (time.perf_counter() - t0) * 1e3
    out["mAP_table (2000 images x 80 classes x 10 thresholds: rn_map_match + rn_map_ap)"] = {
        "wall_ms_incl_host_packing": round(wall, 1), "table_mean": float(np.nanmean(tab)), "launches": 2, "d2h_copies": 1}
    return out


def cpu_baseline_loss(cfg, max_images=None):
    """The CPU oracle on a bounded sample of the same workload (one image per host thread)."""
    import numpy as np

    from tests import synth as syn
    from oracle import oracle as orc

    H, W, C, M = cfg["H"], cfg["W"], cfg["C"], cfg["M"]
    threads = orc.num_threads()
    n = max(2, min(threads, max_images or 32))
    an = orc.anchors(H, W)
    A = an.shape[0]
    rng = np.random.default_rng(1002)
    clas = (1.0 / (1.0 + np.exp(-(rng.standard_normal((n, A, C), dtype=np.float32) - np.float32(4.6))))).astype(np.float32)
    reg = rng.standard_normal((n, A, 4), dtype=np.float32) * np.float32(0.5)
    gb, gc = syn.make_targets(n, M, H, W, C, seed=1002)
    orc.loss(an[:1024], clas[:1, :1024], reg[:1, :1024], gb.numpy()[:1], gc.numpy()[:1])  # warm the library
    gbn, gcn = gb.numpy(), gc.numpy()
    t0 = time.perf_counter()
    passes = 0
    while True:   # bounded sample: whole passes over the batch until ~2 s of wall time (x cores of CPU work)
        orc.loss(an, clas, reg, gbn, gcn)
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= 2.0 or passes >= 64:
            break
    # BASELINE.json configs[0]: the CPU-runnable case, B=2, 512x512, 20 classes, <= 10 GT boxes / image
    an1 = orc.anchors(512, 512)
    gb1, gc1 = syn.make_targets(2, 10, 512, 512, 20, seed=1001)
    c1, r1 = syn.make_train_activations(2, an1.shape[0], 20, seed=1001)
    t1 = time.perf_counter()
    orc.loss(an1, c1.numpy(), r1.numpy(), gb1.numpy(), gc1.numpy())
    cfg1_ms = (time.perf_counter() - t1) * 1e3
    return n * passes / dt, min(threads, n), ("oracle port, %d passes over %d COCO-shaped images (800x1344, C=80) fwd+bwd, 1 image per "
                                              "thread, %.1f s wall; configs[0] (B=2, 512x512, C=20) takes %.1f ms"
                                              % (passes, n, dt, cfg1_ms))


def _reference_loss_inputs(cfg, n, seed=1002):
    """n COCO-shaped images as host tensors (numpy RNG; the same generator the C port's baseline uses)."""
    import numpy as np
    import torch

    from oracle import oracle as orc
    from tests import synth as syn

    H, W, C, M = cfg["H"], cfg["W"], cfg["C"], cfg["M"]
    an = orc.anchors(H, W)
    A = an.shape[0]
    rng = np.random.default_rng(seed)
    clas = (1.0 / (1.0 + np.exp(-(rng.standard_normal((n, A, C), dtype=np.float32) - np.float32(4.6))))).astype(np.float32)
    reg = rng.standard_normal((n, A, 4), dtype=np.float32) * np.float32(0.5)
    gb, gc = syn.make_targets(n, M, H, W, C, seed=seed)
    return torch.from_numpy(an), torch.from_numpy(clas), torch.from_numpy(reg), gb, gc


def reference_available():
    from oracle import ref_shim
    return ref_shim.available()


def reference_cpu_loss(cfg, n, steps, warmup):
    """THE REFERENCE ITSELF (unmodified sources, oracle/_ref or /root/reference): SSD_loss(...)(activ, target) +
    loss.backward() (Applications/Vision.py:1607-1644, General/Learner.py:513-514) on the host cores, torch CPU kernels
    with all intra-op threads, `steps` passes over n COCO-shaped images.  Returns (images/s, threads, seconds)."""
    import torch

    from oracle import ref_shim

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    an, clas, reg, gb, gc = _reference_loss_inputs(cfg, n)
    with ref_shim.cpu_mode():
        _, vis = ref_shim.load()
        f = vis.SSD_loss()

        def step():
            cl, rg = clas.clone().requires_grad_(True), reg.clone().requires_grad_(True)
            loss = f([an, rg, cl], [gb, gc])
            loss.backward()
            return float(loss.item())

        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            last = step()
        dt = time.perf_counter() - t0
    return n * steps / dt, torch.get_num_threads(), dt, last


def reference_cuda_loss(cfg, device, n=4, reps=3):
    """The unmodified reference in its intended mode: the same call, tensors and model outputs on the B200, torch's CUDA
    kernels (per-image Python loop, ~2 host syncs per image).  images/s fwd+bwd."""
    import torch

    from oracle import ref_shim

    _, vis = ref_shim.load()
    an, clas, reg, gb, gc = (t.to(device) for t in _reference_loss_inputs(cfg, n, seed=1009))
    f = vis.SSD_loss()

    def step():
        cl, rg = clas.clone().requires_grad_(True), reg.clone().requires_grad_(True)
        loss = f([an, rg, cl], [gb, gc])
        loss.backward()
        return loss

    step()
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    torch.cuda.synchronize(device)
    dt = time.perf_counter() - t0
    return n * reps / dt, "%d passes over %d COCO-shaped images; the unmodified reference (Vision.py:1607-1644 + autograd) on cuda, torch %s" % (reps, n, torch.__version__)


def parity_check(anchors, cap_set, device):
    """Outside the timed region: the FIRST timed batch (same device tensors the graph replays) against the CPU oracle --
    assignment of every anchor of every image bit-exact, positive counts equal, the three loss scalars rtol 1e-5, and the
    gradients of the first and last image rtol 1e-5 (scaled for dreg, host libm).  Returns a dict for the JSON line."""
    import numpy as np
    import torch

    from neuralnetworklibrary_b200.vision import SSD_loss
    from oracle import oracle as orc
    from tests import synth as syn

    clas, reg, gb, gc = cap_set
    B, A, C = clas.shape
    H, W = anchors._rn_geom.H, anchors._rn_geom.W
    cd, rd = clas.detach().clone().requires_grad_(True), reg.detach().clone().requires_grad_(True)
    f = SSD_loss(keep_matches=True)   # the fused step also writes its dense assignment for this check
    loss = f([anchors, rd, cd], [gb, gc])
    loss.backward()
    matches, npos = f.last_assignment
    got3 = np.array([loss.item(), f.reg_loss.item(), f.clas_loss.item()], np.float32)
    an = orc.anchors(H, W)
    o = orc.loss(an, clas.cpu().numpy(), reg.cpu().numpy(), gb.cpu().numpy(), gc.cpu().numpy(), want_matches=True)
    ok_m = bool(np.array_equal(matches.cpu().numpy(), o["matches"]))
    ok_n = bool(np.array_equal(npos.cpu().numpy(), o["npos"])) if "npos" in o else bool(
        np.array_equal(npos.cpu().numpy(), (o["matches"] >= 0).sum(axis=1)))
    rel3 = float(np.max(np.abs(got3.astype(np.float64) - o["out3"]) / np.abs(o["out3"])))
    ok_g = True
    try:
        for i in (0, B - 1):
            syn.assert_rel(cd.grad[i].cpu().numpy(), o["dclas"][i], what="dclas")
            syn.assert_dreg_close(rd.grad[i:i + 1].cpu().numpy(), o["dreg"][i:i + 1])
    except AssertionError:
        ok_g = False
    del cd, rd
    return {"parity_checked": bool(ok_m and ok_n and rel3 <= 1e-5 and ok_g), "matches_equal": ok_m, "npos_equal": ok_n,
            "loss_max_rel_err": rel3, "grads_rtol_1e-5": ok_g,
            "what": "first timed batch (B=%d, A=%d, C=%d) vs the CPU oracle, outside the timed region" % (B, A, C)}


def torch_cuda_baseline(cfg, device, n=4):
    """The reference's own sequence of torch operations (tests/torch_restatement.py, pinned bit for bit against the
    unmodified reference on CPU) executed with torch's CUDA kernels: what the reference does on a GPU, its intended
    mode.  The reference itself is Python and cannot travel to the GPU box.  images/s fwd+bwd."""
    import torch

    from tests import synth as syn
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    from tests import torch_restatement as tr

    H, W, C, M = cfg["H"], cfg["W"], cfg["C"], cfg["M"]
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=device)).clone()
    clas, reg = device_activations(n, anchors.shape[0], C, 1009, device, mu=-4.6)
    gb, gc = syn.make_targets(n, M, H, W, C, seed=1009)
    gb, gc = gb.to(device), gc.to(device)

    def step():
        cl, rg = clas.clone().requires_grad_(True), reg.clone().requires_grad_(True)
        loss = tr.ssd_loss(anchors, rg, cl, gb, gc)[0]
        loss.backward()
        return loss

    step()
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        step()
    torch.cuda.synchronize(device)
    dt = time.perf_counter() - t0
    return n * reps / dt, "%d passes over %d COCO-shaped images, torch %s CUDA kernels, per-image Python loop as in the reference" % (reps, n, torch.__version__)


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    B = args.per_gpu_batch
    peak, peak_src = peaks()

    anchors, sets = make_loss_sets(COCO, B, device, 2, 1002 + 1000 * rank)
    A = anchors.shape[0]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms, launches, last_loss = time_loss_graph(anchors, sets, args.steps, args.warmup, device, world)
    rank_ms = RANK_MS.get("last")
    exchange_kind = EXCHANGE["kind"]
    clocks = sampler.stop() if rank == 0 else None
    images = B * world * args.steps
    value = images / (total_ms * 1e-3)
    eager_ms, kern_ms, _ = time_loss_eager(anchors, sets, args.steps, args.warmup, device)
    parity = None
    if rank == 0 and not args.no_parity_check:
        try:
            parity = parity_check(anchors, sets[0], device)
        except Exception as exc:
            parity = {"parity_checked": False, "error": repr(exc)}
    del sets

    e2e_ms, h2d, d2h = time_loss_e2e(COCO, B, max(2, min(args.steps, 5)), args.warmup, device, world)
    e2e_steps = max(2, min(args.steps, 5))
    e2e_value = B * world * e2e_steps / (e2e_ms * 1e-3)

    # BASELINE.json configs[4]: a global COCO batch of 256 images sharded by image over the ranks (all ranks take part:
    # every step ends with the loss exchange); one input set per rank, 256/world x 64.5 MB >> L2.
    b256 = None
    if 256 % world == 0 and not args.no_b256:
        try:
            import torch
            torch.cuda.empty_cache()
            an256, sets256 = make_loss_sets(COCO, 256 // world, device, 1, 3002 + 1000 * rank)
            steps256 = max(3, args.steps // 4)
            ms256, _, _ = time_loss_graph(an256, sets256, steps256, 3, device, world)
            del sets256
            torch.cuda.empty_cache()
            b256 = {"workload": "COCO shape, global batch 256 image-sharded over %d GPU(s) (%d images per GPU), loss fwd+bwd "
                                "+ loss exchange (BASELINE configs[4])" % (world, 256 // world),
                    "images_per_s": round(256 * steps256 / (ms256 * 1e-3), 1), "ms_per_step": round(ms256 / steps256, 4),
                    "steps": steps256, "scaling": "strong",
                    "roofline_frac_whole_step": round(loss_bytes(256 // world, A, COCO["C"]) * steps256 / (ms256 * 1e-3) / 1e9 / peak, 4)}
        except Exception as exc:
            b256 = {"error": repr(exc)}

    line = None
    if rank == 0:
        alg = loss_bytes(B, A, COCO["C"])
        achieved = alg / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": "images/sec loss fwd+bwd (COCO shape)", "value": round(value, 1), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "loss_exchange": exchange_kind,
            "ms_per_step_by_rank": rank_ms,   # N > 1: each rank's own device time; ms_per_step is their maximum
            "config": {"workload": "coco_loss_fwd_bwd: assign + focal/smooth-L1 fwd+bwd, B=%d/GPU, 800x1344, A=%d, C=80, M=20"
                                   % (B, A), "global_batch": B * world, "parallelism": "image-sharded dp%d" % world,
                       "l2": "inputs (1.1 GB/step, 2 rotating sets) larger than the 126 MB L2",
                       "api": "SSD_loss.capture(): CUDA-graph replay of rn_loss_step (sparse assignment: fill + one CTA per box, the fused loss fwd/bwd kernel, the final reduction)"},
            "eager": {"api": "SSD_loss()(...) + loss.backward(), call by call", "images_per_s": round(B * args.steps / (eager_ms * 1e-3), 1),
                      "ms_per_step": round(eager_ms / args.steps, 4)},
            "e2e": {"value": round(e2e_value, 1), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "h2d_gb_per_s_per_rank": round(h2d * e2e_steps / (e2e_ms * 1e-3) / 1e9, 2),
                    "limiter": "host->device copy of the step's inputs (pinned memory, PCIe / host fabric): %.1f GB/s per rank; the "
                               "kernels need %.2f ms of the %.1f ms step" % (h2d * e2e_steps / (e2e_ms * 1e-3) / 1e9, total_ms / args.steps, e2e_ms / e2e_steps)},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": measured_traffic("rn_loss_kernel<4,20,true,true,false>", B, 16),
                         "traffic_source": "ncu --set full capture, profiles/r02_traffic.json", "kernel": "rn_loss_kernel<4,20,true,true,false>",
                         "kernel_ms": round(kern_ms, 4), "algorithmic_bytes": alg, "peak_source": peak_src,
                         "timed": "CUDA events around every rn_loss launch (streaming kernel + final reduction) of the eager pass (same inputs, same process)",
                         "whole_step_frac": round(alg * args.steps / (total_ms * 1e-3) / 1e9 / peak, 4)},
            "clocks": clocks, "loss": last_loss,
        }
        if parity is not None:
            line["parity_checked"] = parity.pop("parity_checked")
            line["parity"] = parity
        if b256 is not None:
            line["coco_b256_sharded"] = b256
        # extra workloads (device-timed, single GPU share of the job): post-processing and Pascal loss
        try:
            pms, pwall, ncand, nkept, _ = time_postproc(COCO, args.postproc_batch, max(3, args.steps // 2), 3, device)
            psteps = max(3, args.steps // 2)
            palg = loss_bytes(args.postproc_batch, A, COCO["C"], grad=False)
            line["postproc"] = {"workload": "coco_postproc B=%d 800x1344 C=80 thresh=0.05 top_k=1000 max_boxes=20" % args.postproc_batch,
                                "images_per_s": round(args.postproc_batch * psteps / (pms * 1e-3), 1),
                                "images_per_s_e2e_wall": round(args.postproc_batch * psteps / (pwall * 1e-3), 1),
                                "ms_per_step": round(pms / psteps, 4), "candidates_per_image": ncand, "kept": nkept,
                                "roofline_frac_whole_call": round(palg / (pms / psteps * 1e-3) / 1e9 / peak, 4)}
        except Exception as exc:  # keep the headline line even if an extra fails
            line["postproc"] = {"error": repr(exc)}
        for key, fl in (("postproc_levels", False), ("postproc_levels_logits", True)):
            try:
                import torch
                torch.cuda.empty_cache()
                lms, lay, lcand, lkept = time_postproc_levels(COCO, args.postproc_batch, max(3, args.steps // 2), 3, device, fl)
                palg = loss_bytes(args.postproc_batch, A, COCO["C"], grad=False)
                line[key] = {"workload": "coco_postproc on the heads' NCHW level tensors%s, B=%d 800x1344 C=80"
                                         % (" (logits, sigmoid fused)" if fl else "", args.postproc_batch),
                             "images_per_s": round(args.postproc_batch / (lms * 1e-3), 1), "ms_per_step": round(lms, 4),
                             "candidates_per_image": lcand, "kept": lkept,
                             "roofline_frac_whole_call": round(palg / (lms * 1e-3) / 1e9 / peak, 4),
                             "removed_head_layout_ops_ms": None if lay is None else round(lay, 4)}
            except Exception as exc:
                line[key] = {"error": repr(exc)}
        try:
            pan, psets = make_loss_sets(PASCAL, 32, device, 4, 1003)
            pA = pan.shape[0]
            psteps = max(args.steps, 40)
            pt, _, _ = time_loss_graph(pan, psets, psteps, args.warmup, device, 1)
            pe, pk, _ = time_loss_eager(pan, psets, args.steps, args.warmup, device)
            del psets
            line["pascal"] = {"workload": "pascal_loss_fwd_bwd B=32 512x512 C=20 M=10 (BASELINE configs[1])",
                              "images_per_s": round(32 * psteps / (pt * 1e-3), 1), "ms_per_step": round(pt / psteps, 4),
                              "rn_loss_ms": round(pk, 4),
                              "eager_ms_per_step": round(pe / args.steps, 4),
                              "eager_note": "SSD_loss()(...) + backward() call by call: at this size the step is bound by the "
                                            "Python issue time of the autograd Function, not by the device",
                              "roofline_frac_rn_loss": round(loss_bytes(32, pA, 20) / (pk * 1e-3) / 1e9 / peak, 4),
                              "roofline_frac_whole_step": round(loss_bytes(32, pA, 20) * psteps / (pt * 1e-3) / 1e9 / peak, 4)}
        except Exception as exc:
            line["pascal"] = {"error": repr(exc)}
        try:  # the head's sigmoid fused into the loss kernel (logits in)
            lan, lsets = make_loss_sets(COCO, B, device, 2, 1005, logits=True)
            lt, _, _ = time_loss_graph(lan, lsets, args.steps, args.warmup, device, 1, from_logits=True)
            del lsets
            line["logits"] = {"workload": "coco_loss_fwd_bwd from logits (head sigmoid fused), B=%d 800x1344 C=80" % B,
                              "images_per_s": round(B * args.steps / (lt * 1e-3), 1), "ms_per_step": round(lt / args.steps, 4),
                              "roofline_frac_whole_step": round(loss_bytes(B, A, COCO["C"]) * args.steps / (lt * 1e-3) / 1e9 / peak, 4),
                              "note": "also removes one read+write pass over [B,A,C] from the model's forward and one from its "
                                      "backward (not counted here)"}
        except Exception as exc:
            line["logits"] = {"error": repr(exc)}
        for key, fl in (("levels", False), ("levels_logits", True)):
            try:  # SURVEY.md section 8f row 1, second half: the loss on the heads' NCHW level tensors
                vt, _, lay = time_loss_levels(COCO, B, args.steps, args.warmup, device, fl)
                line[key] = {"workload": "coco_loss_fwd_bwd on the heads' NCHW level tensors%s, B=%d 800x1344 C=80"
                                         % (" (logits, sigmoid fused)" if fl else "", B),
                             "images_per_s": round(B * args.steps / (vt * 1e-3), 1), "ms_per_step": round(vt / args.steps, 4),
                             "roofline_frac_whole_step": round(loss_bytes(B, A, COCO["C"]) * args.steps / (vt * 1e-3) / 1e9 / peak, 4),
                             "removed_head_layout_ops_fwd_ms": None if lay is None else round(lay, 4),
                             "note": "removes the heads' %spermute/contiguous/view and the cat over levels from the model "
                                     "(removed_head_layout_ops_fwd_ms: those torch ops, forward only, same tensors; their "
                                     "backward is removed as well)" % ("sigmoid, " if fl else "")}
            except Exception as exc:
                line[key] = {"error": repr(exc)}
        if world == 1 and not args.no_aux:
            try:
                import torch
                torch.cuda.empty_cache()
                line["aux_kernels"] = time_aux_kernels(device, peak)
            except Exception as exc:
                line["aux_kernels"] = {"error": repr(exc)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                tv, tsample = torch_cuda_baseline(COCO, device)
                line["torch_cuda_baseline"] = {"value": round(tv, 1), "unit": "images/s", "kind": "restatement", "sample": tsample}
            except Exception as exc:
                line["torch_cuda_baseline"] = {"error": repr(exc)}
        if world == 1 and not args.no_cpu_baseline:
            if reference_available():
                try:
                    rv, rsample = reference_cuda_loss(COCO, device)
                    line["reference_cuda"] = {"value": round(rv, 1), "unit": "images/s", "kind": "reference", "sample": rsample}
                except Exception as exc:
                    line["reference_cuda"] = {"error": repr(exc)}
            v, cores, sample = cpu_baseline_loss(COCO)
            port = {"value": round(v, 3), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
            line["cpu_baseline"] = port
            if reference_available():
                try:
                    cv, cthreads, cdt, _ = reference_cpu_loss(COCO, 2, 6, 1)
                    line["cpu_baseline"] = {"value": round(cv, 3), "unit": "images/s", "cores": cthreads, "kind": "reference",
                                            "sample": "the unmodified reference (SSD_loss + backward, torch CPU kernels, %d intra-op "
                                                      "threads): 6 passes over B=2 COCO-shaped images (800x1344, C=80, M=20), %.1f s"
                                                      % (cthreads, cdt)}
                    line["cpu_baseline_port"] = port
                except Exception as exc:
                    line["cpu_baseline_reference_error"] = repr(exc)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


# --------------------------------------------------------------------------------------------------
# reference arm: THE REFERENCE ITSELF on the host cores -- the unmodified sources staged in oracle/_ref (or the checkout
# at /root/reference), SSD_loss + backward with torch's CPU kernels on all threads.  If the sources are not staged the arm
# falls back to the C port of the algorithm (oracle/retina_oracle.c) and says so (`kind`).
# --------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return None
    H, W, C, M = COCO["H"], COCO["W"], COCO["C"], COCO["M"]
    port_v, port_cores, port_sample = cpu_baseline_loss(COCO)
    port = {"value": round(port_v, 3), "unit": "images/s", "cores": port_cores, "kind": "port", "sample": port_sample}
    A = 201600
    if reference_available():
        n = 2   # one step = a bounded sample: B=2 COCO-shaped images (the reference is a per-image Python loop: linear in B)
        value, threads, dt, _ = reference_cpu_loss(COCO, n, args.steps, min(args.warmup, 2))
        kind = "reference"
        sample = ("the unmodified reference (Applications/Vision.py SSD_loss + loss.backward()), torch CPU kernels with %d "
                  "intra-op threads, %d COCO-shaped images (800x1344, A=%d, C=80, M=20) per step" % (threads, n, A))
        cores = threads
        ms_per_step = dt / args.steps * 1e3
    else:
        value, cores, sample, kind = port_v, port_cores, port_sample, "port"
        ms_per_step = None
    line = {
        "impl": "reference", "metric": "images/sec loss fwd+bwd (COCO shape)", "value": round(value, 3), "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None if ms_per_step is None else round(ms_per_step, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "coco_loss_fwd_bwd: assign + focal/smooth-L1 fwd+bwd, 800x1344, A=%d, C=80, M=20" % A,
                   "note": "the reference's own CPU implementation of the path on the host cores; bounded sample per step"},
        "cpu_baseline": {"value": round(value, 3), "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
        "cpu_baseline_port": port,
        "e2e": {"value": round(value, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--per-gpu-batch", type=int, default=16)
    ap.add_argument("--postproc-batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-b256", action="store_true", help="skip the BASELINE configs[4] extra (256 images over the ranks)")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the oracle check of the first timed batch")
    ap.add_argument("--exchange", default="pipelined", choices=["stream", "graph", "nccl", "pipelined"],
                    help="multi-GPU loss exchange: rn_peer_exchange on a side stream / inside the step's graph, NCCL, or the "
                         "pipelined rn_peer_push / rn_peer_collect pair")
    ap.add_argument("--no-aux", action="store_true", help="skip the auxiliary-kernel measurements (section 8f rows 2-4)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    MODE["exchange"] = args.exchange
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
