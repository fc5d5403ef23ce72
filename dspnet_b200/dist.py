"""Image-sharded multi-GPU execution of the multibox path (one process per GPU, torch.distributed).

Every image is independent in all three operators (per-image loops at operator/multibox_target.cc:92 and
multibox_detection.cc:74), so a batch is partitioned into contiguous image slices, one per rank, with the anchors
replicated (regenerated locally by the prior kernel -- no broadcast).  There is no collective on the data path;
the single exchange step is the all-gather of what the consumers read back on the host in the reference:
the surviving detection rows (``det[:, 0] >= 0``, detect/multitask_detector.py:268-271) and the per-image target
statistics (what train/metric.py:35-46 reduces).  Training targets themselves stay on the GPU that owns the image.

``shard_slice`` / ``ShardedMultiBox`` are backend-agnostic host logic (tested on CPU with gloo, world_size 2, with
the oracle injected as the compute provider); ``DetectionGatherer`` is the plain CUDA/NCCL pipeline (compaction
kernel on the compute stream -> all_gather_into_tensor on a side stream, double buffered); ``P2PDetectionGatherer``
is what bench.py uses: one kernel that compacts and stores into every peer's buffer over NVLink, with sequence
numbers and acknowledgements as device-side flow control.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


def shard_slice(batch, world, rank):
    """Contiguous image slice [begin, end) of `rank`; the first ``batch % world`` ranks get one extra image."""
    base, extra = divmod(batch, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


class ShardedMultiBox:
    """Runs MultiBoxTarget / MultiBoxDetection on this rank's image slice and gathers the small results.

    ops: object with ``MultiBoxTarget`` and ``MultiBoxDetection`` callables (default: the CUDA operators of this
    package).  group: torch.distributed process group (default: WORLD).
    """

    def __init__(self, ops=None, group=None):
        if ops is None:
            from . import ops as _ops
            ops = _ops
        self.ops = ops
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def local(self, batch):
        return shard_slice(batch, self.world, self.rank)

    def _gather_rows(self, local, batch):
        """All-gather per-image rows whose slice sizes may differ by one across ranks."""
        if self.world == 1:
            return local
        sizes = [shard_slice(batch, self.world, r) for r in range(self.world)]
        widest = max(e - b for b, e in sizes)
        pad = torch.zeros((widest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(parts, pad, group=self.group)
        return torch.cat([p[: e - b] for p, (b, e) in zip(parts, sizes)], dim=0)

    def detection(self, cls_prob, loc_pred, anchor, max_rows, **params):
        """cls_prob / loc_pred hold the FULL batch (only this rank's slice is read).  Returns
        (rows (B, max_rows, 7), counts (B,)) for the whole batch on every rank, plus this rank's full output."""
        batch = cls_prob.shape[0]
        b, e = self.local(batch)
        out = self.ops.MultiBoxDetection(cls_prob[b:e], loc_pred[b:e], anchor, **params)
        out_t = torch.as_tensor(out)
        rows, counts = compact_rows(out_t, max_rows)
        return self._gather_rows(rows, batch), self._gather_rows(counts, batch), out

    def target(self, anchor, label, cls_pred, **params):
        """Returns this rank's [loc_target, loc_mask, cls_target] and the gathered per-image statistics
        (B, 3): [num_positive, num_negative, num_ignored] as train/metric.py derives them from cls_target."""
        batch = label.shape[0]
        b, e = self.local(batch)
        outs = self.ops.MultiBoxTarget(anchor, label[b:e], cls_pred[b:e], **params)
        ct = torch.as_tensor(outs[2])
        ignore = params.get("ignore_label", -1.0)
        stats = torch.stack([(ct > 0).sum(1), (ct == 0).sum(1), (ct == ignore).sum(1)], dim=1).to(torch.int32)
        return outs, self._gather_rows(stats, batch)


def compact_rows(out, max_rows):
    """Surviving rows (id >= 0) of every image in row order, at most max_rows, padded with -1, and their counts.
    CUDA tensors go through the det_compact kernel; host tensors (gloo tests) through torch indexing."""
    B, A, _ = out.shape
    if out.is_cuda:
        rows = torch.empty((B, max_rows, 7), dtype=torch.float32, device=out.device)
        counts = torch.empty((B,), dtype=torch.int32, device=out.device)
        with torch.cuda.device(out.device):
            _lib.check(_lib.lib().dspmb_detection_compact_f32(
                out.data_ptr(), None, B, A, max_rows, rows.data_ptr(), counts.data_ptr(),
                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return rows, counts
    rows = torch.full((B, max_rows, 7), -1.0, dtype=torch.float32)
    counts = torch.zeros((B,), dtype=torch.int32)
    for i in range(B):
        keep = out[i][out[i, :, 0] >= 0][:max_rows]
        rows[i, : keep.shape[0]] = keep
        counts[i] = keep.shape[0]
    return rows, counts


class DetectionGatherer:
    """Double-buffered compaction + NCCL all-gather of the detections, overlapped with the next step's kernels."""

    launches_per_submit = 1  # det_compact_kernel (the NCCL kernel is not ours)

    def __init__(self, batch, anchors, max_rows, device, world, group=None):
        self.B, self.A, self.K, self.device, self.world, self.group = batch, anchors, max_rows, device, world, group
        self.lib = _lib.lib()
        self.comm = torch.cuda.Stream(device)
        self.rows = [torch.empty((batch, max_rows, 7), dtype=torch.float32, device=device) for _ in range(2)]
        self.counts = [torch.empty((batch,), dtype=torch.int32, device=device) for _ in range(2)]
        self.all_rows = [torch.empty((world * batch, max_rows, 7), dtype=torch.float32, device=device) for _ in range(2)]
        self.all_counts = [torch.empty((world * batch,), dtype=torch.int32, device=device) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.done = [None, None]

    def submit(self, out, step):
        s = step & 1
        compute = torch.cuda.current_stream(self.device)
        if self.done[s] is not None:
            compute.wait_event(self.done[s])  # the previous collective on this buffer has consumed it
        _lib.check(self.lib.dspmb_detection_compact_f32(out.data_ptr(), None, self.B, self.A, self.K,
                                                        self.rows[s].data_ptr(), self.counts[s].data_ptr(),
                                                        ctypes.c_void_p(compute.cuda_stream)))
        self.ready[s].record(compute)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.ready[s])
            dist.all_gather_into_tensor(self.all_rows[s], self.rows[s], group=self.group)
            dist.all_gather_into_tensor(self.all_counts[s], self.counts[s], group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.comm)
            self.done[s] = ev
        return self.all_rows[s], self.all_counts[s]

    def drain(self):
        torch.cuda.current_stream(self.device).wait_stream(self.comm)


class P2PDetectionGatherer:
    """The exchange step without NCCL on the critical path: ``dspmb_detection_gather_f32`` compacts this rank's
    detections (and / or takes the per-image target statistics) and stores them straight into every peer's gather
    buffer over NVLink (CUDA IPC mapped peer memory).  torch.distributed is only used once, at construction, to
    exchange the 64-byte IPC handles.

    Contract (SPMD): every rank calls ``submit(out, step, stats=...)`` for every step, steps counting up by one; the
    tensors handed to submit(step) must stay untouched until submit(step + 2) has been called (the gather kernel reads
    them on a side stream, beside the next step's kernels).
    ``gathered(step)`` / ``gathered_stats(step)`` return this rank's copy of the whole batch for the two most recent
    steps.  The two slots are flow-controlled on the device: a generation is only overwritten after every rank has
    acknowledged it -- ``gathered`` acknowledges after its copy, and ``submit`` acknowledges on the consumer's behalf
    any generation that was never read (``release``), so a rank that runs ahead waits instead of overwriting.  All
    waits are bounded; ``check()`` reports a timeout.
    """

    launches_per_submit = 1  # det_gather_kernel (+ wait / ack kernels of one thread block each)

    def __init__(self, batch, anchors, max_rows, device, world, rank, group=None, stats_width=0):
        assert max_rows % 4 == 0
        self.B, self.A, self.K, self.SW = batch, anchors, max_rows, stats_width
        self.device, self.world, self.rank = torch.device(device), world, rank
        self.lib = _lib.lib()
        self.nbytes = self.lib.dspmb_gather_buffer_bytes(batch, max_rows, stats_width, world)
        base = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dspmb_p2p_alloc(self.nbytes, ctypes.byref(base), handle))
        self.local = base.value
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, handle.raw, group=group)
        else:
            handles[0] = handle.raw
        self.peers = (ctypes.c_void_p * world)()
        self._opened = []
        for r in range(world):
            if r == rank:
                self.peers[r] = self.local
            else:
                p = ctypes.c_void_p()
                with torch.cuda.device(self.device):
                    _lib.check(self.lib.dspmb_p2p_open(handles[r], ctypes.byref(p)))
                self.peers[r] = p.value
                self._opened.append(p.value)
        self.step_of = [None, None]     # step whose data currently lives in the slot
        self.unacked = [False, False]   # ... and has not been acknowledged by this rank yet
        # the exchange runs beside the next step's kernels, on a side stream owned by the library's gather context
        # (submit() is ONE C call: event edges, acknowledgement of an unread generation, gather kernel)
        with torch.cuda.device(self.device):
            self.ctx = self.lib.dspmb_gather_ctx_create()
        if not self.ctx:
            raise _lib.DspmbError(_lib.ERR_CUDA, self.lib.dspmb_last_error().decode())
        self.side = torch.cuda.ExternalStream(self.lib.dspmb_gather_ctx_side_stream(self.ctx), device=self.device)
        self.read = torch.cuda.Event()
        if world > 1:
            dist.barrier(group=group)  # every peer has mapped every buffer before the first store

    def _side(self):
        return ctypes.c_void_p(self.side.cuda_stream)

    def _wait(self, slot, stream):
        _lib.check(self.lib.dspmb_detection_gather_wait(self.local, self.B, self.K, self.SW, self.world, slot,
                                                        self.step_of[slot] + 1, stream))

    def _ack(self, slot):
        """On the side stream: this rank is done with the generation in `slot`."""
        _lib.check(self.lib.dspmb_detection_gather_ack(self.B, self.K, self.SW, self.rank, self.world, self.peers, slot,
                                                       self.step_of[slot] + 1, self._side()))
        self.unacked[slot] = False

    def release(self, slot):
        """Acknowledge the generation in `slot` without reading it (after all of it has arrived here)."""
        if self.step_of[slot] is not None and self.unacked[slot]:
            with torch.cuda.device(self.device):
                self._wait(slot, self._side())
                self._ack(slot)

    def submit(self, out, step, valid_count=None, stats=None):
        slot = step & 1
        prev = self.step_of[slot]
        assert prev is None or step == prev + 2, "submit() must be called for every step"
        compute = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            # one call: compute waits for the gather kernel that read `out` two steps ago; an unread generation in the
            # slot is acknowledged on the consumer's behalf (after it has arrived); side stream waits for compute;
            # gather kernel of this step
            _lib.check(self.lib.dspmb_gather_submit(
                self.ctx, out.data_ptr() if out is not None else None,
                valid_count.data_ptr() if valid_count is not None else None,
                stats.data_ptr() if (stats is not None and self.SW) else None, self.B, self.A, self.K, self.SW,
                self.rank, self.world, self.peers, slot, step + 1, 0 if prev is None else prev + 1,
                1 if (prev is not None and self.unacked[slot]) else 0, ctypes.c_void_p(compute.cuda_stream)))
        self.step_of[slot] = step
        self.unacked[slot] = True

    def _fetch(self, step, want_rows, want_stats):
        slot = step & 1
        assert self.step_of[slot] == step, "only the two most recent steps are held"
        cur = torch.cuda.current_stream(self.device)
        n = self.world * self.B
        rows = counts = stats = None
        with torch.cuda.device(self.device):
            cur.wait_stream(self.side)
            self._wait(slot, ctypes.c_void_p(cur.cuda_stream))
            if want_rows and self.K:
                rows = torch.empty((n, self.K, 7), dtype=torch.float32, device=self.device)
                counts = torch.empty((n,), dtype=torch.int32, device=self.device)
            if want_stats and self.SW:
                stats = torch.empty((n, self.SW), dtype=torch.int32, device=self.device)
            _lib.check(self.lib.dspmb_detection_gather_read(
                self.local, self.B, self.K, self.SW, self.world, slot,
                rows.data_ptr() if rows is not None else None, counts.data_ptr() if counts is not None else None,
                stats.data_ptr() if stats is not None else None, ctypes.c_void_p(cur.cuda_stream)))
            if self.unacked[slot]:  # acknowledged on the side stream, after the copies above
                self.read.record(cur)
                self.side.wait_event(self.read)
                self._ack(slot)
        return rows, counts, stats

    def gathered(self, step):
        """(rows (world*B, K, 7) float32, counts (world*B,) int32) copied out of this rank's buffer."""
        rows, counts, _ = self._fetch(step, True, False)
        return rows, counts

    def gathered_stats(self, step):
        """(world*B, stats_width) int32 target statistics of every image of the batch."""
        return self._fetch(step, False, True)[2]

    def drain(self):
        """The current stream waits until every submitted generation has fully arrived on this rank."""
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            cur.wait_stream(self.side)
            for slot in (0, 1):
                if self.step_of[slot] is not None:
                    self._wait(slot, ctypes.c_void_p(cur.cuda_stream))

    def check(self):
        """False if any bounded wait of the exchange ran out (synchronises the device)."""
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            return self.lib.dspmb_gather_error(self.local, self.B, self.K, self.SW, self.world) == 0

    def close(self):
        with torch.cuda.device(self.device):
            for slot in (0, 1):  # peers may still be waiting for this rank's acknowledgements
                self.release(slot)
            torch.cuda.synchronize(self.device)
            if self.world > 1:
                dist.barrier()
            if self.ctx:
                self.lib.dspmb_gather_ctx_destroy(self.ctx)
                self.ctx = None
            for p in self._opened:
                self.lib.dspmb_p2p_close(p)
            self._opened = []
            if self.local:
                self.lib.dspmb_p2p_free(self.local)
                self.local = None
