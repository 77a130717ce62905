"""Frame-sharded data parallelism: one process per GPU, torch.distributed over NCCL.

Every stage of the hot path is independent per frame except the frame-to-frame link, where the
last frame of rank r needs the boxes of the FIRST frame of rank r+1.  That is the only
data-path collective: one all-gather of each rank's first-frame boxes (<= max_boxes*16 B + a
count per rank -- latency-bound over NVSwitch), issued on a side stream so it overlaps the NMS
kernel; the link kernel waits on it.  The reference has no distributed code at all
(single process), so there is nothing to mirror here beyond the frame order.

The exchange itself is backend-agnostic torch.distributed plumbing (NCCL on GPUs; the CPU
tests drive it with gloo, world_size 2).
"""
import torch
import torch.distributed as dist


def shard_range(n_frames, world_size, rank):
    """Contiguous, balanced frame range [start, stop) of ``rank`` (earlier ranks take the remainder)."""
    base, rem = divmod(int(n_frames), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def bind_rank_cpus(local_rank, local_world):
    """Pin this process to its own slice of the visible CPUs (call BEFORE allocating pinned buffers, so that
    their pages are first touched from that slice).  On a multi-socket host this keeps a rank's staging
    copies and the DMA reads of its GPU on the local memory node; on a single-node VM it only stops the
    ranks' staging threads from migrating over each other.  Returns the CPU list."""
    import os
    cpus = sorted(os.sched_getaffinity(0))
    per = len(cpus) // max(int(local_world), 1)
    if per < 1:
        return cpus
    mine = cpus[local_rank * per:(local_rank + 1) * per]
    os.sched_setaffinity(0, mine)
    return mine


class BoundaryExchange(object):
    """All-gather of first-frame boxes; each rank reads its right neighbour's slot.

    Slot layout per rank: [max_boxes, 4] float32 boxes, then one row whose first word holds the box count as
    int32 BITS -- the link kernel reads the neighbour's count from the device (``halo_count``), so ragged
    shards need no host round trip.  Buffers have fixed addresses: a CUDA graph can hold them."""

    def __init__(self, max_boxes, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.max_boxes = int(max_boxes)
        self.device = torch.device(device)
        self.send = torch.zeros((self.max_boxes + 1, 4), dtype=torch.float32, device=self.device)
        self.recv = torch.zeros((self.world, self.max_boxes + 1, 4), dtype=torch.float32, device=self.device)
        self._send_count = self.send.view(torch.int32)[self.max_boxes, 0]
        self._sent_count = -1

    def start(self, first_frame_boxes):
        """Enqueue the all-gather (async when the backend supports it); returns a handle.  ``first_frame_boxes``
        may be a device tensor or a pinned host tensor (it is copied into the send slot either way)."""
        n = int(first_frame_boxes.shape[0])
        if n > self.max_boxes:
            raise ValueError("first frame has %d boxes > max_boxes %d" % (n, self.max_boxes))
        if n:
            self.send[:n].copy_(first_frame_boxes, non_blocking=True)
        if n != self._sent_count:
            # fill_ passes the value as a kernel argument.  (``send[i, 0] = n`` copies a host scalar from
            # pageable memory: that copy is stream-ordered behind the previous step's kernels and BLOCKS
            # the host until they finish -- measured: host enqueue time == device time per step, 0.53 ms
            # instead of 0.46, profiles/r01_multi_probe.json.)
            self._send_count.fill_(n)
            self._sent_count = n
        if self.world == 1:
            return None
        return dist.all_gather_into_tensor(self.recv.view(-1, 4), self.send, group=self.group, async_op=True)

    def finish(self, handle):
        """Wait and return (halo, halo_count): the next rank's first-frame buffer [max_boxes, 4] and its box
        count as an int32 device tensor [1]; (None, None) on the last rank."""
        if handle is not None:
            handle.wait()
        if self.world == 1 or self.rank == self.world - 1:
            return None, None
        slot = self.recv[self.rank + 1]
        return slot[:self.max_boxes], slot.view(torch.int32)[self.max_boxes, :1]


class ShardedVideoPostProcessor(object):
    """The per-rank step of the multi-GPU pipeline: NMS of the local frames, boundary exchange,
    link (local frames + halo).  Weak scaling: every rank holds up to ``n_frames`` frames."""

    def __init__(self, n_frames, n_boxes, n_classes, nms_thresh=0.3, device=None, group=None, n_chunks=8, n_slots=2,
                 n_stage=None, bind_cpus=False, stage_threads=0):
        from .vdet.video_det import VideoPostProcessor
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if bind_cpus and world > 1:
            import os
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
            local_rank = int(os.environ.get("LOCAL_RANK", dist.get_rank(group) % local_world))
            self.cpus = bind_rank_cpus(local_rank, local_world)
            if not stage_threads:
                stage_threads = max(1, min(8, len(self.cpus)))
        self.pp = VideoPostProcessor(n_frames, n_boxes, n_classes, nms_thresh, device, n_chunks=n_chunks, n_slots=n_slots,
                                     n_stage=n_stage, stage_threads=stage_threads)
        # one exchange buffer set per slot: step k+1's all-gather must not overwrite the halo step k still links against
        self.exchanges = [BoundaryExchange(n_boxes, self.pp.device, group) for _ in self.pp.slots]
        self.exchange = self.exchanges[0]
        self.side = torch.cuda.Stream(device=self.pp.device, priority=-1)
        self.n_boxes = n_boxes
        if self.exchange.world > 1:
            from . import _lib
            _lib.load().vdet_set_reserved_sms(2)       # room for the all-gather next to the NMS grid

    def _exchange(self, d_first_frame):
        """Boundary all-gather on the side stream; returns (halo, halo_count); join the side stream before the link."""
        main = torch.cuda.current_stream()
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            handle = self.exchange.start(d_first_frame)
            return self.exchange.finish(handle)

    def step_device(self, d_boxes, d_scores):
        """Device-resident inputs; returns the result dict of VideoPostProcessor.run_device.
        On a single rank there is no exchange and the step is replayed from a CUDA graph.

        The boundary all-gather is enqueued first, on a side stream, and overlaps the NMS kernel: with
        more than one rank a few SMs are kept out of the persistent NMS grid (vdet_set_reserved_sms)
        so that the NCCL kernel is scheduled immediately; the link kernel then waits on it.  Successors of
        the last frame are ``T*N + index into the neighbour's first frame``.

        Measured alternatives that did not pay (2 x B200, profiles/r01_scaling.md): linking the
        shard's own frames first and only the last frame after the exchange (0.55-0.56 ms/step
        against 0.54), and replaying the sharded step from a CUDA graph -- torch refuses the capture
        of this fork/join around the NCCL all-gather ("capturing stream has unjoined work")."""
        from . import ops
        pp = self.pp
        if self.exchange.world == 1:
            return pp.run_device(d_boxes, d_scores, None, graph=True)
        main = torch.cuda.current_stream()
        halo, halo_count = self._exchange(d_boxes[:self.n_boxes])
        pp.status.zero_()
        out = ops.nms_frames(d_boxes, d_scores, pp.seg_offsets, pp.nms_thresh, pp.N, want_mask=True,
                             status=pp.status, frame_major_out=True, out=(pp.d_idx, pp.d_cnt, pp.d_mask))
        # the link waits for the exchange only, not for the NMS: on the side stream it fills the SMs the persistent
        # NMS grid leaves idle in its last round
        with torch.cuda.stream(self.side):
            succ, link_iou = ops.link_frames(d_boxes, pp.seg_offsets, pp.N, halo, halo_row_base=pp.T * pp.N,
                                             out=(pp.d_succ, pp.d_iou), halo_count=halo_count, ws=pp.slots[0].link_ws)
        main.wait_stream(self.side)
        res = pp._views(out)
        res.update(succ=succ, link_iou=link_iou)
        return res

    def _halo_fn(self, slot_index, h_first_frame):
        """Boundary exchange of one staged step, enqueued on the current (launch) stream ahead of the step:
        the first frame goes up on its own (a few KB from the pinned upload buffer), the all-gather runs, and
        the step -- eager or a CUDA graph -- links against this slot's fixed halo buffer."""
        ex = self.exchanges[slot_index]
        return ex.finish(ex.start(h_first_frame))

    def submit_host(self, boxes, scores, counts=None, graph=True):
        """Enqueue one end-to-end step from host arrays (pageable or not) without waiting for it; returns a
        ticket for :meth:`collect`.  ``n_slots`` steps may be in flight (double-buffered pinned, device and
        result buffers), which keeps the host->device link busy across step boundaries."""
        multi = self.exchange.world > 1
        return self.pp.submit_host(boxes, scores, counts, halo_fn=self._halo_fn if multi else None, graph=graph)

    def submit_staged(self, graph=True):
        """The same from inputs already staged with ``self.pp.stage(...)`` (re-submission of one shard) or written in
        place (``input_buffers`` / ``commit_inputs``)."""
        multi = self.exchange.world > 1
        return self.pp.submit_staged(halo_fn=self._halo_fn if multi else None, graph=graph)

    def input_buffers(self):
        """The pinned upload buffers of the next step, for a producer that writes in place
        (VideoPostProcessor.input_buffers); follow with :meth:`commit_inputs` and :meth:`submit_staged`."""
        return self.pp.input_buffers()

    def commit_inputs(self, counts=None):
        self.pp.commit_inputs(counts)

    def step_host(self, boxes=None, scores=None, counts=None, graph=False):
        """The end-to-end step, synchronous (host arrays, or the staged shard when none are given)."""
        if boxes is None:
            return self.collect(self.submit_staged(graph))
        return self.collect(self.submit_host(boxes, scores, counts, graph))

    def collect(self, ticket):
        """Wait for a submitted step; a StepResult of host views (valid until the slot is reused)."""
        return self.pp.collect(ticket)


def sharded_vid_nms(dets_local, thresh, row_offset, group=None):
    """vid_nms (utils/nms.pyx:71-125) of a video whose rows are sharded by frame over the ranks.

    ``dets_local`` [M_r, 6] float32 CUDA tensor = this rank's rows (whole frames only: a frame must
    not straddle ranks), ``row_offset`` = global index of its first row.  Suppression is per
    frame, hence local; only the reference's GLOBAL descending-score keep order needs the other
    ranks: one all-gather of the kept (score, global row) lists (padded to the longest) and one
    stable sort.  Every rank returns the same int64 tensor of global row indices.
    """
    from . import ops
    keep = ops.vid_nms(dets_local, thresh)                         # local rows, local score order
    scores = dets_local[:, 5].contiguous()[keep]
    rows = keep + int(row_offset)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return rows
    world = dist.get_world_size(group)
    dev = dets_local.device
    n_local = torch.tensor([rows.numel()], dtype=torch.int64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, n_local, group=group)
    counts = counts.cpu().tolist()
    cap = max(max(counts), 1)
    send_s = torch.zeros(cap, dtype=torch.float32, device=dev)
    send_r = torch.zeros(cap, dtype=torch.int64, device=dev)
    send_s[:rows.numel()] = scores
    send_r[:rows.numel()] = rows
    all_s = torch.empty(world * cap, dtype=torch.float32, device=dev)
    all_r = torch.empty(world * cap, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_s, send_s, group=group)
    dist.all_gather_into_tensor(all_r, send_r, group=group)
    cat_s = torch.cat([all_s[r * cap:r * cap + counts[r]] for r in range(world)])
    cat_r = torch.cat([all_r[r * cap:r * cap + counts[r]] for r in range(world)])
    _, merged = ops.sort_by_score_desc(cat_s, cat_r)
    return merged


# ------------------------------------------------------------------------------------------
# temporal stages (BASELINE config 4): rows = (tubelet, class) score sequences
# ------------------------------------------------------------------------------------------
class ShardedTemporalRows(object):
    """Sharding BY TUBELET: every rank owns a contiguous range of rows and runs completion / max-pool /
    conv on it with no communication (the preferred split: rows outnumber ranks by orders of magnitude,
    and score completion needs a tubelet's whole frame axis).  ``gather`` reassembles the full block on
    every rank when a consumer needs it (one all-gather, shards padded to the largest)."""

    def __init__(self, n_rows, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_rows = int(n_rows)
        self.spans = [shard_range(self.n_rows, self.world, r) for r in range(self.world)]
        self.start, self.stop = self.spans[self.rank]

    def local(self, rows):
        """This rank's slice of a full [n_rows, ...] array / tensor."""
        return rows[self.start:self.stop]

    def gather(self, local_rows):
        """[stop-start, L] on every rank -> [n_rows, L] on every rank."""
        if local_rows.shape[0] != self.stop - self.start:
            raise ValueError("gather: expected %d local rows" % (self.stop - self.start))
        if self.world == 1:
            return local_rows
        cap = max(b - a for a, b in self.spans)
        send = local_rows.new_zeros((cap,) + tuple(local_rows.shape[1:]))
        send[:local_rows.shape[0]] = local_rows
        recv = local_rows.new_empty((self.world * cap,) + tuple(local_rows.shape[1:]))
        dist.all_gather_into_tensor(recv, send, group=self.group)
        return torch.cat([recv[r * cap:r * cap + (b - a)] for r, (a, b) in enumerate(self.spans)])


def frame_sharded_window_op(local_cols, half_window, op, pad_value=None, group=None):
    """Sharding BY FRAME for the window stages (temporal max-pool / conv): every rank holds the columns
    (frames) [f0, f1) of ALL rows and needs ``half_window`` columns from each neighbour -- the same
    single boundary all-gather as the link, payload rows x 2h values per rank.

    ``local_cols`` [rows, L_r] (L_r >= half_window on every rank); ``op(ext)`` maps the haloed block
    [rows, h + L_r + h] to an output of the same shape (e.g. ``ops.temporal_maxpool(ext, w)``); the h
    columns beyond the video's two ends are filled with ``pad_value`` when given (max-pool: -1e5, conv
    with zero padding: 0), otherwise with the nearest edge column.  Returns [rows, L_r]."""
    h = int(half_window)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    rows, L = local_cols.shape
    if h == 0:
        return op(local_cols)
    if L < h:
        raise ValueError("frame_sharded_window_op: a shard of %d frames is shorter than the halo %d" % (L, h))
    edge = torch.cat([local_cols[:, :h], local_cols[:, L - h:]], dim=1).contiguous()       # [rows, 2h]
    if world > 1:
        recv = edge.new_empty((world,) + tuple(edge.shape))
        dist.all_gather_into_tensor(recv.view(world * rows, 2 * h), edge, group=group)
    else:
        recv = edge.unsqueeze(0)

    def outside(col):
        return local_cols.new_full((rows, h), pad_value) if pad_value is not None else col.expand(rows, h)

    left = recv[rank - 1][:, h:] if rank > 0 else outside(local_cols[:, :1])
    right = recv[rank + 1][:, :h] if rank < world - 1 else outside(local_cols[:, L - 1:])
    ext = torch.cat([left, local_cols, right], dim=1).contiguous()
    return op(ext)[:, h:h + L]


def frame_sharded_completion_(local_cols, miss_thr=-10.0, group=None, status=None):
    """do_score_completion (vdet/tubelet_cls.py:284-303) for rows sharded BY FRAME: every rank holds the
    columns (frames) [f0, f1) of ALL rows, in place.  A run of missing scores that touches a shard edge is completed
    from the nearest valid score on the other side of it, which may be several ranks away: one all-gather of a
    4-value summary per (row, rank) -- first valid column + value, last valid column + value -- and of the shard
    widths gives every rank the (gap, value) pair it needs on each side; the kernel then works with the positions
    of the whole tubelet (``vdet_score_completion_bounded``).  Bit-identical to completing the unsharded rows."""
    from . import ops
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    rows, L = local_cols.shape
    dt = local_cols.dtype
    valid = local_cols > miss_thr                                   # missing: score <= miss_thr (:286)
    any_valid = valid.any(dim=1)
    idx = torch.arange(L, device=local_cols.device)
    first = torch.where(valid, idx, torch.full_like(idx, L)).min(dim=1).values if L else torch.zeros(rows, dtype=torch.long, device=local_cols.device)
    last = torch.where(valid, idx, torch.full_like(idx, -1)).max(dim=1).values if L else torch.full((rows,), -1, dtype=torch.long, device=local_cols.device)
    safe_first, safe_last = first.clamp(max=max(L - 1, 0)), last.clamp(min=0)
    summ = torch.stack([torch.where(any_valid, first, torch.full_like(first, -1)).to(torch.float64),
                        local_cols.gather(1, safe_first[:, None])[:, 0].to(torch.float64) if L else torch.zeros(rows, dtype=torch.float64, device=local_cols.device),
                        torch.where(any_valid, last, torch.full_like(last, -1)).to(torch.float64),
                        local_cols.gather(1, safe_last[:, None])[:, 0].to(torch.float64) if L else torch.zeros(rows, dtype=torch.float64, device=local_cols.device),
                        torch.full((rows,), float(L), dtype=torch.float64, device=local_cols.device)], dim=1).contiguous()
    if world > 1:
        allsum = summ.new_empty((world, rows, 5))
        dist.all_gather_into_tensor(allsum.view(world * rows, 5), summ, group=group)
    else:
        allsum = summ.unsqueeze(0)
    bounds = torch.full((rows, 4), -1.0, dtype=torch.float64, device=local_cols.device)
    # nearest valid score to the left: walk the ranks below this one, accumulating the frames in between
    gap = torch.zeros(rows, dtype=torch.float64, device=local_cols.device)
    found = torch.zeros(rows, dtype=torch.bool, device=local_cols.device)
    for r in range(rank - 1, -1, -1):
        s_r = allsum[r]
        has = (s_r[:, 2] >= 0) & ~found
        bounds[:, 0] = torch.where(has, gap + (s_r[:, 4] - 1 - s_r[:, 2]), bounds[:, 0])
        bounds[:, 1] = torch.where(has, s_r[:, 3], bounds[:, 1])
        found |= has
        gap = gap + s_r[:, 4]
    gap = torch.zeros(rows, dtype=torch.float64, device=local_cols.device)
    found = torch.zeros(rows, dtype=torch.bool, device=local_cols.device)
    for r in range(rank + 1, world):
        s_r = allsum[r]
        has = (s_r[:, 0] >= 0) & ~found
        bounds[:, 2] = torch.where(has, gap + s_r[:, 0], bounds[:, 2])
        bounds[:, 3] = torch.where(has, s_r[:, 1], bounds[:, 3])
        found |= has
        gap = gap + s_r[:, 4]
    return ops.score_completion_(local_cols, None, miss_thr, status, bounds=bounds.to(dt))
