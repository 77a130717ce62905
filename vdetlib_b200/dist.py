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


class BoundaryExchange(object):
    """All-gather of first-frame boxes; each rank reads its right neighbour's slot."""

    def __init__(self, max_boxes, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.max_boxes = int(max_boxes)
        self.device = torch.device(device)
        # slot layout per rank: [max_boxes, 4] boxes then one row holding the count in [0,0]
        self.send = torch.zeros((self.max_boxes + 1, 4), dtype=torch.float32, device=self.device)
        self.recv = torch.zeros((self.world, self.max_boxes + 1, 4), dtype=torch.float32, device=self.device)
        self._sent_count = -1

    def start(self, first_frame_boxes):
        """Enqueue the all-gather (async when the backend supports it); returns a handle."""
        n = int(first_frame_boxes.shape[0])
        if n > self.max_boxes:
            raise ValueError("first frame has %d boxes > max_boxes %d" % (n, self.max_boxes))
        self.send[:n].copy_(first_frame_boxes)
        if n != self._sent_count:
            # fill_ passes the value as a kernel argument.  (``send[i, 0] = float(n)`` copies a host
            # scalar from pageable memory: that copy is stream-ordered behind the previous step's
            # kernels and BLOCKS the host until they finish -- measured: host enqueue time == device
            # time per step, 0.53 ms instead of 0.46, profiles/r01_multi_probe.json.)
            self.send[self.max_boxes, 0].fill_(float(n))
            self._sent_count = n
        if self.world == 1:
            return None
        return dist.all_gather_into_tensor(self.recv.view(-1, 4), self.send, group=self.group, async_op=True)

    def finish(self, handle, count_hint=None):
        """Wait and return the halo = boxes of the next rank's first frame (None on the last rank).

        ``count_hint``: the neighbour's box count when it is known a priori (uniform frames);
        avoids reading the count back from the device."""
        if self.world == 1 or self.rank == self.world - 1:
            if handle is not None:
                handle.wait()
            return None
        handle.wait()
        slot = self.recv[self.rank + 1]
        n = int(count_hint) if count_hint is not None else int(slot[self.max_boxes, 0].item())
        return slot[:n]


class ShardedVideoPostProcessor(object):
    """The per-rank step of the multi-GPU pipeline: NMS of the local frames, boundary exchange,
    link (local frames + halo).  Weak scaling: every rank holds ``n_frames`` frames."""

    def __init__(self, n_frames, n_boxes, n_classes, nms_thresh=0.3, device=None, group=None, n_chunks=8):
        from .vdet.video_det import VideoPostProcessor
        self.pp = VideoPostProcessor(n_frames, n_boxes, n_classes, nms_thresh, device, n_chunks=n_chunks)
        self.exchange = BoundaryExchange(n_boxes, self.pp.device, group)
        self.side = torch.cuda.Stream(device=self.pp.device, priority=-1)
        self.n_boxes = n_boxes
        if self.exchange.world > 1:
            from . import _lib
            _lib.load().vdet_set_reserved_sms(2)       # room for the all-gather next to the NMS grid

    def _exchange(self, d_first_frame):
        """Boundary all-gather on the side stream; returns (halo, join) -- call join() before the link."""
        main = torch.cuda.current_stream()
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            handle = self.exchange.start(d_first_frame)
            halo = self.exchange.finish(handle, count_hint=self.n_boxes)
        return halo

    def step_device(self, d_boxes, d_scores):
        """Device-resident inputs; returns the result dict of VideoPostProcessor.run_device.
        On a single rank there is no exchange and the step is replayed from a CUDA graph.

        The boundary all-gather is enqueued first, on a side stream, and overlaps the NMS kernel: with
        more than one rank a few SMs are kept out of the persistent NMS grid (vdet_set_reserved_sms)
        so that the NCCL kernel is scheduled immediately; the link kernel then waits on it.

        Measured alternatives that did not pay (2 x B200, profiles/r01_scaling.md): linking the
        shard's own frames first and only the last frame after the exchange (0.55-0.56 ms/step
        against 0.54), and replaying the sharded step from a CUDA graph -- torch refuses the capture
        of this fork/join around the NCCL all-gather ("capturing stream has unjoined work")."""
        from . import ops
        pp = self.pp
        if self.exchange.world == 1:
            return pp.run_device(d_boxes, d_scores, None, graph=True)
        main = torch.cuda.current_stream()
        halo = self._exchange(d_boxes[:self.n_boxes])
        out = ops.nms_frames(d_boxes, d_scores, pp.seg_offsets, pp.nms_thresh, pp.N, want_mask=True,
                             status=pp.status, frame_major_out=True, out=(pp.d_idx, pp.d_cnt, pp.d_mask))
        main.wait_stream(self.side)
        succ, link_iou = ops.link_frames(d_boxes, pp.seg_offsets, pp.N, halo, out=(pp.d_succ, pp.d_iou))
        res = pp._views(out)
        res.update(succ=succ, link_iou=link_iou)
        return res

    def step_host(self, graph=False):
        """The end-to-end step, synchronous: pipelined H2D from the pinned staging buffers (fill them
        with ``self.pp.stage(boxes, scores)``), kernels + boundary exchange, D2H of the results."""
        return self.collect(self.submit_host(graph))

    def submit_host(self, graph=False):
        """Enqueue one end-to-end step without waiting for it; returns a ticket for :meth:`collect`.
        Two steps may be in flight (double-buffered device and result buffers), which keeps the
        host->device link busy across step boundaries.  ``graph`` (single rank only): replay the
        step from one CUDA graph."""
        multi = self.exchange.world > 1
        return self.pp.submit_staged(halo_fn=self._halo_then_join if multi else None, graph=graph and not multi)

    def collect(self, ticket):
        """Wait for a submitted step; host views of its results (valid until the slot is reused)."""
        return self.pp.collect(ticket)

    def _halo_then_join(self, d_first_frame):
        halo = self._exchange(d_first_frame)
        # the compute stream must not start the link before the all-gather finished; the wait is
        # enqueued now (cheap: the exchange is ~20 us and the NMS chunks run in between anyway)
        torch.cuda.current_stream().wait_stream(self.side)
        return halo


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
