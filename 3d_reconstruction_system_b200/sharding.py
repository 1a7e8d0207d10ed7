"""Multi-GPU plumbing of the mapping path: one process per GPU, torch.distributed for the
rendezvous and the one exchange step.  torch is imported lazily and only here; the single-GPU product path never needs it.

* Back-projection / pose transform: frames are independent.  `frame_range` gives rank r the contiguous frames
  [r*n/G, (r+1)*n/G); the merged cloud is the concatenation in rank order, no data-path collective.
* Occupancy (insertPointCloud mode): ray casting is independent per scan, but the clamped float32 log-odds add is not
  associative, so every voxel must see its updates in scan order.  Scans are processed in rounds of G*C consecutive
  scans: rank r ray-casts scans [base + r*C, base + (r+1)*C) into brick-delta records (R3D_DELTA_RECORD_BYTES each),
  the records of the round are all-gathered (counts first, then one padded payload), and every rank applies them in
  global scan order.  Each rank ends with the same tree as a 1-GPU run, bit for bit, for any G.
  With `owner_partition=True` a rank applies only the bricks it owns (hash(brick key) mod G) -- 1/G of the apply work
  and of the map memory per GPU -- and the per-rank trees are disjoint pieces of the same map, merged at the end by
  `gather_bricks`.
"""
import numpy as np

RECORD_BYTES = 136


def frame_range(n_frames, world, rank):
    """Contiguous shard [lo, hi) of rank `rank`; sizes differ by at most one frame."""
    base, rem = divmod(int(n_frames), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def scan_rounds(n_scans, world, scans_per_rank):
    """Yields (base, [(rank, first_scan, n)] ...) for rounds of world*scans_per_rank consecutive scans; the last round may
    be ragged (ranks at the end get fewer or zero scans)."""
    per_round = world * scans_per_rank
    for base in range(0, n_scans, per_round):
        parts = []
        for r in range(world):
            a = min(n_scans, base + r * scans_per_rank)
            b = min(n_scans, a + scans_per_rank)
            parts.append((r, a, b - a))
        yield base, parts


def brick_owner(keys_u64, world):
    """Owner rank of a brick key (the same mixing function the kernels use for their hash tables)."""
    x = np.asarray(keys_u64, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xff51afd7ed558ccd)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xc4ceb9fe1a85ec53)
        x ^= x >> np.uint64(33)
    return ((x >> np.uint64(32)) % np.uint64(world)).astype(np.int64)


def allgather_varlen(payload, counts, group=None):
    """All-gather of one variable-length uint8 payload per rank (utility; the merge below uses fixed slots instead).

    payload: 1-D uint8 tensor (CPU for gloo, CUDA for nccl) holding this rank's records back to back.
    counts:  1-D int64 tensor (same device) with this rank's per-scan record counts (fixed length on every rank).
    Returns (payloads, counts_all): list of G 1-D uint8 tensors (views into one buffer) and a (G, len(counts)) int64 tensor.
    Two collectives: counts, then the payload padded to the round's maximum."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts_all = torch.empty((world, counts.numel()), dtype=torch.int64, device=counts.device)
    dist.all_gather_into_tensor(counts_all.view(-1), counts.contiguous(), group=group)
    nbytes = (counts_all.sum(dim=1) * RECORD_BYTES).cpu()
    maxb = int(nbytes.max().item())
    if maxb == 0:
        return [payload[:0] for _ in range(world)], counts_all
    send = payload
    if send.numel() < maxb:
        send = torch.zeros(maxb, dtype=torch.uint8, device=payload.device)
        send[:payload.numel()] = payload
    elif send.numel() > maxb:
        send = send[:maxb]
    recv = torch.empty(world * maxb, dtype=torch.uint8, device=payload.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    return [recv[r * maxb:r * maxb + int(nbytes[r].item())] for r in range(world)], counts_all


class _Exchange:
    """The one exchange step of the merge: per round every rank contributes ONE fixed-size slot -- a header (payload bytes,
    bytes it would have needed, per-scan record counts) followed by its records -- and one all-gather moves all slots.
    No separate counts collective, no per-round allocation, no zero-fill: send / receive buffers are double-buffered
    (round r + 1 is packed and sent while round r is applied) and kept in `state` between runs.  A rank whose records do
    not fit its slot says so in its header; every rank sees that in the same round and the slots are regrown."""

    def __init__(self, world, scans_per_rank, device, state, group):
        self.world, self.C, self.device, self.group, self.state = world, scans_per_rank, device, group, state
        self.hdr_words = 2 + scans_per_rank
        self.hdr_bytes = ((self.hdr_words * 8 + 255) // 256) * 256        # payload starts 256-byte aligned (device records want 8)
        self.slot = 0
        key = ("xchg", str(device), world, scans_per_rank)
        if key in state:
            self.slot, self.send, self.recv = state[key]
        else:
            self._alloc(max(int(state.get("slot_hint", 0)), int(state.get("slot_min", 1 << 20))))
        self.key = key

    def _alloc(self, payload_bytes):
        import torch
        self.slot = self.hdr_bytes + ((int(payload_bytes) + 255) // 256) * 256
        self.send = [torch.empty(self.slot, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.recv = [torch.empty(self.world * self.slot, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.state[("xchg", str(self.device), self.world, self.C)] = (self.slot, self.send, self.recv)

    def capacity(self):
        return self.slot - self.hdr_bytes

    def start(self, parity, payload, counts):
        """Pack this rank's slot and start the all-gather; returns the work handle."""
        import torch
        import torch.distributed as dist
        nbytes = int(payload.numel())
        fits = nbytes <= self.capacity()
        hdr = torch.zeros(self.hdr_words, dtype=torch.int64)
        hdr[0] = nbytes if fits else 0
        hdr[1] = nbytes
        for i, c in enumerate(counts):
            hdr[2 + i] = int(c)
        send = self.send[parity]
        send[:self.hdr_words * 8].copy_(hdr.view(torch.uint8), non_blocking=True)
        if fits and nbytes:
            send[self.hdr_bytes:self.hdr_bytes + nbytes].copy_(payload, non_blocking=True)
        return dist.all_gather_into_tensor(self.recv[parity], send, group=self.group, async_op=True)

    def headers(self, parity):
        """(world, 2 + C) int64 on the host (one small read-back)."""
        import torch
        recv = self.recv[parity].view(self.world, self.slot)
        return recv[:, :self.hdr_words * 8].contiguous().view(torch.int64).view(self.world, self.hdr_words).cpu().numpy()

    def payload(self, parity, rank, nbytes):
        off = rank * self.slot + self.hdr_bytes
        return self.recv[parity][off:off + nbytes]


def merged_insert(n_scans, rank, world, compute_delta, apply_delta, make_buffer, group=None, scans_per_rank=4, fence=None,
                  compute_round=None, apply_round=None, overlap=True, state=None):
    """Scan-ordered multi-GPU insertPointCloud.  Rounds of world * scans_per_rank consecutive scans; rank r ray-casts scans
    [base + r*C, base + (r+1)*C) of a round, the rounds' records are exchanged (one all-gather of fixed slots, _Exchange)
    and every rank applies them in global scan order.  overlap=True software-pipelines the rounds: round r's records travel
    and are applied while round r + 1 is ray-cast (the exchange is started BEFORE the next round's ray casting is queued,
    its records are applied right after that ray casting has been queued).

    compute_delta(scan_idx, out, offset_bytes) -> n_records | (n_records, new_out): ray-casts one scan and writes its
        records into the uint8 tensor `out` at `offset_bytes` (it may grow the tensor and return the new one).
    apply_delta(records_tensor, n_records, scan_idx): applies one scan's records (a 1-D uint8 view) to the local map.
    compute_round(first_scan, n, out) -> (counts list, out) and apply_round(payload_tensor, counts list, first_scan):
        optional batched forms of the two (one library call per round and rank instead of one per scan).
    make_buffer(nbytes) -> 1-D uint8 tensor on the exchange device (this rank's record buffer).
    fence(): called after a collective completed and before its data is applied (stream hand-over between torch and the
        library).
    state: dict kept by the caller between runs (exchange buffers are allocated once).
    Returns the number of records applied locally."""
    import torch
    applied = 0
    state = state if state is not None else {}
    bufs = [make_buffer(1 << 20), None]          # this rank's records of the round in flight / the round being cast
    xchg = None

    def cast(first, n_mine, buf):
        counts = [0] * scans_per_rank
        off = 0
        if compute_round is not None and n_mine:
            cnts, buf = compute_round(first, n_mine, buf)
            for i, c in enumerate(cnts):
                counts[i] = int(c)
            off = sum(counts) * RECORD_BYTES
        else:
            for i in range(n_mine):
                res = compute_delta(first + i, buf, off)
                if isinstance(res, tuple):
                    n_rec, buf = res
                else:
                    n_rec = res
                counts[i] = int(n_rec)
                off += int(n_rec) * RECORD_BYTES
        return counts, off, buf

    def finish(pending):
        """Wait for a round's exchange and apply it; returns False when some rank's records did not fit (nothing applied)."""
        nonlocal applied
        work, parity, parts = pending
        work.wait()
        if fence is not None:
            fence()
        hdr = xchg.headers(parity)
        if int(hdr[:, 1].max()) > xchg.capacity():
            return False, int(hdr[:, 1].max())
        for r, a, n in parts:
            if n == 0:
                continue
            cnts = [int(c) for c in hdr[r, 2:2 + n]]
            total = sum(cnts)
            if total == 0:
                continue
            seg = xchg.payload(parity, r, total * RECORD_BYTES)
            if apply_round is not None:
                apply_round(seg, cnts, a)
            else:
                o = 0
                for i, c in enumerate(cnts):
                    if c:
                        apply_delta(seg[o:o + c * RECORD_BYTES], c, a + i)
                    o += c * RECORD_BYTES
            applied += total
        return True, 0

    rounds = list(scan_rounds(n_scans, world, scans_per_rank))
    pending = None                                # (work, parity, parts) of the round whose records are travelling
    kept = {}                                     # parity -> (counts, nbytes) of the rounds not yet applied (for a regrow)
    for k, (base, parts) in enumerate(rounds):
        parity = k & 1
        _, first, n_mine = parts[rank]
        if bufs[parity] is None:
            bufs[parity] = make_buffer(1 << 20) if parity == 0 else torch.empty_like(bufs[0])
        counts, nbytes, bufs[parity] = cast(first, n_mine, bufs[parity])
        if xchg is None:
            xchg = _Exchange(world, scans_per_rank, bufs[parity].device, state, group)
        kept[parity] = (counts, nbytes)
        work = xchg.start(parity, bufs[parity][:nbytes], counts)
        cur = (work, parity, parts)
        last = k + 1 == len(rounds)
        # pipelined: apply the round before this one now (its records arrived while this round was ray-cast) and leave
        # this round's exchange in flight; otherwise (and for the last round) apply this round too
        todo = [it for it in ([pending] + ([cur] if (not overlap or last) else [])) if it is not None]
        pending = cur if (overlap and not last) else None
        for n_done, item in enumerate(todo):
            ok, need = finish(item)
            if ok:
                continue
            # Some rank's records did not fit its slot; every rank sees that in this same round.  Let everything in flight
            # land, regrow the slots, and exchange the rounds not yet applied again, in order (their records are kept).
            redo = todo[n_done:] + ([pending] if pending is not None else [])
            for it in redo:
                it[0].wait()
            if fence is not None:
                fence()
            state["slot_hint"] = need + need // 2
            xchg._alloc(state["slot_hint"])
            for it in redo:
                c2, nb2 = kept[it[1]]
                ok2, need2 = finish((xchg.start(it[1], bufs[it[1]][:nb2], c2), it[1], it[2]))
                if not ok2:       # a later round needs even more
                    state["slot_hint"] = need2 + need2 // 2
                    xchg._alloc(state["slot_hint"])
                    ok2, _ = finish((xchg.start(it[1], bufs[it[1]][:nb2], c2), it[1], it[2]))
                    assert ok2, "exchange slot still too small after regrowing"
            pending = None
            break
    if xchg is not None:
        state["exchange"] = ("one all-gather of fixed %d-byte slots per round (header with the per-scan record counts + 136-byte "
                             "brick-delta records), double-buffered, software-pipelined with the next round's ray casting: %s"
                             % (xchg.slot, "on" if overlap else "off"))
    return applied


class OctreeSharder:
    """Glue between merged_insert and the GPU OcTree of this package (device buffers are torch CUDA uint8 tensors)."""

    def __init__(self, tree, get_scan, maxrange=-1.0, owner_partition=False, rank=0, world=1, get_scan_batch=None, state=None):
        """get_scan(s) -> (points, origin).  get_scan_batch(first, n) -> (points buffer with the n scans back to back,
        per-scan point counts, origins (n, 3)): optional, lets a round's scans go through one library call.
        state: a dict the caller keeps between runs (record / exchange buffers are allocated once and reused)."""
        self.tree, self.get_scan, self.maxrange = tree, get_scan, float(maxrange)
        self.owner_partition, self.rank, self.world = owner_partition, rank, world
        self.get_scan_batch = get_scan_batch
        self.state = state if state is not None else {}
        self._buf = self.state.get("buf")
        self._defer = False

    def make_buffer(self, nbytes):
        """Record buffer of this rank, kept across rounds and runs (a too small buffer costs a re-cast of the round)."""
        import torch
        want = max(int(nbytes), 32 << 20)
        if getattr(self, "_buf", None) is None or self._buf.numel() < want:
            self._buf = torch.empty(want, dtype=torch.uint8, device=torch.device("cuda", self.tree._ctx.device))
        self.state["buf"] = self._buf
        return self._buf

    def compute_delta(self, scan_idx, out, offset):
        import torch
        points, origin = self.get_scan(scan_idx)
        n = self.tree.computeScanDeltaOnDevice(points, origin, self.maxrange)
        need = offset + n * RECORD_BYTES
        if need > out.numel():
            grown = torch.empty(max(need, out.numel() * 2), dtype=torch.uint8, device=out.device)
            grown[:offset] = out[:offset]
            torch.cuda.current_stream(out.device).synchronize()
            if out is self._buf:
                self._buf = self.state["buf"] = grown
            out = grown
        if n:
            self.tree.scanDeltaInto(out.data_ptr() + offset, n)
        return n, out

    def apply_delta(self, records, n_records, scan_idx):
        if self.owner_partition:
            self.tree.applyDeltaOwned(records, n_records, self.rank, self.world)
        else:
            self.tree.applyDelta(records, n_records)

    def compute_round(self, first, n, out):
        """All of this rank's scans of a round in one library call when the scans are slices of one device buffer
        (get_scan_batch given), else scan by scan."""
        import torch
        if self.get_scan_batch is None:
            cnts, off = [], 0
            for i in range(n):
                c, out = self.compute_delta(first + i, out, off)
                cnts.append(c)
                off += c * RECORD_BYTES
            return cnts, out
        points, counts, origins = self.get_scan_batch(first, n)
        while True:
            try:
                rec = self.tree.computeScanDeltasInto(points, counts, origins, self.maxrange, out.data_ptr(), out.numel() // RECORD_BYTES)
                return [int(c) for c in rec], out
            except MemoryError:
                grown = torch.empty(out.numel() * 2, dtype=torch.uint8, device=out.device)
                if out is self._buf:
                    self._buf = self.state["buf"] = grown
                out = grown

    def apply_round(self, payload, counts, first_scan):
        # noted only: every rank's share of the round is applied in ONE sorted, scan-ordered pass (three launches instead of one
        # per scan and rank) by the next round's ray-casting call, or by the end of run()
        fn = self.tree.deferDeltasOwned if self._defer else self.tree.applyDeltasOwned
        fn(payload.data_ptr(), counts, self.rank if self.owner_partition else 0, self.world if self.owner_partition else 1)

    def fence(self):
        import torch
        torch.cuda.current_stream().synchronize()

    def run(self, n_scans, group=None, scans_per_rank=8, overlap=None):
        """overlap (default on; R3D_MERGE_OVERLAP=0 turns it off): a round's exchange and apply run beside the next round's
        ray casting."""
        import os
        if overlap is None:
            overlap = os.environ.get("R3D_MERGE_OVERLAP", "1") != "0"
        self._defer = self.get_scan_batch is not None
        try:
            return self._run(n_scans, group, scans_per_rank, overlap)
        finally:
            if self._defer:
                self.tree.flushDeferred()

    def _run(self, n_scans, group, scans_per_rank, overlap):
        return merged_insert(n_scans, self.rank, self.world, self.compute_delta, self.apply_delta, self.make_buffer, group=group,
                             scans_per_rank=scans_per_rank, fence=self.fence, compute_round=self.compute_round, apply_round=self.apply_round,
                             overlap=overlap, state=self.state)


def gather_bricks(tree, group=None, state=None):
    """Merge owner-partitioned maps: every rank exports its bricks (key, 512 log-odds, 512 known bits), the bricks are
    all-gathered and imported, so that every rank ends with the whole map (disjoint union; bit-identical to a 1-GPU run).
    state: a dict the caller keeps between runs (the export / receive buffers are allocated once)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = torch.device("cuda", tree._ctx.device)
    state = state if state is not None else {}
    n = tree.numBricks()
    rec = tree.BRICK_RECORD_BYTES
    counts = torch.tensor([n], dtype=torch.int64, device=dev)
    counts_all = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_all, counts, group=group)
    ch = counts_all.cpu().numpy()
    maxn = int(ch.max())
    if maxn == 0:
        return 0
    need = maxn * rec
    if state.get("gather_send") is None or state["gather_send"].numel() < need:
        state["gather_send"] = torch.empty(need + need // 4, dtype=torch.uint8, device=dev)
        state["gather_recv"] = torch.empty(world * (need + need // 4), dtype=torch.uint8, device=dev)
    send = state["gather_send"][:need]
    recv = state["gather_recv"][:world * need]
    if n:
        tree.exportBricks(send.data_ptr(), n)          # (the padding behind a rank's bricks is never read)
    tree.reserve(int(ch.sum()))                         # one regrowth at most, before the imports
    dist.all_gather_into_tensor(recv, send, group=group)
    torch.cuda.current_stream().synchronize()
    me = dist.get_rank(group)
    total = 0
    for r in range(world):
        if r == me or ch[r] == 0:
            continue
        tree.importBricks(recv.data_ptr() + r * need, int(ch[r]))
        total += int(ch[r])
    return total
