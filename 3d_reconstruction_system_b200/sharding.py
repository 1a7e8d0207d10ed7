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
    """All-gather of one variable-length uint8 payload per rank.

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


def _start_gather(payload, counts, group, sync_stream):
    """Counts all-gather (small, blocking) then the payload all-gather, asynchronous: returns (work, recv, counts_host, maxb).
    The payload is first copied into a private send buffer (the caller reuses its own for the next round)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts_all = torch.empty((world, counts.numel()), dtype=torch.int64, device=counts.device)
    dist.all_gather_into_tensor(counts_all.view(-1), counts.contiguous(), group=group)
    counts_host = counts_all.cpu().numpy()
    maxb = int(counts_host.sum(axis=1).max()) * RECORD_BYTES
    if maxb == 0:
        return None, None, counts_host, 0
    send = torch.zeros(maxb, dtype=torch.uint8, device=payload.device)
    send[:payload.numel()] = payload
    if sync_stream is not None:
        sync_stream()              # the copy above must have read `payload` before the caller overwrites it
    recv = torch.empty(world * maxb, dtype=torch.uint8, device=payload.device)
    work = dist.all_gather_into_tensor(recv, send, group=group, async_op=True)
    return (work, send), recv, counts_host, maxb


def merged_insert(n_scans, rank, world, compute_delta, apply_delta, make_buffer, group=None, scans_per_rank=4, fence=None,
                  compute_round=None, apply_round=None, overlap=False):
    """Scan-ordered multi-GPU insertPointCloud.  overlap=True software-pipelines the rounds: while round r's records
    travel (asynchronous all-gather), round r-1 is applied and round r+1 is ray-cast.  Measured on 2 x B200 it is SLOWER
    than the plain sequence (1 340 vs 1 620 scans/s): the persistent ray-casting kernel fills every SM, the NCCL kernel
    waits for CTA slots and then spins on its peer while holding them, so the default applies a round as soon as it has
    been gathered.

    compute_delta(scan_idx, out, offset_bytes) -> n_records | (n_records, new_out): ray-casts one scan and writes its
        records into the uint8 tensor `out` at `offset_bytes` (it may grow the tensor and return the new one).
    apply_delta(records_tensor, n_records, scan_idx): applies one scan's records (a 1-D uint8 view) to the local map.
    compute_round(first_scan, n, out) -> (counts list, out) and apply_round(payload_tensor, counts list, first_scan):
        optional batched forms of the two (one library call per round and rank instead of one per scan).
    make_buffer(nbytes) -> 1-D uint8 tensor on the exchange device.
    fence(): called after a collective completed and before its data is applied (stream hand-over between torch and the
        library); also used to make sure a send buffer has been read.
    Returns the number of records applied locally."""
    import torch
    applied = 0
    buf = make_buffer(1 << 20)

    def finish(pending):
        nonlocal applied
        handle, recv, counts_host, maxb, parts = pending
        if handle is not None:
            handle[0].wait()
            if fence is not None:
                fence()
        for r, a, n in parts:
            if n == 0 or maxb == 0:
                continue
            cnts = [int(c) for c in counts_host[r, :n]]
            total = sum(cnts)
            if total == 0:
                continue
            seg = recv[r * maxb:r * maxb + total * RECORD_BYTES]
            if apply_round is not None:
                apply_round(seg, cnts, a)
            else:
                o = 0
                for i, c in enumerate(cnts):
                    if c:
                        apply_delta(seg[o:o + c * RECORD_BYTES], c, a + i)
                    o += c * RECORD_BYTES
            applied += total

    pending = None
    for base, parts in scan_rounds(n_scans, world, scans_per_rank):
        _, first, n_mine = parts[rank]
        counts = torch.zeros(scans_per_rank, dtype=torch.int64)
        off = 0
        if compute_round is not None and n_mine:
            cnts, buf = compute_round(first, n_mine, buf)
            for i, c in enumerate(cnts):
                counts[i] = int(c)
            off = int(sum(int(c) for c in cnts)) * RECORD_BYTES
        else:
            for i in range(n_mine):
                res = compute_delta(first + i, buf, off)
                if isinstance(res, tuple):
                    n_rec, buf = res
                else:
                    n_rec = res
                counts[i] = n_rec
                off += int(n_rec) * RECORD_BYTES
        counts = counts.to(buf.device)
        handle, recv, counts_host, maxb = _start_gather(buf[:off], counts, group, fence)
        if pending is not None:
            finish(pending)
        pending = (handle, recv, counts_host, maxb, parts)
        if not overlap:
            finish(pending)
            pending = None
    if pending is not None:
        finish(pending)
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
            out = self._buf = grown
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
                out = self._buf = torch.empty(out.numel() * 2, dtype=torch.uint8, device=out.device)

    def apply_round(self, payload, counts, first_scan):
        self.tree.applyDeltasOwned(payload.data_ptr(), counts, self.rank if self.owner_partition else 0, self.world if self.owner_partition else 1)

    def fence(self):
        import torch
        torch.cuda.current_stream().synchronize()

    def run(self, n_scans, group=None, scans_per_rank=4, overlap=False):
        return merged_insert(n_scans, self.rank, self.world, self.compute_delta, self.apply_delta, self.make_buffer, group=group,
                             scans_per_rank=scans_per_rank, fence=self.fence, compute_round=self.compute_round, apply_round=self.apply_round,
                             overlap=overlap)


def gather_bricks(tree, group=None):
    """Merge owner-partitioned maps: every rank exports its bricks (key, 512 log-odds, 512 known bits), the bricks are
    all-gathered and imported, so that every rank ends with the whole map (disjoint union; bit-identical to a 1-GPU run)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = torch.device("cuda", tree._ctx.device)
    n = tree.numBricks()
    rec = tree.BRICK_RECORD_BYTES
    mine = torch.empty(max(n, 1) * rec, dtype=torch.uint8, device=dev)
    if n:
        tree.exportBricks(mine.data_ptr(), n)
    counts = torch.tensor([n], dtype=torch.int64, device=dev)
    counts_all = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_all, counts, group=group)
    ch = counts_all.cpu().numpy()
    maxn = int(ch.max())
    if maxn == 0:
        return 0
    send = torch.zeros(maxn * rec, dtype=torch.uint8, device=dev)
    send[:n * rec] = mine[:n * rec]
    recv = torch.empty(world * maxn * rec, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    torch.cuda.current_stream().synchronize()
    me = dist.get_rank(group)
    total = 0
    for r in range(world):
        if r == me or ch[r] == 0:
            continue
        tree.importBricks(recv.data_ptr() + r * maxn * rec, int(ch[r]))
        total += int(ch[r])
    return total
