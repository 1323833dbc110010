"""Multi-GPU plumbing for the OCR path: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  The reference has no distributed code at all; this is new surface.

What crosses the fabric - never inside a forward pass:
  * `broadcast_state_dict`   one-time weight broadcast from rank 0 (one flat buffer per dtype);
  * `exchange_groups`        the crop scatter: whole reference mini-batches ("groups": the unit whose rows must stay
                             together because the AR loop stops per group) move between ranks so that every GPU gets a
                             balanced number of encoder tokens; packed u8 canvases + int32 descriptors travel with
                             all_to_all_single and split sizes;
  * `return_results`         the result gather: (ids int32, probs fp32) per crop back to the owning rank.
Pages are sharded by rank for detection (page p -> rank p // pages_per_rank); that needs no communication.
"""
import numpy as np
import torch
import torch.distributed as dist


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_state_dict(sd, device=None):
    """Broadcast every floating tensor of rank 0's state_dict (one flat fp32 buffer), returns a CPU state_dict."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return sd
    device = _dev()
    keys = [k for k in sd if torch.is_floating_point(sd[k])]
    flat = torch.cat([sd[k].detach().reshape(-1).to(torch.float32) for k in keys]).to(device)
    dist.broadcast(flat, src=0)
    flat = flat.cpu()
    out, off = dict(sd), 0
    for k in keys:
        n = sd[k].numel()
        out[k] = flat[off:off + n].reshape(sd[k].shape).clone()
        off += n
    return out


def balance_groups(costs_by_rank, world, tolerance=0.03):
    """Deterministic group -> rank assignment.  costs_by_rank[r] = list of costs (encoder tokens) of the groups rank r
    owns.  Groups stay with their owner unless moving them lowers the maximum load by more than `tolerance`
    (longest-processing-time greedy over the groups of overloaded ranks).  Returns assign[r][g] = destination rank."""
    if len(costs_by_rank) != world:
        costs_by_rank = list(costs_by_rank) + [[] for _ in range(world - len(costs_by_rank))]
    load = [float(sum(c)) for c in costs_by_rank]
    assign = [[r] * len(c) for r, c in enumerate(costs_by_rank)]
    total = sum(load)
    if world == 1 or total == 0:
        return assign
    target = total / world
    # candidates: groups of ranks above target, largest first
    order = sorted(((c, r, g) for r, cs in enumerate(costs_by_rank) for g, c in enumerate(cs)), key=lambda t: (-t[0], t[1], t[2]))
    for c, r, g in order:
        if load[r] <= target * (1 + tolerance):
            continue
        dst = min(range(world), key=lambda k: (load[k], k))
        if dst != r and load[dst] + c < load[r] - 1e-9 and max(load[dst] + c, load[r] - c) < load[r]:
            assign[r][g] = dst
            load[r] -= c
            load[dst] += c
    return assign


def _all_to_all_bytes(send_chunks):
    """send_chunks[k]: 1-D uint8 torch tensor for rank k.  Returns the list of received uint8 tensors."""
    world = dist.get_world_size()
    device = _dev()
    sizes = torch.tensor([c.numel() for c in send_chunks], dtype=torch.int64, device=device)
    recv_sizes = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(recv_sizes, sizes)
    rs = recv_sizes.tolist()
    send = torch.cat([c.to(device) for c in send_chunks]) if sum(c.numel() for c in send_chunks) else \
        torch.empty(0, dtype=torch.uint8, device=device)
    recv = torch.empty(int(sum(rs)), dtype=torch.uint8, device=device)
    dist.all_to_all_single(recv, send, output_split_sizes=rs, input_split_sizes=sizes.tolist())
    out, off = [], 0
    for n in rs:
        out.append(recv[off:off + n].cpu())
        off += n
    return out


STATS = {"exchange_calls": 0, "exchange_bytes_sent": 0, "exchange_bytes_received": 0, "exchange_ms": 0.0}


def exchange_canvases_dev(canv, byte_splits, metas):
    """The crop scatter without touching the host: GPU-to-GPU all_to_all over NCCL / NVLink.

    canv: flat uint8 tensor ON THE COMPUTE DEVICE holding the canvases of every group that leaves this rank, ordered by
    destination rank; byte_splits[k] = bytes of it that go to rank k; metas[k] = int32 numpy array describing those
    groups ([gid, n, (w, wp, byte offset in the chunk) * n] per group).  Returns (recv, recv_splits, recv_metas): one flat uint8 tensor on the same
    device with the canvases this rank received (ordered by source rank), the bytes per source and the parsed
    descriptors per source.  Under gloo (CPU tests) the tensors are CPU tensors and the same code runs."""
    import time
    world = dist.get_world_size()
    device = canv.device
    t0 = time.perf_counter()
    meta_t = [torch.from_numpy(np.ascontiguousarray(m, dtype=np.int32).view(np.uint8).copy()) for m in metas]
    sizes = torch.tensor([[int(byte_splits[k]), int(meta_t[k].numel())] for k in range(world)], dtype=torch.int64,
                         device=device)
    recv_sizes = torch.empty_like(sizes)
    dist.all_to_all_single(recv_sizes, sizes)
    rs = recv_sizes.cpu().tolist()
    recv_splits = [int(r[0]) for r in rs]
    meta_splits = [int(r[1]) for r in rs]
    recv = torch.empty(int(sum(recv_splits)), dtype=torch.uint8, device=device)
    dist.all_to_all_single(recv, canv[: int(sum(byte_splits))], output_split_sizes=recv_splits,
                           input_split_sizes=[int(b) for b in byte_splits])
    msend = torch.cat(meta_t).to(device) if sum(m.numel() for m in meta_t) else torch.empty(0, dtype=torch.uint8,
                                                                                           device=device)
    mrecv = torch.empty(int(sum(meta_splits)), dtype=torch.uint8, device=device)
    dist.all_to_all_single(mrecv, msend, output_split_sizes=meta_splits,
                           input_split_sizes=[int(m.numel()) for m in meta_t])
    mh = mrecv.cpu().numpy()
    recv_metas, off = [], 0
    for n in meta_splits:
        recv_metas.append(mh[off:off + n].view(np.int32).copy())
        off += n
    if device.type == "cuda":
        torch.cuda.current_stream(device).synchronize()
    STATS["exchange_calls"] += 1
    STATS["exchange_bytes_sent"] += int(sum(byte_splits))
    STATS["exchange_bytes_received"] += int(sum(recv_splits))
    STATS["exchange_ms"] += (time.perf_counter() - t0) * 1e3
    return recv, recv_splits, recv_metas


def parse_group_meta(meta):
    """int32 [gid, n, (w, wp, byte offset inside the sender's chunk) * n] * groups -> list of (gid, widths, padded
    widths, offsets)."""
    out, off = [], 0
    while off < len(meta):
        gid, n = int(meta[off]), int(meta[off + 1])
        wm = meta[off + 2: off + 2 + 3 * n].reshape(n, 3)
        out.append((gid, wm[:, 0].astype(np.int64), wm[:, 1].astype(np.int64), wm[:, 2].astype(np.int64)))
        off += 2 + 3 * n
    return out


def _pack_group(canvases, padded, gid):
    """-> uint8 tensor: header int32 [n, gid], per crop int32 [w, wp], then the canvases back to back."""
    n = len(canvases)
    meta = np.array([n, gid] + [v for c, p in zip(canvases, padded) for v in (c.shape[1], p)], dtype=np.int32)
    body = [np.ascontiguousarray(c).reshape(-1) for c in canvases]
    return torch.from_numpy(np.concatenate([meta.view(np.uint8)] + body)) if n else torch.from_numpy(meta.view(np.uint8).copy())


def _unpack_groups(buf, height=32):
    """Inverse of a concatenation of _pack_group blobs -> list of (gid, canvases, padded)."""
    a = buf.numpy()
    out, off = [], 0
    while off < a.size:
        n, gid = a[off:off + 8].view(np.int32)
        off += 8
        wm = a[off:off + 8 * n].view(np.int32).reshape(n, 2)
        off += 8 * n
        canv = []
        for w, _ in wm:
            sz = height * int(w) * 3
            canv.append(a[off:off + sz].reshape(height, int(w), 3))
            off += sz
        out.append((int(gid), canv, [int(p) for _, p in wm]))
    return out


def exchange_groups(groups, assign_row, height=32):
    """groups: list of (canvases, padded_widths) this rank owns; assign_row[g] = destination rank.
    Returns work = list of (owner_rank, owner_gid, canvases, padded) this rank must recognise (own groups first)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    chunks = [[] for _ in range(world)]
    for g, (canv, padded) in enumerate(groups):
        if assign_row[g] != rank:
            chunks[assign_row[g]].append(_pack_group(canv, padded, g))
    send = [torch.cat(c) if c else torch.empty(0, dtype=torch.uint8) for c in chunks]
    recv = _all_to_all_bytes(send)
    work = [(rank, g, canv, padded) for g, (canv, padded) in enumerate(groups) if assign_row[g] == rank]
    for src, blob in enumerate(recv):
        for gid, canv, padded in _unpack_groups(blob, height):
            work.append((src, gid, canv, padded))
    return work


def return_results(work, results, n_groups_local, S=101):
    """work as returned by exchange_groups; results[i] = (ids int32 (n,S), probs f32 (n,S)) for work[i].
    Sends every foreign group's result back to its owner; returns out[g] = (ids, probs) for the local groups."""
    world, rank = dist.get_world_size(), dist.get_rank()
    out = [None] * n_groups_local
    chunks = [[] for _ in range(world)]
    for (owner, gid, canv, _), (ids, probs) in zip(work, results):
        if owner == rank:
            out[gid] = (ids, probs)
        else:
            head = np.array([gid, ids.shape[0]], dtype=np.int32).view(np.uint8)
            chunks[owner].append(torch.from_numpy(np.concatenate(
                [head, np.ascontiguousarray(ids, dtype=np.int32).view(np.uint8).reshape(-1),
                 np.ascontiguousarray(probs, dtype=np.float32).view(np.uint8).reshape(-1)])))
    send = [torch.cat(c) if c else torch.empty(0, dtype=torch.uint8) for c in chunks]
    for blob in _all_to_all_bytes(send):
        a, off = blob.numpy(), 0
        while off < a.size:
            gid, n = a[off:off + 8].view(np.int32)
            off += 8
            ids = a[off:off + 4 * n * S].view(np.int32).reshape(n, S).copy()
            off += 4 * n * S
            probs = a[off:off + 4 * n * S].view(np.float32).reshape(n, S).copy()
            off += 4 * n * S
            out[int(gid)] = (ids, probs)
    return out


def gather_costs(local_costs):
    """all_gather of the per-group costs -> costs_by_rank (list of lists), identical on every rank."""
    world = dist.get_world_size()
    device = _dev()
    n = torch.tensor([len(local_costs)], dtype=torch.int64, device=device)
    ns = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(ns, n)
    m = int(max(t.item() for t in ns))
    mine = torch.zeros(max(m, 1), dtype=torch.int64, device=device)
    if local_costs:
        mine[:len(local_costs)] = torch.tensor(local_costs, dtype=torch.int64, device=device)
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine)
    return [allc[r][:int(ns[r].item())].tolist() for r in range(world)]
