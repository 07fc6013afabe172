"""Multi-GPU plumbing: the path shards by INDEPENDENT SIGNALS - contiguous index ranges
per rank, no per-iteration collective, one gather of the sparse codes at the end.  The reference
has no distributed code at all (its only fan-out is multiprocessing.Pool over a hyper-parameter
list, scripts/scale_weight_effect_mlcsc.py:165); this is the new partition the north star names.

Works with any torch.distributed backend: `nccl` (device tensors) on the GPU box, `gloo` (host
tensors) in the CPU tests.
"""
import numpy as np


def shard_range(n_items, rank, world_size):
    """Contiguous [lo, hi) of rank `rank`: sizes differ by at most one, earlier ranks get the extras."""
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_events(pos_list, idx_list, coef_list, dtype=np.float64):
    """Per-signal event lists -> (counts[S], pos[n], idx[n], coef[n]) flat arrays."""
    counts = np.array([len(p) for p in pos_list], dtype=np.int64)
    cat = lambda xs, dt: (np.concatenate(xs).astype(dt) if len(xs) and counts.sum() else np.zeros(0, dt))  # noqa: E731
    return counts, cat(pos_list, np.int32), cat(idx_list, np.int32), cat(coef_list, dtype)


def unpack_events(counts, pos, idx, coef):
    out_p, out_i, out_c = [], [], []
    o = 0
    for n in counts:
        n = int(n)
        out_p.append(pos[o:o + n])
        out_i.append(idx[o:o + n])
        out_c.append(coef[o:o + n])
        o += n
    return out_p, out_i, out_c


def gather_events(counts, pos, idx, coef, dst=0, group=None, device=None, stream=None):
    """Gathers every rank's flat events on rank `dst`.  One all_gather of the sizes, then one padded
    gather per array (3 in all).  Returns on dst: list over ranks of (counts, pos, idx, coef) numpy
    arrays, in rank order (= global signal order under shard_range); None elsewhere.
    `stream` (nccl): a side CUDA stream for the staging copies and the collectives, so that the gather of one batch
    does not wait for the kernels of the next batch already queued on the compute stream (host pipelines)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [(counts, pos, idx, coef)]
    if stream is not None and dist.get_backend(group) == 'nccl':
        with torch.cuda.stream(stream):
            return gather_events(counts, pos, idx, coef, dst=dst, group=group, device=device, stream=None)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device('cpu') if backend == 'gloo' else (device or torch.device('cuda', torch.cuda.current_device()))
    sizes = torch.tensor([len(counts), len(pos)], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    max_s, max_n = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())

    def padded(a, n, dt):
        t = torch.zeros((max(n, 1),), dtype=dt, device=dev)
        if len(a):
            t[:len(a)] = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        return t

    coef_dt = torch.float64 if np.asarray(coef).dtype == np.float64 else torch.float32
    payload = [padded(counts, max_s, torch.int64), padded(pos, max_n, torch.int32), padded(idx, max_n, torch.int32),
               padded(coef, max_n, coef_dt)]
    gathered = []
    for t in payload:
        if backend == 'nccl':
            bufs = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(bufs, t, group=group)
        else:
            bufs = [torch.zeros_like(t) for _ in range(world)] if rank == dst else None
            dist.gather(t, bufs, dst=dst, group=group)
        gathered.append(bufs)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        ns, nn = int(all_sizes[r, 0]), int(all_sizes[r, 1])
        out.append((gathered[0][r][:ns].cpu().numpy(), gathered[1][r][:nn].cpu().numpy(),
                    gathered[2][r][:nn].cpu().numpy(), gathered[3][r][:nn].cpu().numpy()))
    return out


def gather_device_events(dev, dst=0, group=None, host_out=None):
    """NCCL gather of the sparse codes straight from the device (the one collective of the path; no host bounce).
    `dev` = dict(offsets=int64[S+1], pos, idx, coef flat device tensors, total=n) as Engine.encode_host_pipelined hands to
    its on_device_events hook (compacted by hsc_b200_mp_compact_events).  One all_gather of the offsets (every rank learns
    every rank's sizes: (S+1)*8 bytes each), then one gather per array padded to the largest rank.  Runs on the current
    stream.  Returns on dst: dict(offsets=int64 [world,S+1] numpy, pos/idx/coef = lists over ranks of numpy arrays when
    `host_out` is True (device-to-host read of exactly the gathered atoms), else device tensors [world,n_max]); None
    elsewhere.  Without an initialised process group: the local codes."""
    import torch
    import torch.distributed as dist
    n = int(dev['total'])
    if not (dist.is_available() and dist.is_initialized()):
        off = dev['offsets'].cpu().numpy()[None]
        return dict(offsets=off, pos=[dev['pos'][:n].cpu().numpy()], idx=[dev['idx'][:n].cpu().numpy()], coef=[dev['coef'][:n].cpu().numpy()])
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    offsets = dev['offsets'].contiguous()
    offs = [torch.empty_like(offsets) for _ in range(world)]
    dist.all_gather(offs, offsets, group=group)
    all_off_h = torch.stack(offs).cpu().numpy()             # tiny; also tells every rank the padded size
    n_max = max(int(all_off_h[:, -1].max()), 1)
    out = {}
    for key in ('pos', 'idx', 'coef'):
        send = dev[key][:n_max].contiguous()
        if send.numel() < n_max:                            # a buffer shorter than the largest rank's code: pad
            send = torch.cat([send, torch.zeros((n_max - send.numel(),), dtype=send.dtype, device=send.device)])
        recv = torch.empty((world, n_max), dtype=send.dtype, device=send.device) if rank == dst else None
        dist.gather(send, list(recv.unbind(0)) if rank == dst else None, dst=dst, group=group)
        out[key] = recv
    if rank != dst:
        return None
    out['offsets'] = all_off_h
    if host_out:
        for key in ('pos', 'idx', 'coef'):
            out[key] = [out[key][r, :int(all_off_h[r, -1])].cpu().numpy() for r in range(world)]
    return out
