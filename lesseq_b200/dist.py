"""Gene-range sharding over torch.distributed (one process per GPU).

The reference's only parallelism is "run N processes with different gene_begin_idx/gene_end_idx
and concatenate the outputs" (count/count.cpp:113-114, 202-215).  Here rank r of P takes the
contiguous slice [b + r*n/P, b + (r+1)*n/P) of the selected gene ranks; every rank filters reads
with the covered regions of the WHOLE selection so results do not depend on P.
No collective runs inside the loop; per-event results are gathered to rank 0 once at the end
(NCCL for CUDA tensors, gloo for CPU tensors in the tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Slice of n gene ranks owned by `rank` (the reference's begin/end mechanism, end exclusive)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def gather_rows(local, root=0):
    """Concatenate per-rank 1-D/2-D numpy arrays (rows = events or forms, rank order = gene-rank
    order) on `root`.  Returns the concatenation on root, None elsewhere.  Variable lengths are
    handled by a size exchange followed by one padded all_gather per array."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    a = np.ascontiguousarray(local)
    width = int(np.prod(a.shape[1:])) if a.ndim > 1 else 1  # explicit: an empty shard has zero rows
    flat = a.reshape(a.shape[0], width)
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[rank] = flat.shape[0]
    dist.all_reduce(sizes)
    maxn = int(sizes.max().item())
    nbytes = width * a.dtype.itemsize  # explicit: a rank whose shard is empty has zero rows
    raw = torch.from_numpy(np.ascontiguousarray(flat).view(np.uint8).reshape(flat.shape[0], nbytes).copy())
    pad = torch.zeros((maxn, raw.shape[1]), dtype=torch.uint8)
    pad[: raw.shape[0]] = raw
    pad = pad.to(dev)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    if rank != root:
        return None
    parts = [o[: int(sizes[r].item())].cpu().numpy() for r, o in enumerate(out)]
    cat = np.ascontiguousarray(np.concatenate(parts, axis=0)).view(a.dtype).reshape(-1, width)
    if a.ndim == 1:
        return cat.reshape(-1)
    return cat.reshape((-1,) + a.shape[1:])


def gather_results(count, solve, root=0):
    """Gather the dicts returned by Context.get_count / get_solve.  n_valid and bases are
    [M, E_local]: they are gathered along the event axis."""
    out = {}
    for k, v in count.items():
        if v.ndim == 2:
            g = gather_rows(np.ascontiguousarray(v.T), root)
            out[k] = None if g is None else np.ascontiguousarray(g.T)
        else:
            out[k] = gather_rows(v, root)
    for k, v in (solve or {}).items():
        out[k] = gather_rows(v, root)
    return out


class DeviceGather:
    """One NCCL all_gather per step of every rank's result arrays, device to device.

    Sizes (events, forms per shard) are exchanged once at construction.  `gather(ctx)` packs the
    context's device result arrays into one byte buffer, all-gathers it over NVLink and leaves the
    bytes of all shards in rank order on every rank's device; `unpack()` (rank 0, outside any
    timed region if only the device-resident state matters) reassembles numpy arrays in gene-rank
    order."""

    FIELDS = (("n_valid", 8, "M*E", np.uint64), ("bases", 8, "M*E", np.uint64), ("iso_count", 8, "F", np.uint64),
              ("theta", 8, "F", np.float64), ("rpkm", 8, "F", np.float64), ("loglik", 8, "E", np.float64),
              ("ll_per_read", 8, "E", np.float64), ("iters", 4, "E", np.uint32))

    def __init__(self, E_local, F_local, M):
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        dev = torch.device("cuda", torch.cuda.current_device())
        sizes = torch.zeros(self.world, 2, dtype=torch.int64, device=dev)
        sizes[self.rank, 0], sizes[self.rank, 1] = E_local, F_local
        dist.all_reduce(sizes)
        self.sizes = sizes.cpu().numpy()
        self.M = M
        self.nbytes = [self._bytes(int(e), int(f)) for e, f in self.sizes]
        self.maxb = (max(self.nbytes) + 15) // 16 * 16
        self.send = torch.zeros(self.maxb, dtype=torch.uint8, device=dev)
        self.recv = torch.zeros(self.world * self.maxb, dtype=torch.uint8, device=dev)

    def _count(self, kind, E, F):
        return {"M*E": self.M * E, "F": F, "E": E}[kind]

    def _bytes(self, E, F):
        return sum(w * self._count(kind, E, F) for _, w, kind, _ in self.FIELDS)

    def gather(self, ctx):
        res = ctx.device_results()
        off = 0
        for name, w, _, _ in self.FIELDS:
            b = res[name].view(torch.uint8)
            self.send[off:off + b.numel()] = b
            off += b.numel()
        dist.all_gather_into_tensor(self.recv, self.send)
        return self.recv

    def unpack(self):
        raw = self.recv.cpu().numpy()
        parts = {name: [] for name, _, _, _ in self.FIELDS}
        for r in range(self.world):
            E, F = int(self.sizes[r, 0]), int(self.sizes[r, 1])
            off = r * self.maxb
            for name, w, kind, dt in self.FIELDS:
                n = self._count(kind, E, F)
                a = raw[off:off + n * w].view(dt)
                parts[name].append(a.reshape(self.M, E) if kind == "M*E" else a)
                off += n * w
        return {k: np.concatenate(v, axis=1 if v[0].ndim == 2 else 0) for k, v in parts.items()}
