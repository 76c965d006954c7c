"""Gene-range shards over several GPUs, one process (rank) per GPU.

The reference's only parallelism is "run N processes with different gene_begin_idx / gene_end_idx, each reading the
whole read file, and concatenate the outputs" (count/count.cpp:113-114, 202-215).  Here rank r of P owns the gene
ranks [n*r/P, n*(r+1)/P) of the selection; every rank holds an arbitrary slice of the read set, buckets it per owning
rank on the device (lsq_route_kernel), exchanges the buckets over NCCL, counts and solves its own events, and one
NCCL send/recv group gathers the result rows on rank 0 in gene-rank order.  The C ABI does all of it
(lsq_router_build / lsq_shard_reads / lsq_gather); torch.distributed is only used to hand the NCCL id around.
"""
import numpy as np

from . import api


def shard_boundaries(n_events, world):
    """Event index boundaries of the P gene-rank shards (the reference's begin/end mechanism, end exclusive)."""
    return np.array([(n_events * r) // world for r in range(world + 1)], dtype=np.uint32)


class ShardedJob:
    """full: context with the tables of the whole selection (routes reads); shard: this rank's events."""

    def __init__(self, forms_path, map_path, device, rank, world, begin=0, end=api.UINT64_MAX, stream=None, uid=None):
        self.rank, self.world = rank, world
        self.full_ann = api.Annotation(forms_path, map_path, begin, end)
        self.shard_begin = shard_boundaries(self.full_ann.n_events, world)
        sb, se = int(self.shard_begin[rank]), int(self.shard_begin[rank + 1])
        self.ann = api.Annotation(forms_path, map_path, begin, end, shard=(sb, se))
        assert self.ann.n_events == se - sb
        if uid is None:
            uid = broadcast_unique_id(rank)
        self.comm = api.Comm(uid, world, rank, device)
        self.full = api.Context(device, stream=stream)
        self.full.upload_events(self.full_ann)
        api.router_build(self.full, self.shard_begin)
        self.shard = api.Context(device, stream=stream)
        self.shard.upload_events(self.ann)
        self.stats = None

    def step(self, reads, read_type, read_length, first_line, total_read_bases=None, m=0):
        """One pass: declare this rank's slice, bucket + exchange, count (+ solve), gather on rank 0.
        Returns (count, solve) for the whole selection on rank 0, (None, None) elsewhere."""
        self.full.set_reads(m, read_type, read_length, reads)
        self.stats = api.shard_reads(self.full, m, self.shard, m, self.comm, first_line)
        self.shard.count()
        if total_read_bases is not None:
            self.shard.solve(total_read_bases)
        return api.gather(self.full, self.shard, self.comm, 0, solved=total_read_bases is not None)

    def close(self):
        self.shard.close()
        self.full.close()
        self.comm.close()


def broadcast_unique_id(rank):
    """NCCL id of rank 0, handed to every rank through the default torch.distributed group (gloo or nccl)."""
    import torch.distributed as dist
    box = [api.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]
