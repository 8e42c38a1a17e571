"""Multi-GPU film accumulation: one process per GPU, sample-index sharding, one
collective.

Replaces the reference's block scheduler + TCP/SSH RemoteWorker merge
(src/librender/renderproc.cpp:142-148 `m_film->put(block)`, src/libcore/sched_remote.cpp) with:
  sample s of every pixel  ->  rank (s mod world)       (every rank sees the whole image: no halo,
                                                          identical load wherever the medium projects)
  film = sum over ranks of the partial [R,G,B,alpha,weight] films    (ImageBlock::put is linear)
The sum is ONE torch.distributed reduce (NCCL over NVLink on the GPU box; gloo in the CPU tests).
"""
import torch.distributed as dist


def shard(rank, world):
    """(sample_begin, sample_stride) of a rank"""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of size %d" % (rank, world))
    return rank, world


def local_spp(spp_total, rank, world):
    """number of sample indices s in [0, spp_total) with s mod world == rank"""
    return (spp_total - rank + world - 1) // world if rank < spp_total else 0


def render_sharded(render_local, film, dst=0, all_ranks=False):
    """`render_local(sample_begin, sample_stride)` accumulates this rank's partial film into the
    tensor `film` (H x W x 5, or H x W x (3*frames+2) for a transient film; zeroed by the caller) and returns its stats dict; the partial films
    are then summed onto rank `dst` (or onto every rank).  Counters are summed with a second,
    tiny all-reduce by `reduce_stats`."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    begin, stride = shard(rank, world)
    stats = render_local(begin, stride)
    if world > 1:
        if all_ranks:
            dist.all_reduce(film, op=dist.ReduceOp.SUM)
        else:
            dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return stats


def reduce_stats(stats, device=None, keys=("samples", "ray_steps", "scatter_events", "null_collisions",
                                           "boundary_exits", "nonfinite_dropped", "connections", "connections_failed",
                                           "connection_steps")):
    import torch
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return dict(stats)
    t = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out = dict(stats)
    out.update({k: int(v) for k, v in zip(keys, t.tolist())})
    return out
