"""Multi-GPU decomposition of the path integrator: scene replicated, Sobol sample numbers
dealt round-robin, films summed by ONE NCCL reduce — which lives in the library (ptrs_film_reduce /
ptrs_multi_render, csrc/multi_gpu.cu).  This module is the host-side glue for the one-process-per-GPU launch
(torchrun): who renders which sample numbers, and how the communicator's id reaches every rank.

The reference has no distributed code; the decomposition relies on two facts of its design: the global
Sobol index is a pure function of (pixel, sample number) (sampler/sobol.rs:169-175) and the film is a
plain sum of per-sample contributions (film.rs:102-103, 223-226)."""
import os


def sample_shard(rank, world_size):
    """(stride, phase): rank renders sample numbers s with s % stride == phase.  Interleaving keeps every
    rank's work statistically identical (the same pixels, neighbouring Sobol points)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return world_size, rank


def shard_sample_counts(spp, world_size):
    """How many of the spp sample numbers each rank renders under sample_shard()."""
    return [len(range(r, spp, world_size)) for r in range(world_size)]


def strong_scaling_plan(total_spp, world_size):
    """A fixed total of `total_spp` samples per pixel over world_size ranks: [(stride, phase, n_samples)] per rank.
    The union over ranks is exactly {0 .. total_spp-1}; counts differ by at most one."""
    counts = shard_sample_counts(total_spp, world_size)
    return [(world_size, r, counts[r]) for r in range(world_size)]


def exchange_comm_id(make_id, rank, world_size, store=None):
    """Ship rank 0's 128-byte communicator id (gpu.Comm.unique_id) to every rank through the rendezvous store that
    torchrun / torch.distributed already runs (a TCPStore at MASTER_ADDR:MASTER_PORT) — any other channel would do.
    `store` may be any object with set(key, bytes) / get(key) -> bytes (tests pass a dict-backed one)."""
    if world_size == 1:
        return make_id()
    if store is None:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            box = [make_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        from datetime import timedelta

        store = dist.TCPStore(os.environ["MASTER_ADDR"], int(os.environ["MASTER_PORT"]) + 1, world_size, rank == 0, timedelta(seconds=120))
    if rank == 0:
        uid = make_id()
        store.set("ptrs_comm_id", uid)
        return uid
    return bytes(store.get("ptrs_comm_id"))


def reduce_film(film_tensor, dst=0):
    """Host-side stand-in for ptrs_film_reduce on CPU tensors (gloo): used by the world-size-2 CPU test of the
    decomposition, where there is no device film to reduce.  No-op without an initialised process group."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film_tensor, dst=dst, op=dist.ReduceOp.SUM)
    return film_tensor
