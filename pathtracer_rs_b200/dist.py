"""Multi-GPU plumbing for the path integrator: one process per GPU, scene replicated, Sobol sample
numbers sharded across ranks, films combined with ONE collective (sum) — SURVEY.md §8e.

The reference has no distributed code; the decomposition relies on two facts of its design: the global
Sobol index is a pure function of (pixel, sample number) (sampler/sobol.rs:169-175) and the film is a
plain sum of per-sample contributions (film.rs:102-103, 223-226)."""


def sample_shard(rank, world_size):
    """(stride, phase): rank renders sample numbers s with s % stride == phase.  Interleaving keeps every
    rank's work statistically identical (the same pixels, neighbouring Sobol points)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return world_size, rank


def shard_sample_counts(spp, world_size):
    """How many of the spp sample numbers each rank renders under sample_shard()."""
    return [len(range(r, spp, world_size)) for r in range(world_size)]


def reduce_film(film_tensor, dst=0):
    """Sum the per-rank films into rank `dst` (NCCL on GPUs, gloo on CPU tensors).  No-op without an
    initialised process group or with a single rank."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film_tensor, dst=dst, op=dist.ReduceOp.SUM)
    return film_tensor
