"""Sample-split multi-GPU rendering.

The scene and BVH are replicated; rank r of G renders its own block of sample indices of EVERY
pixel into a private fp32 film (counter-based RNG keyed by the global sample index, so the union
of the streams is the same for any G); ONE reduce(sum) of the film goes to rank 0 — NCCL over
NVLink on GPUs, gloo in the CPU tests of this host logic — which then applies the single
film::rescale(1/processed) of the reference (renderer_pt.cpp:307-309).

torch.distributed is plumbing only: the film tensor wraps the device buffer liblmcu rendered into.
"""


def sample_block(total_spp, rank, world):
    """Contiguous block [offset, offset+count) of the sample indices owned by `rank`; blocks tile [0,total_spp)."""
    base, rem = divmod(int(total_spp), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def film_as_tensor(torch, ptr, n, device_index):
    """Zero-copy torch view of liblmcu's device film."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(h, device=torch.device("cuda", device_index))


def reduce_and_rescale(film, total_spp, rank, world, dist=None):
    """Film reduce to rank 0 + the reference's single rescale.  `film` holds per-pixel SUMS on every rank."""
    if world > 1:
        dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        film.mul_(1.0 / float(total_spp))
    return film


def render_pt_sample_split(render_sums, total_spp, rank, world, dist=None):
    """render_sums(sample_offset, spp) -> torch tensor of per-pixel sums for that block (rescale=0).
    Returns the final image on rank 0 (partial sums elsewhere)."""
    offset, count = sample_block(total_spp, rank, world)
    film = render_sums(offset, count)
    return reduce_and_rescale(film, total_spp, rank, world, dist)
