"""Sample-split multi-GPU rendering.

The scene and BVH are replicated; rank r of G renders its own block of sample indices of EVERY
pixel into a private fp32 film (counter-based RNG keyed by the global sample index, so the union
of the streams is the same for any G); ONE reduce(sum) of the film goes to rank 0 — NCCL over
NVLink on GPUs, gloo in the CPU tests of this host logic — which then applies the single
film::rescale(1/processed) of the reference (renderer_pt.cpp:307-309).

Two ways to combine the films:
  reduce_and_rescale        torch.distributed reduce (NCCL) + one scale          — the baseline
  PeerFilmReducer           liblmcu's own kernel: the root reads every peer film over NVLink peer memory
                            (CUDA IPC mappings of the peers' device films) and fuses the sum with
                            film::rescale in ONE pass (lmcu_film_reduce_peers)     — the default on GPUs

torch.distributed is plumbing only (handle exchange, barriers): the film tensor wraps the device buffer
liblmcu rendered into.
"""


def sample_block(total_spp, rank, world):
    """Contiguous block [offset, offset+count) of the sample indices owned by `rank`; blocks tile [0,total_spp)."""
    base, rem = divmod(int(total_spp), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def intersect_ray_split(intersect, rays, rank, world, dist=None):
    """Accel batches on several GPUs (SURVEY.md §8e "accel batches"): the BVH is replicated, rank r traces the contiguous
    block of the batch that `sample_block` gives it, and the hit records are gathered on rank 0 in ray order — there is no
    other exchange.  intersect(rays_block) -> hits_block (numpy structured arrays).  Returns all hits on rank 0, else None."""
    import numpy as np
    offset, count = sample_block(len(rays), rank, world)
    mine = intersect(rays[offset:offset + count])
    if world == 1:
        return mine
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    return np.concatenate(parts) if rank == 0 else None


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


class PeerFilmReducer:
    """Fused film reduce + rescale over peer memory, one process per GPU.

    `film_handle()`   -> bytes: an opaque handle of this rank's film (Scene.film_ipc_handle)
    `open_handle(h)`  -> a pointer-like object the root's reduce accepts (lmcu.ipc_open on the root's device)
    `reduce(ptrs, s)` -> root only: film = (film + sum(peer films)) * s      (Scene.film_reduce_peers)
    `sync()`          -> waits for this rank's render (torch.cuda.synchronize)
    Opened mappings are cached per handle: the film buffer survives across renders of the same size.
    """

    def __init__(self, dist, rank, world, film_handle, open_handle, reduce, sync):
        self.dist, self.rank, self.world = dist, rank, world
        self.film_handle, self.open_handle, self.reduce, self.sync = film_handle, open_handle, reduce, sync
        self._opened = {}
        self._last = None

    def __call__(self, total_spp):
        scale = 1.0 / float(total_spp)
        if self.world == 1:
            self.sync()
            self.reduce([], scale)
            return
        mine = self.film_handle()
        # exchange handles only when somebody's film buffer changed (first step, or a new film size)
        changed = [None] * self.world
        self.dist.all_gather_object(changed, mine if mine != self._last else None)
        self._last = mine
        if self.rank == 0:
            for r in range(1, self.world):
                if changed[r] is not None:
                    self._opened[r] = self.open_handle(changed[r])
        self.sync()
        self.dist.barrier()                 # every rank's samples are in its film
        if self.rank == 0:
            self.reduce([self._opened[r] for r in range(1, self.world)], scale)
        self.dist.barrier()                 # peers may overwrite their films again
