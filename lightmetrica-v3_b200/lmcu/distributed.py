"""Sample-split multi-GPU rendering.

The scene and BVH are replicated; rank r of G renders its own block of sample indices of EVERY
pixel into a private fp32 film (counter-based RNG keyed by the global sample index, so the union
of the streams is the same for any G); ONE reduce(sum) of the film goes to rank 0 — NCCL over
NVLink on GPUs, gloo in the CPU tests of this host logic — which then applies the single
film::rescale(1/processed) of the reference (renderer_pt.cpp:307-309).

Two ways to combine the films:
  reduce_and_rescale        torch.distributed reduce (NCCL) + one scale          — the baseline
  PeerFilmReducer           liblmcu's own kernels over NVLink peer memory (CUDA IPC mappings of the ranks' device
                            films), the sum fused with film::rescale in ONE pass: slice-parallel (every rank
                            combines its slice of all films into the root's film, lmcu_film_reduce_slice — the
                            default on GPUs) or root-gather (lmcu_film_reduce_peers)

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


def render_spi_sample_split(render_sums, num_samples, num_pixels, rank, world, dist=None):
    """The same split for the renderers on the image-space scheduler (scheduler::spi::*: renderer::lt, bdpt, the
    single-strategy path renderers, pt with primary_ray_sampling_mode = image): a sample is one path whose contributions
    land on arbitrary pixels, rank r renders the contiguous block of the `num_samples` sample indices that sample_block
    gives it (LmcuRenderParams: num_samples = count, sample_offset = offset, rescale = 0), the films are summed on rank 0
    and scaled ONCE by num_pixels / num_samples (renderer_lt.cpp:162, renderer_bdpt.cpp:390-391).
    render_sums(sample_offset, count) -> torch tensor of per-pixel sums.  A rank whose block is empty contributes zeros."""
    offset, count = sample_block(num_samples, rank, world)
    film = render_sums(offset, count)
    return reduce_and_rescale(film, float(num_samples) / float(num_pixels), rank, world, dist)


class SceneFilmOps:
    """The `ops` of PeerFilmReducer for an lmcu.Scene (or, with `scene` a callable, for whichever scene it returns at the
    time of the call: a host that re-creates its scene per render keeps ONE reducer and its cached mappings as long as
    the film buffer is recycled at the same address)."""

    def __init__(self, lmcu, scene, device, sync):
        self._lmcu, self._scene, self._device, self.sync = lmcu, scene, device, sync

    def _sc(self):
        return self._scene() if callable(self._scene) else self._scene

    def film_handle(self):
        return self._sc().film_ipc_handle()

    def film_ptr(self):
        return self._sc().film_device()

    def open_handle(self, h):
        return self._lmcu.ipc_open(self._device, h)

    def close_handle(self, ptr):
        self._lmcu.ipc_close(self._device, ptr)

    def reduce_root(self, ptrs, scale):
        self._sc().film_reduce_peers(ptrs, scale)

    def reduce_slice(self, ptrs, out_ptr, begin, end, scale):
        self._sc().film_reduce_slice(ptrs, out_ptr, begin, end, scale)


def film_slices(num_floats, world):
    """Slice [begin, end) of the film each rank combines in the slice-parallel reduce: contiguous, tiling [0, num_floats),
    every begin a multiple of 4 floats (the kernel moves float4s); the last rank takes the ragged tail."""
    quads = int(num_floats) // 4
    out = []
    for r in range(world):
        off, cnt = sample_block(quads, r, world)
        out.append((off * 4, int(num_floats) if r == world - 1 else (off + cnt) * 4))
    return out


class PeerFilmReducer:
    """Fused film reduce + rescale over peer memory, one process per GPU.  `ops` is the rank's scene seen through five
    callables (Scene methods on GPUs, file-backed stand-ins in the CPU tests of this orchestration):

    ops.film_handle()               -> bytes: opaque handle of this rank's film           (Scene.film_ipc_handle)
    ops.film_ptr()                  -> (pointer, number of floats) of this rank's film    (Scene.film_device)
    ops.open_handle(h)              -> pointer valid on THIS rank's device                (lmcu.ipc_open)
    ops.close_handle(ptr)                                                               (lmcu.ipc_close)
    ops.reduce_root(ptrs, s)        -> root only: film = (film + sum(peer films)) * s     (Scene.film_reduce_peers)
    ops.reduce_slice(ptrs, out, b, e, s) -> out[b:e] = s * sum(films[k][b:e])           (Scene.film_reduce_slice)
    ops.sync()                      -> waits for this rank's render                       (torch.cuda.synchronize)

    mode "slice" (default): every rank combines its slice of ALL films and stores it into the root's film — all NVLink
    ports busy at once, (G-1)/G of a film inbound per GPU.  mode "root": the root pulls G-1 whole films.
    Mappings are cached per handle and re-exchanged only when some rank's film buffer changed; `any_changed(flag)` is an
    all-reduce(MAX) of that flag over the ranks, which is also the "every rank's samples are in its film" barrier.
    """

    def __init__(self, dist, rank, world, ops, any_changed, mode="slice"):
        self.dist, self.rank, self.world, self.ops, self.any_changed, self.mode = dist, rank, world, ops, any_changed, mode
        self._opened = {}           # rank -> (handle, pointer)
        self._last = None
        self.exchanges = 0

    def _exchange(self, mine):
        handles = [None] * self.world
        self.dist.all_gather_object(handles, mine)
        want = range(self.world) if self.mode == "slice" else (range(1, self.world) if self.rank == 0 else ())
        for r in want:
            if r == self.rank:
                continue
            old = self._opened.get(r)
            if old is not None and old[0] == handles[r]:
                continue
            if old is not None:
                self.ops.close_handle(old[1])
            self._opened[r] = (handles[r], self.ops.open_handle(handles[r]))
        self.exchanges += 1

    def __call__(self, total_spp):
        scale = 1.0 / float(total_spp)
        if self.world == 1:
            self.ops.sync()
            self.ops.reduce_root([], scale)
            return
        mine = self.ops.film_handle()
        changed = mine != self._last
        self._last = mine
        self.ops.sync()                                 # this rank's samples are in its film ...
        if self.any_changed(changed):                   # ... and after this collective, so are everybody's
            self._exchange(mine)
        if self.mode == "slice":
            own, n = self.ops.film_ptr()
            films = [own if r == self.rank else self._opened[r][1] for r in range(self.world)]
            begin, end = film_slices(n, self.world)[self.rank]
            self.ops.reduce_slice(films, films[0], begin, end, scale)
        elif self.rank == 0:
            self.ops.reduce_root([self._opened[r][1] for r in range(1, self.world)], scale)
        self.dist.barrier()                             # the root's film is complete; peers may overwrite theirs again

    def close(self):
        """Unmaps the peers' films (before any rank frees its film)."""
        for _, ptr in self._opened.values():
            self.ops.close_handle(ptr)
        self._opened = {}
        self._last = None
        if self.world > 1:
            self.dist.barrier()
