"""Host logic of the multi-GPU sample split on CPU: world_size 2 and 3 over gloo.
The per-rank renderer is a deterministic stand-in (sum over the block's sample indices of a per-sample
image), so the reduced, rescaled film has a closed form; the real renderer is covered by the -m gpu tests
(test_sample_split_is_additive) and by bench.py --gpus N."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lmcu import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_sample_image(i, npix):
    g = torch.Generator().manual_seed(1000 + i)
    return torch.rand(npix * 3, generator=g)


def _worker(rank, world, port, total_spp, npix, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def render_sums(offset, count):
        film = torch.zeros(npix * 3)
        for i in range(offset, offset + count):
            film += _fake_sample_image(i, npix)
        return film

    film = D.render_pt_sample_split(render_sums, total_spp, rank, world, dist)
    if rank == 0:
        torch.save(film, out)
    dist.barrier()
    dist.destroy_process_group()


def _peer_worker(rank, world, port, total_spp, npix, tmp, out, mode):
    """PeerFilmReducer orchestration with files standing in for CUDA IPC mappings: a "pointer" is the path of a rank's
    film file, a peer store is a read-modify-write of that file."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    film = torch.zeros(npix * 3)
    opened, closed = [], []
    mine = os.path.join(tmp, "film%d.pt" % rank)

    class Ops:
        @staticmethod
        def film_handle():
            return mine.encode()

        @staticmethod
        def film_ptr():
            return mine, npix * 3

        @staticmethod
        def sync():
            torch.save(film, mine)                      # "render complete": the film becomes readable by the peers

        @staticmethod
        def open_handle(h):
            opened.append(h)
            return h.decode()

        @staticmethod
        def close_handle(ptr):
            closed.append(ptr)

        @staticmethod
        def reduce_root(ptrs, scale):
            acc = film.clone()
            for path in ptrs:
                acc += torch.load(path)
            film.copy_(acc * scale)

        @staticmethod
        def reduce_slice(ptrs, out_ptr, begin, end, scale):
            assert begin % 4 == 0 and ptrs[rank] == mine
            acc = torch.zeros(end - begin)
            for path in ptrs:
                acc += torch.load(path)[begin:end]
            torch.save(acc * scale, os.path.join(tmp, "slice%d.pt" % rank))   # the peer store, applied once all have read

    def any_changed(flag):
        t = torch.tensor([1 if flag else 0])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return bool(t.item())

    red = D.PeerFilmReducer(dist, rank, world, Ops, any_changed, mode=mode)
    for step in range(2):                                # second step must reuse the cached mappings
        offset, count = D.sample_block(total_spp, rank, world)
        film.zero_()
        for i in range(offset, offset + count):
            film += _fake_sample_image(i, npix)
        red(total_spp)
        if mode == "slice" and rank == 0:                # assemble the stores into the root's film
            for r, (b, e) in enumerate(D.film_slices(npix * 3, world)):
                film[b:e] = torch.load(os.path.join(tmp, "slice%d.pt" % r))
        dist.barrier()
    assert red.exchanges == 1
    assert len(opened) == (world - 1 if (mode == "slice" or rank == 0) else 0)
    red.close()
    assert len(closed) == len(opened)
    if rank == 0:
        torch.save(film, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["slice", "root"])
@pytest.mark.parametrize("world,total_spp", [(2, 8), (3, 7)])
def test_peer_film_reducer(tmp_path, world, total_spp, mode):
    npix = 13                                            # 39 floats: ragged against the 4-float slices
    out = str(tmp_path / "film.pt")
    mp.spawn(_peer_worker, args=(world, _free_port(), total_spp, npix, str(tmp_path), out, mode), nprocs=world, join=True)
    got = torch.load(out)
    want = sum(_fake_sample_image(i, npix) for i in range(total_spp)) / total_spp
    assert torch.allclose(got, want, atol=1e-6)


def test_film_slices_tile_the_film():
    for n in (3, 4, 39, 1920 * 1080 * 3, 37 * 23 * 3):
        for world in (1, 2, 3, 8):
            sl = D.film_slices(n, world)
            assert sl[0][0] == 0 and sl[-1][1] == n
            for (b0, e0), (b1, _) in zip(sl, sl[1:]):
                assert e0 == b1 and b1 % 4 == 0 and b0 <= e0


@pytest.mark.parametrize("world,total_spp", [(2, 8), (3, 7)])
def test_sample_split_reduce(tmp_path, world, total_spp):
    npix = 12
    out = str(tmp_path / "film.pt")
    mp.spawn(_worker, args=(world, _free_port(), total_spp, npix, out), nprocs=world, join=True)
    got = torch.load(out)
    want = sum(_fake_sample_image(i, npix) for i in range(total_spp)) / total_spp
    assert torch.allclose(got, want, atol=1e-6)


def _spi_worker(rank, world, port, num_samples, npix, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def render_sums(offset, count):                      # stand-in light tracer: sample i splats onto an arbitrary pixel
        film = torch.zeros(npix * 3)
        for i in range(offset, offset + count):
            px = (i * 7 + 3) % npix
            film[px * 3:px * 3 + 3] += _fake_sample_image(i, 1)
        return film

    film = D.render_spi_sample_split(render_sums, num_samples, npix, rank, world, dist)
    if rank == 0:
        torch.save(film, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,num_samples", [(2, 50), (3, 2)])
def test_image_space_sample_split(tmp_path, world, num_samples):
    """lt / bdpt style split: blocks of path indices, arbitrary-pixel splats, one rescale by pixels / samples on rank 0
    (with 2 samples on 3 ranks one rank owns an empty block)."""
    npix = 12
    out = str(tmp_path / "film.pt")
    mp.spawn(_spi_worker, args=(world, _free_port(), num_samples, npix, out), nprocs=world, join=True)
    got = torch.load(out)
    want = torch.zeros(npix * 3)
    for i in range(num_samples):
        px = (i * 7 + 3) % npix
        want[px * 3:px * 3 + 3] += _fake_sample_image(i, 1)
    want *= npix / num_samples
    assert torch.allclose(got, want, atol=1e-6)


def test_sample_blocks_tile_the_range():
    for total in (1, 7, 16, 1024, 16384):
        for world in (1, 2, 3, 4, 8):
            blocks = [D.sample_block(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (o0, c0), (o1, _) in zip(blocks, blocks[1:]):
                assert o0 + c0 == o1


def _ray_worker(rank, world, port, n, out):
    import numpy as np
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rays = np.arange(n, dtype=np.float64)

    def intersect(block):                                # stand-in accel: a pure function of each ray
        return np.stack([block * 2.0, np.full(len(block), float(rank))], -1)

    hits = D.intersect_ray_split(intersect, rays, rank, world, dist)
    if rank == 0:
        np.save(out, hits)
    else:
        assert hits is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1001), (3, 10), (3, 2)])
def test_accel_batch_ray_split(tmp_path, world, n):
    """Ray batches split by contiguous ranges across ranks and gathered in ray order on rank 0 (ragged and tiny batches:
    a rank may own zero rays)."""
    import numpy as np
    out = str(tmp_path / "hits.npy")
    mp.spawn(_ray_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
    hits = np.load(out)
    assert hits.shape == (n, 2)
    assert np.array_equal(hits[:, 0], np.arange(n) * 2.0)
    owners = hits[:, 1]
    assert (np.diff(owners) >= 0).all() and set(owners.astype(int)) <= set(range(world))     # contiguous blocks in rank order
