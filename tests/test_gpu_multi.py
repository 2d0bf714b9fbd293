"""The multi-GPU film combine, asserted on hardware.

One GPU is enough for everything but the wire: two or three lmcu.Scenes on device 0 stand in for the peers of the
root-gather kernel (k_film_reduce_peers) and of the slice-parallel kernel (k_film_reduce_slice); two PROCESSES on
device 0 exercise the CUDA-IPC export / open / close path and the PeerFilmReducer orchestration exactly as bench.py
--gpus N runs it.  With >= 2 GPUs the same two-process test runs across devices (NVLink peer loads and stores)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import lmcu
from lmcu import abi, scenes
from lmcu import distributed as D

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _render_sums(sb, w, h, offset, count):
    sc = lmcu.Scene(sb)
    sc.render("pt", abi.render_params(w, h, spp=count, max_verts=8, sample_offset=offset, rescale=0), download=False)
    return sc, sc.film_download(w, h).astype(np.float64)


@pytest.mark.parametrize("size", [(64, 36), (37, 23), (5, 3), (1, 1)])     # 37*23*3 and 5*3*3 floats leave a tail (n & 3)
@pytest.mark.parametrize("peers", [1, 3])
def test_film_reduce_peers_equals_sum_of_downloads(size, peers):
    """root = (root + sum of peers) * scale must equal the same expression over the downloaded films."""
    w, h = size
    sb = scenes.cornell_spheres(8)
    spp = 4
    parts = [_render_sums(sb, w, h, k * spp, spp) for k in range(peers + 1)]
    want = sum(f for _, f in parts) / (spp * (peers + 1))
    root = parts[0][0]
    root.film_reduce_peers([sc.film_device()[0] for sc, _ in parts[1:]], 1.0 / (spp * (peers + 1)))
    got = root.film_download(w, h)
    assert np.allclose(got, want, rtol=1e-6, atol=1e-7)
    for sc, f in parts[1:]:                                   # the peers' films are read, never written
        assert np.array_equal(sc.film_download(w, h).astype(np.float64), f)
    # and the combined film is the single-GPU render of all the sample indices (same counter-RNG streams)
    ref, _ = root.render("pt", abi.render_params(w, h, spp=spp * (peers + 1), max_verts=8))
    assert np.allclose(got, ref, rtol=1e-4, atol=1e-5)
    for sc, _ in parts:
        sc.close()


@pytest.mark.parametrize("size", [(64, 36), (37, 23), (2, 1)])
@pytest.mark.parametrize("world", [2, 3])
def test_film_reduce_slice_equals_sum_of_downloads(size, world):
    """Every "rank" combines its slice of all films into the root's film; together they produce (sum of films) * scale."""
    w, h = size
    sb = scenes.cornell_spheres(8)
    spp = 3
    parts = [_render_sums(sb, w, h, k * spp, spp) for k in range(world)]
    want = sum(f for _, f in parts) / (spp * world)
    ptrs = [sc.film_device()[0] for sc, _ in parts]
    n = w * h * 3
    for r, (b, e) in enumerate(D.film_slices(n, world)):
        parts[r][0].film_reduce_slice(ptrs, ptrs[0], b, e, 1.0 / (spp * world))
    got = parts[0][0].film_download(w, h)
    assert np.allclose(got, want, rtol=1e-6, atol=1e-7)
    for sc, _ in parts:
        sc.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run_ranks(mode, devices):
    port = _free_port()
    procs = []
    for rank in range(len(devices)):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(len(devices)), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "mp_peer_reduce.py"), mode] + [str(d) for d in devices],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for p, out in zip(procs, outs):
        assert p.returncode == 0, out[-3000:]
    assert "OK" in outs[0], outs[0][-3000:]
    return outs[0]


@pytest.mark.parametrize("mode", ["slice", "root"])
def test_peer_film_reducer_two_processes_cuda_ipc(mode):
    """Two ranks, two processes, CUDA IPC film mappings, the reducer bench.py uses — on ONE device (on a multi-GPU box the
    next test repeats it across devices)."""
    _run_ranks(mode, [0, 0])


@pytest.mark.parametrize("mode", ["slice", "root"])
def test_peer_film_reducer_across_gpus(mode):
    n = lmcu.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    _run_ranks(mode, list(range(min(n, 4))))
