"""Worker of tests/test_gpu_multi.py: one rank of a sample-split render whose films are combined by PeerFilmReducer over
REAL CUDA IPC mappings.  Every rank uses the CUDA device given on the command line (all ranks on one GPU is allowed: the
mappings, the peer loads / stores and the orchestration are the same as across GPUs; only the wire differs), so the
process group is gloo — NCCL refuses two ranks on one device and carries no film data here anyway.

usage: mp_peer_reduce.py <mode: slice|root> <device of rank 0> <device of rank 1> ... ; RANK / WORLD_SIZE / MASTER_* in env.
Rank 0 checks the combined film against its own single-GPU render of ALL sample indices and prints "OK <max rel err>".
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "lightmetrica-v3_b200"))
import lmcu  # noqa: E402
from lmcu import abi, scenes  # noqa: E402
from lmcu import distributed as D  # noqa: E402


def main():
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    device = int(sys.argv[2 + rank])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(device)

    def any_changed(flag):
        t = torch.tensor([1 if flag else 0])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return bool(t.item())

    sb = scenes.cornell_spheres(16)
    sc = lmcu.Scene(sb, device=device)
    red = D.PeerFilmReducer(dist, rank, world, D.SceneFilmOps(lmcu, sc, device, torch.cuda.synchronize), any_changed, mode=mode)
    worst = 0.0
    # two film sizes (the second ragged against float4 slices, and a NEW film buffer => handles are exchanged again),
    # two steps each (the second must reuse the cached mappings)
    for (w, h, total_spp) in ((96, 54, 24), (37, 23, 9)):
        for step in range(2):
            offset, count = D.sample_block(total_spp, rank, world)
            sc.render("pt", abi.render_params(w, h, spp=count, max_verts=8, seed=7 + step, sample_offset=offset, rescale=0), download=False)
            red(total_spp)
            if rank == 0:
                got = sc.film_download(w, h).astype(np.float64)
                want, _ = sc.render("pt", abi.render_params(w, h, spp=total_spp, max_verts=8, seed=7 + step))
                # same Philox streams for any split: only the fp32 summation order differs
                err = np.abs(got - want).max() / max(1e-9, np.abs(want).max())
                assert np.allclose(got, want, rtol=1e-4, atol=1e-5), "combined film differs from the single-GPU render: %g" % err
                worst = max(worst, err)
            dist.barrier()
    assert red.exchanges == 2, red.exchanges
    red.close()
    sc.close()
    if rank == 0:
        print("OK %.3g" % worst, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
