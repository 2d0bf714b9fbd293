#!/usr/bin/env python
"""bench.py — headline benchmark of the ray-tracing hot path (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            our CUDA path on N B200s
  python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU path (oracle/_ref)

Workload (config.workload): BASELINE.json configs[1] = C2, `renderer::pt` of the procedural
Cornell box with a diffuse and a glass UV sphere (32 268 triangles), area-light NEE + MIS,
1920x1080, max_verts 20, seed 42 — one STEP renders the whole image at SPP samples per pixel
(1024 on the GPU arm).  Metric: Mpaths/s = width*height*spp / elapsed / 1e6.

N > 1 (torchrun, one rank per GPU): sample split.  The scene and BVH are replicated, rank r renders
its own `spp` samples of every pixel with disjoint counter-RNG sample indices (sample_offset =
r*spp); rank 0 then combines the fp32 films with liblmcu's own kernel, which reads the peers' films
over NVLink peer memory and fuses the sum with the 1/(N*spp) rescale (--reduce nccl: dist.reduce).
Per-GPU work is fixed as N grows => "scaling": "weak".  value = paths of all ranks / max-over-ranks
device time (the reduce is inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v3_b200"))

W, H, MAX_VERTS, SEED = 1920, 1080, 20, 42
GPU_SPP = 1024
REF_SPP_PER_STEP = 16         # reference arm / cpu_baseline: bounded sample of the same workload (~9 s of CPU work per step)
METRIC = "Mpaths/s renderer::pt (C2: Cornell box + diffuse/glass spheres, 1920x1080, max_verts 20, MIS)"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def reference_rate(spp, steps, warmup):
    """The reference's renderer::pt + accel::sahbvh under parallel::openmp on every host core."""
    from oracle import ref
    from lmcu import abi, scenes
    threads = ref.init(0)
    sb = scenes.cornell_spheres(64)
    rs = ref.RefScene(sb)
    p = abi.render_params(W, H, spp=spp, max_verts=MAX_VERTS, seed=SEED)
    times = []
    for i in range(warmup + steps):
        _, elapsed, _ = rs.render("pt", p)
        if i >= warmup:
            times.append(elapsed)
    rs.close()
    t = sum(times) / len(times)
    return W * H * spp / t / 1e6, t, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/liblmref.so was not built"}))
        return
    value, t, threads = reference_rate(REF_SPP_PER_STEP, args.steps, args.warmup)
    sample = "%dx%d @ %d spp per step (full C2 image, bounded spp; rate is spp-independent)" % (W, H, REF_SPP_PER_STEP)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 renderer::pt Cornell box + diffuse/glass spheres 1920x1080 max_verts=20 mis, "
                               "reference CPU path (accel::sahbvh, parallel::openmp)", "spp_per_step": REF_SPP_PER_STEP},
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def make_reducer(torch, dist, lmcu, D, sc, rank, world, local_rank, mode):
    """Film combine for one scene: 'p2p' = liblmcu's fused peer-memory reduce+rescale kernel, 'nccl' = dist.reduce."""
    if mode == "p2p":
        return D.PeerFilmReducer(dist, rank, world, sc.film_ipc_handle, lambda h: lmcu.ipc_open(local_rank, h),
                                 sc.film_reduce_peers, torch.cuda.synchronize)

    def nccl_reduce(total_spp):
        ptr, n = sc.film_device()
        film = D.film_as_tensor(torch, ptr, n, local_rank)
        D.reduce_and_rescale(film, total_spp, rank, world, dist)
    return nccl_reduce


def accel_rates(torch, lmcu, scenes, stream):
    """BASELINE.json configs[2] (C3): closest-hit throughput on the 9 999 392-triangle heightfield with DEVICE-resident
    ray batches — coherent pinhole rays and incoherent cosine-bounce rays.  Informational (not the headline)."""
    import numpy as np
    sb = scenes.heightfield_scene(2236)
    sc = lmcu.Scene(sb)
    sc.set_stream(stream.cuda_stream)
    out = {"triangles": int(sc.build_stats.num_triangles), "build_ms": sc.build_stats.elapsed * 1e3,
           "bvh_bytes": int(sc.build_stats.node_bytes + sc.build_stats.triangle_bytes)}

    def run(rays, reps=3):
        d_rays = torch.from_numpy(rays.view(np.float64).reshape(-1, 8)).cuda()
        d_hits = torch.empty((len(rays), 4), dtype=torch.float64, device="cuda")
        best = 1e9
        for _ in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sc.intersect_device(d_rays.data_ptr(), len(rays), d_hits.data_ptr(), stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        return len(rays) / best / 1e6, d_hits.cpu().numpy().view(scenes.HIT_DTYPE).reshape(-1)

    rays = scenes.camera_rays(sb.camera, 2048, 2048)
    out["coherent_mrays_per_s"], hits = run(rays)
    p = sb.meshes[0][0]
    gn = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    gn /= np.linalg.norm(gn, axis=1, keepdims=True)
    bounce = scenes.bounce_rays(rays, hits, lambda h: gn[h["face"]].copy())
    out["incoherent_mrays_per_s"], _ = run(bounce)
    sc.intersect(rays[:1 << 18], counted=True)
    out["nodes_per_ray"], out["tris_per_ray"] = sc.nodes_per_ray, sc.tris_per_ray
    out["rays"] = {"coherent": int(len(rays)), "incoherent": int(len(bounce))}
    sc.close()
    return out


def run_ours(args):
    import numpy as np
    import torch
    import lmcu
    from lmcu import abi, scenes
    from lmcu import distributed as D

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL announces itself on STDOUT ("NCCL version ...") when its communicator comes up; stdout carries exactly one
        # JSON line, so file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    spp = args.spp
    sb = scenes.cornell_spheres(64)
    sc = lmcu.Scene(sb, device=local_rank)
    stream = torch.cuda.Stream(device=local_rank)    # handle 0 (legacy default stream) would mean "the scene's own stream"
    torch.cuda.set_stream(stream)
    sc.set_stream(stream.cuda_stream)
    npix = W * H
    reducer = None

    def step():
        """One pass of the hot path: render this rank's samples of every pixel (sums), combine the films, rescale."""
        nonlocal reducer
        p = abi.render_params(W, H, spp=spp, max_verts=MAX_VERTS, seed=SEED, sample_offset=rank * spp, rescale=0)
        _, st = sc.render("pt", p, download=False)
        if reducer is None:
            reducer = make_reducer(torch, dist, lmcu, D, sc, rank, world, local_rank, args.reduce)
        reducer(world * spp)
        return st

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats = []
    e0.record(stream)
    for _ in range(args.steps):
        stats.append(step())
    e1.record(stream)
    sync()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    ms_per_step = ms_total / args.steps
    paths_per_step = npix * spp * world
    value = paths_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the public API with HOST buffers: scene description (host arrays) -> H2D upload ->
    # BVH build -> render -> film combine -> film D2H into a host array, every step, wall clock, max over ranks
    desc_bytes = sum(p.nbytes + (n.nbytes if n is not None else 0) for p, n in sb.meshes)
    sync()
    e2e_steps = max(1, min(args.steps, 2))

    def e2e_step():
        sc2 = lmcu.Scene(sb, device=local_rank)
        sc2.set_stream(stream.cuda_stream)
        p = abi.render_params(W, H, spp=spp, max_verts=MAX_VERTS, seed=SEED, sample_offset=rank * spp, rescale=0)
        sc2.render("pt", p, download=False)
        make_reducer(torch, dist, lmcu, D, sc2, rank, world, local_rank, "nccl" if world > 1 else args.reduce)(world * spp)
        film = sc2.film_download(W, H) if rank == 0 else None           # D2H of the result
        torch.cuda.synchronize()
        sc2.close()
        return film

    # one untimed pass: the first scene of this kind pays the process-wide one-off allocation of a second set of path
    # pools (the resident scene above keeps its own; the pools are sized by the job, so the warm-up runs the same job),
    # exactly like the warm-up of the timed loop above
    e2e_step()
    sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_film = e2e_step()
    t = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = {"value": paths_per_step / float(t.item()) / 1e6, "unit": "Mpaths/s",
           "h2d_bytes_per_step": int(desc_bytes), "d2h_bytes_per_step": int(npix * 3 * 4),
           "what": "lmcu.Scene(host scene description) [H2D + BVH build] + render('pt') + film combine + film D2H, "
                   "wall clock, max over ranks"}

    if rank == 0:
        assert host_film is not None and np.isfinite(host_film).all() and host_film.mean() > 0.01
        # ---- roofline of the dominant kernel (k_pt_extend_q), SURVEY.md 8(d)
        # counters pass (separate, untimed): mean nodes / triangles fetched per closest-hit ray
        pc = abi.render_params(W, H, spp=8, max_verts=MAX_VERTS, seed=SEED, count_traversal=True)
        _, stc = sc.render("pt", pc, download=False)
        n_nodes, n_tris = stc["bvh_nodes_per_ray"], stc["bvh_tris_per_ray"]
        b_ray = 32 + 16 + n_nodes * 80 + n_tris * 48
        # In the timed region TWO wavefront loops run side by side (LmcuRenderParams::wavefronts), so the CUDA events
        # around an extend launch also see the other loop's shade / shadow kernels on the same SMs: that duration is
        # reported as `overlapped`, but it is not the kernel's own.  The kernel's launch duration is measured live here
        # in a single-wavefront render of the same workload (kernels of one loop run one at a time), 256 spp.
        ps = abi.render_params(W, H, spp=256, max_verts=MAX_VERTS, seed=SEED, rescale=0, wavefronts=1)
        sc.render("pt", abi.render_params(W, H, spp=16, max_verts=MAX_VERTS, seed=SEED, rescale=0, wavefronts=1), download=False)
        _, solo = sc.render("pt", ps, download=False)
        ext_s, ext_rays, ext_launches, total_s = solo["extend_seconds"], solo["extend_rays"], solo["extend_launches"], solo["elapsed"]
        stats_solo = [solo]
        ov_s = sum(s["extend_seconds"] for s in stats)
        ov_rays = sum(s["extend_rays"] for s in stats)
        peak, peak_src = load_peaks()
        achieved = ext_rays * b_ray / ext_s / 1e9
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "r01_extend_traffic.json")) as f:
                tr = json.load(f)
            # the captured launch traced tr["rays"] rays; a launch of this run traces rays_per_launch: same bytes per ray
            traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) / tr["rays"] * (ext_rays / max(1, ext_launches))
            traffic_src = tr["source"] + "; %.1f DRAM bytes per ray, scaled to this run's rays per launch" % (
                (tr["dram_bytes_read"] + tr["dram_bytes_write"]) / tr["rays"])
        except Exception:
            pass
        roofline = {
            "bound": "hbm", "kernel": "k_pt_extend_q (closest-hit BVH traversal over the path queue)", "achieved": achieved,
            "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch",
            "traffic_source": traffic_src, "peak_source": peak_src,
            "bytes_per_ray": b_ray, "nodes_per_ray": n_nodes, "tris_per_ray": n_tris,
            "algorithmic_bytes_per_launch": ext_rays / max(1, ext_launches) * b_ray,
            "rays_per_launch": ext_rays / max(1, ext_launches), "avg_launch_ms": ext_s / max(1, ext_launches) * 1e3,
            "share_of_step": {"extend": ext_s / total_s, "shade": sum(s["shade_seconds"] for s in stats_solo) / total_s,
                              "shadow": sum(s["shadow_seconds"] for s in stats_solo) / total_s},
            "measured_in": "single-wavefront render of the same workload inside this run (256 spp, %.1f Mpaths/s): the kernel timed "
                           "with nothing beside it" % (solo["paths"] / solo["elapsed"] / 1e6),
            "overlapped": {"achieved": ov_rays * b_ray / ov_s / 1e9, "frac": ov_rays * b_ray / ov_s / 1e9 / peak,
                           "what": "same figure from the events of the timed region, where a launch shares the SMs with the other "
                                   "wavefront's kernels"},
            "note": "BVH + triangles of C2 are 1.9 MB and stay L1/L2-resident: the algorithmic bytes are served by cache "
                    "(measured DRAM traffic is the ray queue only), so frac is algorithmic bytes over the HBM peak, not "
                    "DRAM utilisation; the kernel is bound by instruction issue of divergent warps (profiles/)",
        }
        # ---- CPU baseline: the reference on this box's host cores, bounded sample
        cpu = None
        try:
            from oracle import ref
            if ref.available():
                v, t, threads = reference_rate(REF_SPP_PER_STEP, 1, 1)
                cpu = {"value": v, "unit": "Mpaths/s", "cores": threads, "kind": "reference",
                       "sample": "%dx%d @ %d spp (full C2 image, bounded spp), 1 warm-up + 1 timed render" % (W, H, REF_SPP_PER_STEP)}
        except Exception as ex:  # the baseline must never take the bench line down
            cpu = {"value": None, "unit": "Mpaths/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % ex}
        accel = None
        if world == 1 and not args.no_accel:
            try:
                accel = accel_rates(torch, lmcu, scenes, stream)
            except Exception as ex:
                accel = {"failed": str(ex)}
        rays = sum(s["rays"] for s in stats)
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 shading / f64 ray-triangle", "data": "synthetic",
            "config": {"workload": "C2 renderer::pt Cornell box + diffuse/glass UV spheres (32268 tris) 1920x1080 "
                                   "max_verts=20 sampling_mode=mis seed=42", "spp_per_gpu": spp, "total_spp": spp * world,
                       "parallelism": "sample split x%d, film combine = %s" % (
                           world, "fused peer-memory reduce+rescale kernel over NVLink (CUDA IPC)" if args.reduce == "p2p" else "NCCL reduce"),
                       "wavefronts": 2,
                       "l2": "inputs larger than L2: every wavefront iteration streams its 16 Mi-path SoA queues "
                             "(2.4 GB) through HBM, and every step regenerates all paths"},
            "mrays_per_s": rays * world / (ms_total * 1e-3) / 1e6,
            "e2e": e2e, "gpu_launches": int(sum(s["kernel_launches"] for s in stats)) + args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "accel_c3": accel, "clocks": clocks,
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp", type=int, default=GPU_SPP)
    ap.add_argument("--reduce", default="p2p", choices=["p2p", "nccl"], help="film combine across GPUs")
    ap.add_argument("--no-accel", action="store_true", help="skip the informational C3 accel batch rates")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
