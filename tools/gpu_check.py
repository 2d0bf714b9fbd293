"""Diagnostic run on a GPU box: parity vs the reference + first throughput numbers (not a bench)."""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v3_b200"))
import numpy as np  # noqa: E402
import lmcu  # noqa: E402
from lmcu import abi, scenes  # noqa: E402
from oracle import ref  # noqa: E402


def rmse(a, b):
    return float(np.sqrt(np.mean((np.asarray(a, dtype=np.float64) - b) ** 2)))


def compare_hits(name, hr, hs):
    bad = (hr["primitive"] != hs["primitive"]) | (hr["face"] != hs["face"])
    ties = bad & (hr["t"] == hs["t"])
    nontie = bad & ~ties
    same = ~bad & (hr["primitive"] >= 0)
    print("  [%s] n=%d hit=%.4f mismatch=%d ties=%d non-tie=%d t_bitexact=%s uv_bitexact=%s" % (
        name, len(hr), (hr["primitive"] >= 0).mean(), bad.sum(), ties.sum(), nontie.sum(),
        bool((hr["t"][same] == hs["t"][same]).all()),
        bool(((hr["u"][same] == hs["u"][same]) & (hr["v"][same] == hs["v"][same])).all())), flush=True)
    for i in np.where(nontie)[0][:4]:
        print("     ref", hr[i], "gpu", hs[i])


def section(fn):
    try:
        t = time.time()
        fn()
        print("  (%.1fs)" % (time.time() - t), flush=True)
    except Exception:
        traceback.print_exc()
        sys.stdout.flush()


def main():
    print("devices", lmcu.device_count(), "ref", ref.available(), flush=True)
    have_ref = ref.available()
    if have_ref:
        print("ref threads", ref.init(0), flush=True)

    def c1():
        print("== C1 cornell 36 tris, 1920x1080 camera rays")
        sb = scenes.cornell_box()
        sc = lmcu.Scene(sb)
        print("  build", sc.build_stats.elapsed, "nodes", sc.build_stats.num_wide_nodes, "fp32", sc.build_stats.triangles_fp32)
        rays = scenes.camera_rays(sb.camera, 1920, 1080)
        t = time.time(); hs = sc.intersect(rays, counted=True); dt = time.time() - t
        print("  gpu intersect e2e %.3fs nodes/ray %.2f tris/ray %.2f" % (dt, sc.nodes_per_ray, sc.tris_per_ray))
        t = time.time(); hs = sc.intersect(rays); dt = time.time() - t
        print("  gpu intersect e2e (2nd) %.3fs = %.1f Mrays/s" % (dt, len(rays) / dt / 1e6))
        if have_ref:
            rs = ref.RefScene(sb)
            hr = rs.intersect(rays)
            print("  ref intersect %.3fs = %.2f Mrays/s" % (rs.last_seconds, len(rays) / rs.last_seconds / 1e6))
            compare_hits("C1 cam", hr, hs)
            p = abi.render_params(1920, 1080)
            fr, el, _ = rs.render("raycast", p)
            fg, st = sc.render("raycast", p)
            d = np.abs(fr - fg).max(axis=2)
            tie = ((hr["primitive"] != hs["primitive"]) | (hr["face"] != hs["face"])).reshape(1080, 1920)
            print("  raycast ref %.3fs gpu %.5fs rmse %.3e rmse(non-tie) %.3e max(non-tie) %.3e" % (
                el, st["elapsed"], rmse(fg, fr), rmse(fg[~tie], fr[~tie]), d[~tie].max()))
            rs.close()
        sc.close()

    def c1x():
        print("== C1 with transformed boxes (FP64 triangles)")
        sb = scenes.cornell_box(True)
        sc = lmcu.Scene(sb)
        print("  fp32", sc.build_stats.triangles_fp32)
        rays = np.concatenate([scenes.camera_rays(sb.camera, 640, 360), scenes.random_rays(200000, np.array([-1, 0, -1.]), np.array([1, 2, 1.]))])
        hs = sc.intersect(rays)
        if have_ref:
            rs = ref.RefScene(sb); compare_hits("C1x", rs.intersect(rays), hs); rs.close()
        sc.close()

    def c2():
        print("== C2 cornell + spheres 32k tris")
        sb = scenes.cornell_spheres(64)
        sc = lmcu.Scene(sb)
        print("  tris", sb.num_triangles(), "build %.4fs" % sc.build_stats.elapsed, "nodes", sc.build_stats.num_wide_nodes)
        rays = np.concatenate([scenes.camera_rays(sb.camera, 960, 540), scenes.random_rays(500000, np.array([-1, 0, -1.]), np.array([1, 2, 1.]))])
        hs = sc.intersect(rays, counted=True)
        print("  nodes/ray %.2f tris/ray %.2f" % (sc.nodes_per_ray, sc.tris_per_ray))
        if have_ref:
            rs = ref.RefScene(sb)
            print("  ref build %.2fs" % rs.build_seconds)
            hr = rs.intersect(rays)
            print("  ref %.2f Mrays/s" % (len(rays) / rs.last_seconds / 1e6))
            compare_hits("C2", hr, hs)
            for mode in (2, 1, 0):
                p = abi.render_params(480, 270, spp=64, max_verts=20, sampling_mode=mode)
                fr, el, _ = rs.render("pt", p)
                fg, st = sc.render("pt", p)
                print("  pt mode %d ref mean %s (%.2fs, %.2f Mpaths/s) gpu mean %s (%.4fs) rel %.4f rmse %.4f" % (
                    mode, fr.mean(axis=(0, 1)), el, 480 * 270 * 64 / el / 1e6, fg.mean(axis=(0, 1)), st["elapsed"],
                    abs(fg.mean() - fr.mean()) / fr.mean(), rmse(fg, fr)))
            p = abi.render_params(480, 270, spp=16)
            fr, el, _ = rs.render("ao", p)
            fg, st = sc.render("ao", p)
            print("  ao ref mean %.5f gpu mean %.5f rmse %.4f" % (fr.mean(), fg.mean(), rmse(fg, fr)))
            rs.close()
        for spp in (16, 64):
            p = abi.render_params(1920, 1080, spp=spp, max_verts=20)
            fg, st = sc.render("pt", p, download=False)
            print("  pt 1080p spp=%d: %.3fs  %.1f Mpaths/s  %.1f Mrays/s  rays/path %.2f launches %d" % (
                spp, st["elapsed"], st["paths"] / st["elapsed"] / 1e6, st["rays"] / st["elapsed"] / 1e6,
                st["rays"] / st["paths"], st["kernel_launches"]), flush=True)
        sc.close()

    def c3(n):
        print("== C3 heightfield n=%d" % n)
        sb = scenes.heightfield_scene(n)
        t = time.time(); sc = lmcu.Scene(sb); dt = time.time() - t
        print("  tris", sb.num_triangles(), "build device %.3fs (e2e %.2fs) nodes %d" % (sc.build_stats.elapsed, dt, sc.build_stats.num_wide_nodes))
        rays = scenes.camera_rays(sb.camera, 2048, 2048)
        hs = sc.intersect(rays, counted=True)
        print("  coherent nodes/ray %.2f tris/ray %.2f hit %.3f" % (sc.nodes_per_ray, sc.tris_per_ray, (hs["primitive"] >= 0).mean()))
        t = time.time(); hs = sc.intersect(rays); dt = time.time() - t
        print("  coherent e2e %.1f Mrays/s" % (len(rays) / dt / 1e6))
        P = sb.meshes[0][0]

        def nrm(h):
            tri = P[h["face"]]
            g = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
            return g / np.linalg.norm(g, axis=1, keepdims=True)
        br = scenes.bounce_rays(rays, hs, nrm)
        hb = sc.intersect(br, counted=True)
        print("  incoherent n=%d nodes/ray %.2f tris/ray %.2f hit %.3f" % (len(br), sc.nodes_per_ray, sc.tris_per_ray, (hb["primitive"] >= 0).mean()))
        t = time.time(); hb = sc.intersect(br); dt = time.time() - t
        print("  incoherent e2e %.1f Mrays/s" % (len(br) / dt / 1e6))
        if have_ref and n <= 512:
            rs = ref.RefScene(sb)
            print("  ref build %.2fs" % rs.build_seconds)
            sub = np.concatenate([rays[::16], br[::16]])
            hr = rs.intersect(sub)
            print("  ref %.2f Mrays/s" % (len(sub) / rs.last_seconds / 1e6))
            compare_hits("C3", hr, np.concatenate([hs[::16], hb[::16]]))
            rs.close()
        sc.close()

    section(c1)
    section(c1x)
    section(c2)
    section(lambda: c3(512))
    section(lambda: c3(2236))


if __name__ == "__main__":
    main()
