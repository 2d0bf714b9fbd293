"""Parity tests proper: the CUDA path (through the C ABI of liblmcu.so) against
  (1) the committed golden outputs of the unmodified reference (tests/golden/reference_outputs.npz),
  (2) the reference itself (oracle/_ref) run side by side on the same seeded inputs, when it travelled,
  (3) size-independent properties at BASELINE.json's full sizes.
Gates G1-G3 of SURVEY.md §8c; tolerances are written next to each assert."""
import os
import sys

import numpy as np
import pytest

import lmcu
from lmcu import abi, scenes
from parity import assert_hit_parity, assert_pt_parity, assert_pt_rmse, classify_hits, rmse

pytestmark = pytest.mark.gpu

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs.npz"))


def golden_hits(name):
    h = np.zeros(len(GOLDEN[name + "_t"]), dtype=scenes.HIT_DTYPE)
    h["t"], h["u"], h["v"] = GOLDEN[name + "_t"], GOLDEN[name + "_u"], GOLDEN[name + "_v"]
    h["primitive"], h["face"] = GOLDEN[name + "_prim"], GOLDEN[name + "_face"]
    return h


def golden_rays(name, sb):
    from golden.make_golden import golden_rays as g
    return g(name, sb)


CASES = {
    "cornell": lambda: scenes.cornell_box(),
    "cornell_fp64": lambda: scenes.cornell_box(True),
    "spheres12": lambda: scenes.cornell_spheres(12),
    "heightfield48": lambda: scenes.heightfield_scene(48),
}


@pytest.mark.parametrize("records", ["default", "float"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_g1_closest_hit_vs_golden(name, records, monkeypatch):
    """(primitive, face) equal to accel::sahbvh for every ray; t,u,v bit-identical (gate: 1e-5 relative).  Both triangle
    records: small scenes keep 80-byte FP64 records by default; `float` forces the 48-byte records a large
    float-representable scene (C3) gets, which promote to the same FP64 values on load."""
    if records == "float":
        monkeypatch.setenv("LMCU_TRI64", "0")
    sb = CASES[name]()
    sc = lmcu.Scene(sb)
    assert bool(sc.build_stats.triangles_fp32) == (records == "float" and name != "cornell_fp64")
    hits = sc.intersect(golden_rays(name, sb))
    # measured: the symmetric camera of the fixtures meets shared edges exactly on 22 / 21 / 16 / 0 of the 26 736 rays
    assert_hit_parity(golden_hits(name), hits, max_special={"cornell": 22, "cornell_fp64": 21, "spheres12": 16, "heightfield48": 0}[name])
    if name == "spheres12":
        film, _ = sc.render("pt", abi.render_params(96, 54, spp=8, max_verts=8))
        assert np.isfinite(film).all() and film.mean() > 0.01
    sc.close()


def test_g1_full_c1_vs_reference(reference):
    """C1: all 2 073 600 pixel-centre rays of the 1920x1080 Cornell box."""
    sb = scenes.cornell_box()
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    rays = scenes.camera_rays(sb.camera, 1920, 1080)
    c = assert_hit_parity(rs.intersect(rays), sc.intersect(rays), max_special=167)      # measured: 151 ties + 16 boundary misses
    print("C1 special cases:", {k: v for k, v in c.items() if k != "bad_mask"})
    rs.close(); sc.close()


def test_g1_asymmetric_camera_has_no_special_cases(reference):
    """Off the symmetry plane no pixel-centre ray meets a scene edge exactly: zero mismatches of any kind."""
    sb = scenes.cornell_spheres(32)
    sb.set_camera(position=(0.13, 1.07, 4.9), center=(0.02, 0.97, 0), up=(0, 1, 0), vfov=31, aspect=16 / 9)
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    rays = scenes.camera_rays(sb.camera, 960, 540)
    c = classify_hits(rs.intersect(rays), sc.intersect(rays))
    assert c["mismatch"] == 0, c
    rs.close(); sc.close()


def test_g1_incoherent_and_window(reference):
    sb = scenes.cornell_spheres(32)
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    rays = scenes.random_rays(400000, np.array([-1, 0, -1.]), np.array([1, 2, 1.]), seed=5)
    assert_hit_parity(rs.intersect(rays), sc.intersect(rays), max_special=0)
    rays["tmin"], rays["tmax"] = 0.25, 0.8
    hr, hs = rs.intersect(rays), sc.intersect(rays)
    assert_hit_parity(hr, hs, max_special=0)
    ok = hs["primitive"] >= 0
    assert ((hs["t"][ok] >= 0.25) & (hs["t"][ok] <= 0.8)).all()
    rs.close(); sc.close()


def test_g1_tiny_and_ragged_scenes(reference):
    for nf in (1, 2, 3, 4, 5, 9, 25):
        sb = scenes.SceneBuilder()
        rng = np.random.RandomState(nf)
        tri = scenes._f32(rng.rand(nf, 3, 3) * 2 - 1)
        sb.add_primitive(sb.add_mesh(tri, scenes.flat_normals(tri)), sb.diffuse((.5, .5, .5)))
        sb.set_camera(position=(0, 0, 4), center=(0, 0, 0), up=(0, 1, 0), vfov=40, aspect=1)
        sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
        rays = scenes.camera_rays(sb.camera, 128, 128)
        assert_hit_parity(rs.intersect(rays), sc.intersect(rays), max_special=0)
        rs.close(); sc.close()


def test_g1_instanced_transforms(reference):
    """func_render_instancing pattern: one mesh, many rotated/scaled instances (FP64 triangles)."""
    sb = scenes.instanced_spheres(grid=3, num_theta=12)
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    assert not sc.build_stats.triangles_fp32
    rays = np.concatenate([scenes.camera_rays(sb.camera, 480, 270),
                           scenes.random_rays(100000, np.array([-2, 0, -2.]), np.array([2, 3, 2.]))])
    assert_hit_parity(rs.intersect(rays), sc.intersect(rays), max_special=0)
    rs.close(); sc.close()


def test_accel_save_load_round_trip(tmp_path):
    """SURVEY.md 8f N4 (Component::save / load of an Accel): the built BVH written by lmcu_accel_save and restored by
    lmcu_accel_load (a) into a scene with the same description: hits identical to the built scene's and the same pt render, no
    builder run; (b) into a fresh accel-only scene: identical hits; (c) a corrupted, truncated or foreign blob is refused."""
    for sb in (scenes.cornell_spheres(16), scenes.cornell_box(True), scenes.textured_scene(tmp_path, 8, "obj")):
        rays = np.concatenate([scenes.camera_rays(sb.camera, 128, 72), scenes.random_rays(4000, np.array([-1, 0, -1.]), np.array([1, 2, 1.]))])
        a = lmcu.Scene(sb)
        blob = a.save_accel()
        ha = a.intersect(rays)
        p = abi.render_params(96, 54, spp=8, max_verts=8)
        fa, _ = a.render("pt", p)
        a.close()
        b = lmcu.Scene(sb, accel_blob=blob)
        assert b.build_stats.elapsed == 0.0
        hb = b.intersect(rays)
        fb, _ = b.render("pt", p)
        b.close()
        c = lmcu.Scene(None, accel_blob=blob)
        hc = c.intersect(rays)
        c.close()
        for h in (hb, hc):
            assert all(np.array_equal(ha[k], h[k]) for k in ("t", "u", "v", "primitive", "face"))
        assert np.allclose(fa, fb, rtol=1e-4, atol=1e-5)   # same seed and paths; only the order of the film's atomic adds varies
    bad = bytearray(blob)
    bad[len(bad) // 2] ^= 0x40
    for broken in (bytes(bad), blob[:len(blob) // 2], b"not a blob" * 20):
        with pytest.raises(lmcu.LmcuError) as e:
            lmcu.Scene(None, accel_blob=broken)
        assert e.value.code == 1
    with pytest.raises(lmcu.LmcuError) as e:           # a blob of another scene
        lmcu.Scene(scenes.cornell_box(), accel_blob=blob)
    assert e.value.code == 1
    # same triangle / primitive counts, other geometry or other transforms: refused by the description digest (the
    # world-space triangles would come from the blob, materials and primitives from the upload)
    base = scenes.instanced_spheres(grid=2, num_theta=8)
    blob = lmcu.Scene(base).save_accel()
    moved = scenes.instanced_spheres(grid=2, num_theta=8)
    moved.primitives[1]["M"] = moved.primitives[1]["M"].copy()
    moved.primitives[1]["M"][0, 3] += 0.25
    bent = scenes.instanced_spheres(grid=2, num_theta=8)
    bent.meshes[0][0][3, 1, 2] += 0.125
    for other in (moved, bent):
        with pytest.raises(lmcu.LmcuError) as e:
            lmcu.Scene(other, accel_blob=blob)
        assert e.value.code == 1 and "other geometry" in str(e.value)
    lmcu.Scene(scenes.instanced_spheres(grid=2, num_theta=8), accel_blob=blob).close()
    # a mesh without vertex normals (the geometric normal is used): its unused normal slots must not leak into the digest
    tri = scenes._f32(np.random.RandomState(4).rand(40, 3, 3) * 2 - 1)
    def flat():
        sb = scenes.SceneBuilder()
        sb.add_primitive(sb.add_mesh(tri), sb.diffuse((.5, .5, .5)))
        sb.set_camera(position=(0, 0, 4), center=(0, 0, 0), up=(0, 1, 0), vfov=40, aspect=1)
        return sb
    keep = [lmcu.Scene(scenes.cornell_spheres(8)) for _ in range(2)]          # churn the allocator between the two scenes
    blob2 = lmcu.Scene(flat()).save_accel()
    for k in keep:
        k.close()
    lmcu.Scene(flat(), accel_blob=blob2).close()
    # a blob whose checksum is right but whose indices point outside its own arrays (a writer of another version):
    # every index is range-checked on the host before anything reaches the device
    import struct
    raw = bytearray(blob)
    hdr = 8 + 6 * 4 + 3 * 8 + 6 * 4 + 2 * 4             # AccelBlobHeader: magic, 6 counts, payload bytes, checksum, digest, bounds, pad
    struct.pack_into("<I", raw, hdr + 16, 0x7fffffff)       # WideNode 0: child_base
    _rechecksum(raw, hdr)
    with pytest.raises(lmcu.LmcuError) as e:
        lmcu.Scene(None, accel_blob=bytes(raw))
    assert e.value.code == 1 and "outside the node array" in str(e.value)


def _rechecksum(raw, hdr):
    """checksum64 of lmcu.cu over the payload, written back into the header (tests only)."""
    import struct
    M = (1 << 64) - 1
    h = 1469598103934665603
    body = memoryview(raw)[hdr:]
    n8 = len(body) // 8
    for (w,) in struct.iter_unpack("<Q", body[:n8 * 8]):
        h = ((h ^ w) * 1099511628211) & M
        h ^= h >> 29
    for b in body[n8 * 8:]:
        h = ((h ^ b) * 1099511628211) & M
    struct.pack_into("<Q", raw, 8 + 6 * 4 + 8, h)


def test_film_accumulates_in_fp64():
    """the film used to be ONE fp32 sum per channel until the final rescale; past ~2^16 a float's ulp is 0.008
    and small contributions vanish (a darkening bias the <= 256 spp parity tests never saw).  Now the fp32 film is folded
    into FP64 sums every few dozen samples per pixel (k_film_fold): 65 536 spp in ONE call must equal the FP64 sum, on the
    host, of 64 calls of 1024 spp over the same sample indices (same counter-RNG streams => same paths)."""
    sb = scenes.cornell_spheres(8)
    sc = lmcu.Scene(sb)
    w, h, n, chunk = 32, 18, 65536, 1024
    one, st = sc.render("pt", abi.render_params(w, h, spp=n, max_verts=6))
    acc = np.zeros((h, w, 3))
    for k in range(n // chunk):
        f, _ = sc.render("pt", abi.render_params(w, h, spp=chunk, max_verts=6, sample_offset=k * chunk, rescale=0))
        acc += f
    acc /= n
    sc.close()
    assert st["processed"] == n
    assert np.allclose(one, acc, rtol=2e-4, atol=1e-6), np.abs(one / np.maximum(acc, 1e-9) - 1).max()


def test_pt_progress_hook():
    """LmcuRenderParams::progress (the progress::update hook of the reference's schedulers, src/scheduler.cpp:45): called
    between wavefront batches with non-decreasing path counts that never exceed the job."""
    sb = scenes.cornell_spheres(8)
    sc = lmcu.Scene(sb)
    seen = []
    cb = abi.PROGRESS_FN(lambda user, done, total, elapsed: seen.append((done, total, elapsed)))
    w, h, spp = 640, 360, 256
    sc.render("pt", abi.render_params(w, h, spp=spp, max_verts=8, progress=cb), download=False)
    sc.close()
    assert len(seen) >= 2
    done = [d for d, _, _ in seen]
    assert all(t == w * h * spp for _, t, _ in seen) and done == sorted(done) and done[-1] <= w * h * spp and done[-1] > 0
    assert all(b[2] >= a[2] for a, b in zip(seen, seen[1:]))


def test_accel_scalar_and_concurrent_batches():
    """lmcu_accel_intersect1 (the scalar contract) answers exactly like the batch entry; batches launched back to back on
    two CUDA streams do not share a work cursor."""
    import torch
    sb = scenes.cornell_spheres(16)
    sc = lmcu.Scene(sb)
    rays = scenes.random_rays(200000, np.array([-1, 0, -1.]), np.array([1, 2, 1.]), seed=3)
    want = sc.intersect(rays)
    for i in range(0, 2000, 7):
        one = sc.intersect1(rays[i])
        assert all(one[k] == want[k][i] for k in ("t", "u", "v", "primitive", "face"))
    d_rays = torch.from_numpy(rays.view(np.float64).reshape(-1, 8)).cuda()
    half = len(rays) // 2
    outs = [torch.empty((half, 4), dtype=torch.float64, device="cuda") for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(3):
        for k in range(2):
            sc.intersect_device(d_rays[k * half:(k + 1) * half].data_ptr(), half, outs[k].data_ptr(), streams[k].cuda_stream)
    torch.cuda.synchronize()
    got = np.concatenate([o.cpu().numpy().view(scenes.HIT_DTYPE).reshape(-1) for o in outs])
    assert all(np.array_equal(got[k], want[k][:2 * half]) for k in ("t", "u", "v", "primitive", "face"))
    sc.close()


def test_pool_traversal_variant_is_bit_identical(monkeypatch):
    """trace2.cuh (LMCU_TRACE=pool: K = 2 rays per lane in a shared-memory pool, one phase per warp iteration) is an opt-in
    scheduling variant of the same arithmetic: hits identical to the default kernel bit for bit, same pt image up to the
    order of the film's atomic adds.  (Measured slower on B200 — profiles/r02_ncu_summary.md — hence opt-in.)"""
    sb = scenes.cornell_spheres(24)
    rays = np.concatenate([scenes.camera_rays(sb.camera, 320, 180), scenes.random_rays(60000, np.array([-1, 0, -1.]), np.array([1, 2, 1.]))])
    p = abi.render_params(160, 90, spp=32, max_verts=12)
    a = lmcu.Scene(sb)
    ha = a.intersect(rays)
    fa, _ = a.render("pt", p)
    a.close()
    monkeypatch.setenv("LMCU_TRACE", "pool")
    b = lmcu.Scene(sb)
    hb = b.intersect(rays)
    fb, _ = b.render("pt", p)
    b.close()
    assert all(np.array_equal(ha[k], hb[k]) for k in ("t", "u", "v", "primitive", "face"))
    assert np.allclose(fa, fb, rtol=1e-4, atol=1e-5)


def test_g1_edge_cases(reference):
    """The edge-case batches of tests/edge_cases.py through the CUDA queue traversal (k_intersect_q)."""
    import edge_cases
    for name, (sb, rays, special) in edge_cases.cases().items():
        rs, sc = reference.RefScene(sb), lmcu.Scene(sb)
        hr, hg = rs.intersect(rays), sc.intersect(rays)
        assert_hit_parity(hr, hg, max_special=special)
        # idempotence: a hit found at t is found again with tmax = t (the upper bound is inclusive).  Checked against our
        # own first answer, not the reference's: for an axis-aligned triangle the reference's FP64 slab test and its
        # triangle test round the same t differently, so accel::sahbvh itself loses such a hit when tmax = t
        ok = hg["primitive"] >= 0
        if ok.any():
            at = rays[ok].copy(); at["tmax"] = hg["t"][ok]
            again = sc.intersect(at)
            assert all(np.array_equal(again[k], hg[k][ok]) for k in ("t", "u", "v", "primitive", "face")), name
        rs.close(); sc.close()


def test_empty_batch_and_errors():
    sb = scenes.cornell_box()
    sc = lmcu.Scene(sb)
    assert len(sc.intersect(np.zeros(0, dtype=scenes.RAY_DTYPE))) == 0
    with pytest.raises(lmcu.LmcuError):
        sc.render("pt", abi.render_params(64, 64, spp=0))
    with pytest.raises(lmcu.LmcuError):
        sc.render("pt", abi.render_params(64, 64, spp=1, max_verts=1))
    sc.close()
    empty = scenes.SceneBuilder()
    with pytest.raises(lmcu.LmcuError) as e:
        lmcu.Scene(empty)
    assert e.value.code == 2          # lm::Error::Unsupported, "Missing primitives" (include/lm/scene.h:232-238)


@pytest.mark.parametrize("name", ["cornell", "spheres12"])
def test_g2_raycast_vs_golden(name):
    """Deterministic rays: image RMSE <= 1e-6 outside tie pixels (FP32 shading noise only)."""
    sb = CASES[name]()
    sc = lmcu.Scene(sb)
    film, _ = sc.render("raycast", abi.render_params(192, 108))
    hits = sc.intersect(golden_rays(name, sb)[:192 * 108])
    bad = classify_hits(golden_hits(name)[:192 * 108], hits)["bad_mask"].reshape(108, 192)
    assert bad.mean() < 3e-3
    assert rmse(film[~bad], GOLDEN[name + "_raycast"][~bad]) <= 1e-6
    sc.close()


def test_g2_raycast_full_c1_vs_reference(reference):
    sb = scenes.cornell_box()
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    rays = scenes.camera_rays(sb.camera, 1920, 1080)
    bad = classify_hits(rs.intersect(rays), sc.intersect(rays))["bad_mask"].reshape(1080, 1920)
    for kw in ({}, {"visualize_normal": True}, {"use_constant_color": True}, {"color": (.2, .4, .6)}, {"bg_color": (.1, .2, .3)}):
        p = abi.render_params(1920, 1080, **kw)
        fr, _, _ = rs.render("raycast", p)
        fg, _ = sc.render("raycast", p)
        assert rmse(fg[~bad], fr[~bad]) <= 1e-6, kw
        assert np.abs(fg[~bad] - fr[~bad]).max() <= 1e-5, kw
    rs.close(); sc.close()


@pytest.mark.parametrize("tag,mode", [("mis", abi.SAMPLING_MIS), ("naive", abi.SAMPLING_NAIVE)])
def test_g3_pt_vs_golden(tag, mode):
    """Equal spp, different RNG streams: means within 1 %, no structure in the difference image."""
    sb = CASES["spheres12"]()
    sc = lmcu.Scene(sb)
    film, st = sc.render("pt", abi.render_params(96, 54, spp=1024, max_verts=12, sampling_mode=mode))
    assert st["processed"] == 1024 and st["paths"] == 96 * 54 * 1024
    assert_pt_parity(GOLDEN["spheres12_pt_" + tag], film, 1024, rel_mean=0.01)
    sc.close()


def test_g3_ao_vs_golden():
    sb = CASES["spheres12"]()
    sc = lmcu.Scene(sb)
    film, _ = sc.render("ao", abi.render_params(96, 54, spp=256))
    assert_pt_parity(GOLDEN["spheres12_ao"], film, 256, rel_mean=0.01)
    sc.close()


@pytest.mark.parametrize("mode", [abi.SAMPLING_MIS, abi.SAMPLING_NEE, abi.SAMPLING_NAIVE])
def test_g3_pt_c2_vs_reference(reference, mode):
    """C2 scene (32 268 triangles, diffuse + glass spheres, area light) at reduced resolution."""
    sb = scenes.cornell_spheres(64)
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    # 256 spp as 4 independent estimates of 64 spp each (reference: 4 seeds; ours: 4 disjoint blocks of sample indices)
    B, spp = 4, 64
    R = np.stack([rs.render("pt", abi.render_params(240, 136, spp=spp, max_verts=20, sampling_mode=mode, seed=42 + b))[0] for b in range(B)])
    G = np.stack([sc.render("pt", abi.render_params(240, 136, spp=spp, max_verts=20, sampling_mode=mode, sample_offset=b * spp))[0] for b in range(B)])
    fr, fg = R.mean(axis=0), G.mean(axis=0)
    assert_pt_parity(fr, fg, B * spp, rel_mean=0.01)
    ratio = assert_pt_rmse(R, G, k=1.25)                 # per-channel RMSE <= 1.25 x the Monte-Carlo prediction
    # calibration: a second set of reference runs with other seeds passes the same gates
    R2 = np.stack([rs.render("pt", abi.render_params(240, 136, spp=spp, max_verts=20, sampling_mode=mode, seed=7 + b))[0] for b in range(B)])
    assert_pt_parity(fr, R2.mean(axis=0), B * spp, rel_mean=0.01)
    print("rmse / predicted: ours", ratio, "reference vs reference", assert_pt_rmse(R, R2, k=1.25))
    rs.close(); sc.close()


def test_g3_pt_high_spp_bias(reference):
    """What FP32 shading + FP32 rays cost against the all-FP64 reference, asserted: at 8192 spp on a 96x54 film the Monte-Carlo
    error of the difference of the image means is ~6e-4 relative (measured with the CPU build of the device code: -2e-4 +-
    3e-4 at 32 768 spp), so the 0.3 % gate below is a gate on BIAS."""
    sb = scenes.cornell_spheres(64)
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    p = abi.render_params(96, 54, spp=8192, max_verts=20)
    fr, _, _ = rs.render("pt", p)
    fg, _ = sc.render("pt", p)
    c = assert_pt_parity(fr, fg, 8192, rel_mean=0.003)
    print("relative difference of the image means at 8192 spp:", c["rel_mean"])
    assert (c["rel_mean"] <= 0.003).all()
    rs.close(); sc.close()


def test_g3_pt_all_materials_env_light(reference):
    """C4-style materials (glossy / glass / mirror / diffuse) under light::envconst, instanced."""
    sb = scenes.instanced_spheres(grid=2, num_theta=24)
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    B, spp = 4, 64
    R = np.stack([rs.render("pt", abi.render_params(192, 108, spp=spp, max_verts=20, seed=42 + b))[0] for b in range(B)])
    G = np.stack([sc.render("pt", abi.render_params(192, 108, spp=spp, max_verts=20, sample_offset=b * spp))[0] for b in range(B)])
    assert_pt_parity(R.mean(axis=0), G.mean(axis=0), B * spp, rel_mean=0.015)
    assert_pt_rmse(R, G, k=1.25)
    rs.close(); sc.close()


def test_g3_pt_env_image_light(reference, tmp_path):
    """C4 variant: light::env with a procedurally written 512x256 .hdr (sun lobe + sky gradient), rot and scale."""
    env = scenes.env_light(str(tmp_path / "env.hdr"), w=512, h=256, rot=75.0, scale=0.8)
    sb = scenes.instanced_spheres(grid=2, num_theta=24, env=env)
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    p = abi.render_params(192, 108)
    fr, _, _ = rs.render("raycast", p)
    fg, _ = sc.render("raycast", p)
    d = np.abs(fg - fr).max(axis=2)
    assert (d > 1e-4).mean() < 3e-3 and rmse(fg[d <= 1e-4], fr[d <= 1e-4]) <= 1e-6
    for mode in (abi.SAMPLING_MIS, abi.SAMPLING_NEE, abi.SAMPLING_NAIVE):
        p = abi.render_params(160, 90, spp=256, max_verts=12, sampling_mode=mode)
        fr, _, _ = rs.render("pt", p)
        fg, _ = sc.render("pt", p)
        assert_pt_parity(fr, fg, 256, rel_mean=0.015)
    rs.close(); sc.close()


@pytest.mark.parametrize("kind", ["obj", "mix"])
def test_g3_textured_and_mixture_materials(reference, tmp_path, kind):
    """SURVEY.md 8f N2, the OBJ-import material set: mapKd textures (texture::bitmap from a PNG, nearest texel),
    material::mixture_wavefrontobj with and without an alpha mask, material::mask, and the constant-weight mixtures (rr and
    marginalized; diffuse / glossy / mirror / mask entries).  Raycast shows Material::reflectance (deterministic: G2
    tolerances), pt in all three sampling modes passes the statistical gate G3."""
    sb = scenes.textured_scene(tmp_path, 24, kind)
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    p = abi.render_params(320, 180)
    fr, _, _ = rs.render("raycast", p)
    fg, _ = sc.render("raycast", p)
    d = np.abs(fg - fr).max(axis=2)
    assert (d > 1e-4).mean() < 3e-3                    # pixels whose ray meets an edge exactly (ties) or a texel border
    assert rmse(fg[d <= 1e-4], fr[d <= 1e-4]) <= 1e-6
    for mode in (abi.SAMPLING_MIS, abi.SAMPLING_NEE, abi.SAMPLING_NAIVE):
        p = abi.render_params(160, 90, spp=256, max_verts=12, sampling_mode=mode)
        fr, _, _ = rs.render("pt", p)
        fg, _ = sc.render("pt", p)
        assert_pt_parity(fr, fg, 256, rel_mean=0.015)
    rs.close(); sc.close()


@pytest.mark.parametrize("kind", ["obj", "mix"])
def test_textured_and_mixture_materials_vs_golden(tmp_path, kind):
    """The same material set against the COMMITTED outputs of the reference (tests/golden/reference_outputs_materials.npz):
    needs no reference on the box."""
    golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs_materials.npz"))
    sb = scenes.textured_scene(tmp_path, 12, kind)
    sc = lmcu.Scene(sb)
    fg, _ = sc.render("raycast", abi.render_params(192, 108))
    fr = golden[kind + "_raycast"]
    d = np.abs(fg - fr).max(axis=2)
    assert (d > 1e-4).mean() < 3e-3 and rmse(fg[d <= 1e-4], fr[d <= 1e-4]) <= 1e-6
    fg, _ = sc.render("pt", abi.render_params(96, 54, spp=1024, max_verts=12))
    assert_pt_parity(golden[kind + "_pt_mis"], fg, 1024, rel_mean=0.01)
    sc.close()


def test_composite_material_validation():
    """Mixture entries must be single-component lobes (the reference hands an entry component 0 whatever it is,
    material_mixture.cpp:67-78), texture indices must exist."""
    sb = scenes.cornell_box()
    g = sb.glass(1.5)
    sb.primitives[0]["material"] = sb.mixture_rr([(g, 1.0)])
    with pytest.raises(lmcu.LmcuError) as e:
        lmcu.Scene(sb)
    assert e.value.code == 2
    sb = scenes.cornell_box()
    sb.primitives[0]["material"] = sb.diffuse((1, 1, 1), map_kd=3)
    with pytest.raises(lmcu.LmcuError) as e:
        lmcu.Scene(sb)
    assert e.value.code == 1


def test_g3_pt_point_and_directional_lights(reference):
    """light::point + light::directional (NEE only, MIS weight 1), mixed with an area light in a second scene."""
    sb = scenes.delta_lights_scene(2, 24)
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    for mode in (abi.SAMPLING_MIS, abi.SAMPLING_NEE):
        p = abi.render_params(160, 90, spp=256, max_verts=12, sampling_mode=mode)
        fr, _, _ = rs.render("pt", p)
        fg, _ = sc.render("pt", p)
        assert_pt_parity(fr, fg, 256, rel_mean=0.01)
    rs.close(); sc.close()
    sb = scenes.cornell_spheres(16)
    sb.add_point_light((3, 3, 3), (0.3, 1.5, 0.2))
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    p = abi.render_params(160, 90, spp=256, max_verts=12)
    fr, _, _ = rs.render("pt", p)
    fg, _ = sc.render("pt", p)
    assert_pt_parity(fr, fg, 256, rel_mean=0.01)
    rs.close(); sc.close()


def test_g3_pt_image_mode_vs_reference(reference):
    """primary_ray_sampling_mode=image + scheduler::spi::sample: same number of samples over the whole film."""
    sb = scenes.cornell_spheres(32)
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    w, h, spp = 256, 144, 128
    p = abi.render_params(w, h, max_verts=20, primary_ray_sampling_mode=abi.PRIMARY_IMAGE, num_samples=w * h * spp)
    fr, _, pr = rs.render("pt", p)
    fg, st = sc.render("pt", p)
    assert pr == st["processed"] == st["paths"] == w * h * spp
    assert_pt_parity(fr, fg, spp, rel_mean=0.01)
    # image mode and pixel mode estimate the same image
    fp, _ = sc.render("pt", abi.render_params(w, h, spp=spp, max_verts=20))
    assert_pt_parity(fp, fg, spp, rel_mean=0.01)
    # sample split over image-mode sample indices is additive and seed-stable
    half = w * h * spp // 2
    pa = abi.render_params(w, h, max_verts=20, primary_ray_sampling_mode=abi.PRIMARY_IMAGE, num_samples=half, rescale=0)
    pb = abi.render_params(w, h, max_verts=20, primary_ray_sampling_mode=abi.PRIMARY_IMAGE, num_samples=half, rescale=0, sample_offset=half)
    fa, _ = sc.render("pt", pa)
    fb, _ = sc.render("pt", pb)
    assert np.allclose((fa + fb) * (w * h / (2.0 * half)), fg, rtol=2e-3, atol=2e-4)
    rs.close(); sc.close()


@pytest.mark.parametrize("image", [False, True])
def test_pt_time_schedulers(reference, image):
    """scheduler::spp::time / scheduler::spi::time (src/scheduler.cpp:75-103,162-187): whole passes until render_time
    has elapsed; `processed` reports what was rendered and the film is scaled by it.  Not deterministic => checked
    by its contract and statistically against a fixed-spp reference image."""
    import time
    sb = scenes.cornell_spheres(16)
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    w, h = 160, 90
    sc.render("pt", abi.render_params(w, h, spp=4))                        # allocate the pools outside the budget
    kw = dict(primary_ray_sampling_mode=abi.PRIMARY_IMAGE, samples_per_iter=5000) if image else {}
    budget = 0.12
    p = abi.render_params(w, h, max_verts=12, scheduler=abi.SCHED_TIME, render_time=budget, **kw)
    t0 = time.perf_counter()
    fg, st = sc.render("pt", p)
    wall = time.perf_counter() - t0
    assert budget <= wall < 1.0
    assert st["processed"] >= 1 and st["paths"] == st["processed"] * (1 if image else w * h)
    if image:
        assert st["processed"] % 5000 == 0
    spp_equiv = st["paths"] / (w * h)
    assert spp_equiv > 256                                                 # a B200 renders far more than that in 0.12 s
    # the passes walk consecutive sample indices: the result is the fixed-count render of the same `processed`
    # (same counter-RNG streams; fp32 summation order aside)
    kf = dict(kw, num_samples=st["processed"]) if image else dict(spp=st["processed"])
    kf.pop("samples_per_iter", None)
    ff, _ = sc.render("pt", abi.render_params(w, h, max_verts=12, **kf))
    assert np.allclose(ff, fg, rtol=2e-3, atol=2e-4)
    # against the reference: its 1024-spp image.  The block z-test assumes comparable noise in both images; the budget
    # buys tens of thousands of spp here, so the difference image is the reference's own (skewed) 1024-spp noise and the
    # z-test is widened accordingly — the sharp check is the fixed-count equality above, and fixed-count renders are held
    # to the full gate in the tests before this one
    fr, _, _ = rs.render("pt", abi.render_params(w, h, spp=1024, max_verts=12))
    assert_pt_parity(fr, fg, 1024, rel_mean=0.01, z_frac=0.04)
    # render_time = 0: exactly one pass (the reference checks the clock after each pass)
    fz, sz = sc.render("pt", abi.render_params(w, h, max_verts=12, scheduler=abi.SCHED_TIME, render_time=0.0, **kw))
    assert sz["processed"] >= 1 and np.isfinite(fz).all() and fz.mean() > 0
    rs.close(); sc.close()


def test_wavefront_count_does_not_change_the_image():
    """1, 2, 3 and 4 staggered wavefronts trace the same paths (the RNG is keyed by pixel and sample index, not by the
    loop that carries the path): identical images up to the order of the film's atomic adds; processed / paths / rays
    counters identical.  Includes a job smaller than one queue and image mode."""
    sb = scenes.cornell_spheres(12)
    sc = lmcu.Scene(sb)
    for kw in (dict(spp=24), dict(spp=1), dict(primary_ray_sampling_mode=abi.PRIMARY_IMAGE, num_samples=160 * 90 * 16)):
        ref_film, ref_st = None, None
        for waves in (1, 2, 3, 4):
            film, st = sc.render("pt", abi.render_params(160, 90, max_verts=10, wavefronts=waves, **kw))
            if ref_film is None:
                ref_film, ref_st = film, st
                continue
            assert np.allclose(film, ref_film, rtol=2e-4, atol=1e-5), (waves, kw)
            assert all(st[k] == ref_st[k] for k in ("processed", "paths", "rays", "extend_rays", "shadow_rays")), (waves, kw)
    with pytest.raises(lmcu.LmcuError):
        sc.render("pt", abi.render_params(64, 36, spp=1, wavefronts=9))
    sc.close()


def test_sample_split_is_additive():
    """Sample split (multi-GPU partition): two half renders with disjoint sample offsets sum to the full render."""
    sb = scenes.cornell_spheres(16)
    sc = lmcu.Scene(sb)
    full, _ = sc.render("pt", abi.render_params(128, 72, spp=64, max_verts=8, rescale=0))
    a, _ = sc.render("pt", abi.render_params(128, 72, spp=32, max_verts=8, rescale=0, sample_offset=0))
    b, _ = sc.render("pt", abi.render_params(128, 72, spp=32, max_verts=8, rescale=0, sample_offset=32))
    # same counter-RNG streams; only the fp32 atomic summation order differs
    assert np.allclose(a + b, full, rtol=1e-4, atol=1e-4)
    sc.close()


def test_full_size_properties_c3():
    """C3 at full size (9 999 392 triangles): properties that need no oracle.
    - every hit lies inside [tmin,tmax] and has valid barycentrics
    - a second-hit query started just behind the first hit never returns the same triangle
    - re-tracing with tmax = t_hit returns the same hit (idempotence, inclusive upper bound)"""
    sb = scenes.heightfield_scene(2236)
    sc = lmcu.Scene(sb)
    assert sc.build_stats.num_triangles == 9999392 and sc.build_stats.triangles_fp32     # 48-byte float records at this size
    rays = scenes.camera_rays(sb.camera, 1024, 1024)
    h = sc.intersect(rays)
    ok = h["primitive"] >= 0
    assert ok.mean() > 0.3
    assert (h["t"][ok] >= 1e-3).all() and (h["u"][ok] >= 0).all() and (h["v"][ok] >= 0).all()
    assert (h["u"][ok] + h["v"][ok] <= 1 + 1e-12).all()
    r2 = rays[ok].copy()
    r2["tmax"] = h["t"][ok]
    h2 = sc.intersect(r2)
    assert (h2["face"] == h["face"][ok]).all() and (h2["t"] == h["t"][ok]).all()
    r3 = rays[ok].copy()
    r3["tmin"] = h["t"][ok] * (1 + 1e-9)
    h3 = sc.intersect(r3)
    both = h3["primitive"] >= 0
    assert (h3["face"][both] != h["face"][ok][both]).all()
    assert (h3["t"][both] > h["t"][ok][both]).all()
    sc.close()


def test_full_size_properties_c4():
    """C4 at full size: 1000 rotated/scaled instances of a 49 728-triangle sphere + ground = 49 728 002 FP64 triangles,
    glossy / glass / mirror / diffuse, environment light.  No oracle can hold this scene (the reference needs > 15 GB and
    an O(n log^2 n) build), so the checks are size-independent: re-trace idempotence, second-hit monotonicity, hit
    points on the instanced spheres' surfaces, and a finite, energy-bounded pt image under a unit-radiance sky."""
    sb = scenes.instanced_spheres(grid=10, num_theta=112)
    sc = lmcu.Scene(sb)
    assert sc.build_stats.num_triangles == 49728002 and not sc.build_stats.triangles_fp32
    rays = scenes.camera_rays(sb.camera, 640, 360)
    h = sc.intersect(rays)
    ok = h["primitive"] >= 0
    assert ok.mean() > 0.5
    assert (h["u"][ok] >= 0).all() and (h["v"][ok] >= 0).all() and (h["u"][ok] + h["v"][ok] <= 1 + 1e-12).all()
    # hit points of sphere instances lie on a sphere of radius 0.4 * scale around the instance origin (faceted: within 0.1 %)
    sph = ok & (h["primitive"] < 1000)
    P = rays["o"][sph] + rays["d"][sph] * h["t"][sph, None]
    M = np.stack([np.asarray(sb.primitives[i]["M"]) for i in range(1000)])
    Mi = M[h["primitive"][sph]]
    centre = Mi[:, :3, 3]
    scale = np.linalg.norm(Mi[:, :3, 0], axis=1)
    r = np.linalg.norm(P - centre, axis=1) / (0.4 * scale)
    assert sph.sum() > 1000 and r.max() <= 1 + 1e-9 and r.min() >= np.cos(np.pi / 112) - 1e-6
    r2 = rays[ok].copy()
    r2["tmax"] = h["t"][ok]
    h2 = sc.intersect(r2)
    assert (h2["primitive"] == h["primitive"][ok]).all() and (h2["face"] == h["face"][ok]).all() and (h2["t"] == h["t"][ok]).all()
    r3 = rays[ok].copy()
    r3["tmin"] = h["t"][ok] * (1 + 1e-9)
    h3 = sc.intersect(r3)
    both = h3["primitive"] >= 0
    assert (h3["t"][both] > h["t"][ok][both]).all()
    film, st = sc.render("pt", abi.render_params(320, 180, spp=16, max_verts=20))
    assert np.isfinite(film).all() and st["paths"] == 320 * 180 * 16
    assert 0.05 < film.mean() < 1.0 + 1e-3          # Le = 1 sky, all albedos < 1 (glass eta^2 factors cancel on exit)
    sc.close()


@pytest.mark.parametrize("scene", ["spheres", "env", "delta", "obj"])
def test_g3_lt_vs_reference(reference, tmp_path, scene):
    """renderer::lt on the device (lmcu_render_lt) against the reference's renderer::lt at equal sample
    counts — area light + glass / diffuse, constant environment + glossy / mirror, point + directional lights, the
    OBJ-import material set.  G3 gates: image means, block z-test, RMSE against the Monte-Carlo prediction."""
    sb = {"spheres": lambda: scenes.cornell_spheres(32), "env": lambda: scenes.instanced_spheres(grid=2, num_theta=24),
          "delta": lambda: scenes.delta_lights_scene(2, 16), "obj": lambda: scenes.textured_scene(tmp_path, 12, "obj")}[scene]()
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    w, h, B, spp = 160, 90, 4, 64
    n = w * h * spp
    img = dict(max_verts=12, num_samples=n, primary_ray_sampling_mode=abi.PRIMARY_IMAGE)
    R = np.stack([rs.render("renderer::lt", abi.render_params(w, h, seed=42 + b, **img))[0] for b in range(B)])
    G = np.stack([sc.render("lt", abi.render_params(w, h, sample_offset=b * n, **img))[0] for b in range(B)])
    assert G.mean() > 0.005
    assert_pt_parity(R.mean(axis=0), G.mean(axis=0), B * spp, rel_mean=0.02)
    assert_pt_rmse(R, G, k=1.25)
    # sample-split additivity: two halves rendered as sums add up to the whole (same counter-RNG streams)
    whole, st = sc.render("lt", abi.render_params(w, h, **img))
    assert st["processed"] == n
    parts = [sc.render("lt", abi.render_params(w, h, max_verts=12, num_samples=n // 2, sample_offset=k * (n // 2), rescale=0,
                                               primary_ray_sampling_mode=abi.PRIMARY_IMAGE))[0] for k in range(2)]
    assert np.allclose((parts[0] + parts[1]) * (w * h / n), whole, rtol=1e-4, atol=1e-5)
    # time scheduler: a budget buys at least the first full pass and the image stays the same image
    ft, st = sc.render("lt", abi.render_params(w, h, max_verts=12, scheduler=abi.SCHED_TIME, render_time=0.05, primary_ray_sampling_mode=abi.PRIMARY_IMAGE))
    assert st["processed"] > 0 and st["elapsed"] >= 0.04
    assert abs(ft.mean() - whole.mean()) < 0.03 * whole.mean()
    rs.close(); sc.close()


@pytest.mark.parametrize("scene", ["spheres", "env", "delta", "envmap", "obj", "mix"])
def test_g3_bdpt_vs_reference(reference, tmp_path, scene):
    """SURVEY.md 8f N4 (the other integrators): renderer::bdpt on the device (lmcu_render_bdpt) against the reference's
    renderer::bdpt at equal sample counts — area light + glass / diffuse (caustic paths through s >= 2 strategies), constant
    environment + glossy / mirror / glass (lights at infinity are not connectable: joint primary-ray densities), point +
    directional lights (degenerated and specular endpoints), an environment map, the OBJ-import material set (textures, alpha
    mask, mixture_wavefrontobj) and the constant-weight mixtures.  G3 gates as for pt / lt."""
    import os
    sb = {"spheres": lambda: scenes.cornell_spheres(32), "env": lambda: scenes.instanced_spheres(grid=2, num_theta=24),
          "delta": lambda: scenes.delta_lights_scene(2, 16),
          "envmap": lambda: scenes.instanced_spheres(grid=2, num_theta=16, env=scenes.env_light(os.path.join(str(tmp_path), "e.hdr"), 64, 32)),
          "obj": lambda: scenes.textured_scene(tmp_path, 12, "obj"), "mix": lambda: scenes.textured_scene(tmp_path, 12, "mix")}[scene]()
    sc, rs = lmcu.Scene(sb), reference.RefScene(sb)
    w, h, B, spp = 128, 72, 4, 32
    n = w * h * spp
    img = dict(max_verts=7, num_samples=n, primary_ray_sampling_mode=abi.PRIMARY_IMAGE)
    R = np.stack([rs.render("renderer::bdpt", abi.render_params(w, h, seed=42 + b, **img))[0] for b in range(B)])
    G = np.stack([sc.render("bdpt", abi.render_params(w, h, sample_offset=b * n, **img))[0] for b in range(B)])
    assert G.mean() > 0.005
    assert_pt_parity(R.mean(axis=0), G.mean(axis=0), B * spp, rel_mean=0.02)
    assert_pt_rmse(R, G, k=1.25)
    # sample-split additivity (counter RNG keyed by the global sample index)
    whole, st = sc.render("bdpt", abi.render_params(w, h, **img))
    assert st["processed"] == n and st["rays"] > 2 * n
    parts = [sc.render("bdpt", abi.render_params(w, h, max_verts=7, num_samples=n // 2, sample_offset=k * (n // 2), rescale=0,
                                                 primary_ray_sampling_mode=abi.PRIMARY_IMAGE))[0] for k in range(2)]
    assert np.allclose((parts[0] + parts[1]) * (w * h / n), whole, rtol=1e-4, atol=1e-5)
    # time scheduler
    ft, st = sc.render("bdpt", abi.render_params(w, h, max_verts=7, scheduler=abi.SCHED_TIME, render_time=0.05, primary_ray_sampling_mode=abi.PRIMARY_IMAGE))
    assert st["processed"] > 0 and st["processed"] % 100000 == 0 and st["elapsed"] >= 0.04
    assert abs(ft.mean() - whole.mean()) < 0.03 * whole.mean()
    rs.close(); sc.close()


def test_bdpt_agrees_with_pt_and_rejects_bad_arguments():
    """Bidirectional and unidirectional estimators of the same integral: image means of bdpt and pt (MIS) agree; path length
    selection by min_verts / max_verts partitions the image; bad arguments fail loudly."""
    sb = scenes.cornell_spheres(24)
    sc = lmcu.Scene(sb)
    w, h, spp = 160, 90, 256
    n = w * h * spp
    img = dict(num_samples=n, primary_ray_sampling_mode=abi.PRIMARY_IMAGE)
    fb, _ = sc.render("bdpt", abi.render_params(w, h, max_verts=8, **img))
    fp, _ = sc.render("pt", abi.render_params(w, h, max_verts=8, **img))
    assert np.abs(fb.mean(axis=(0, 1)) - fp.mean(axis=(0, 1))).max() < 0.01 * fp.mean()
    short, _ = sc.render("bdpt", abi.render_params(w, h, max_verts=4, **img))
    long_, _ = sc.render("bdpt", abi.render_params(w, h, max_verts=8, min_verts=5, **img))
    assert np.allclose(short + long_, fb, rtol=2e-3, atol=2e-4)          # same subpaths, disjoint sets of connections
    with pytest.raises(lmcu.LmcuError, match="max_verts"):
        sc.render("bdpt", abi.render_params(w, h, max_verts=21, **img))
    with pytest.raises(lmcu.LmcuError, match="num_samples"):
        sc.render("bdpt", abi.render_params(w, h, max_verts=8, primary_ray_sampling_mode=abi.PRIMARY_IMAGE))
    with pytest.raises(lmcu.LmcuError, match="min_verts"):
        sc.render("bdpt", abi.render_params(w, h, max_verts=4, min_verts=5, **img))
    sc.close()


@pytest.mark.parametrize("scene", ["spheres", "env", "obj"])
def test_bdpt_kernels_equal_the_one_thread_form(scene, tmp_path):
    """The device runs bdpt as three kernels (subpaths -> strategy-major connections -> any-hit queue); tests/hostsim runs
    the same functions as one loop per sample on the CPU.  Same counter-RNG draws, so the images agree to float rounding."""
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    import sim
    sb = {"spheres": lambda: scenes.cornell_spheres(12), "env": lambda: scenes.instanced_spheres(grid=2, num_theta=10),
          "obj": lambda: scenes.textured_scene(tmp_path, 8, "obj")}[scene]()
    sc, ss = lmcu.Scene(sb), sim.SimScene(sb)
    w, h = 48, 27
    for mv, mn in ((7, 0), (4, 3)):
        p = abi.render_params(w, h, max_verts=mv, min_verts=mn, num_samples=w * h * 8, primary_ray_sampling_mode=abi.PRIMARY_IMAGE)
        fg, st = sc.render("bdpt", p)
        fs = ss.render("bdpt", p)
        assert abs(st["rays"] - ss.rays) <= 1e-3 * ss.rays          # a vertex on a decision boundary may round differently (libm vs CUDA)
        assert np.allclose(fg, fs, rtol=1e-3, atol=1e-4 * float(fs.max()))
    sc.close()


@pytest.mark.parametrize("cap_gb", [0.02, 24.0])
def test_block_cache_cap_and_lru(cap_gb):
    """The library recycles device blocks (path pools, films, builder temporaries) through a per-device cache: never more than
    LMCU_CACHE_GB parked (the blocks parked longest ago make room), the most recent sizes stay available, results do not depend
    on it, lmcu_cache_trim() empties it."""
    import json
    import subprocess
    env = dict(os.environ, LMCU_CACHE_GB=str(cap_gb), PYTHONPATH=os.pathsep.join(sys.path))
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cache_child.py"), str(cap_gb)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    for b, n in out["parked"]:
        assert 0 < b <= out["cap"] and n > 0
    assert out["reuse"][0] == out["reuse"][1]                 # the second build of the last scene found every block parked
    assert out["after_trim"] == [0, 0]
    assert all(0.02 < m < 1.0 for m in out["film_means"])
    if cap_gb < 1:
        assert max(b for b, _ in out["parked"]) > 0.5 * out["cap"]          # the small cap was actually reached
