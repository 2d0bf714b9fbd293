"""Procedural scene generators produce the geometry BASELINE.json's configs name."""
import numpy as np

from lmcu import scenes


def test_triangle_counts():
    assert scenes.cornell_box().num_triangles() == 36
    assert scenes.cornell_spheres(64).num_triangles() == 32268
    assert 2 * 2236 * 2236 == 9999392
    assert scenes.heightfield_scene(64).num_triangles() == 2 * 64 * 64
    p, n = scenes.uv_sphere(112)
    assert p.shape[0] == 49728


def test_sphere_winding_is_outward_and_normals_unit():
    p, n = scenes.uv_sphere(16)
    gn = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    assert (np.einsum("ij,ij->i", gn, p.mean(1)) > 0).all()
    assert np.allclose(np.linalg.norm(n, axis=-1), 1.0)


def test_cornell_coordinates_are_float_representable():
    sb = scenes.cornell_spheres(8)
    for p, _ in sb.meshes:
        assert (p.astype(np.float32).astype(np.float64) == p).all()


def test_light_faces_down():
    sb = scenes.cornell_box()
    light = [q for q in sb.primitives if q["light"] >= 0]
    assert len(light) == 1
    p = sb.meshes[light[0]["mesh"]][0]
    gn = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    assert (gn[:, 1] < 0).all()


def test_camera_rays_are_unit_and_cover_image():
    sb = scenes.cornell_box()
    r = scenes.camera_rays(sb.camera, 32, 18)
    assert r.shape == (32 * 18,)
    assert np.allclose(np.linalg.norm(r["d"], axis=1), 1.0)
    assert (r["d"][:, 2] < 0).all()
