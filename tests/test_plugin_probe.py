"""Host logic of the Lightmetrica plugin, no GPU needed: plugin/lm_cuda.so is loaded by the UNMODIFIED reference
(lm::comp::load_plugin) and asked what it recovers from an lm::Scene through the public component interfaces —
material / texture parameters are private in the reference, so the plugin probes them (lm_cuda.cpp probe_material,
probe_mixture).  The answers must equal what the scene was built from."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from lmcu import abi, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "lightmetrica-v3_b200", "plugin", "lm_cuda")


@pytest.fixture(scope="module")
def probe(reference):
    if not os.path.exists(PLUGIN + ".so"):
        pytest.skip("plugin not built")
    reference.load_plugin(PLUGIN)
    so = C.CDLL(PLUGIN + ".so")

    def run(sb):
        rs = reference.RefScene(sb)
        buf = C.create_string_buffer(1 << 20)
        n = so.lm_cuda_probe_scene(rs.loc().encode(), buf, len(buf))
        text = buf.value.decode()
        rs.close()
        assert n > 0, text
        return json.loads(text)
    return run


def by_primitive(out, sb):
    """probed material of each primitive, in primitive order"""
    return [out["materials"][i] for i in out["primitive_materials"]]


def test_plain_materials_and_texcoords(probe, tmp_path):
    sb = scenes.instanced_spheres(grid=2, num_theta=8)
    sb.materials[0]["ay"] = 0.3
    out = probe(sb)
    for got, prim in zip(by_primitive(out, sb), sb.primitives):
        want = sb.materials[prim["material"]]
        assert got["type"] == want["type"]
        if want["type"] in (abi.MAT_DIFFUSE, abi.MAT_GLOSSY):
            assert np.allclose(got["color"], want["color"], rtol=1e-9)
        if want["type"] == abi.MAT_GLOSSY:
            assert np.allclose([got["ax"], got["ay"]], [want["ax"], want["ay"]], rtol=1e-6)
        if want["type"] == abi.MAT_GLASS:
            assert abs(got["Ni"] - want["Ni"]) < 1e-9


def test_obj_material_set(probe, tmp_path):
    """mixture_wavefrontobj: Kd / mapKd / Ks / roughness from its underlying lobes, the alpha mask from pdf_component()."""
    sb = scenes.textured_scene(tmp_path, 8, "obj")
    out = probe(sb)
    assert [(t["w"], t["h"], t["c"]) for t in out["textures"]] == [(64, 32, 3), (64, 32, 4)]
    for t, want in zip(out["textures"], sb.textures):
        assert abs(t["sum"] - float(want["data"].astype(np.float64).sum())) < 1e-3 * want["data"].size * 1e-3 + 1e-2
    want_t = sum(float(t.sum()) for t in sb.mesh_t.values())
    assert abs(out["texcoord_sum"] - want_t) < 1e-9 * max(1.0, abs(want_t))
    for got, prim in zip(by_primitive(out, sb), sb.primitives):
        want = sb.materials[prim["material"]]
        assert got["type"] == want["type"]
        tex = want.get("map_kd")
        assert got["map_kd"] == (0 if tex is None else tex + 1)
        if tex is None:
            assert np.allclose(got["color"], want["color"], rtol=1e-9)
        if want["type"] == abi.MAT_MIXTURE_OBJ:
            assert np.allclose(got["Ks"], want["Ks"], rtol=1e-9)
            assert np.allclose([got["ax"], got["ay"]], [want["ax"], want["ay"]], rtol=1e-6)
            has_alpha = tex is not None and sb.textures[tex]["data"].shape[2] == 4
            assert got["alpha_mask"] == (1 if has_alpha else 0)
    # no_alpha_mask=True on an RGBA texture must be seen as "no mask"
    sb2 = scenes.textured_scene(tmp_path, 8, "obj")
    for m in sb2.materials:
        if m["type"] == abi.MAT_MIXTURE_OBJ:
            m["alpha_mask"] = False
    assert all(m["alpha_mask"] == 0 for m in probe(sb2)["materials"])


def test_constant_weight_mixtures(probe, tmp_path):
    """entries through foreach_underlying(), weights from sample_direction() (rr, specular groups) and a least-squares fit of
    eval() (the marginalized non-specular group)."""
    sb = scenes.textured_scene(tmp_path, 8, "mix")
    out = probe(sb)
    mats = out["materials"]
    for got, prim in zip(by_primitive(out, sb), sb.primitives):
        want = sb.materials[prim["material"]]
        assert got["type"] == want["type"]
        if want["type"] not in (abi.MAT_MIXTURE_RR, abi.MAT_MIXTURE_MARGINALIZED):
            continue
        assert len(got["entries"]) == len(want["entries"])
        for (gi, gw), (wi, ww) in zip(got["entries"], want["entries"]):
            assert mats[gi]["type"] == sb.materials[wi]["type"]
            assert abs(gw - ww) < 1e-6, (got["entries"], want["entries"])


def test_film_cuda_behaves_like_film_bitmap(reference, probe, monkeypatch):
    """film::cuda (the plugin's bulk-write film) under the reference's OWN CPU renderers: set_pixel
    (renderer::raycast) and concurrent splat (renderer::pt) must give what film::bitmap gives, through Film::buffer()."""
    sb = scenes.cornell_spheres(8)
    rs = reference.RefScene(sb)
    p = abi.render_params(96, 54)
    fb, _, _ = rs.render("raycast", p)
    pp = abi.render_params(64, 36, spp=64, max_verts=8)
    pb, _, _ = rs.render("pt", pp)
    monkeypatch.setenv("LMREF_FILM", "film::cuda")
    fc, _, _ = rs.render("raycast", p)
    pc, _, nproc = rs.render("pt", pp)
    rs.close()
    assert np.abs(fc - fb).max() <= 1e-7 * max(1.0, np.abs(fb).max())      # fp32 pixels vs fp64 pixels
    assert nproc == 64
    from parity import assert_pt_parity
    assert_pt_parity(pb, pc, 64, rel_mean=0.03)
