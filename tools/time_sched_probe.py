"""Probe: why does a long time-budget render differ from the reference more than equal-spp renders do?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lightmetrica-v3_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import lmcu
from lmcu import abi, scenes
from oracle import ref


def zfrac(a, b):
    d = (np.asarray(a, np.float64) - np.asarray(b, np.float64)).mean(axis=2)
    h, w = d.shape
    bl = d[:h // 8 * 8, :w // 8 * 8].reshape(h // 8, 8, w // 8, 8)
    z = bl.mean(axis=(1, 3)) / (bl.std(axis=(1, 3)) / 8 + 1e-9)
    return float((np.abs(z) > 4).mean())


ref.init(0)
sb = scenes.cornell_spheres(16)
sc, rs = lmcu.Scene(sb), ref.RefScene(sb)
w, h = 160, 90
fr, _, _ = rs.render("pt", abi.render_params(w, h, spp=256, max_verts=12))
fr2, _, _ = rs.render("pt", abi.render_params(w, h, spp=2048, max_verts=12, seed=7))
out = {}
for spp in (256, 2048, 8192, 32768):
    out[spp], _ = sc.render("pt", abi.render_params(w, h, spp=spp, max_verts=12))
ft, st = sc.render("pt", abi.render_params(w, h, max_verts=12, scheduler=abi.SCHED_TIME, render_time=0.4))
print("time processed", st["processed"])
print("ref256 vs ref2048: z", zfrac(fr, fr2), "means", fr.mean(), fr2.mean())
for spp, f in out.items():
    print("spp", spp, "mean", f.mean(), "z vs ref256", zfrac(fr, f), "z vs ref2048", zfrac(fr2, f))
print("time mean", ft.mean(), "z vs ref256", zfrac(fr, ft), "z vs ref2048", zfrac(fr2, ft), "z vs ours32768", zfrac(out[32768], ft))
