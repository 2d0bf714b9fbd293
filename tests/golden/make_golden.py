"""Generates the golden fixtures from the UNMODIFIED reference (oracle/_ref): run in the build
container after `make -C oracle ref`.  Inputs are regenerated from seeds by lmcu.scenes, so only
the reference's outputs are stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lightmetrica-v3_b200")]
from lmcu import abi, scenes  # noqa: E402
from oracle import ref  # noqa: E402

CASES = {
    "cornell": lambda: scenes.cornell_box(),
    "cornell_fp64": lambda: scenes.cornell_box(True),
    "spheres12": lambda: scenes.cornell_spheres(12),
    "heightfield48": lambda: scenes.heightfield_scene(48),
}


def golden_rays(name, sb):
    lo, hi = (np.array([-1, 0, -1.]), np.array([1, 2, 1.])) if "height" not in name else (np.array([-1, -.2, -1.]), np.array([1, .6, 1.]))
    return np.concatenate([scenes.camera_rays(sb.camera, 192, 108), scenes.random_rays(6000, lo, hi, seed=11)])


def main():
    ref.init(0)
    out = {}
    for name, make in CASES.items():
        sb = make()
        rs = ref.RefScene(sb)
        h = rs.intersect(golden_rays(name, sb))
        out[name + "_t"], out[name + "_u"], out[name + "_v"] = h["t"], h["u"], h["v"]
        out[name + "_prim"], out[name + "_face"] = h["primitive"].astype(np.int16), h["face"].astype(np.int32)
        if name in ("cornell", "spheres12"):
            f, _, _ = rs.render("raycast", abi.render_params(192, 108))
            out[name + "_raycast"] = f.astype(np.float32)
        if name == "spheres12":
            for mode, tag in ((abi.SAMPLING_MIS, "mis"), (abi.SAMPLING_NAIVE, "naive")):
                f, _, _ = rs.render("pt", abi.render_params(96, 54, spp=1024, max_verts=12, sampling_mode=mode))
                out["spheres12_pt_" + tag] = f.astype(np.float32)
            f, _, _ = rs.render("ao", abi.render_params(96, 54, spp=256))
            out["spheres12_ao"] = f.astype(np.float32)
        rs.close()
    # single-threaded reference pt: deterministic mt19937 stream (seed + thread 0); run in a fresh process because the
    # reference's per-thread Rng is a thread_local that survives across render() calls (renderer_pt.cpp:91)
    import pickle
    import subprocess
    code = ("import sys,pickle; sys.path[:0]=[%r,%r]\n"
            "from oracle import ref; from lmcu import abi, scenes\n"
            "ref.init(1); sb=scenes.cornell_spheres(8); rs=ref.RefScene(sb)\n"
            "f,_,_=rs.render('pt', abi.render_params(48,27,spp=4,max_verts=10,seed=42))\n"
            "sys.stdout.buffer.write(pickle.dumps(f))\n") % (ROOT, os.path.join(ROOT, "lightmetrica-v3_b200"))
    out["spheres8_pt_1thread_seed42"] = pickle.loads(subprocess.check_output([sys.executable, "-c", code]))
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
