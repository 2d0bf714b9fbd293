"""Golden fixtures of the OBJ-import material set from the UNMODIFIED reference (oracle/_ref): textured
diffuse, material::mixture_wavefrontobj with alpha mask, material::mask, the constant-weight mixtures.  Kept in their own
file so that reference_outputs.npz stays byte-identical (the reference's multi-threaded pt is not reproducible run to run).

    python tests/golden/make_golden_materials.py
"""
import os
import pickle
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lightmetrica-v3_b200")]
from lmcu import abi, scenes  # noqa: E402
from oracle import ref  # noqa: E402

KINDS = ("obj", "mix")
RAYCAST = dict(width=192, height=108)
PT = dict(width=96, height=54, spp=1024, max_verts=12)
PT1 = dict(width=40, height=24, spp=4, max_verts=12, seed=13)


def main():
    ref.init(0)
    tmp = tempfile.mkdtemp()
    out = {}
    for kind in KINDS:
        sb = scenes.textured_scene(tmp, 12, kind)
        rs = ref.RefScene(sb)
        f, _, _ = rs.render("raycast", abi.render_params(**RAYCAST))
        out[kind + "_raycast"] = f.astype(np.float32)
        f, _, _ = rs.render("pt", abi.render_params(**PT))
        out[kind + "_pt_mis"] = f.astype(np.float32)
        rs.close()
        code = ("import sys,pickle; sys.path[:0]=[%r,%r]\n"
                "from oracle import ref; from lmcu import abi, scenes\n"
                "ref.init(1); sb=scenes.textured_scene(%r, 12, %r); rs=ref.RefScene(sb)\n"
                "f,_,_=rs.render('pt', abi.render_params(**%r))\n"
                "sys.stdout.buffer.write(pickle.dumps(f))\n") % (ROOT, os.path.join(ROOT, "lightmetrica-v3_b200"), tmp, kind, PT1)
        out[kind + "_pt_1thread_seed13"] = pickle.loads(subprocess.check_output([sys.executable, "-c", code]))
    np.savez_compressed(os.path.join(HERE, "reference_outputs_materials.npz"), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
