"""Probe: how much does ray ORDER matter to the queue traversal kernel?  C2 diffuse-bounce rays in (a) pixel order
(what a wavefront queue roughly holds after compaction), (b) random order, (c) sorted by origin cell + direction octant."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v3_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402
import lmcu  # noqa: E402
from lmcu import scenes  # noqa: E402
from trav_bench import device_batch  # noqa: E402


def morton3(q, bits):
    out = np.zeros(len(q), dtype=np.uint64)
    for b in range(bits):
        for a in range(3):
            out |= ((q[:, a].astype(np.uint64) >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + a)
    return out


def main():
    sb = scenes.cornell_spheres(64)
    sc = lmcu.Scene(sb)
    cam = scenes.camera_rays(sb.camera, 1920, 1080)
    h = sc.intersect(cam)
    # geometric normals per hit from the triangle soup
    gn = []
    for p, _ in sb.meshes:
        n = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
        gn.append(n / np.linalg.norm(n, axis=1, keepdims=True))
    prim_mesh = np.array([q["mesh"] for q in sb.primitives])

    def normals(hh):
        out = np.zeros((len(hh), 3))
        for m in range(len(gn)):
            sel = prim_mesh[hh["primitive"]] == m
            out[sel] = gn[m][hh["face"][sel]]
        return out

    rays = scenes.bounce_rays(cam, h, normals)
    # second bounce too (more incoherent), concatenated
    h2 = sc.intersect(rays)
    rays2 = scenes.bounce_rays(rays, h2, normals, seed=11)
    rays = np.concatenate([rays, rays2])
    print("bounce rays:", len(rays))
    rng = np.random.RandomState(0)
    orders = {"pixel order": np.arange(len(rays)), "random": rng.permutation(len(rays))}
    for cell_bits in (3, 5):
        lo, hi = np.array([-1, 0, -1.]), np.array([1, 2, 1.])
        q = np.clip(((rays["o"] - lo) / (hi - lo) * (1 << cell_bits)).astype(np.int64), 0, (1 << cell_bits) - 1)
        octant = ((rays["d"][:, 0] < 0) * 4 + (rays["d"][:, 1] < 0) * 2 + (rays["d"][:, 2] < 0)).astype(np.uint64)
        key_oct_first = (octant << np.uint64(3 * cell_bits)) | morton3(q, cell_bits)
        key_cell_first = (morton3(q, cell_bits) << np.uint64(3)) | octant
        orders["octant then %d-bit cell" % cell_bits] = np.argsort(key_oct_first, kind="stable")
        orders["%d-bit cell then octant" % cell_bits] = np.argsort(key_cell_first, kind="stable")
    for name, o in orders.items():
        mr, _ = device_batch(sc, rays[o])
        print("  %-28s %8.1f Mrays/s" % (name, mr), flush=True)


if __name__ == "__main__":
    main()
