"""Shared parity checks (SURVEY.md §8c gates G1-G3)."""
import numpy as np


def rmse(a, b):
    return float(np.sqrt(np.mean((np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)) ** 2)))


def classify_hits(ref_hits, hits):
    """G1 bookkeeping.  Returns dict(mismatch, ties, ref_boundary_miss, other).

    tie                 both hit, t equal to 1e-12 relative (bit-equal in all but coplanar-overlap cases), different
                        (primitive, face): coincident / shared-edge geometry.  The reference keeps the LAST triangle of
                        its own traversal order (accel_sahbvh.cpp:286-290), and for coplanar overlapping triangles whose
                        t differ by an ulp its box test (two roundings) can cull the nearer one
    ref_boundary_miss   reference misses, we hit exactly ON a triangle edge (a barycentric is 0 or u+v is 1 to
                        1e-12): the reference's box test (math.h:143-158) rejects a hit lying on a slab boundary
                        by one ulp although its own triangle test accepts it
    other               anything else — must be zero
    """
    bad = (ref_hits["primitive"] != hits["primitive"]) | (ref_hits["face"] != hits["face"])
    ties = bad & (ref_hits["primitive"] >= 0) & (hits["primitive"] >= 0) & \
        (np.abs(ref_hits["t"] - hits["t"]) <= 1e-12 * np.maximum(1.0, np.abs(ref_hits["t"])))
    edge = np.minimum(np.minimum(hits["u"], hits["v"]), np.abs(1.0 - hits["u"] - hits["v"])) <= 1e-12
    ref_miss = bad & (ref_hits["primitive"] < 0) & (hits["primitive"] >= 0) & edge
    other = bad & ~ties & ~ref_miss
    return {"mismatch": int(bad.sum()), "ties": int(ties.sum()), "ref_boundary_miss": int(ref_miss.sum()),
            "other": int(other.sum()), "bad_mask": bad}


def assert_hit_parity(ref_hits, hits, max_special=None):
    """G1.  `max_special`: how many special rays (ties + reference boundary misses) the batch may hold.  Every fixture
    states the count MEASURED for it (the batches are deterministic: the numbers were taken with the CPU build of the
    traversal, tests/hostsim, against the reference); without one the allowance is 2e-4 of the batch — C1's 1920x1080
    symmetric camera has 167 of 2 073 600 = 8e-5.  Anything that is neither kind fails whatever the allowance."""
    c = classify_hits(ref_hits, hits)
    assert c["other"] == 0, "non-tie closest-hit mismatches: %d" % c["other"]
    allowed = int(2e-4 * len(hits)) if max_special is None else max_special
    assert c["ties"] + c["ref_boundary_miss"] <= allowed, {k: v for k, v in c.items() if k != "bad_mask"}
    same = ~c["bad_mask"] & (ref_hits["primitive"] >= 0)
    # FP64 leaf test with the reference's operation order: t, u, v are bit-identical, far inside the 1e-5 gate
    assert (ref_hits["t"][same] == hits["t"][same]).all()
    assert (ref_hits["u"][same] == hits["u"][same]).all() and (ref_hits["v"][same] == hits["v"][same]).all()
    miss = (ref_hits["primitive"] < 0) & (hits["primitive"] < 0)
    assert miss.sum() + same.sum() + c["mismatch"] == len(hits)
    return c


def assert_pt_parity(ref_film, film, spp, rel_mean=0.01, z_frac=2e-3, block=8):
    """G3: statistical image parity at equal spp (the RNG streams differ by design).

    - difference of the image means per channel <= max(rel_mean * mean, 4.5 standard errors of that difference):
      rel_mean is the stated bias tolerance; the standard error (per-pixel difference std / sqrt(#pixels), both
      images being independent estimates) keeps the gate meaningful for high-variance settings such as naive
      sampling at low spp, where Monte-Carlo noise alone exceeds rel_mean (the reference is itself
      non-deterministic run to run: per-thread RNG + dynamic scheduling, renderer_pt.cpp:91)
    - 8x8 block means of the difference image, z-scored by the block's own noise: the fraction of |z| > 4 stays
      <= z_frac (both images are unbiased estimates of the same radiance, so the difference is zero-mean noise;
      a shading or sampling error shows up as block means far from zero)
    """
    ref_film = np.asarray(ref_film, dtype=np.float64)
    film = np.asarray(film, dtype=np.float64)
    mr, mg = ref_film.mean(axis=(0, 1)), film.mean(axis=(0, 1))
    rel = np.abs(mg - mr) / np.maximum(mr, 1e-12)
    se = (film - ref_film).reshape(-1, film.shape[2]).std(axis=0) / np.sqrt(film.shape[0] * film.shape[1])
    tol = np.maximum(rel_mean * np.maximum(mr, 1e-12), 4.5 * se)
    assert (np.abs(mg - mr) <= tol).all(), "image means differ: ref %s ours %s (tolerance %s)" % (mr, mg, tol)
    d = (film - ref_film).mean(axis=2)
    h, w = d.shape
    hb, wb = h // block, w // block
    blocks = d[:hb * block, :wb * block].reshape(hb, block, wb, block)
    bm = blocks.mean(axis=(1, 3))
    bs = blocks.std(axis=(1, 3)) / np.sqrt(block * block) + 1e-9
    z = bm / bs
    frac = float((np.abs(z) > 4).mean())
    assert frac <= max(z_frac, 2.0 / z.size), "structured difference: %.4f of %dx%d blocks have |z|>4" % (frac, block, block)
    return {"rel_mean": rel, "z_frac": frac}


def assert_pt_rmse(ref_batches, our_batches, k=1.25):
    """G3, RMSE gate: both renderers deliver B independent equal-spp estimates of the same image (different
    seeds / disjoint sample blocks).  The per-channel RMSE between the two batch MEANS must not exceed k times what
    Monte-Carlo noise alone predicts,

        rmse  <=  k * sqrt( mean over pixels of (var_ref + var_ours) / B ),

    where var_* is the per-pixel sample variance ACROSS the batches — so the tolerance is measured on the spot for every
    scene and sampling mode, not assumed.  E[rmse^2] equals the right-hand side without k when both estimators are unbiased;
    any bias adds its mean square on top.  k = 1.25 covers the sampling error of the variance estimate itself
    (ref-vs-ref runs with different seeds land at 0.97 .. 1.04 on these scenes; tests/test_hostsim.py calibrates it).
    Returns rmse / predicted per channel."""
    ref_batches = np.asarray(ref_batches, dtype=np.float64)
    our_batches = np.asarray(our_batches, dtype=np.float64)
    B = ref_batches.shape[0]
    assert B >= 4 and our_batches.shape == ref_batches.shape
    mr, mo = ref_batches.mean(axis=0), our_batches.mean(axis=0)
    noise2 = (ref_batches.var(axis=0, ddof=1) + our_batches.var(axis=0, ddof=1)) / B
    got = np.sqrt(((mr - mo) ** 2).mean(axis=(0, 1)))
    want = np.sqrt(noise2.mean(axis=(0, 1)))
    ratio = got / np.maximum(want, 1e-30)
    assert (ratio <= k).all(), "per-channel RMSE %s exceeds %.2f x the Monte-Carlo prediction %s" % (got, k, want)
    return ratio


def render_batches(render, batches, spp):
    """B estimates of spp samples each: render(batch_index) -> image."""
    return np.stack([np.asarray(render(b), dtype=np.float64) for b in range(batches)])
