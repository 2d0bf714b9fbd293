"""Procedural scenes of BASELINE.json's configs as plain arrays.

A SceneBuilder mirrors the calls a Lightmetrica user makes (load mesh/material/light, add a
(transformed) primitive: /root/reference/include/lm/scene.h:120-140) and produces the
LmcuSceneDesc that liblmcu.so, the reference harness and the C port all consume.  All coordinates
are snapped to float32-representable doubles unless stated, so device triangles can be stored as
float without changing a single bit of the FP64 geometry the reference sees.
"""
import ctypes as C
import numpy as np

from . import abi


def _f32(a):
    return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)


def translate(v):
    M = np.eye(4)
    M[:3, 3] = v
    return M


def scale(v):
    return np.diag([v[0], v[1], v[2], 1.0])


def rotate(angle_deg, axis):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    t = np.radians(angle_deg)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) + np.sin(t) * K + (1 - np.cos(t)) * (K @ K)
    M = np.eye(4)
    M[:3, :3] = R
    return M


class SceneBuilder:
    def __init__(self):
        self.meshes = []        # (p[nf,3,3], n[nf,3,3] or None)
        self.materials = []     # dict
        self.lights = []        # dict
        self.primitives = []    # dict
        self.camera = None
        self.textures = []      # dict(data=float32[h,w,c] in texture::bitmap row order, path=str or None)
        self.mesh_t = {}        # mesh index -> t[nf,3,2] texture coordinates
        self._keep = []

    # -- assets -------------------------------------------------------------------------------
    def add_mesh(self, p, n=None, t=None):
        p = np.ascontiguousarray(p, dtype=np.float64).reshape(-1, 3, 3)
        if n is not None:
            n = np.ascontiguousarray(n, dtype=np.float64).reshape(-1, 3, 3)
            assert n.shape == p.shape
        self.meshes.append((p, n))
        if t is not None:
            t = np.ascontiguousarray(t, dtype=np.float64).reshape(-1, 3, 2)
            assert t.shape[0] == p.shape[0]
            self.mesh_t[len(self.meshes) - 1] = t
        return len(self.meshes) - 1

    def add_texture(self, data, path=None):
        """texture::bitmap: data = float32 [h,w,3|4] in the row order the reference holds after loading `path`."""
        data = np.ascontiguousarray(data, dtype=np.float32)
        assert data.ndim == 3 and data.shape[2] in (3, 4)
        self.textures.append(dict(data=data, path=path))
        return len(self.textures) - 1

    def add_material(self, type, color=(1, 1, 1), ax=0.0, ay=0.0, Ni=1.0, **extra):
        self.materials.append(dict(type=type, color=tuple(color), ax=ax, ay=ay, Ni=Ni, **extra))
        return len(self.materials) - 1

    def diffuse(self, Kd, map_kd=None):
        return self.add_material(abi.MAT_DIFFUSE, Kd, map_kd=map_kd)

    def mask(self):
        return self.add_material(abi.MAT_MASK, (0, 0, 0))

    def mixture_obj(self, Kd, Ks, ax, ay, map_kd=None, alpha_mask=True):
        """material::mixture_wavefrontobj: what the OBJ loader creates per .mtl entry (model_wavefrontobj.cpp:393-413)."""
        return self.add_material(abi.MAT_MIXTURE_OBJ, Kd, ax, ay, Ks=tuple(Ks), map_kd=map_kd, alpha_mask=alpha_mask)

    def mixture_rr(self, entries):
        """material::constant_weight_mixture_rr; entries = [(material index, weight), ...]"""
        return self.add_material(abi.MAT_MIXTURE_RR, (0, 0, 0), entries=list(entries))

    def mixture_marginalized(self, entries):
        """material::constant_weight_mixture_marginalized; entries = [(material index, weight), ...]"""
        return self.add_material(abi.MAT_MIXTURE_MARGINALIZED, (0, 0, 0), entries=list(entries))

    def glossy(self, Ks, ax, ay):
        return self.add_material(abi.MAT_GLOSSY, Ks, ax, ay)

    def glass(self, Ni):
        return self.add_material(abi.MAT_GLASS, (0, 0, 0), Ni=Ni)

    def mirror(self):
        return self.add_material(abi.MAT_MIRROR, (0, 0, 0))

    # -- scene graph (flattened: one entry per primitive node with a mesh) -----------------------
    def add_primitive(self, mesh, material, Ke=None, M=None):
        """Ke != None attaches a light::area to the primitive (light_area.cpp:68-73)."""
        light = -1
        if Ke is not None:
            assert all(l["type"] == abi.LIGHT_AREA for l in self.lights), "area lights first: the reference harness adds the others after the meshes"
            self.lights.append(dict(type=abi.LIGHT_AREA, primitive=len(self.primitives), Ke=tuple(Ke)))
            light = len(self.lights) - 1
        M = np.eye(4) if M is None else np.asarray(M, dtype=np.float64)
        self.primitives.append(dict(node_index=len(self.primitives), mesh=mesh, material=material,
                                    light=light, M=M))
        return len(self.primitives) - 1

    def set_envconst(self, Le):
        self.lights.append(dict(type=abi.LIGHT_ENVCONST, primitive=-1, Ke=tuple(Le)))

    def set_env(self, rgb, rot=0.0, scale=1.0, path=None):
        rgb = np.ascontiguousarray(rgb, dtype=np.float32)
        self.lights.append(dict(type=abi.LIGHT_ENV, primitive=-1, Ke=(0, 0, 0), rgb=rgb, rot=rot,
                                scale=scale, path=path))

    def add_point_light(self, Le, position):
        self.lights.append(dict(type=abi.LIGHT_POINT, primitive=-1, Ke=tuple(Le), position=tuple(position)))

    def add_directional_light(self, Le, direction):
        self.lights.append(dict(type=abi.LIGHT_DIRECTIONAL, primitive=-1, Ke=tuple(Le), position=tuple(direction)))

    def set_camera(self, position, center, up, vfov, aspect):
        self.camera = dict(position=tuple(position), center=tuple(center), up=tuple(up),
                           vfov=float(vfov), aspect=float(aspect))

    # -- statistics ---------------------------------------------------------------------------
    def num_triangles(self):
        return sum(self.meshes[p["mesh"]][0].shape[0] for p in self.primitives)

    # -- ctypes view --------------------------------------------------------------------------
    def desc(self):
        """Returns an LmcuSceneDesc whose pointers stay valid while this builder is alive."""
        dp = C.POINTER(C.c_double)
        meshes = (abi.LmcuMesh * max(1, len(self.meshes)))()
        for i, (p, n) in enumerate(self.meshes):
            meshes[i].num_faces = p.shape[0]
            meshes[i].p = p.ctypes.data_as(dp)
            meshes[i].n = n.ctypes.data_as(dp) if n is not None else None
            meshes[i].t = self.mesh_t[i].ctypes.data_as(dp) if i in self.mesh_t else None
        prims = (abi.LmcuPrimitive * max(1, len(self.primitives)))()
        for i, q in enumerate(self.primitives):
            prims[i].node_index, prims[i].mesh = q["node_index"], q["mesh"]
            prims[i].material, prims[i].light = q["material"], q["light"]
            prims[i].M = (C.c_double * 16)(*np.asarray(q["M"]).T.reshape(-1))   # column-major
        mats = (abi.LmcuMaterial * max(1, len(self.materials)))()
        for i, m in enumerate(self.materials):
            mats[i].type, mats[i].color = m["type"], abi.c_double3(*m["color"])
            mats[i].ax, mats[i].ay, mats[i].Ni = m["ax"], m["ay"], m["Ni"]
            mats[i].map_kd = 0 if m.get("map_kd") is None else m["map_kd"] + 1
            mats[i].Ks = abi.c_double3(*m.get("Ks", (0, 0, 0)))
            mats[i].alpha_mask = 1 if m.get("alpha_mask") else 0
            if m.get("entries"):
                em = (C.c_int32 * len(m["entries"]))(*[e[0] for e in m["entries"]])
                ew = (C.c_double * len(m["entries"]))(*[float(e[1]) for e in m["entries"]])
                mats[i].num_entries, mats[i].entry_material, mats[i].entry_weight = len(m["entries"]), em, ew
                self._keep.append((em, ew))
        texs = (abi.LmcuTexture * max(1, len(self.textures)))()
        for i, t in enumerate(self.textures):
            texs[i].h, texs[i].w, texs[i].c = t["data"].shape
            texs[i].data = t["data"].ctypes.data_as(C.POINTER(C.c_float))
            texs[i].path = t["path"].encode() if t["path"] else None
        lights = (abi.LmcuLight * max(1, len(self.lights)))()
        for i, l in enumerate(self.lights):
            lights[i].type, lights[i].primitive = l["type"], l["primitive"]
            lights[i].node_index = l["primitive"] if l["primitive"] >= 0 else len(self.primitives)
            lights[i].Ke = abi.c_double3(*l["Ke"])
            lights[i].position = abi.c_double3(*l.get("position", (0, 0, 0)))
            if l["type"] == abi.LIGHT_ENV:
                lights[i].env_h, lights[i].env_w = l["rgb"].shape[:2]
                lights[i].env_rgb = l["rgb"].ctypes.data_as(C.POINTER(C.c_float))
                lights[i].env_rot, lights[i].env_scale = l["rot"], l["scale"]
                lights[i].env_path = l["path"].encode() if l["path"] else None
        d = abi.LmcuSceneDesc()
        d.num_meshes, d.num_primitives = len(self.meshes), len(self.primitives)
        d.num_materials, d.num_lights = len(self.materials), len(self.lights)
        d.meshes, d.primitives, d.materials, d.lights = meshes, prims, mats, lights
        d.num_textures, d.textures = len(self.textures), texs
        if self.camera is not None:
            cam = abi.LmcuCamera()
            cam.position = abi.c_double3(*self.camera["position"])
            cam.center = abi.c_double3(*self.camera["center"])
            cam.up = abi.c_double3(*self.camera["up"])
            cam.vfov, cam.aspect = self.camera["vfov"], self.camera["aspect"]
            d.camera = C.pointer(cam)
            self._keep.append(cam)
        self._keep.append((meshes, prims, mats, lights, texs))
        return d


# ------------------------------------------------------------------------------------------------
# Geometry generators
# ------------------------------------------------------------------------------------------------
def quad(a, b, c, d):
    """Two CCW triangles (a,b,c), (a,c,d); the normal is (b-a)x(c-a)."""
    a, b, c, d = (np.asarray(v, dtype=np.float64) for v in (a, b, c, d))
    return np.stack([np.stack([a, b, c]), np.stack([a, c, d])])


def flat_normals(p):
    p = p.reshape(-1, 3, 3)
    gn = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    gn /= np.linalg.norm(gn, axis=1, keepdims=True)
    return np.repeat(gn[:, None, :], 3, axis=1)


def box(lo, hi):
    """Axis-aligned box, 12 outward-facing triangles."""
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    q = [
        quad((x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)),   # +z
        quad((x1, y0, z0), (x0, y0, z0), (x0, y1, z0), (x1, y1, z0)),   # -z
        quad((x1, y0, z1), (x1, y0, z0), (x1, y1, z0), (x1, y1, z1)),   # +x
        quad((x0, y0, z0), (x0, y0, z1), (x0, y1, z1), (x0, y1, z0)),   # -x
        quad((x0, y1, z1), (x1, y1, z1), (x1, y1, z0), (x0, y1, z0)),   # +y
        quad((x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1)),   # -y
    ]
    return np.concatenate(q)


def uv_sphere(num_theta, num_phi=None, r=1.0):
    """The UV sphere of /root/reference/functest/func_render_instancing.py:60-92 (same vertex
    formula and face order), returned as soup positions + smooth normals, outward winding."""
    num_phi = 2 * num_theta if num_phi is None else num_phi
    i, j = np.meshgrid(np.arange(num_theta + 1), np.arange(num_phi), indexing="ij")
    theta = i * np.pi / num_theta
    phi = j * 2 * np.pi / num_phi
    ns = np.stack([np.sin(theta) * np.sin(phi), np.cos(theta), np.sin(theta) * np.cos(phi)], -1).reshape(-1, 3)
    vs = r * ns
    fs = []
    for ii in range(1, num_theta + 1):
        jj = np.arange(1, num_phi + 1)
        p00 = (ii - 1) * num_phi + jj - 1
        p01 = (ii - 1) * num_phi + jj % num_phi
        p10 = ii * num_phi + jj - 1
        p11 = ii * num_phi + jj % num_phi
        rows = []
        if ii > 1:
            rows.append(np.stack([p10, p01, p00], -1))
        if ii < num_theta:
            rows.append(np.stack([p11, p01, p10], -1))
        # interleave exactly like the reference loop (per j: first the i>1 face, then the i<n face)
        fs.append(np.stack(rows, 1).reshape(-1, 3))
    fs = np.concatenate(fs)
    p = vs[fs]
    n = ns[fs]
    # make the winding outward (geometric normal on the same side as the shading normals)
    gn = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    flip = np.einsum("ij,ij->i", gn, n.mean(1)) < 0
    p[flip] = p[flip][:, [0, 2, 1]]
    n[flip] = n[flip][:, [0, 2, 1]]
    return p, n


def _fbm(x, z, octaves=6, seed=1):
    """Value-noise fBm on a lattice hashed from (seed, octave): deterministic, numpy only."""
    rng = np.random.RandomState(seed)
    y = np.zeros_like(x)
    amp, freq = 1.0, 2.0
    for _ in range(octaves):
        n = int(freq) + 2
        lattice = rng.rand(n, n)
        fx, fz = (x * 0.5 + 0.5) * freq, (z * 0.5 + 0.5) * freq
        ix, iz = np.clip(fx.astype(np.int64), 0, n - 2), np.clip(fz.astype(np.int64), 0, n - 2)
        tx, tz = fx - ix, fz - iz
        tx, tz = tx * tx * (3 - 2 * tx), tz * tz * (3 - 2 * tz)
        v = (lattice[ix, iz] * (1 - tx) * (1 - tz) + lattice[ix + 1, iz] * tx * (1 - tz)
             + lattice[ix, iz + 1] * (1 - tx) * tz + lattice[ix + 1, iz + 1] * tx * tz)
        y += amp * (v - 0.5)
        amp *= 0.5
        freq *= 2.0
    return y


def heightfield(n, height=0.15, seed=1):
    """(n+1)^2 grid over [-1,1]^2 -> 2 n^2 triangles (C3: n=2236 -> 9 999 392)."""
    g = np.linspace(-1.0, 1.0, n + 1)
    x, z = np.meshgrid(g, g, indexing="ij")
    y = height * _fbm(x, z, 6, seed)
    v = _f32(np.stack([x, y, z], -1))
    a, b, c, d = v[:-1, :-1], v[1:, :-1], v[1:, 1:], v[:-1, 1:]
    t1 = np.stack([a, c, b], -2)      # normal +y
    t2 = np.stack([a, d, c], -2)
    return np.stack([t1, t2], 2).reshape(-1, 3, 3)


# ------------------------------------------------------------------------------------------------
# Configs
# ------------------------------------------------------------------------------------------------
CORNELL_CAMERA = dict(position=(0, 1, 5), center=(0, 1, 0), up=(0, 1, 0), vfov=30, aspect=16 / 9)


def _cornell_shell(sb):
    white = sb.diffuse((.725, .725, .725))
    red = sb.diffuse((.63, .065, .05))
    green = sb.diffuse((.14, .45, .091))
    black = sb.diffuse((0, 0, 0))
    floor = quad((-1, 0, 1), (1, 0, 1), (1, 0, -1), (-1, 0, -1))         # +y
    ceil_ = quad((-1, 2, -1), (1, 2, -1), (1, 2, 1), (-1, 2, 1))         # -y
    back = quad((-1, 0, -1), (1, 0, -1), (1, 2, -1), (-1, 2, -1))        # +z
    left = quad((-1, 0, 1), (-1, 0, -1), (-1, 2, -1), (-1, 2, 1))        # +x
    right = quad((1, 0, -1), (1, 0, 1), (1, 2, 1), (1, 2, -1))           # -x
    lq = quad((-.25, 1.9921875, -.25), (.25, 1.9921875, -.25), (.25, 1.9921875, .25), (-.25, 1.9921875, .25))  # -y
    for geo, mat in ((np.concatenate([floor, ceil_, back]), white), (left, red), (right, green)):
        geo = _f32(geo)
        sb.add_primitive(sb.add_mesh(geo, flat_normals(geo)), mat)
    lq = _f32(lq)
    sb.add_primitive(sb.add_mesh(lq, flat_normals(lq)), black, Ke=(17, 12, 4))
    sb.set_camera(**CORNELL_CAMERA)
    return white


def cornell_box(transformed_boxes=False):
    """C1: 18 quads = 36 triangles in [-1,1]x[0,2]x[-1,1]."""
    sb = SceneBuilder()
    white = _cornell_shell(sb)
    tall = translate((-.35, 0, -.3)) @ rotate(17, (0, 1, 0)) @ scale((.3, .6, .3)) @ translate((0, 1, 0))
    short = translate((.35, 0, .3)) @ rotate(-17, (0, 1, 0)) @ scale((.3, .3, .3)) @ translate((0, 1, 0))
    unit = box((-1, -1, -1), (1, 1, 1))
    for M in (tall, short):
        if transformed_boxes:
            sb.add_primitive(sb.add_mesh(unit, flat_normals(unit)), white, M=M)
        else:
            w = _f32(unit.reshape(-1, 3) @ M[:3, :3].T + M[:3, 3]).reshape(-1, 3, 3)
            sb.add_primitive(sb.add_mesh(w, flat_normals(w)), white)
    return sb


def cornell_spheres(num_theta=64):
    """C2/C5: Cornell shell + diffuse and glass UV spheres (2 x 16 128 tris at num_theta=64)."""
    sb = SceneBuilder()
    _cornell_shell(sb)
    p, n = uv_sphere(num_theta, 2 * num_theta, 1.0)
    for centre, mat in (((-.4, .35, -.3), sb.diffuse((.8, .8, .8))), ((.45, .35, .3), sb.glass(1.5))):
        w = _f32(p * .35 + np.asarray(centre))
        sb.add_primitive(sb.add_mesh(w, n), mat)
    return sb


def heightfield_scene(n=2236):
    """C3: n x n heightfield, single diffuse material; camera of the coherent batch."""
    sb = SceneBuilder()
    geo = heightfield(n)
    sb.add_primitive(sb.add_mesh(geo, flat_normals(geo)), sb.diffuse((.7, .7, .7)))
    sb.set_camera(position=(0, 1.2, 2.2), center=(0, 0, 0), up=(0, 1, 0), vfov=40, aspect=1.0)
    return sb


def instanced_spheres(grid=10, num_theta=112, seed=3, env=None):
    """C4: grid^3 transformed instances of one UV sphere (num_theta=112 -> 49 728 tris each),
    materials cycling glossy/glass/mirror/diffuse, ground quad, constant (or image) env light."""
    sb = SceneBuilder()
    p, n = uv_sphere(num_theta, 2 * num_theta, 0.4)
    mesh = sb.add_mesh(_f32(p), n)
    mats = [sb.glossy((.8, .6, .3), .15, .15), sb.glass(1.5), sb.mirror(), sb.diffuse((.7, .7, .7))]
    rng = np.random.RandomState(seed)
    k = 0
    half = (grid - 1) / 2
    for ix in range(grid):
        for iy in range(grid):
            for iz in range(grid):
                axis = rng.randn(3)
                M = (translate((ix - half, iy + .5, iz - half)) @ rotate(rng.rand() * 360, axis)
                     @ scale([.7 + .3 * rng.rand()] * 3))
                sb.add_primitive(mesh, mats[k % 4], M=M)
                k += 1
    g = grid
    ground = _f32(quad((-g, 0, g), (g, 0, g), (g, 0, -g), (-g, 0, -g)))
    sb.add_primitive(sb.add_mesh(ground, flat_normals(ground)), sb.diffuse((.5, .5, .5)))
    if env is None:
        sb.set_envconst((1, 1, 1))
    else:
        sb.set_env(**env)
    sb.set_camera(position=(grid * 1.1, grid * .9, grid * 1.6), center=(0, grid * .45, 0), up=(0, 1, 0),
                  vfov=35, aspect=16 / 9)
    return sb


def affine_instances(seed=0, n_inst=40, num_theta=10):
    """Two meshes (a UV sphere, a box) under `n_inst` general affine transforms: rotation + shear (a random 3x3), anisotropic
    scale 0.02 .. 30 per axis, every third one mirrored (negative determinant).  The stress fixture of the two-level layout:
    object-space node tests under transforms that are far from rigid."""
    rng = np.random.RandomState(seed)
    sb = SceneBuilder()
    p, n = uv_sphere(num_theta, 2 * num_theta, 0.5)
    m1 = sb.add_mesh(_f32(p), n)
    bx = _f32(box((-.5, -.3, -.2), (.5, .3, .2)))
    m2 = sb.add_mesh(bx, flat_normals(bx))
    mat = sb.diffuse((.6, .6, .6))
    for i in range(n_inst):
        A = rng.randn(3, 3) @ np.diag(np.exp(rng.uniform(np.log(0.02), np.log(30), 3)))
        if i % 3 == 0:
            A[:, 0] *= -1
        M = np.eye(4)
        M[:3, :3] = A * 0.5
        M[:3, 3] = rng.uniform(-4, 4, 3)
        sb.add_primitive(m1 if i % 2 else m2, mat, M=M)
    sb.set_envconst((1, 1, 1))
    sb.set_camera(position=(0, 0, 14), center=(0, 0, 0), up=(0, 1, 0), vfov=40, aspect=1)
    return sb


# ------------------------------------------------------------------------------------------------
# Procedural environment map (C4 variant): sun lobe + sky gradient, stored as Radiance RGBE so that the
# reference (texture::bitmap -> stb_image) and the device see bit-identical texels
# ------------------------------------------------------------------------------------------------
def _rgbe_encode(rgb):
    """float RGB -> RGBE bytes (Ward's float2rgbe: mantissa scaled by 256 / 2^e, truncated)."""
    rgb = np.asarray(rgb, dtype=np.float64)
    v = rgb.max(axis=-1)
    m, e = np.frexp(v)
    scale = np.where(v < 1e-32, 0.0, m * 256.0 / np.maximum(v, 1e-300))
    out = np.zeros(rgb.shape[:-1] + (4,), dtype=np.uint8)
    out[..., :3] = np.clip(rgb * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.where(v < 1e-32, 0, e + 128).astype(np.uint8)
    return out


def _rgbe_decode(rgbe):
    """stb_image's stbi__hdr_convert: rgb * 2^(e - 136), zero when e == 0; float32 like stbi_loadf."""
    e = rgbe[..., 3].astype(np.int32)
    f = np.where(e != 0, np.ldexp(np.float32(1.0), e - 136), np.float32(0)).astype(np.float32)
    return (rgbe[..., :3].astype(np.float32) * f[..., None]).astype(np.float32)


def procedural_env(w=1024, h=512, seed=5):
    """Returns the RGBE image in TEXTURE order (row 0 = v 0 = straight up after the reference's vertical flip)."""
    rng = np.random.RandomState(seed)
    v = (np.arange(h) + .5) / h
    u = (np.arange(w) + .5) / w
    theta, phi = np.pi * v[:, None], 2 * np.pi * u[None, :]
    d = np.stack([np.sin(theta) * np.sin(phi), np.cos(theta) * np.ones_like(phi), np.sin(theta) * np.cos(phi)], -1)
    sky = np.array([.35, .55, 1.0]) * (0.25 + 0.75 * np.clip(d[..., 1], 0, 1))[..., None]
    ground = np.array([.18, .15, .12]) * np.clip(-d[..., 1], 0, 1)[..., None] * .5
    sun_dir = np.array([.45, .65, .6]) + .05 * rng.randn(3)
    sun_dir /= np.linalg.norm(sun_dir)
    c = np.clip((d * sun_dir).sum(-1), 0, 1)
    sun = np.array([1.0, .9, .7]) * (40.0 * c ** 400 + 0.6 * c ** 12)[..., None]
    return _rgbe_encode(sky + ground + sun)


def write_hdr(path, rgbe_tex):
    """Writes the RGBE texels as a flat (non-RLE) Radiance .hdr whose FILE rows are the texture rows reversed:
    texture::bitmap loads with flip=true (src/texture/texture_bitmap.cpp:55-57)."""
    h, w = rgbe_tex.shape[:2]
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n" % (h, w))
        f.write(np.ascontiguousarray(rgbe_tex[::-1]).tobytes())


def env_light(path, w=1024, h=512, seed=5, rot=0.0, scale=1.0):
    """kwargs for SceneBuilder.set_env: the file for the reference, the decoded texels for the C ABI."""
    rgbe = procedural_env(w, h, seed)
    write_hdr(path, rgbe)
    return dict(rgb=_rgbe_decode(rgbe), rot=rot, scale=scale, path=path)


def write_png(path, img):
    """8-bit RGB / RGBA PNG from uint8 [h,w,3|4], row 0 = top (stdlib only)."""
    import struct
    import zlib
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, c = img.shape

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, {3: 2, 4: 6}[c], 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def png_texels(img):
    """What texture::bitmap holds after stbi_loadf(flip=true) of that PNG: colour channels through stb's LDR->HDR ramp
    pow(v/255, 2.2), alpha linear, rows bottom-up (src/texture/texture_bitmap.cpp:55-66)."""
    img = np.asarray(img, dtype=np.uint8)
    lin = img.astype(np.float32) / np.float32(255.0)
    out = np.power(lin.astype(np.float64), float(np.float32(2.2))).astype(np.float32)
    if img.shape[2] == 4:
        out[..., 3] = lin[..., 3]
    return np.ascontiguousarray(out[::-1])


def checker_texture(path, w=64, h=32, cells=8, alpha=False, seed=11):
    """Procedural colour checker (+ an alpha channel with holes) written as a PNG; returns kwargs for add_texture."""
    rng = np.random.RandomState(seed)
    ys, xs = np.mgrid[0:h, 0:w]
    cell = (xs * cells // w) + (ys * cells // h) * cells
    pal = rng.randint(40, 256, size=(cells * cells, 3)).astype(np.uint8)
    img = pal[cell]
    if alpha:
        a = np.where(((xs * cells // w) + (ys * cells // h)) % 3 == 0, 0, 255).astype(np.uint8)
        a[(ys * cells // h) % 4 == 1] = 128
        img = np.concatenate([img, a[..., None]], axis=2)
    write_png(path, img)
    return dict(data=png_texels(img), path=path)


def planar_uv(p, scale=1.0):
    """Texture coordinates from the xz (or xy for vertical faces) position of each vertex."""
    p = np.asarray(p, dtype=np.float64).reshape(-1, 3, 3)
    n = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    vertical = np.abs(n[:, 1]) < np.maximum(np.abs(n[:, 0]), np.abs(n[:, 2]))
    t = np.where(vertical[:, None, None], p[:, :, [0, 1]] + p[:, :, [2, 2]] * [1.0, 0.0], p[:, :, [0, 2]])
    return np.ascontiguousarray(t * scale)


def textured_scene(tmpdir, num_theta=12, kind="obj"):
    """Cornell shell whose floor / back wall / spheres carry the OBJ-import material set:
    kind = "obj": mixture_wavefrontobj with a mapKd checker (floor), with an RGBA mapKd whose alpha masks the surface
                  (a screen in front of the back wall), plain textured diffuse (back wall), untextured mixture_obj sphere
    kind = "mix": constant-weight mixtures (rr on one sphere, marginalized on the other, entries diffuse/glossy/mirror/mask)
    """
    import os
    sb = SceneBuilder()
    _cornell_shell(sb)
    tex = sb.add_texture(**checker_texture(os.path.join(str(tmpdir), "checker.png")))
    tex_a = sb.add_texture(**checker_texture(os.path.join(str(tmpdir), "checker_a.png"), alpha=True, seed=12))
    sp, sn = uv_sphere(num_theta)
    sp = _f32(sp * 0.35)
    th = np.arccos(np.clip(sn[..., 1], -1, 1)) / np.pi
    ph = (np.arctan2(sn[..., 0], sn[..., 2]) / (2 * np.pi)) % 1.0
    st = np.stack([ph, th], -1)
    sphere = sb.add_mesh(sp, sn, st)
    fl = _f32(quad((-0.8, 0.01, 0.8), (0.8, 0.01, 0.8), (0.8, 0.01, -0.8), (-0.8, 0.01, -0.8)))
    floor = sb.add_mesh(fl, flat_normals(fl), planar_uv(fl, 0.5))
    scr = _f32(quad((-0.9, 0.2, -0.5), (0.9, 0.2, -0.5), (0.9, 1.6, -0.5), (-0.9, 1.6, -0.5)))
    screen = sb.add_mesh(scr, flat_normals(scr), planar_uv(scr, 0.6))
    if kind == "obj":
        sb.add_primitive(floor, sb.mixture_obj((.5, .5, .5), (.3, .3, .3), .2, .1, map_kd=tex))
        sb.add_primitive(screen, sb.mixture_obj((.5, .5, .5), (.05, .05, .05), .3, .3, map_kd=tex_a))
        sb.add_primitive(sphere, sb.mixture_obj((.7, .3, .2), (.4, .4, .4), .15, .15), M=translate((-.4, .36, .1)))
        sb.add_primitive(sphere, sb.diffuse((1, 1, 1), map_kd=tex), M=translate((.45, .36, .3)))
    else:
        d, g, m, k = sb.diffuse((.6, .5, .2)), sb.glossy((.7, .7, .7), .1, .2), sb.mirror(), sb.mask()
        d2 = sb.diffuse((1, 1, 1), map_kd=tex)
        sb.add_primitive(floor, sb.mixture_marginalized([(d2, .7), (g, .3)]))
        sb.add_primitive(screen, sb.mixture_rr([(k, .6), (d, .4)]))
        sb.add_primitive(sphere, sb.mixture_rr([(d, .5), (g, .3), (m, .2)]), M=translate((-.4, .36, .1)))
        sb.add_primitive(sphere, sb.mixture_marginalized([(d, .4), (g, .3), (m, .2), (k, .1)]), M=translate((.45, .36, .3)))
    sb.set_camera(position=(0, 1, 5), center=(0, 1, 0), up=(0, 1, 0), vfov=30, aspect=16 / 9)
    return sb


def delta_lights_scene(grid=2, num_theta=12):
    """Open scene lit ONLY by a directional and a point light (func_lights.py pattern): every contribution comes
    through next-event estimation with MIS weight 1 (renderer_pt.cpp:176-180)."""
    sb = instanced_spheres(grid=grid, num_theta=num_theta)
    sb.lights = []
    sb.add_directional_light((3, 2.5, 2), (-0.3, -1, -0.2))
    sb.add_point_light((30, 30, 40), (2.0, 4.0, 3.0))
    return sb


# ------------------------------------------------------------------------------------------------
# Ray batches (host, FP64, float-representable origins) — SURVEY.md §8d C1/C3
# ------------------------------------------------------------------------------------------------
def camera_rays(camera, width, height):
    """Pixel-centre primary rays, the arithmetic of camera_pinhole.cpp:90-95,128-132."""
    import math

    def dot3(a, b):     # glm::dot: (x*x' + y*y') + z*z'
        return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]

    def cross3(x, y):   # glm::cross
        return np.stack([x[1] * y[2] - y[1] * x[2], x[2] * y[0] - y[2] * x[0], x[0] * y[1] - y[0] * x[1]])

    pos, center, up = (np.asarray(camera[k], dtype=np.float64) for k in ("position", "center", "up"))
    w = pos - center
    w = w * (1.0 / np.sqrt(dot3(w, w)))          # glm::normalize = v * inversesqrt(dot(v, v))
    u = cross3(up, w)
    u = u * (1.0 / np.sqrt(dot3(u, u)))
    v = cross3(w, u)
    tf = math.tan(camera["vfov"] * math.pi / 180 * .5)
    x = (np.arange(width) + .5) / width
    y = (np.arange(height) + .5) / height
    rx, ry = np.meshgrid(2 * x - 1, 2 * y - 1, indexing="xy")
    d = np.stack([camera["aspect"] * tf * rx, tf * ry, -np.ones_like(rx)], -1)
    d = d * (1.0 / np.sqrt(dot3(d, d)))[..., None]
    dirs = (d[..., 0:1] * u + d[..., 1:2] * v) + d[..., 2:3] * w
    rays = np.zeros(width * height, dtype=RAY_DTYPE)
    rays["o"] = pos
    rays["d"] = dirs.reshape(-1, 3)
    rays["tmin"] = 1e-3
    rays["tmax"] = 1e10
    return rays


RAY_DTYPE = np.dtype([("o", "<f8", 3), ("d", "<f8", 3), ("tmin", "<f8"), ("tmax", "<f8")])
HIT_DTYPE = np.dtype([("t", "<f8"), ("u", "<f8"), ("v", "<f8"), ("primitive", "<i4"), ("face", "<i4")])


def random_rays(n, lo, hi, seed=7):
    """Incoherent rays: origins uniform in the box [lo,hi], uniform directions."""
    rng = np.random.RandomState(seed)
    rays = np.zeros(n, dtype=RAY_DTYPE)
    rays["o"] = _f32(lo + (np.asarray(hi) - np.asarray(lo)) * rng.rand(n, 3))
    d = rng.randn(n, 3)
    rays["d"] = d / np.sqrt((d * d).sum(-1, keepdims=True))
    rays["tmin"] = 1e-3
    rays["tmax"] = 1e10
    return rays


def bounce_rays(rays, hits, tri_normal_fn, seed=7):
    """C3 (ii): one cosine-weighted bounce ray per hit of a primary batch."""
    ok = hits["primitive"] >= 0
    o = rays["o"][ok] + rays["d"][ok] * hits["t"][ok, None]
    n = tri_normal_fn(hits[ok])
    flip = (n * rays["d"][ok]).sum(-1) > 0
    n[flip] = -n[flip]
    rng = np.random.RandomState(seed)
    u0, u1 = rng.rand(len(o)), rng.rand(len(o))
    r, phi = np.sqrt(u0), 2 * np.pi * u1
    lx, ly = r * np.cos(phi), r * np.sin(phi)
    lz = np.sqrt(np.maximum(0, 1 - lx * lx - ly * ly))
    a = np.where(np.abs(n[:, 0:1]) > .9, np.array([[0., 1, 0]]), np.array([[1., 0, 0]]))
    t = np.cross(a, n)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    b = np.cross(n, t)
    d = lx[:, None] * t + ly[:, None] * b + lz[:, None] * n
    out = np.zeros(len(o), dtype=RAY_DTYPE)
    out["o"] = _f32(o)
    out["d"] = d
    out["tmin"] = 1e-3
    out["tmax"] = 1e10
    return out
