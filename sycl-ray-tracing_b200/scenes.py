"""Procedural workloads of the named shapes (BASELINE.json configs; exact definitions in SURVEY.md Appendix A).

The reference ships no Dragon OBJ and no HDR skysphere (.MISSING_LARGE_BLOBS), so C2..C5 use seeded procedural
stand-ins. All generators evaluate in float64 and round to float32 once, as Appendix A states.
"""
from __future__ import annotations

import numpy as np

from .api import Camera, Translation

# shipped data/OBJs/pbrt_dragon.mtl: Material.001 Kd (1, .71, .29) Pr .25 Pm 1; Material.002 Kd .8 Pr .4 Pm 1
DEFAULT_MATERIAL = [1.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 1.0]          # slot 0 of parse_obj (utils.cpp:75)


def _hash01(i, j, s):
    """Appendix A hash(i, j, s) -> [-1, 1), vectorised over uint32 arrays."""
    i = i.astype(np.uint32); j = j.astype(np.uint32)
    with np.errstate(over="ignore"):
        h = (i * np.uint32(73856093)) ^ (j * np.uint32(19349663)) ^ np.uint32((s * 83492791) & 0xFFFFFFFF)
        h ^= h >> np.uint32(16); h *= np.uint32(0x7feb352d)
        h ^= h >> np.uint32(15); h *= np.uint32(0x846ca68b)
        h ^= h >> np.uint32(16)
    return (h.astype(np.float64) * (2.0 ** -32) * 2.0 - 1.0).astype(np.float32).astype(np.float64)


def displaced_sphere(nu: int = 1000, nv: int = 500, seed: int = 1234, outward: bool = False, radius: float = 3.0,
                     center=(0.0, 0.0, 0.0)) -> np.ndarray:
    """Displaced lat-long sphere, 2*nu*nv triangles as (n, 9) float32.

    outward=False: Triangle(a,c,d), Triangle(a,d,b) (the C2 mesh); outward=True: Triangle(a,d,c), Triangle(a,b,d) (C3/C4)."""
    i = np.arange(nv + 1)[:, None]
    j = np.arange(nu)[None, :]
    theta = np.pi * (0.02 + 0.96 * i / nv)
    phi = 2.0 * np.pi * j / nu
    ii = np.broadcast_to(i, (nv + 1, nu)); jj = np.broadcast_to(j, (nv + 1, nu))
    r = radius + 0.15 * np.sin(7.0 * theta) * np.sin(5.0 * phi) + 0.05 * _hash01(ii, jj, seed)
    P = np.stack([r * np.sin(theta) * np.cos(phi), r * np.cos(theta) * np.broadcast_to(np.ones_like(phi), r.shape),
                  r * np.sin(theta) * np.sin(phi)], axis=-1)
    P = (P + np.asarray(center, np.float64)).astype(np.float32)
    a = P[:-1, :, :]
    b = np.roll(P, -1, axis=1)[:-1, :, :]
    c = P[1:, :, :]
    d = np.roll(P, -1, axis=1)[1:, :, :]
    if outward:
        t0 = np.concatenate([a, d, c], axis=-1); t1 = np.concatenate([a, b, d], axis=-1)
    else:
        t0 = np.concatenate([a, c, d], axis=-1); t1 = np.concatenate([a, d, b], axis=-1)
    tris = np.stack([t0, t1], axis=2).reshape(-1, 9)          # quad (i, j) -> its two triangles, i outer, j inner
    return np.ascontiguousarray(tris, np.float32)


def ground_quad(g: float = -3.3, half: float = 40.0) -> np.ndarray:
    return np.array([[-half, g, -half, -half, g, half, half, g, half],
                     [-half, g, -half, half, g, half, half, g, -half]], np.float32)


def procedural_sky(w: int = 2048, h: int = 1024) -> np.ndarray:
    """Sky gradient + Gaussian sun (Appendix A); (h, w, 4) float32, alpha 0 like read_image_float (utils.cpp:119)."""
    y = (np.arange(h)[:, None] + 0.5); x = (np.arange(w)[None, :] + 0.5)
    theta = np.pi * y / h; phi = 2.0 * np.pi * x / w
    d = np.stack([-np.sin(theta) * np.cos(phi), -np.cos(theta) * np.ones_like(phi), -np.sin(theta) * np.sin(phi)], axis=-1)
    el = np.deg2rad(35.0)
    sun = np.array([0.6 * np.cos(el), np.sin(el), 0.8 * np.cos(el)])
    ang = np.arccos(np.clip((d * sun).sum(-1), -1.0, 1.0))
    base = 0.3 + 1.2 * np.maximum(0.0, d[..., 1])
    s = 5.0e4 * np.exp(-0.5 * (ang / np.deg2rad(1.5)) ** 2)
    img = np.zeros((h, w, 4), np.float64)
    img[..., 0] = 0.6 * base + s
    img[..., 1] = 0.75 * base + 0.95 * s
    img[..., 2] = base + 0.85 * s
    return img.astype(np.float32)


def c2_scene(nu: int = 1000, nv: int = 500):
    """C2: primary-ray microbenchmark mesh (1 000 000 triangles at the default size), camera Camera(45, T(0,0,10.5))."""
    tri = displaced_sphere(nu, nv, seed=1234, outward=False)
    mats = np.array([DEFAULT_MATERIAL, [0, 0, 0, 1, 0.8, 0.8, 0.8, 1, 0.0, 1.0]], np.float32)
    return dict(tri9=tri, mat_idx=np.ones(len(tri), np.int32), mats10=mats, emissive=np.zeros(0, np.int32),
                camera=Camera(45.0, Translation(0.0, 0.0, 10.5)))


def c3_scene(roughness: float = 0.25, metalness: float = 1.0, nu: int = 1000, nv: int = 500, sky_w: int = 2048, sky_h: int = 1024):
    """C3/C4: Dragon-class stand-in: outward-wound displaced sphere (material 1 = Material.001 of pbrt_dragon.mtl) on a
    two-triangle ground (material 2 = Material.002), procedural sun+sky, PBRT_DRAGON_CAMERA."""
    mesh = displaced_sphere(nu, nv, seed=1234, outward=True)
    tri = np.concatenate([mesh, ground_quad()], axis=0)
    mat_idx = np.concatenate([np.ones(len(mesh), np.int32), np.full(2, 2, np.int32)])
    mats = np.array([DEFAULT_MATERIAL,
                     [0, 0, 0, 1, 1.0, 0.71, 0.29, 1, metalness, max(roughness, 1.0e-2)],     # clamp of utils.cpp:82
                     [0, 0, 0, 1, 0.8, 0.8, 0.8, 1, 1.0, 0.4]], np.float32)
    return dict(tri9=tri, mat_idx=mat_idx, mats10=mats, emissive=np.zeros(0, np.int32), env=procedural_sky(sky_w, sky_h),
                camera=Camera.PBRT_DRAGON_CAMERA)


def c5_scene(n_instances: int = 200, nu: int = 500, nv: int = 100, seed: int = 7, sky_w: int = 2048, sky_h: int = 1024):
    """C5: n_instances baked copies of a 2*nu*nv-triangle displaced sphere on a jittered grid + ground
    (200 x 100 000 = 20 M triangles at the default size)."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(np.sqrt(n_instances)))
    parts, mat = [], []
    spacing = 2.2
    for k in range(n_instances):
        gx, gz = k % side, k // side
        cx = (gx - (side - 1) / 2) * spacing + rng.uniform(-0.3, 0.3)
        cz = (gz - (side - 1) / 2) * spacing + rng.uniform(-0.3, 0.3) - 6.0
        cy = -3.3 + 1.0 + rng.uniform(0.0, 0.4)
        m = displaced_sphere(nu, nv, seed=seed + k, outward=True, radius=0.9, center=(cx, cy, cz))
        parts.append(m); mat.append(np.full(len(m), 1 + (k % 2), np.int32))
    parts.append(ground_quad()); mat.append(np.full(2, 2, np.int32))
    mats = np.array([DEFAULT_MATERIAL, [0, 0, 0, 1, 1.0, 0.71, 0.29, 1, 1.0, 0.25], [0, 0, 0, 1, 0.8, 0.8, 0.8, 1, 1.0, 0.4]], np.float32)
    return dict(tri9=np.concatenate(parts), mat_idx=np.concatenate(mat), mats10=mats, emissive=np.zeros(0, np.int32),
                env=procedural_sky(sky_w, sky_h), camera=Camera.PBRT_DRAGON_CAMERA)


def torus_knot(nu: int = 2000, nv: int = 250, seed: int = 42, p: int = 2, q: int = 3, scale: float = 1.15, tube: float = 0.52,
               center=(0.0, 0.2, 0.0)) -> np.ndarray:
    """Displaced (p, q) torus-knot tube, 2*nu*nv triangles (1 000 000 at the default size), outward wound: the concave
    Dragon-class stand-in SURVEY §8d prefers (self-occlusion and inter-reflection between the strands, unlike the near-convex
    displaced sphere). Curve C(t) = ((2 + cos q t) cos p t, sin q t, (2 + cos q t) sin p t) * scale; tube radius
    tube * (1 + 0.18 sin(9 t) sin(4 phi)) + 0.025 hash(i, j, seed); frame by parallel transport of the curve's normal."""
    t = 2.0 * np.pi * np.arange(nu) / nu
    c = np.stack([(2.0 + np.cos(q * t)) * np.cos(p * t), np.sin(q * t), (2.0 + np.cos(q * t)) * np.sin(p * t)], axis=-1) * scale
    d = np.stack([-q * np.sin(q * t) * np.cos(p * t) - p * (2.0 + np.cos(q * t)) * np.sin(p * t), q * np.cos(q * t),
                  -q * np.sin(q * t) * np.sin(p * t) + p * (2.0 + np.cos(q * t)) * np.cos(p * t)], axis=-1)
    tan = d / np.linalg.norm(d, axis=-1, keepdims=True)
    # a smooth normal field: remove the tangent component from the vector pointing away from the knot's axis (never parallel to it)
    radial = np.stack([np.cos(p * t), np.zeros_like(t), np.sin(p * t)], axis=-1)
    n1 = radial - (radial * tan).sum(-1, keepdims=True) * tan
    n1 /= np.linalg.norm(n1, axis=-1, keepdims=True)
    n2 = np.cross(tan, n1)
    phi = 2.0 * np.pi * np.arange(nv) / nv
    ii = np.broadcast_to(np.arange(nu)[:, None], (nu, nv)); jj = np.broadcast_to(np.arange(nv)[None, :], (nu, nv))
    r = tube * (1.0 + 0.18 * np.sin(9.0 * t)[:, None] * np.sin(4.0 * phi)[None, :]) + 0.025 * _hash01(ii, jj, seed)
    P = c[:, None, :] + r[..., None] * (np.cos(phi)[None, :, None] * n1[:, None, :] + np.sin(phi)[None, :, None] * n2[:, None, :])
    P = (P + np.asarray(center, np.float64)).astype(np.float32)
    a = P
    b = np.roll(P, -1, axis=1)
    cc = np.roll(P, -1, axis=0)
    dd = np.roll(np.roll(P, -1, axis=0), -1, axis=1)
    t0 = np.concatenate([a, b, dd], axis=-1); t1 = np.concatenate([a, dd, cc], axis=-1)
    return np.ascontiguousarray(np.stack([t0, t1], axis=2).reshape(-1, 9), np.float32)


def c3_knot_scene(roughness: float = 0.25, metalness: float = 1.0, nu: int = 2000, nv: int = 250, sky_w: int = 2048, sky_h: int = 1024):
    """C3 with the concave stand-in: the displaced torus knot (material 1) over the same ground, sky, camera and materials."""
    mesh = torus_knot(nu, nv)
    tri = np.concatenate([mesh, ground_quad()], axis=0)
    mat_idx = np.concatenate([np.ones(len(mesh), np.int32), np.full(2, 2, np.int32)])
    mats = np.array([DEFAULT_MATERIAL, [0, 0, 0, 1, 1.0, 0.71, 0.29, 1, metalness, max(roughness, 1.0e-2)],
                     [0, 0, 0, 1, 0.8, 0.8, 0.8, 1, 1.0, 0.4]], np.float32)
    return dict(tri9=tri, mat_idx=mat_idx, mats10=mats, emissive=np.zeros(0, np.int32), env=procedural_sky(sky_w, sky_h),
                camera=Camera.PBRT_DRAGON_CAMERA)


# the reference's own default assets (source/main.cpp:34-35); both are missing from the repository (.MISSING_LARGE_BLOBS) and are picked
# up when somebody drops them in
DRAGON_OBJ = "data/OBJs/pbrt_dragon.obj"
DRAGON_SKY = "data/Skyspheres/evening_road_01_puresky_2k.hdr"


def find_real_dragon(roots) -> tuple | None:
    import os
    for r in roots:
        if r and os.path.exists(os.path.join(r, DRAGON_OBJ)) and os.path.exists(os.path.join(r, DRAGON_SKY)):
            return os.path.join(r, DRAGON_OBJ), os.path.join(r, DRAGON_SKY)
    return None


def real_dragon_scene(obj_path: str, sky_path: str):
    """C3 on the reference's real assets, ingested by the library itself (b200rt_obj_load / b200rt_hdr_load)."""
    from .api import parse_obj, read_image_float
    o = parse_obj(obj_path)
    o["env"] = read_image_float(sky_path)          # (h, w, 3): expanded to RGBA on the device
    o["camera"] = Camera.PBRT_DRAGON_CAMERA
    return o


def algorithmic_bytes_per_ray(n_tri: int) -> int:
    """one root-to-leaf descent over the reference's data shapes (octree fan-out 8, 56-byte 7-slab
    volumes, 8 x 36-byte triangles per leaf): 24 + ceil(log8(max(N,8)/8)) * 8*56 + 8*36 + 8."""
    import math
    levels = math.ceil(math.log(max(n_tri, 8) / 8.0, 8) - 1e-12) if n_tri > 8 else 0
    levels = max(levels, 1) if n_tri > 8 else 0
    return 24 + levels * 8 * 56 + 8 * 36 + 8
