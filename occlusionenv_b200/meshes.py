"""Scene meshes for the occlusion environment (host side, numpy).

A *scene* is what the reference calls ``self.meshes`` (``environment.py:294``): a list
``[full_mesh, obj_1, obj_2, ...]``.  Here it is one packed vertex array, one packed face array and
the face/vertex ranges of the objects inside it; the "full" mesh of the reference
(``environment.py:65-66,82-86`` / ``join_meshes_as_scene`` at ``:191``) is simply the whole range.

Reference behaviour reproduced:
  * ``load_default_meshes`` (``environment.py:53-88``): teapot + the same teapot shifted by
    ``(+2, 0, 0)``.  (The reference returns 3 meshes but indexes 4 -- SURVEY Appendix B-1 -- so the
    reward is generalised to ``sum_{i<j} A_i A_j`` and works for 2 or 3 objects.)
  * ``load_shapenet_meshes`` scene *layout* (``environment.py:147-148,171``): offsets ``(x2,0,1)``
    and ``(-x2,0,2)``; ShapeNet itself is not available offline, so procedural displaced
    icospheres of ShapeNet-like size stand in (BASELINE config 3).
  * ``pytorch3d.io.load_obj`` for ``v`` / ``f a//n`` records (``environment.py:56-57``).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


@dataclass
class SceneMesh:
    """Packed scene: ``verts`` (V,3) f32 world coordinates, ``faces`` (F,3) i32 indices into
    ``verts``, ``obj_face_start`` (n_obj+1,) i32 face ranges of the objects."""

    verts: np.ndarray
    faces: np.ndarray
    obj_face_start: np.ndarray
    obj_vert_start: np.ndarray

    @property
    def n_obj(self) -> int:
        return len(self.obj_face_start) - 1

    def object(self, i: int):
        """(verts, faces) of object ``i`` as a stand-alone mesh (local indices)."""
        v0, v1 = self.obj_vert_start[i], self.obj_vert_start[i + 1]
        f0, f1 = self.obj_face_start[i], self.obj_face_start[i + 1]
        return self.verts[v0:v1].copy(), (self.faces[f0:f1] - v0).astype(np.int32)

    @property
    def max_object_faces(self) -> int:
        return int(np.max(np.diff(self.obj_face_start)))


def pack_scene(objects: Sequence[tuple]) -> SceneMesh:
    """Concatenate ``[(verts, faces), ...]`` the way the reference builds its full mesh."""
    verts, faces, fstart, vstart = [], [], [0], [0]
    for v, f in objects:
        v = np.ascontiguousarray(v, dtype=np.float32)
        f = np.ascontiguousarray(f, dtype=np.int32)
        faces.append(f + vstart[-1])
        verts.append(v)
        vstart.append(vstart[-1] + v.shape[0])
        fstart.append(fstart[-1] + f.shape[0])
    return SceneMesh(
        np.ascontiguousarray(np.concatenate(verts, 0)),
        np.ascontiguousarray(np.concatenate(faces, 0)),
        np.asarray(fstart, np.int32),
        np.asarray(vstart, np.int32),
    )


def load_obj(path: str):
    """Minimal OBJ reader: ``v x y z`` and ``f a b c`` / ``a/t/n`` / ``a//n`` (1-based, negative
    indices allowed); polygons are fan-triangulated like ``pytorch3d.io.load_obj``."""
    verts: List[List[float]] = []
    faces: List[List[int]] = []
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith("v "):
                p = line.split()
                verts.append([float(p[1]), float(p[2]), float(p[3])])
            elif line.startswith("f "):
                idx = []
                for tok in line.split()[1:]:
                    i = int(tok.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(verts) + i)
                for k in range(1, len(idx) - 1):
                    faces.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(verts, np.float32), np.asarray(faces, np.int32)


def load_teapot():
    """The reference's ``data/teapot.obj`` (1292 verts, 2464 faces), shipped as a binary fixture."""
    d = np.load(os.path.join(_DATA_DIR, "teapot.npz"))
    return d["verts"].astype(np.float32), d["faces"].astype(np.int32)


def make_box(center=(2.0, 0.5, 0.0), half=(0.35, 0.45, 0.55)):
    """Axis-aligned box occluder, 8 verts / 12 faces, outward counter-clockwise winding (so
    ``cull_backfaces`` keeps the faces turned towards the camera).  Not in the reference: BASELINE
    config 1/2 name a "box occluder"; this is its definition."""
    cx, cy, cz = center
    hx, hy, hz = half
    v = np.array(
        [[cx + sx * hx, cy + sy * hy, cz + sz * hz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)],
        np.float32,
    )
    # vertex id = 4*ix + 2*iy + iz
    quads = [
        (0, 1, 3, 2),  # -x
        (4, 6, 7, 5),  # +x
        (0, 4, 5, 1),  # -y
        (2, 3, 7, 6),  # +y
        (0, 2, 6, 4),  # -z
        (1, 5, 7, 3),  # +z
    ]
    f = []
    for a, b, c, d in quads:
        f += [[a, b, c], [a, c, d]]
    f = np.asarray(f, np.int32)
    # make every face outward: flip when the normal points towards the centre
    ctr = np.asarray(center, np.float32)
    for i in range(len(f)):
        p0, p1, p2 = v[f[i, 0]], v[f[i, 1]], v[f[i, 2]]
        n = np.cross(p1 - p0, p2 - p0)
        if np.dot(n, (p0 + p1 + p2) / 3 - ctr) < 0:
            f[i, 1], f[i, 2] = f[i, 2], f[i, 1]
    return v, f


def default_scene(occluder: str = "teapot") -> SceneMesh:
    """Target teapot + occluder.  ``occluder='teapot'`` is the reference scene
    (``environment.py:55,63``: the same teapot shifted by +2 in x); ``'box'`` is the BASELINE box."""
    tv, tf = load_teapot()
    if occluder == "teapot":
        ov = tv + np.array([2.0, 0.0, 0.0], np.float32)
        of = tf
    elif occluder == "box":
        ov, of = make_box()
    else:
        raise ValueError(f"unknown occluder {occluder!r}")
    return pack_scene([(tv, tf), (ov, of)])


def icosphere(subdiv: int):
    """Unit icosphere, 20*4**subdiv faces (subdiv=5 -> 20480 faces / 10242 verts)."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t),
         (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4),
         (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8),
         (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.asarray(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache = {}
        nf = []

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.asarray(v, np.float32), np.asarray(f, np.int32)


def procedural_object(seed: int, subdiv: int = 5, scale: float = 0.5):
    """ShapeNet-sized stand-in: icosphere radially displaced by seeded low-frequency noise."""
    v, f = icosphere(subdiv)
    rng = np.random.default_rng(seed)
    disp = np.ones(len(v), np.float64)
    for _ in range(6):
        k = rng.normal(size=3) * 2.0
        ph = rng.uniform(0, 2 * np.pi)
        disp += 0.12 * np.sin(v.astype(np.float64) @ k + ph)
    axes = rng.uniform(0.6, 1.0, size=3)
    out = v.astype(np.float64) * disp[:, None] * axes[None, :] * scale
    return out.astype(np.float32), f


def procedural_scene(seed: int, n_obj: int = 3, subdiv: int = 5) -> SceneMesh:
    """Three procedural objects in the reference's ShapeNet layout (``environment.py:147-148,171``):
    object 2 offset ``(x2, 0, 1)``, object 3 offset ``(-x2, 0, 2)``, ``x2 ~ N(0,1)``."""
    rng = np.random.default_rng(seed + 7919)
    x2 = float(rng.normal())
    if not 1 <= n_obj <= 4:
        raise ValueError("procedural_scene: n_obj must be 1..4 (OCCL_MAX_OBJ)")
    # a fourth object (not in the reference's layout) goes behind the others on the axis
    offs = [(0.0, 0.0, 0.0), (x2, 0.0, 1.0), (-x2, 0.0, 2.0), (0.5 * x2, 0.0, -1.5)][:n_obj]
    objs = []
    for i, o in enumerate(offs):
        v, f = procedural_object(seed * 3 + i, subdiv)
        objs.append((v + np.asarray(o, np.float32), f))
    return pack_scene(objs)
