"""CPU oracle for the OcclusionEnv transition -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl reference``
legs may import this module.  The product package (``occlusionenv_b200``) never does.

It wraps ``libocclusion_oracle.so`` (``occl_oracle.c``: the strict-fp32 restatement of the pytorch3d
arithmetic) and restates the reference's state machine:

  * constants            /root/reference/environment.py:234-284  (``createRenderers``)
  * ``reset``            /root/reference/environment.py:286-328
  * ``step``             /root/reference/environment.py:352-396
  * vectorised loop      /root/reference/SubProcVecEnv.py:203-220 (sequential ``SimpleVecEnv``)

PARITY UNPINNED: the reference has no tests or golden vectors and pytorch3d (its arithmetic) is not
installable offline; see ``occl_oracle.c`` header and DESIGN.md.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_f = ctypes.POINTER(ctypes.c_float)
c_i = ctypes.POINTER(ctypes.c_int32)
c_d = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libocclusion_oracle.so")
    src = os.path.join(_HERE, "occl_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libocclusion_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _fp(a):
    return a.ctypes.data_as(c_f)


def _ip(a):
    return a.ctypes.data_as(c_i)


# ------------------------------------------------------------------------------------------------
# constants of createRenderers (environment.py:234-284), computed here independently of the product
# ------------------------------------------------------------------------------------------------
SIGMA = np.float32(1e-4)                                   # BlendParams(sigma=1e-4)      :242
BLUR_RADIUS = np.float32(np.log(1.0 / 1e-4 - 1.0) * 1e-4)  # RasterizationSettings        :251
FACES_PER_PIXEL = 100                                      #                              :252
LIGHT = np.array([2.0, 2.0, -2.0], np.float32)             # PointLights location         :275
STEP_SIZE = np.float32(0.05)                               #                              :219


def fov_scale(fov_deg: float = 60.0, znear: float = 1.0) -> np.float32:
    """K00 of FoVPerspectiveCameras (defaults fov=60 deg, znear=1), computed in fp32 like pytorch3d:
    fov*(pi/180) -> tan(fov/2) -> max_y = tan*znear -> K00 = 2*znear/(max_x-min_x)."""
    fov = np.float32(np.float32(fov_deg) * np.float32(np.pi / 180.0))
    t = np.float32(np.tan(np.float32(fov / np.float32(2.0))))
    max_y = np.float32(t * np.float32(znear))
    return np.float32(np.float32(2.0 * znear) / np.float32(max_y - (-max_y)))


PROJ_SCALE = fov_scale()


# ------------------------------------------------------------------------------------------------
# thin wrappers
# ------------------------------------------------------------------------------------------------
# Switches for the discretionary choices of the restatement (see occl_oracle.c); defaults = what the kernels implement.
OPT_TRIG_FP32, OPT_PROJ_MATRIX, OPT_NEIGHBOR_TOPK = 0, 1, 2
OPTIONS = {"trig_fp32": OPT_TRIG_FP32, "proj_matrix": OPT_PROJ_MATRIX, "neighbor_topk": OPT_NEIGHBOR_TOPK}
_PY_OPTIONS = {"clip_lerp_ndc": 0, "specular_center_inverse": 0}


def set_option(name: str, value: int) -> None:
    """``trig_fp32``: sin/cos in fp32 libm instead of double-rounded-once; ``proj_matrix``: one composed 4x4 applied to
    (x, y, z, 1) then / w instead of (s x_view) / z_view; ``neighbor_topk``: cut-quadrilateral triangles matched only
    among the current <= K nearest hits; ``clip_lerp_ndc``: clip_faces interpolates the cut vertex as
    p1 + w (p - p1) on (x z, y z) in one fused expression order (a + w*(b - a)) instead of a*(1-w) + b*w;
    ``specular_center_inverse``: camera centre for the specular term recovered as -T R^-1 (4x4 inverse route of
    cameras.get_camera_center()) instead of the look-at position C."""
    if name in OPTIONS:
        if lib().occl_oracle_set_option(OPTIONS[name], int(value)) != 0:
            raise ValueError(name)
    elif name in _PY_OPTIONS:
        _PY_OPTIONS[name] = int(value)
    else:
        raise ValueError(f"unknown oracle option {name!r}")


def get_options() -> dict:
    d = {k: int(lib().occl_oracle_get_option(v)) for k, v in OPTIONS.items()}
    d.update(_PY_OPTIONS)
    return d


def pose_step(action, el, az, radius, step_size=STEP_SIZE):
    """environment.py:356-368.  Returns (el', az', C(3), R(3,3), T(3)) in fp32."""
    a = np.ascontiguousarray(action, np.float32)
    e = ctypes.c_float(float(el))
    z = ctypes.c_float(float(az))
    C = np.zeros(3, np.float32)
    R = np.zeros(9, np.float32)
    T = np.zeros(3, np.float32)
    lib().occl_oracle_pose_step(_fp(a), ctypes.c_float(float(step_size)), ctypes.c_float(float(radius)),
                                ctypes.byref(e), ctypes.byref(z), _fp(C), _fp(R), _fp(T))
    return np.float32(e.value), np.float32(z.value), C, R.reshape(3, 3), T


def pose_lookat(radius, el, az):
    """environment.py:308 (look_at_view_transform, radians)."""
    C = np.zeros(3, np.float32)
    R = np.zeros(9, np.float32)
    T = np.zeros(3, np.float32)
    lib().occl_oracle_pose_lookat(ctypes.c_float(float(radius)), ctypes.c_float(float(el)),
                                  ctypes.c_float(float(az)), _fp(C), _fp(R), _fp(T))
    return C, R.reshape(3, 3), T


def look_at(C):
    C = np.ascontiguousarray(C, np.float32)
    R = np.zeros(9, np.float32)
    T = np.zeros(3, np.float32)
    lib().occl_oracle_look_at(_fp(C), _fp(R), _fp(T))
    return R.reshape(3, 3), T


def project(verts, R, T, s=PROJ_SCALE):
    verts = np.ascontiguousarray(verts, np.float32)
    R = np.ascontiguousarray(R, np.float32).reshape(-1)
    T = np.ascontiguousarray(T, np.float32)
    out = np.zeros_like(verts)
    lib().occl_oracle_project(_fp(verts), verts.shape[0], _fp(R), _fp(T), ctypes.c_float(float(s)), _fp(out))
    return out


@dataclass
class Fragments:
    pix_to_face: np.ndarray  # (S,S,K) int32
    zbuf: np.ndarray         # (S,S,K)
    bary: np.ndarray         # (S,S,K,3)
    dists: np.ndarray        # (S,S,K)
    nhits: np.ndarray        # (S,S) hits before the K cut
    straddles: bool = False  # set by rasterize_clipped


def rasterize(vproj, faces, S, blur_radius, K, cull_backfaces=True, perspective_correct=True,
              clip_barycentric_coords=None) -> Fragments:
    """MeshRasterizer.forward flags: clip_barycentric_coords defaults to blur>0."""
    if clip_barycentric_coords is None:
        clip_barycentric_coords = blur_radius > 0
    vproj = np.ascontiguousarray(vproj, np.float32)
    faces = np.ascontiguousarray(faces, np.int32)
    p2f = np.empty((S, S, K), np.int32)
    zb = np.empty((S, S, K), np.float32)
    ba = np.empty((S, S, K, 3), np.float32)
    di = np.empty((S, S, K), np.float32)
    nh = np.zeros((S, S), np.int32)
    lib().occl_oracle_rasterize(_fp(vproj), _ip(faces), faces.shape[0], S, ctypes.c_float(float(blur_radius)),
                                K, int(perspective_correct), int(clip_barycentric_coords),
                                int(cull_backfaces), _ip(p2f), _fp(zb), _fp(ba), _fp(di), _ip(nh))
    return Fragments(p2f, zb, ba, di, nh)


def silhouette(frag: Fragments, sigma=SIGMA) -> np.ndarray:
    S, _, K = frag.pix_to_face.shape
    alpha = np.empty((S, S), np.float32)
    lib().occl_oracle_silhouette(_ip(frag.pix_to_face), _fp(frag.dists), S, K, ctypes.c_float(float(sigma)),
                                 _fp(alpha))
    return alpha


def flat_shade(verts, faces, frag: Fragments, cam, light=LIGHT) -> np.ndarray:
    """(4,S,S) observation: flat-shaded RGB + depth channel (environment.py:375-378)."""
    S = frag.pix_to_face.shape[0]
    verts = np.ascontiguousarray(verts, np.float32)
    faces = np.ascontiguousarray(faces, np.int32)
    obs = np.empty((4, S, S), np.float32)
    p2f = np.ascontiguousarray(frag.pix_to_face[..., 0])
    bary = np.ascontiguousarray(frag.bary[..., 0, :])
    zb = np.ascontiguousarray(frag.zbuf[..., 0])
    cam = np.ascontiguousarray(cam, np.float32)
    light = np.ascontiguousarray(light, np.float32)
    lib().occl_oracle_flat_shade(_fp(verts), _ip(faces), _ip(p2f), _fp(bary), _fp(zb), S, _fp(cam), _fp(light),
                                 _fp(obs))
    return obs


def rasterize_backward(vproj, faces, frag: Fragments, grad_dists) -> np.ndarray:
    vproj = np.ascontiguousarray(vproj, np.float32)
    faces = np.ascontiguousarray(faces, np.int32)
    S, _, K = frag.pix_to_face.shape
    g = np.ascontiguousarray(grad_dists, np.float32)
    out = np.zeros((vproj.shape[0], 3), np.float64)
    lib().occl_oracle_rasterize_backward(_fp(vproj), _ip(faces), vproj.shape[0], S, K, 1,
                                         _ip(frag.pix_to_face), _fp(g), out.ctypes.data_as(c_d))
    return out


# ------------------------------------------------------------------------------------------------
# one render of the whole scene from (C, R, T): what reset() and step() share
# ------------------------------------------------------------------------------------------------
Z_CLIP = np.float32(0.5)  # MeshRasterizer: z_clip_value = znear / 2 for perspective cameras


def rasterize_fv(face_verts, neighbor, S, blur_radius, K, cull_backfaces=True, perspective_correct=True,
                 clip_barycentric_coords=None) -> Fragments:
    """rasterize() on explicit face vertices (F,3,3) with pytorch3d's clipped_faces_neighbor_idx rule."""
    if clip_barycentric_coords is None:
        clip_barycentric_coords = blur_radius > 0
    fv = np.ascontiguousarray(face_verts, np.float32).reshape(-1, 9)
    F = fv.shape[0]
    nb = None if neighbor is None else np.ascontiguousarray(neighbor, np.int32)
    p2f = np.empty((S, S, K), np.int32)
    zb = np.empty((S, S, K), np.float32)
    ba = np.empty((S, S, K, 3), np.float32)
    di = np.empty((S, S, K), np.float32)
    nh = np.zeros((S, S), np.int32)
    lib().occl_oracle_rasterize_fv(_fp(fv), _ip(nb) if nb is not None else None, F, S, ctypes.c_float(float(blur_radius)),
                                   K, int(perspective_correct), int(clip_barycentric_coords), int(cull_backfaces),
                                   _ip(p2f), _fp(zb), _fp(ba), _fp(di), _ip(nh))
    return Fragments(p2f, zb, ba, di, nh)


def clip_faces(face_verts, z_clip=Z_CLIP):
    """pytorch3d renderer/mesh/clip.py::clip_faces for the frustum MeshRasterizer builds (only z_clip_value set,
    perspective_correct=True, cull=False), on face_verts (F,3,3) = (x_ndc, y_ndc, z_view).  Per face, by the number
    of vertices with z < z_clip:  0 -> kept;  3 -> removed;  1 -> the quadrilateral in front of the plane, as the
    two triangles t1 = (p4, p2, p5), t2 = (p5, p2, p3);  2 -> the triangle (p1, p4, p5).  p1 is the isolated vertex,
    p2, p3 follow it in the face's own order, p4 / p5 are the intersections of the edges p1p2 / p1p3 with z = z_clip:
        w = (p1.z - z_clip) / (p1.z - p.z);   xy = (p1.xy * p1.z * (1 - w) + p.xy * p.z * w) / z_clip
    (interpolated in view space -- x_ndc * z is s * x_view -- and projected again).  [P3D-recalled; the operation
    order of the interpolation is this restatement's choice: parity unpinned.]
    Returns (clipped face_verts (Fc,3,3), clipped->unclipped index (Fc,), neighbour index (Fc,) or -1,
    barycentric conversion (Fc,3,3): rows = barycentrics of the clipped face's vertices in the unclipped face)."""
    fv = np.ascontiguousarray(face_verts, np.float32)
    zc = np.float32(z_clip)
    out_v, out_idx, out_nb, out_conv = [], [], [], []
    eye = np.eye(3, dtype=np.float32)
    one = np.float32(1.0)
    for f in range(fv.shape[0]):
        v = fv[f]
        behind = v[:, 2] < zc
        n = int(behind.sum())
        if n == 0:
            out_v.append(v); out_idx.append(f); out_nb.append(-1); out_conv.append(eye)
            continue
        if n == 3:
            continue
        i = int(np.argmax(behind)) if n == 1 else int(np.argmax(~behind))
        j, k = (i + 1) % 3, (i + 2) % 3
        p1, p2, p3 = v[i], v[j], v[k]

        def cut(p):
            w = np.float32((p1[2] - zc) / (p1[2] - p[2]))
            a1 = np.float32(one - w)
            if _PY_OPTIONS["clip_lerp_ndc"]:  # a + w (b - a)
                ax, bx = np.float32(p1[0] * p1[2]), np.float32(p[0] * p[2])
                ay, by = np.float32(p1[1] * p1[2]), np.float32(p[1] * p[2])
                x = np.float32(np.float32(ax + np.float32(w * np.float32(bx - ax))) / zc)
                y = np.float32(np.float32(ay + np.float32(w * np.float32(by - ay))) / zc)
                return np.array([x, y, zc], np.float32), w
            x = np.float32(np.float32(np.float32(p1[0] * p1[2]) * a1 + np.float32(np.float32(p[0] * p[2]) * w)) / zc)
            y = np.float32(np.float32(np.float32(p1[1] * p1[2]) * a1 + np.float32(np.float32(p[1] * p[2]) * w)) / zc)
            return np.array([x, y, zc], np.float32), w

        p4, w2 = cut(p2)
        p5, w3 = cut(p3)
        b4 = np.zeros(3, np.float32); b4[i] = one - w2; b4[j] = w2
        b5 = np.zeros(3, np.float32); b5[i] = one - w3; b5[k] = w3
        if n == 1:
            base = len(out_v)
            out_v += [np.stack([p4, p2, p5]), np.stack([p5, p2, p3])]
            out_idx += [f, f]
            out_nb += [base + 1, base]
            out_conv += [np.stack([b4, eye[j], b5]), np.stack([b5, eye[j], eye[k]])]
        else:
            out_v.append(np.stack([p1, p4, p5])); out_idx.append(f); out_nb.append(-1)
            out_conv.append(np.stack([eye[i], b4, b5]))
    if not out_v:
        return (np.zeros((0, 3, 3), np.float32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 3, 3), np.float32))
    return (np.stack(out_v).astype(np.float32), np.asarray(out_idx, np.int32), np.asarray(out_nb, np.int32),
            np.stack(out_conv).astype(np.float32))


def rasterize_clipped(vproj, faces, S, blur_radius, K, **kw) -> "Fragments":
    """clip_faces -> rasterize_meshes -> convert_clipped_rasterization_to_original_faces: face indices are mapped
    back to the unclipped mesh and the barycentrics of cut faces are expressed in the unclipped face."""
    vproj = np.ascontiguousarray(vproj, np.float32)
    faces = np.ascontiguousarray(faces, np.int32)
    fv = vproj[faces]
    if not (fv[:, :, 2] < Z_CLIP).any():       # clip_faces returns its input unchanged
        return rasterize(vproj, faces, S, blur_radius, K, **kw)
    cv, idx, nb, conv = clip_faces(fv)
    fr = rasterize_fv(cv, nb, S, blur_radius, K, **kw)
    p2f = fr.pix_to_face
    hit = p2f >= 0
    safe = np.maximum(p2f, 0)
    if len(idx):
        cut = hit & (idx[safe] >= 0) & ~np.all(conv[safe] == np.eye(3, dtype=np.float32), axis=(-1, -2))
        nb_ = np.einsum("...k,...kj->...j", fr.bary, conv[safe]).astype(np.float32)
        fr.bary = np.where(cut[..., None], nb_, fr.bary).astype(np.float32)
        fr.pix_to_face = np.where(hit, idx[safe], p2f).astype(np.int32)
    fr.straddles = bool((nb >= 0).any() or len(idx) != len(np.unique(idx)) or
                        (((fv[:, :, 2] < Z_CLIP).sum(1) % 3) != 0).any())
    return fr


@dataclass
class RenderOut:
    obs: np.ndarray            # (4,S,S)
    alphas: np.ndarray         # (n_obj,S,S) soft silhouettes
    occl: np.ndarray           # (S,S)  sum_{i<j} A_i A_j  (alpha channel of self.image)
    loss: np.float32
    objects_sq: np.float32     # sum (sum_i A_i)^2   (normWithObjectSize branch, :324)
    pix_to_face: np.ndarray    # (S,S) scene K=1
    zbuf: np.ndarray           # (S,S)
    bary: np.ndarray           # (S,S,3)
    nhits: np.ndarray          # (n_obj,S,S)
    n_covered: np.ndarray      # (n_obj,) px hard-covered by object i rendered alone
    n_visible: np.ndarray      # (n_obj,) px whose nearest scene face belongs to object i
    vproj: np.ndarray          # (V,3)
    frags: list                # per-object K=100 fragments (face ids local to the object)
    zclip_straddle: bool = False  # a face straddles z_clip: pytorch3d would cut it; this frame is not comparable


def render_scene(verts, faces, obj_face_start, obj_vert_start, S, C, R, T, s=PROJ_SCALE,
                 blur=BLUR_RADIUS, sigma=SIGMA, K=FACES_PER_PIXEL, light=LIGHT) -> RenderOut:
    n_obj = len(obj_face_start) - 1
    vproj = project(verts, R, T, s)
    alphas, nhits, n_cov, frags = [], [], [], []
    for i in range(n_obj):
        v0 = obj_vert_start[i]
        f0, f1 = obj_face_start[i], obj_face_start[i + 1]
        of = faces[f0:f1] - v0
        ov = vproj[v0:obj_vert_start[i + 1]]
        fr = rasterize_clipped(ov, of, S, blur, K)               # silhouette settings :249-255
        frags.append(fr)
        alphas.append(silhouette(fr, sigma))
        nhits.append(fr.nhits)
        hard = rasterize_clipped(ov, of, S, 0.0, 1)              # object alone, hard coverage
        n_cov.append(int((hard.pix_to_face[..., 0] >= 0).sum()))
    alphas = np.stack(alphas)
    occl = np.zeros((S, S), np.float32)
    for i in range(n_obj):
        for j in range(i + 1, n_obj):
            occl = occl + alphas[i] * alphas[j]                  # environment.py:373 generalised
    loss = np.float32(np.sum((occl.astype(np.float64)) ** 2))
    objs = np.sum(alphas.astype(np.float32), axis=0)
    objects_sq = np.float32(np.sum(objs.astype(np.float64) ** 2))
    scene = rasterize_clipped(vproj, faces, S, 0.0, 1)           # observation settings :267-273
    cam_center = C
    if _PY_OPTIONS["specular_center_inverse"]:
        # cameras.get_camera_center(): the world-to-view 4x4 [[R, 0], [T, 1]] inverted in fp32, centre = last row
        M = np.eye(4, dtype=np.float32)
        M[:3, :3] = np.asarray(R, np.float32)
        M[3, :3] = np.asarray(T, np.float32)
        cam_center = np.linalg.inv(M).astype(np.float32)[3, :3]
    obs = flat_shade(verts, faces, scene, cam_center, light)
    p2f = scene.pix_to_face[..., 0]
    n_vis = [int(((p2f >= obj_face_start[i]) & (p2f < obj_face_start[i + 1])).sum()) for i in range(n_obj)]
    return RenderOut(obs, alphas, occl, loss, objects_sq, p2f, scene.zbuf[..., 0], scene.bary[..., 0, :],
                     np.stack(nhits), np.asarray(n_cov), np.asarray(n_vis), vproj, frags,
                     any(getattr(f, "straddles", False) for f in frags) or bool(getattr(scene, "straddles", False)))


class OracleOcclusionEnv:
    """Reference-faithful single environment (environment.py:201-402) on the CPU oracle.

    Differences that are deliberate: n_obj-generic reward instead of the
    hard-wired ``meshes[1..3]`` (B-1); no bare ``except`` retry loop; the scene is passed in.
    """

    def __init__(self, verts, faces, obj_face_start, obj_vert_start, img_size=512):
        self.verts = np.ascontiguousarray(verts, np.float32)
        self.faces = np.ascontiguousarray(faces, np.int32)
        self.obj_face_start = np.asarray(obj_face_start, np.int32)
        self.obj_vert_start = np.asarray(obj_vert_start, np.int32)
        self.img_size = img_size
        self.step_size = STEP_SIZE
        self.normWithObjectSize = False
        self.last = None

    def _render(self, C, R, T):
        self.last = render_scene(self.verts, self.faces, self.obj_face_start, self.obj_vert_start,
                                 self.img_size, C, R, T)
        return self.last

    def reset(self, radius=4.0, azimuth=0.0, elevation=0.0):
        self.camera_position = np.zeros(3, np.float32)            # :302
        self.radius = np.float32(radius)
        self.elevation = np.float32(elevation)
        self.azimuth = np.float32(azimuth)
        C, R, T = pose_lookat(self.radius, self.elevation, self.azimuth)  # :308
        out = self._render(C, R, T)
        self.fullReward = out.loss                                # :323
        self.objectMass = np.float32((out.objects_sq if self.normWithObjectSize else out.loss) + np.float32(1))
        return out.obs[None]

    def step(self, action):
        el, az, C, R, T = pose_step(action, self.elevation, self.azimuth, self.radius, self.step_size)
        self.elevation, self.azimuth, self.camera_position = el, az, C
        out = self._render(C, R, T)
        loss = out.loss
        reward = np.float32(self.fullReward - loss)               # :382
        self.fullReward = loss                                    # :384
        finished = bool(self.fullReward < np.float32(0.1))        # :386
        reward = np.float32(reward / self.objectMass)             # :387
        reward = np.float32(reward + np.float32(5)) if finished else np.float32(reward - np.float32(0.2))
        info = {"full_state": out.occl, "position": C, "full_reward": self.fullReward}
        return out.obs[None], reward, finished, info


class OracleSimpleVecEnv:
    """SubProcVecEnv.py:189-235: a sequential in-process loop (auto-reset with defaults on done)."""

    def __init__(self, envs):
        self.envs = envs
        self.num_envs = len(envs)

    def reset(self, azimuths=None):
        return np.stack([e.reset(azimuth=0.0 if azimuths is None else azimuths[i]) for i, e in enumerate(self.envs)])

    def step(self, actions):
        obs, rews, dones, infos = [], [], [], []
        for i, e in enumerate(self.envs):
            o, r, d, info = e.step(actions[i])
            if d:
                info["terminal_observation"] = o
                o = e.reset()
            obs.append(o[0])
            rews.append(r)
            dones.append(d)
            infos.append(info)
        return np.stack(obs), np.asarray(rews, np.float32), np.asarray(dones), infos
