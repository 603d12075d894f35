"""Oracle (b): dense PyTorch formulation of the same transition -- TEST INFRASTRUCTURE ONLY.

Written independently of ``occl_oracle.c`` (broadcast pixel x face tensors instead of scalar loops) from
SURVEY.md Appendix A, for two purposes:
  1. cross-check the C restatement (forward values), and
  2. provide the gradient oracle: d reward / d action by autograd through pose -> projection ->
     point/triangle distances -> sigmoid blend -> occlusion loss (the route of ``demo.py:85-86``),
     in float64 so that it can also be validated against finite differences.

Reference lines restated: ``environment.py:356-368`` (pose), ``:373,381-392`` (reward) and the pytorch3d
pieces listed in ``occl_oracle.c``.  PARITY UNPINNED (no pytorch3d offline), see DESIGN.md.
"""
from __future__ import annotations

import math

import numpy as np
import torch

K_EPS = 1e-8


def _normalize(v, eps):
    return v / v.norm().clamp_min(eps)


def look_at(C):
    up = torch.tensor([0.0, 1.0, 0.0], dtype=C.dtype)
    z = _normalize(-C, 1e-5)
    x = _normalize(torch.linalg.cross(up, z), 1e-5)
    y = _normalize(torch.linalg.cross(z, x), 1e-5)
    if bool((x.abs() <= 5e-3).all()):
        x = _normalize(torch.linalg.cross(y, z), 1e-5)
    R = torch.stack([x, y, z], dim=1)
    T = -(R.t() @ C)
    return R, T


def pose_step(action, el, az, radius, step_size=0.05):
    n = action.norm()
    na = action / n if float(n.detach()) != 0.0 else action
    el = el + na[0] * step_size
    az = az + na[1] * step_size
    C = torch.stack([radius * torch.sin(az) * torch.cos(el), radius * torch.sin(az) * torch.sin(el),
                     radius * torch.cos(az)])
    R, T = look_at(C)
    return el, az, C, R, T


def pose_lookat(radius, el, az):
    C = torch.stack([radius * torch.cos(el) * torch.sin(az), radius * torch.sin(el),
                     radius * torch.cos(el) * torch.cos(az)])
    R, T = look_at(C)
    return C, R, T


def project(verts, R, T, s):
    vv = verts @ R + T
    return torch.stack([s * vv[:, 0] / vv[:, 2], s * vv[:, 1] / vv[:, 2], vv[:, 2]], dim=1)


def _edge(px, py, ax, ay, bx, by):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def _seg(px, py, ax, ay, bx, by):
    bax, bay = bx - ax, by - ay
    l2 = bax * bax + bay * bay
    degenerate = l2 <= K_EPS
    t = (bax * (px - ax) + bay * (py - ay)) / torch.where(degenerate, torch.ones_like(l2), l2)
    t = t.clamp(0.0, 1.0)
    t = torch.where(degenerate, torch.ones_like(t), t)
    qx, qy = ax + t * bax, ay + t * bay
    return (px - qx) ** 2 + (py - qy) ** 2


def pixel_centers(S, dtype):
    i = torch.arange(S, dtype=dtype)
    c = -1.0 + (2.0 * (S - 1 - i) + 1.0) / S  # centre of output index i (rows and columns alike)
    return c


Z_CLIP = 0.5  # MeshRasterizer: z_clip_value = znear / 2 for perspective cameras


def clip_faces(fv, z_clip=Z_CLIP):
    """pytorch3d ``clip_faces`` (renderer/mesh/clip.py, the frustum MeshRasterizer builds) on face vertices
    ``fv`` (F,3,3) = (x_ndc, y_ndc, z_view), DIFFERENTIABLY: the cut vertices p4 / p5 are functions of the face's own
    vertices (interpolation in view space, re-projection), so autograd carries the gradient of a hit on a cut
    triangle back to the uncut vertices -- what pytorch3d's autograd does through its torch ops.  Same cases and vertex
    orders as ``oracle/oracle.py::clip_faces``.  Returns (clipped fv (Fc,3,3), neighbour index (Fc,) long, -1 = none)."""
    behind = fv[:, :, 2].detach() < z_clip
    n = behind.sum(dim=1)
    out, nb, key = [], [], []

    def cut(p1, q):
        w = (p1[2] - z_clip) / (p1[2] - q[2])
        x = ((p1[0] * p1[2]) * (1.0 - w) + (q[0] * q[2]) * w) / z_clip
        y = ((p1[1] * p1[2]) * (1.0 - w) + (q[1] * q[2]) * w) / z_clip
        return torch.stack([x, y, torch.as_tensor(z_clip, dtype=fv.dtype)])

    keep = (n == 0).nonzero()[:, 0]
    out.append(fv[keep])
    key.append(2 * keep)
    mates = []  # (key of a triangle, key of its neighbour)
    for f in ((n == 1) | (n == 2)).nonzero()[:, 0].tolist():
        v, b = fv[f], behind[f]
        i = int(torch.argmax(b.to(torch.int8))) if int(n[f]) == 1 else int(torch.argmax((~b).to(torch.int8)))
        j, k = (i + 1) % 3, (i + 2) % 3
        p1, p2, p3 = v[i], v[j], v[k]
        p4, p5 = cut(p1, p2), cut(p1, p3)
        if int(n[f]) == 1:
            out.append(torch.stack([torch.stack([p4, p2, p5]), torch.stack([p5, p2, p3])]))
            key.append(torch.tensor([2 * f, 2 * f + 1]))
            mates += [(2 * f, 2 * f + 1), (2 * f + 1, 2 * f)]
        else:
            out.append(torch.stack([p1, p4, p5])[None])
            key.append(torch.tensor([2 * f]))
    cv, key = torch.cat(out, dim=0), torch.cat(key)
    order = torch.argsort(key)            # the reference's order: faces in mesh order, t1 before t2
    cv, key = cv[order], key[order]
    where = {int(k): i for i, k in enumerate(key.tolist())}
    nbv = torch.full((len(key),), -1, dtype=torch.long)
    for a, b in mates:
        nbv[where[a]] = where[b]
    return cv, nbv


def frozen_hit_masks(scene, S, action, el0, az0, radius, blur, K):
    """Per object, the (S*S, Fc) boolean matrix "clipped face c is one of the kept hits of pixel p" as the fp32 oracle
    decides it (culls, hit tests, the neighbour rule of cut quadrilaterals, the nearest-K cut).  Where a decision
    is a tie that only rounding breaks -- a pixel inside one triangle of a cut quadrilateral whose nearest edge is the
    shared diagonal sees the SAME distance from both triangles -- float64 would decide differently from the fp32
    reference; the gradient oracle therefore differentiates through the reference's own discrete decisions, which
    is also what autograd does in pytorch3d."""
    from . import oracle as O
    _, _, _, R, T = O.pose_step(np.asarray(action, np.float32), np.float32(el0), np.float32(az0), np.float32(radius))
    vproj = O.project(scene.verts, R, T)
    masks = []
    for i in range(scene.n_obj):
        v0, v1 = int(scene.obj_vert_start[i]), int(scene.obj_vert_start[i + 1])
        f0, f1 = int(scene.obj_face_start[i]), int(scene.obj_face_start[i + 1])
        fv = vproj[v0:v1][scene.faces[f0:f1] - v0]
        if (fv[:, :, 2] < O.Z_CLIP).any():
            cv, _, nb, _ = O.clip_faces(fv)
            fr = O.rasterize_fv(cv, nb, S, blur, K)
        else:
            cv = fv
            fr = O.rasterize_fv(fv, None, S, blur, K)
        m = np.zeros((S * S, len(cv)), bool)
        p2f = fr.pix_to_face.reshape(S * S, -1)
        pix, k = np.nonzero(p2f >= 0)
        m[pix, p2f[pix, k]] = True
        masks.append(torch.from_numpy(m))
    return masks


def soft_alpha_rows(vproj, faces, S, rows, blur, sigma, K, cull=True, hit_mask=None):
    """Soft silhouette alpha for the pixel rows ``rows`` (1-D long tensor): (len(rows), S).  ``hit_mask`` (S*S, Fc):
    the kept hits are given (``frozen_hit_masks``) instead of being decided here."""
    dtype = vproj.dtype
    fv = vproj[faces]  # (F,3,3)
    nb = None
    if bool((fv[:, :, 2].detach() < Z_CLIP).any()):
        fv, nb = clip_faces(fv)
    x0, y0, z0 = fv[:, 0, 0], fv[:, 0, 1], fv[:, 0, 2]
    x1, y1, z1 = fv[:, 1, 0], fv[:, 1, 1], fv[:, 1, 2]
    x2, y2, z2 = fv[:, 2, 0], fv[:, 2, 1], fv[:, 2, 2]
    face_area = _edge(x0, y0, x1, y1, x2, y2)
    zmin = torch.minimum(torch.minimum(z0, z1), z2)
    valid = (face_area.abs() > K_EPS) & ~(zmin < K_EPS)
    if cull:
        valid &= ~(face_area < 0)
    r = math.sqrt(blur)
    xmin = torch.minimum(torch.minimum(x0, x1), x2) - r
    xmax = torch.maximum(torch.maximum(x0, x1), x2) + r
    ymin = torch.minimum(torch.minimum(y0, y1), y2) - r
    ymax = torch.maximum(torch.maximum(y0, y1), y2) + r
    c = pixel_centers(S, dtype)
    py = c[rows]
    # drop faces that cannot touch these rows (pure speed-up, no semantic effect)
    keep = valid & (ymax.detach() >= py.min()) & (ymin.detach() <= py.max())
    idx = keep.nonzero()[:, 0]
    if idx.numel() == 0:
        return torch.zeros(len(rows), S, dtype=dtype)
    sel = lambda a: a[idx][None, :]  # noqa: E731
    x0, y0, z0, x1, y1, z1, x2, y2, z2 = map(sel, (x0, y0, z0, x1, y1, z1, x2, y2, z2))
    xmin, xmax, ymin, ymax = map(sel, (xmin, xmax, ymin, ymax))
    PX = c[None, :].expand(len(rows), S).reshape(-1, 1)
    PY = py[:, None].expand(len(rows), S).reshape(-1, 1)
    in_box = ~((PX > xmax) | (PX < xmin) | (PY > ymax) | (PY < ymin))
    area = _edge(x2, y2, x0, y0, x1, y1) + K_EPS
    w0 = _edge(PX, PY, x1, y1, x2, y2) / area
    w1 = _edge(PX, PY, x2, y2, x0, y0) / area
    w2 = _edge(PX, PY, x0, y0, x1, y1) / area
    t0, t1, t2 = w0 * z1 * z2, z0 * w1 * z2, z0 * z1 * w2
    den = (t0 + t1 + t2).clamp_min(K_EPS)
    b0, b1, b2 = t0 / den, t1 / den, t2 / den
    inside = (b0 > 0) & (b1 > 0) & (b2 > 0)
    c0, c1, c2 = b0.clamp_min(0), b1.clamp_min(0), b2.clamp_min(0)
    sm = (c0 + c1 + c2).clamp_min(1e-5)
    pz = (c0 * z0 + c1 * z1 + c2 * z2) / sm
    dist = torch.minimum(torch.minimum(_seg(PX, PY, x0, y0, x1, y1), _seg(PX, PY, x0, y0, x2, y2)),
                         _seg(PX, PY, x1, y1, x2, y2))
    hit = in_box & (pz >= 0) & (inside | (dist < blur))
    if hit_mask is not None:
        pix = (rows[:, None] * S + torch.arange(S)[None, :]).reshape(-1)
        hit = hit_mask[pix][:, idx]
        nb = None
    if nb is not None:
        # clipped_faces_neighbor_idx: of the two triangles of a cut quadrilateral hitting one pixel only the one with
        # the smaller distance is kept (the first on ties), as the face loop of the reference does
        pos = torch.full((nb.numel(),), -1, dtype=torch.long)
        pos[idx] = torch.arange(idx.numel())
        mate = torch.where(nb[idx] >= 0, pos[nb[idx].clamp_min(0)], torch.full_like(idx, -1))
        has = mate >= 0
        if bool(has.any()):
            m = mate.clamp_min(0)
            d_self, d_mate = dist.detach(), dist.detach()[:, m]
            first = (torch.arange(idx.numel()) < m)[None, :]
            lose = hit[:, m] & has[None, :] & torch.where(first, d_mate < d_self, d_mate <= d_self)
            hit = hit & ~lose
    nh = hit.sum(dim=1)
    if hit_mask is None and int(nh.max()) > K:
        # keep the K nearest by (pz, face index): stable sort on pz keeps the lower index first on ties
        key = torch.where(hit, pz.detach(), torch.full_like(pz, float("inf")))
        order = torch.sort(key, dim=1, stable=True).indices
        rank = torch.empty_like(order)
        rank.scatter_(1, order, torch.arange(order.shape[1]).expand_as(order))
        hit = hit & (rank < K)
    sd = torch.where(inside, -dist, dist)
    prob = torch.sigmoid(-sd / sigma) * hit.to(dtype)
    alpha = 1.0 - torch.prod(1.0 - prob, dim=1)
    return alpha.reshape(len(rows), S)


def occlusion_loss(verts, faces, obj_face_start, obj_vert_start, S, R, T, s, blur, sigma, K, row_chunk=16,
                   backward=False, hit_masks=None):
    """loss = sum_px (sum_{i<j} A_i A_j)^2 (``environment.py:373,381``), evaluated in row chunks.  With
    ``backward=True`` every chunk is back-propagated immediately (the graph of one chunk is alive at a
    time) and the float value is returned; otherwise a differentiable tensor is returned."""
    n_obj = len(obj_face_start) - 1
    total = torch.zeros((), dtype=verts.dtype)
    total_val = 0.0
    alphas = []
    for r0 in range(0, S, row_chunk):
        rows = torch.arange(r0, min(S, r0 + row_chunk))
        vproj = project(verts, R, T, s)
        A = []
        for i in range(n_obj):
            v0, v1 = int(obj_vert_start[i]), int(obj_vert_start[i + 1])
            f0, f1 = int(obj_face_start[i]), int(obj_face_start[i + 1])
            A.append(soft_alpha_rows(vproj[v0:v1], faces[f0:f1] - v0, S, rows, blur, sigma, K,
                                     hit_mask=None if hit_masks is None else hit_masks[i]))
        occl = torch.zeros_like(A[0])
        for i in range(n_obj):
            for j in range(i + 1, n_obj):
                occl = occl + A[i] * A[j]
        part = (occl ** 2).sum()
        alphas.append(torch.stack([a.detach() for a in A]))
        if backward:
            if part.requires_grad:
                part.backward(retain_graph=True)
            total_val += float(part.detach())
        else:
            total = total + part
    alphas = torch.cat(alphas, dim=1)
    return (total_val if backward else total), alphas


def reward_and_grad(scene, S, action, el0, az0, radius, prev_loss, mass, s, blur, sigma, K=100, step_size=0.05,
                    dtype=torch.float64, freeze_hits=False):
    """d reward / d action for one env by autograd (``environment.py:352-392`` in one differentiable
    graph).  Returns (reward, loss, grad_action (2,), alphas).  ``freeze_hits``: take the discrete decisions from the
    fp32 oracle (needed when faces are cut at z_clip, see ``frozen_hit_masks``)."""
    verts = torch.tensor(scene.verts, dtype=dtype)
    faces = torch.tensor(scene.faces, dtype=torch.long)
    a = torch.tensor(np.asarray(action), dtype=dtype, requires_grad=True)
    el, az, C, R, T = pose_step(a, torch.tensor(float(el0), dtype=dtype), torch.tensor(float(az0), dtype=dtype),
                                torch.tensor(float(radius), dtype=dtype), step_size)
    masks = frozen_hit_masks(scene, S, action, el0, az0, radius, blur, K) if freeze_hits else None
    loss_val, alphas = occlusion_loss(verts, faces, scene.obj_face_start, scene.obj_vert_start, S, R, T, s, blur,
                                      sigma, K, backward=True, hit_masks=masks)
    # reward = (prev - loss)/mass + const  ->  d reward / d a = -(d loss / d a) / mass
    g = a.grad if a.grad is not None else torch.zeros_like(a)
    reward = (prev_loss - loss_val) / mass
    return reward, loss_val, (-g / mass).numpy(), alphas.numpy()


def loss_and_pose_grad(scene, S, el0, az0, radius, s, blur, sigma, K=100, dtype=torch.float64):
    """loss and d loss / d (elevation, azimuth) at the step-convention pose (``environment.py:363-368``) by autograd."""
    verts = torch.tensor(scene.verts, dtype=dtype)
    faces = torch.tensor(scene.faces, dtype=torch.long)
    el = torch.tensor(float(el0), dtype=dtype, requires_grad=True)
    az = torch.tensor(float(az0), dtype=dtype, requires_grad=True)
    _, _, C, R, T = pose_step(torch.zeros(2, dtype=dtype), el, az, torch.tensor(float(radius), dtype=dtype))
    loss_val, _ = occlusion_loss(verts, faces, scene.obj_face_start, scene.obj_vert_start, S, R, T, s, blur, sigma, K,
                                 backward=True)
    g = [float(t.grad) if t.grad is not None else 0.0 for t in (el, az)]
    return loss_val, np.asarray(g, np.float64)
