/*
 * occl_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked or imported by the product path).
 *
 * CPU restatement, in strict fp32 (compile with -ffp-contract=off, no fast-math), of the arithmetic
 * that the reference's environment transition reaches through pytorch3d:
 *   reference call sites:  /root/reference/environment.py:234-284 (renderer constants),
 *                          :308-324 (reset render), :363-392 (step render + reward)
 *   third-party algorithm: pytorch3d (pinned ==0.6.2 in requirements.txt:50, =0.7.0 in
 *                          conda_environment.yml:61; NOT vendored, NOT installed here), restated from
 *                          its published sources as summarised in SURVEY.md Appendix A:
 *                            csrc/rasterize_meshes/rasterize_meshes_cpu.cpp::RasterizeMeshesNaiveCpu
 *                            csrc/utils/geometry_utils.h (edge function, barycentrics, perspective
 *                              correction, clip, point-segment distance)
 *                            renderer/cameras.py::look_at_rotation / look_at_view_transform
 *                            renderer/blending.py::sigmoid_alpha_blend, hard_rgb_blend
 *                            renderer/mesh/shading.py::flat_shading, renderer/lighting.py
 *
 * PARITY UNPINNED: the reference ships no tests / golden vectors and pytorch3d cannot be installed
 * offline, so this file is pinned only by (i) SURVEY.md Appendix A, (ii) an independently written
 * dense PyTorch formulation (oracle/dense_torch.py) and (iii) the survey's float64 probe values.
 *
 * Every function is a scalar loop in the same operation order as the pytorch3d CPU path.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* pytorch3d: `const auto kEpsilon = 1e-8;` -> a double constant. */
static const double kEpsilon = 1e-8;

/* ------------------------------------------------------------------------------------------ */
/* A.1 cameras                                                                                  */
/* ------------------------------------------------------------------------------------------ */

/* torch.nn.functional.normalize as called by pytorch3d renderer/cameras.py::look_at_rotation (eps=1e-5)
 * and renderer/lighting.py::diffuse/specular (eps=1e-6). */
static void normalize3(const float v[3], float eps, float out[3]) {
  /* torch.nn.functional.normalize: v / max(||v||_2, eps) */
  float n = sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
  float d = n > eps ? n : eps;
  out[0] = v[0] / d;
  out[1] = v[1] / d;
  out[2] = v[2] / d;
}

static void cross3(const float a[3], const float b[3], float o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

/* pytorch3d renderer/cameras.py::look_at_rotation, called at /root/reference/environment.py:334,367 and
 * (through look_at_view_transform) :308; followed by T = -bmm(R^T, C) of environment.py:335,368.
 * look_at_rotation(C, at=0, up=(0,1,0)) followed by T = -R^T C   (environment.py:367-368).
 * R is row-major 3x3 with the axes as COLUMNS (R[i][0]=x[i], R[i][1]=y[i], R[i][2]=z[i]). */
void occl_oracle_look_at(const float C[3], float R[9], float T[3]) {
  const float up[3] = {0.f, 1.f, 0.f};
  float mC[3] = {0.f - C[0], 0.f - C[1], 0.f - C[2]};
  float x[3], y[3], z[3], t[3];
  normalize3(mC, 1e-5f, z);
  cross3(up, z, t);
  normalize3(t, 1e-5f, x);
  cross3(z, x, t);
  normalize3(t, 1e-5f, y);
  if (fabsf(x[0]) <= 5e-3f && fabsf(x[1]) <= 5e-3f && fabsf(x[2]) <= 5e-3f) {
    cross3(y, z, t);
    normalize3(t, 1e-5f, x);
  }
  for (int i = 0; i < 3; ++i) {
    R[i * 3 + 0] = x[i];
    R[i * 3 + 1] = y[i];
    R[i * 3 + 2] = z[i];
  }
  T[0] = -((x[0] * C[0] + x[1] * C[1]) + x[2] * C[2]);
  T[1] = -((y[0] * C[0] + y[1] * C[1]) + y[2] * C[2]);
  T[2] = -((z[0] * C[0] + z[1] * C[1]) + z[2] * C[2]);
}

/* ------------------------------------------------------------------------------------------ */
/* Switches for the restatement's DISCRETIONARY choices: places where pytorch3d gives no bit-level definition     */
/* (or where its definition cannot run on both CPU and GPU bit-identically).  The default of every switch is what */
/* the CUDA kernels implement; tools/pin_with_pytorch3d.py flips them one at a time on a machine that has the real */
/* library to tell WHICH choice differs from it (DESIGN.md section 3 lists them).                                  */
/* ------------------------------------------------------------------------------------------ */
enum {
  OCCL_OPT_TRIG_FP32 = 0,        /* 0: sin/cos in double, rounded once (default) ; 1: fp32 sinf/cosf like torch.sin on CPU */
  OCCL_OPT_PROJ_MATRIX = 1,      /* 0: x_ndc = (s x_view) / z_view with ((xR00 + yR10) + zR20) + T0 (default) ;
                                    1: one composed 4x4 (world->view->ndc) applied to (x, y, z, 1), then / w, as
                                       Transform3d.compose(...).transform_points does (row-vector bmm, k ascending) */
  OCCL_OPT_NEIGHBOR_TOPK = 2,    /* 0: a cut quadrilateral's triangles are matched among ALL hits of the pixel (default) ;
                                    1: only among the current <= K nearest, as pytorch3d's per-pixel queue does */
  OCCL_OPT_COUNT = 3
};
static int g_opt[OCCL_OPT_COUNT] = {0, 0, 0};
int occl_oracle_set_option(int key, int value) {
  if (key < 0 || key >= OCCL_OPT_COUNT) return -1;
  g_opt[key] = value;
  return 0;
}
int occl_oracle_get_option(int key) { return key >= 0 && key < OCCL_OPT_COUNT ? g_opt[key] : -1; }

/* Trigonometry is evaluated in double and rounded once to fp32 so that the CPU restatement and
 * the device path agree bit-for-bit (fp32 sinf/cosf differ between libm and CUDA by ulps). */
static float sin32(float a) { return g_opt[OCCL_OPT_TRIG_FP32] ? sinf(a) : (float)sin((double)a); }
static float cos32(float a) { return g_opt[OCCL_OPT_TRIG_FP32] ? cosf(a) : (float)cos((double)a); }

/* environment.py:356-365: normalise action, integrate the angles, step-convention camera centre. */
void occl_oracle_pose_step(const float action[2], float step_size, float radius, float* el,
                           float* az, float C[3], float R[9], float T[3]) {
  float n = sqrtf(action[0] * action[0] + action[1] * action[1]);
  float a0 = action[0], a1 = action[1];
  if (n != 0.f) {
    a0 = a0 / n;
    a1 = a1 / n;
  }
  *el = *el + a0 * step_size;
  *az = *az + a1 * step_size;
  float rs = radius * sin32(*az);
  C[0] = rs * cos32(*el);
  C[1] = rs * sin32(*el);
  C[2] = radius * cos32(*az);
  occl_oracle_look_at(C, R, T);
}

/* environment.py:308: look_at_view_transform(radius, elevation, azimuth, degrees=False). */
void occl_oracle_pose_lookat(float dist, float elev, float azim, float C[3], float R[9],
                             float T[3]) {
  float dc = dist * cos32(elev);
  C[0] = dc * sin32(azim);
  C[1] = dist * sin32(elev);
  C[2] = dc * cos32(azim);
  occl_oracle_look_at(C, R, T);
}

/* pytorch3d renderer/mesh/rasterizer.py::MeshRasterizer.transform with FoVPerspectiveCameras defaults
 * (environment.py:238): X_view = X_world R + T; ndc.xy = s * view.xy / view.z; ndc.z := view.z. */
void occl_oracle_project(const float* verts, int V, const float R[9], const float T[3], float s,
                         float* out) {
  if (g_opt[OCCL_OPT_PROJ_MATRIX]) {
    /* M = W2V * P (row vectors): W2V = [[R, 0], [T, 1]], P columns x' = s x, y' = s y, w = z; z is read from the
     * world-to-view transform alone (MeshRasterizer.transform overwrites ndc z with view z). */
    float M[4][3]; /* columns: x', y', w */
    for (int r = 0; r < 3; ++r) {
      M[r][0] = R[r * 3 + 0] * s;
      M[r][1] = R[r * 3 + 1] * s;
      M[r][2] = R[r * 3 + 2];
    }
    M[3][0] = T[0] * s; M[3][1] = T[1] * s; M[3][2] = T[2];
    for (int v = 0; v < V; ++v) {
      const float x = verts[v * 3 + 0], y = verts[v * 3 + 1], z = verts[v * 3 + 2];
      const float xp = ((x * M[0][0] + y * M[1][0]) + z * M[2][0]) + M[3][0];
      const float yp = ((x * M[0][1] + y * M[1][1]) + z * M[2][1]) + M[3][1];
      const float w = ((x * M[0][2] + y * M[1][2]) + z * M[2][2]) + M[3][2];
      out[v * 3 + 0] = xp / w;
      out[v * 3 + 1] = yp / w;
      out[v * 3 + 2] = ((x * R[2] + y * R[5]) + z * R[8]) + T[2];
    }
    return;
  }
  for (int v = 0; v < V; ++v) {
    const float x = verts[v * 3 + 0], y = verts[v * 3 + 1], z = verts[v * 3 + 2];
    float xv = ((x * R[0] + y * R[3]) + z * R[6]) + T[0];
    float yv = ((x * R[1] + y * R[4]) + z * R[7]) + T[1];
    float zv = ((x * R[2] + y * R[5]) + z * R[8]) + T[2];
    out[v * 3 + 0] = (s * xv) / zv;
    out[v * 3 + 1] = (s * yv) / zv;
    out[v * 3 + 2] = zv;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* A.3 / A.4 naive rasteriser                                                                   */
/* ------------------------------------------------------------------------------------------ */

/* pytorch3d csrc/rasterize_meshes/rasterization_utils.h::PixToNonSquareNdc (square image). */
static float pix_to_ndc(int i, int S) {
  /* PixToNonSquareNdc for square images: -offset + (range*i + offset)/S, range=2, offset=1 */
  return -1.0f + (2.0f * (float)i + 1.0f) / (float)S;
}

/* pytorch3d csrc/utils/geometry_utils.h::EdgeFunctionForward(p, a, b). */
static float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
  return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}

/* pytorch3d csrc/utils/geometry_utils.h::PointLineDistanceForward(p, a, b) (squared distance to the segment). */
static float point_segment_dist(float px, float py, float ax, float ay, float bx, float by,
                                float* t_out) {
  const float bax = bx - ax, bay = by - ay;
  const float l2 = bax * bax + bay * bay;
  if ((double)l2 <= kEpsilon) {
    const float dx = px - bx, dy = py - by;
    if (t_out) *t_out = 1.0f;
    return dx * dx + dy * dy;
  }
  const float t = (bax * (px - ax) + bay * (py - ay)) / l2;
  const float tt = fminf(fmaxf(t, 0.0f), 1.0f);
  const float qx = ax + tt * bax, qy = ay + tt * bay;
  const float dx = px - qx, dy = py - qy;
  if (t_out) *t_out = tt;
  return dx * dx + dy * dy;
}

typedef struct {
  float pz;
  int f;
  float dist;
  float b0, b1, b2;
} hit_t;

static int hit_cmp(const void* pa, const void* pb) {
  /* std::tuple<float,int,float,float,float,float> lexicographic order */
  const hit_t* a = (const hit_t*)pa;
  const hit_t* b = (const hit_t*)pb;
  if (a->pz != b->pz) return a->pz < b->pz ? -1 : 1;
  if (a->f != b->f) return a->f < b->f ? -1 : 1;
  if (a->dist != b->dist) return a->dist < b->dist ? -1 : 1;
  if (a->b0 != b->b0) return a->b0 < b->b0 ? -1 : 1;
  if (a->b1 != b->b1) return a->b1 < b->b1 ? -1 : 1;
  if (a->b2 != b->b2) return a->b2 < b->b2 ? -1 : 1;
  return 0;
}

/* Per-face quantities the reference precomputes once per call (ComputeFaceAreas /
 * ComputeFaceBoundingBoxes) plus the pixel-independent skip tests of the A.4 rule. */
typedef struct {
  float v[9];   /* v0 v1 v2 (x_ndc, y_ndc, z_view) */
  float xmin, xmax, ymin, ymax; /* bbox expanded by sqrt(blur_radius) */
  int skip;     /* zmax<0 | back face (when culling) | |area|<=eps | zmin<eps */
} face_t;

static void setup_face(const float* fv, float bbox_r, int cull, face_t* o) {
  memcpy(o->v, fv, sizeof(float) * 9);
  const float x0 = fv[0], y0 = fv[1], z0 = fv[2];
  const float x1 = fv[3], y1 = fv[4], z1 = fv[5];
  const float x2 = fv[6], y2 = fv[7], z2 = fv[8];
  const float face_area = edge_fn(x0, y0, x1, y1, x2, y2); /* EdgeFunctionForward(v0, v1, v2) */
  const float zmax = fmaxf(fmaxf(z0, z1), z2);
  const float zmin = fminf(fminf(z0, z1), z2);
  o->xmin = fminf(fminf(x0, x1), x2) - bbox_r;
  o->xmax = fmaxf(fmaxf(x0, x1), x2) + bbox_r;
  o->ymin = fminf(fminf(y0, y1), y2) - bbox_r;
  o->ymax = fmaxf(fmaxf(y0, y1), y2) + bbox_r;
  int skip = 0;
  if (zmax < 0.f) skip = 1;
  if (cull && face_area < 0.f) skip = 1;
  if ((double)face_area <= kEpsilon && (double)face_area >= -1.0 * kEpsilon) skip = 1;
  if ((double)zmin < kEpsilon) skip = 1; /* z_invalid inside CheckPointOutsideBoundingBox */
  o->skip = skip;
}

/* One evaluation of the A.4 rule = the body of the face loop of RasterizeMeshesNaiveCpu
 * (csrc/rasterize_meshes/rasterize_meshes_cpu.cpp) / CheckPixelInsideFace (rasterize_meshes.cu), with
 * BarycentricCoordinatesForward, BarycentricPerspectiveCorrectionForward, BarycentricClipForward and
 * PointTriangleDistanceForward of geometry_utils.h inlined.  Settings: environment.py:249-255 (K=100, blur,
 * cull_backfaces) and :267-273 (K=1, blur 0).  Returns 1 and fills *h when (pixel, face) is a hit. */
static int eval_pixel_face(const face_t* fc, float px, float py, float blur_radius, int persp,
                           int clip_bary, hit_t* h) {
  if (fc->skip) return 0;
  if (px > fc->xmax || px < fc->xmin || py > fc->ymax || py < fc->ymin) return 0;
  const float* fv = fc->v;
  const float x0 = fv[0], y0 = fv[1], z0 = fv[2];
  const float x1 = fv[3], y1 = fv[4], z1 = fv[5];
  const float x2 = fv[6], y2 = fv[7], z2 = fv[8];

  /* BarycentricCoordsForward */
  const float area = (float)((double)edge_fn(x2, y2, x0, y0, x1, y1) + kEpsilon);
  const float w0 = edge_fn(px, py, x1, y1, x2, y2) / area;
  const float w1 = edge_fn(px, py, x2, y2, x0, y0) / area;
  const float w2 = edge_fn(px, py, x0, y0, x1, y1) / area;
  float b0 = w0, b1 = w1, b2 = w2;
  if (persp) {
    const float t0 = w0 * z1 * z2;
    const float t1 = z0 * w1 * z2;
    const float t2 = z0 * z1 * w2;
    const float den = fmaxf(t0 + t1 + t2, (float)kEpsilon);
    b0 = t0 / den;
    b1 = t1 / den;
    b2 = t2 / den;
  }
  float c0 = b0, c1 = b1, c2 = b2;
  if (clip_bary) {
    c0 = b0 > 0.f ? b0 : 0.f;
    c1 = b1 > 0.f ? b1 : 0.f;
    c2 = b2 > 0.f ? b2 : 0.f;
    const float s = fmaxf(c0 + c1 + c2, 1e-5f);
    c0 = c0 / s;
    c1 = c1 / s;
    c2 = c2 / s;
  }
  const float pz = c0 * z0 + c1 * z1 + c2 * z2;
  if (pz < 0.f) return 0;
  const float d01 = point_segment_dist(px, py, x0, y0, x1, y1, NULL);
  const float d02 = point_segment_dist(px, py, x0, y0, x2, y2, NULL);
  const float d12 = point_segment_dist(px, py, x1, y1, x2, y2, NULL);
  const float dist = fminf(fminf(d01, d02), d12);
  const int inside = b0 > 0.f && b1 > 0.f && b2 > 0.f;
  if (!inside && dist >= blur_radius) return 0;
  h->pz = pz;
  h->dist = inside ? -dist : dist;
  h->b0 = c0;
  h->b1 = c1;
  h->b2 = c2;
  return 1;
}

/* RasterizeMeshesNaiveCpu for one mesh, on explicit face vertices: face_verts (F,3,3) = (x_ndc, y_ndc, z_view).
 * neighbor (F) or NULL: clipped_faces_neighbor_idx of pytorch3d renderer/mesh/clip.py -- the two triangles a
 * face with one vertex nearer than z_clip is cut into name each other; when both hit a pixel only the one with
 * the smaller |dist| is kept (the later face replaces the earlier only if strictly closer), as in the face loop
 * of rasterize_meshes_cpu.cpp.  Outputs are (S,S,K) [bary (S,S,K,3)], -1 filled.  nhits (S,S), optional: hits
 * before the K cut. */
void occl_oracle_rasterize_fv(const float* face_verts, const int32_t* neighbor, int F, int S,
                              float blur_radius, int K, int persp, int clip_bary, int cull,
                              int32_t* pix_to_face, float* zbuf, float* bary, float* dists,
                              int32_t* nhits) {
  const size_t npix = (size_t)S * S;
  for (size_t i = 0; i < npix * K; ++i) {
    pix_to_face[i] = -1;
    zbuf[i] = -1.f;
    dists[i] = -1.f;
  }
  for (size_t i = 0; i < npix * K * 3; ++i) bary[i] = -1.f;

  const float bbox_r = sqrtf(blur_radius);
  face_t* fc = (face_t*)malloc(sizeof(face_t) * (size_t)(F > 0 ? F : 1));
  for (int f = 0; f < F; ++f) setup_face(face_verts + (size_t)f * 9, bbox_r, cull, &fc[f]);
  size_t cap = 4096;
  hit_t* q = (hit_t*)malloc(sizeof(hit_t) * cap);

  for (int yi = 0; yi < S; ++yi) {
    const float yf = pix_to_ndc(S - 1 - yi, S);
    for (int xi = 0; xi < S; ++xi) {
      const float xf = pix_to_ndc(S - 1 - xi, S);
      size_t n = 0, n_all = 0;
      for (int f = 0; f < F; ++f) {
        hit_t h;
        if (!eval_pixel_face(&fc[f], xf, yf, blur_radius, persp, clip_bary, &h)) continue;
        h.f = f;
        if (neighbor && neighbor[f] >= 0) {
          int found = -1;
          for (size_t i = 0; i < n; ++i)
            if (q[i].f == neighbor[f]) { found = (int)i; break; }
          if (found >= 0) {
            if (fabsf(h.dist) < fabsf(q[found].dist)) q[found] = h;
            continue;
          }
        }
        if (n == cap) {
          cap *= 2;
          q = (hit_t*)realloc(q, sizeof(hit_t) * cap);
        }
        q[n++] = h;
        ++n_all;
        if (g_opt[OCCL_OPT_NEIGHBOR_TOPK] && n > (size_t)K) {
          /* pytorch3d's per-pixel queue: never more than K entries -- the farthest is dropped at once, so a later
           * neighbour triangle is only matched against what is still among the K nearest */
          qsort(q, n, sizeof(hit_t), hit_cmp);
          n = (size_t)K;
        }
      }
      if (nhits) nhits[(size_t)yi * S + xi] = (int32_t)n_all;
      if (n == 0) continue;
      /* keeping the K smallest tuples == the reference's sort-and-pop_back deque */
      qsort(q, n, sizeof(hit_t), hit_cmp);
      const size_t m = n < (size_t)K ? n : (size_t)K;
      const size_t base = ((size_t)yi * S + xi) * K;
      for (size_t k = 0; k < m; ++k) {
        pix_to_face[base + k] = q[k].f;
        zbuf[base + k] = q[k].pz;
        dists[base + k] = q[k].dist;
        bary[(base + k) * 3 + 0] = q[k].b0;
        bary[(base + k) * 3 + 1] = q[k].b1;
        bary[(base + k) * 3 + 2] = q[k].b2;
      }
    }
  }
  free(q);
  free(fc);
}

/* RasterizeMeshesNaiveCpu for one mesh.  vproj: (V,3) = (x_ndc, y_ndc, z_view); faces (F,3). */
void occl_oracle_rasterize(const float* vproj, const int32_t* faces, int F, int S,
                           float blur_radius, int K, int persp, int clip_bary, int cull,
                           int32_t* pix_to_face, float* zbuf, float* bary, float* dists,
                           int32_t* nhits) {
  float* fv = (float*)malloc(sizeof(float) * 9 * (size_t)(F > 0 ? F : 1));
  for (int f = 0; f < F; ++f)
    for (int k = 0; k < 3; ++k)
      for (int c = 0; c < 3; ++c) fv[(size_t)f * 9 + k * 3 + c] = vproj[faces[f * 3 + k] * 3 + c];
  occl_oracle_rasterize_fv(fv, NULL, F, S, blur_radius, K, persp, clip_bary, cull, pix_to_face, zbuf, bary, dists, nhits);
  free(fv);
}

/* ------------------------------------------------------------------------------------------ */
/* A.5 soft silhouette                                                                          */
/* ------------------------------------------------------------------------------------------ */

/* pytorch3d renderer/blending.py::sigmoid_alpha_blend as used by SoftSilhouetteShader (environment.py:263,
 * BlendParams(sigma=1e-4) at :242): alpha = 1 - prod_k (1 - sigmoid(-d_k/sigma) * [f_k >= 0]). */
void occl_oracle_silhouette(const int32_t* pix_to_face, const float* dists, int S, int K,
                            float sigma, float* alpha) {
  const size_t npix = (size_t)S * S;
  for (size_t p = 0; p < npix; ++p) {
    float prod = 1.0f;
    for (int k = 0; k < K; ++k) {
      const size_t i = p * K + k;
      float prob = 0.f;
      if (pix_to_face[i] >= 0) {
        const float x = -dists[i] / sigma;
        prob = 1.0f / (1.0f + expf(-x));
      }
      prod = prod * (1.0f - prob);
    }
    alpha[p] = 1.0f - prod;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* A.6 hard flat shading + hard_rgb_blend + depth splice (environment.py:375-378)               */
/* ------------------------------------------------------------------------------------------ */

/* pytorch3d renderer/mesh/shading.py::flat_shading + renderer/lighting.py (PointLights at (2,2,-2),
 * environment.py:275) + renderer/blending.py::hard_rgb_blend + the depth splice of environment.py:376-378.
 * verts: world (V,3); faces: scene faces (F,3); pix_to_face/bary/zbuf from a K=1, blur=0 raster.
 * obs: (4,S,S) planar = RGB + depth (-1 on background). */
void occl_oracle_flat_shade(const float* verts, const int32_t* faces, const int32_t* pix_to_face,
                            const float* bary, const float* zbuf, int S, const float cam[3],
                            const float light[3], float* obs) {
  const size_t npix = (size_t)S * S;
  for (size_t p = 0; p < npix; ++p) {
    const int f = pix_to_face[p];
    if (f < 0) {
      obs[0 * npix + p] = 1.f;
      obs[1 * npix + p] = 1.f;
      obs[2 * npix + p] = 1.f;
      obs[3 * npix + p] = zbuf[p];
      continue;
    }
    const float* v0 = verts + 3 * faces[f * 3 + 0];
    const float* v1 = verts + 3 * faces[f * 3 + 1];
    const float* v2 = verts + 3 * faces[f * 3 + 2];
    float e1[3], e2[3], n[3], nn[3], ctr[3], dir[3], ndir[3], view[3], nview[3];
    for (int c = 0; c < 3; ++c) {
      e1[c] = v1[c] - v0[c];
      e2[c] = v2[c] - v0[c];
      ctr[c] = ((v0[c] + v1[c]) + v2[c]) / 3.0f;
    }
    cross3(e1, e2, n);
    normalize3(n, 1e-6f, nn); /* face_areas_normals: n / max(|n|, 1e-6) */
    normalize3(nn, 1e-6f, n); /* lighting.diffuse re-normalises          */
    for (int c = 0; c < 3; ++c) {
      dir[c] = light[c] - ctr[c];
      view[c] = cam[c] - ctr[c];
    }
    normalize3(dir, 1e-6f, ndir);
    normalize3(view, 1e-6f, nview);
    const float cosang = (n[0] * ndir[0] + n[1] * ndir[1]) + n[2] * ndir[2];
    const float diffuse = 0.3f * (cosang > 0.f ? cosang : 0.f);
    float refl[3];
    for (int c = 0; c < 3; ++c) refl[c] = -ndir[c] + 2.0f * (cosang * n[c]);
    float a = (nview[0] * refl[0] + nview[1] * refl[1]) + nview[2] * refl[2];
    a = (a > 0.f ? a : 0.f) * (cosang > 0.f ? 1.0f : 0.0f);
    const float spec = 0.2f * powf(a, 64.0f);
    const float texel = (bary[p * 3 + 0] + bary[p * 3 + 1]) + bary[p * 3 + 2];
    const float rgb = (0.5f + diffuse) * texel + spec;
    obs[0 * npix + p] = rgb;
    obs[1 * npix + p] = rgb;
    obs[2 * npix + p] = rgb;
    obs[3 * npix + p] = zbuf[p];
  }
}

/* ------------------------------------------------------------------------------------------ */
/* A.7 rasteriser backward on the silhouette route (grad of dists only)                         */
/* ------------------------------------------------------------------------------------------ */

/* pytorch3d csrc/rasterize_meshes/rasterize_meshes_cpu.cpp::RasterizeMeshesBackwardCpu restricted to the
 * grad_dists input (the route of reward.backward(), demo.py:85-86): PointTriangleDistanceBackward.
 * grad_dists (S,S,K) -> grad_vproj (V,3) (xy only; z receives nothing on this route).          */
void occl_oracle_rasterize_backward(const float* vproj, const int32_t* faces, int V, int S, int K,
                                    int persp, const int32_t* pix_to_face,
                                    const float* grad_dists, double* grad_vproj) {
  memset(grad_vproj, 0, sizeof(double) * 3 * (size_t)V);
  for (int yi = 0; yi < S; ++yi) {
    const float py = pix_to_ndc(S - 1 - yi, S);
    for (int xi = 0; xi < S; ++xi) {
      const float px = pix_to_ndc(S - 1 - xi, S);
      for (int k = 0; k < K; ++k) {
        const size_t i = ((size_t)yi * S + xi) * K + k;
        const int f = pix_to_face[i];
        if (f < 0) continue;
        const float g = grad_dists[i];
        if (g == 0.f) continue;
        const int i0 = faces[f * 3 + 0], i1 = faces[f * 3 + 1], i2 = faces[f * 3 + 2];
        const float x0 = vproj[i0 * 3], y0 = vproj[i0 * 3 + 1], z0 = vproj[i0 * 3 + 2];
        const float x1 = vproj[i1 * 3], y1 = vproj[i1 * 3 + 1], z1 = vproj[i1 * 3 + 2];
        const float x2 = vproj[i2 * 3], y2 = vproj[i2 * 3 + 1], z2 = vproj[i2 * 3 + 2];
        const float area = (float)((double)edge_fn(x2, y2, x0, y0, x1, y1) + kEpsilon);
        const float w0 = edge_fn(px, py, x1, y1, x2, y2) / area;
        const float w1 = edge_fn(px, py, x2, y2, x0, y0) / area;
        const float w2 = edge_fn(px, py, x0, y0, x1, y1) / area;
        float b0 = w0, b1 = w1, b2 = w2;
        if (persp) {
          const float t0 = w0 * z1 * z2, t1 = z0 * w1 * z2, t2 = z0 * z1 * w2;
          const float den = fmaxf(t0 + t1 + t2, (float)kEpsilon);
          b0 = t0 / den;
          b1 = t1 / den;
          b2 = t2 / den;
        }
        const int inside = b0 > 0.f && b1 > 0.f && b2 > 0.f;
        const float sign = inside ? -1.f : 1.f;
        float t01, t02, t12;
        const float d01 = point_segment_dist(px, py, x0, y0, x1, y1, &t01);
        const float d02 = point_segment_dist(px, py, x0, y0, x2, y2, &t02);
        const float d12 = point_segment_dist(px, py, x1, y1, x2, y2, &t12);
        int ia, ib;
        float ax, ay, bx, by, t;
        if (d01 <= d02 && d01 <= d12) {
          ia = i0; ib = i1; ax = x0; ay = y0; bx = x1; by = y1; t = t01;
        } else if (d02 <= d01 && d02 <= d12) {
          ia = i0; ib = i2; ax = x0; ay = y0; bx = x2; by = y2; t = t02;
        } else {
          ia = i1; ib = i2; ax = x1; ay = y1; bx = x2; by = y2; t = t12;
        }
        const double qx = (double)ax + (double)t * ((double)bx - ax);
        const double qy = (double)ay + (double)t * ((double)by - ay);
        const double gx = (double)sign * g * 2.0 * (qx - px);
        const double gy = (double)sign * g * 2.0 * (qy - py);
        grad_vproj[ia * 3 + 0] += (1.0 - t) * gx;
        grad_vproj[ia * 3 + 1] += (1.0 - t) * gy;
        grad_vproj[ib * 3 + 0] += t * gx;
        grad_vproj[ib * 3 + 1] += t * gy;
      }
    }
  }
}

int occl_oracle_version(void) { return 1; }
