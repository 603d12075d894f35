// occl_b200.cu -- hand-written sm_100a kernels + the C-ABI of include/occl_b200.h.
//
// Hot path of MILAB-IIT-CV/OcclusionEnv (environment.py:352-396 step, :302-328 reset render) for N
// environments at once.  Six launches per transition, all on the caller's stream:
//
//   pose_kernel        one thread per env: action -> (el, az) -> camera centre -> look-at R, T, with
//                      forward-mode tangents d/d_el, d/d_az carried as dual numbers       (north-star 1)
//   project_kernel     one thread per (env, vertex): world -> view -> NDC, z := view z    (north-star 1)
//   face_setup_kernel  one CTA per env: culls (back face, degenerate, z, z_clip), exact pixel ranges, per-face
//                      lighting, warp-ballot compaction of the live faces, mask of non-empty tiles  (2)
//   raster_kernel      one CTA per (env, image tile): warps stage the faces of the tile in shared memory and
//                      scatter soft-silhouette factors and the nearest-z key into per-pixel accumulators in
//                      shared memory (3,4); pixels with more than K hits get the nearest-K rule; the epilogue
//                      blends, shades, writes RGBD / occlusion map with coalesced stores and reduces loss,
//                      gradient and pixel counts with warp shuffles (5,6)
//   raster_clip_kernel one CTA per env, only for envs with faces cut at z_clip (clip_faces): the same tile
//                      rasteriser in its clip-capable instantiation
//   finalize_kernel    one thread per env: tile partials -> loss, reward, done, state, d reward/d action
//
// Bit-exactness: every expression that decides a hit, its sign, the nearest face or the K-nearest
// set is evaluated in the operation order of pytorch3d's rasteriser with
// IEEE fp32 add/mul/div/sqrt and NO fused multiply-add: this translation unit is compiled with
// -fmad=false (see occlusionenv_b200/build.py).  Nothing here falls back to a CPU path.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "occl_b200.h"

#ifndef OCCL_THREADS
#define OCCL_THREADS 256
#endif
#ifndef OCCL_TILE_W
#define OCCL_TILE_W 32         // compile-time tile of the fast instantiation (other shapes: generic kernel)
#endif
#ifndef OCCL_TILE_H
#define OCCL_TILE_H 32
#endif
#ifndef OCCL_TILE2_W
#define OCCL_TILE2_W 32        // second compile-time tile: half the pixels, for scenes with 3-4 objects or dense meshes
#endif                         // (smaller accumulators -> 4 CTAs/SM instead of 3, more tiles to balance: config 3 +37 %,
#ifndef OCCL_TILE2_H           //  config 2 -14 %)
#define OCCL_TILE2_H 16
#endif
#ifndef OCCL_TILE3_W
#define OCCL_TILE3_W 128       // third compile-time tile: dense meshes (>= OCCL_DENSE_FACES faces): wide and flat
#endif                         // (config 3: 128x4 13.1 k, 64x8 12.6 k, 32x16 11.8 k env-steps/s; four teapots: 303 k / 318 k / 333 k)
#ifndef OCCL_TILE3_H
#define OCCL_TILE3_H 4
#endif
#define OCCL_DENSE_FACES 16384 // default tile: the third one from this many faces on, the second one for 3-4 objects
#define OCCL_WARPS (OCCL_THREADS / 32)
#ifndef SETUP_THREADS
#define SETUP_THREADS 256     // face_setup_kernel: one CTA per env (1024 measured slower: 1 CTA per SM)
#endif
#define SETUP_WARPS (SETUP_THREADS / 32)
#ifndef SETUP_CTAS
#define SETUP_CTAS 4          // resident CTAs per SM the setup kernel is compiled for (64 registers, no spills: 32 warps hide its gather latency; 3 measured 1 % slower on C2)
#endif
#define WBUF_RECS 48          // 16-word records per warp of the tile's scratch region (24 KB per CTA): the round buffers of the main
                              // phase, the difference image of the evaluate-once pre-pass and the K-overflow hit buffers alias it
#ifndef OCCL_CTAS_FWD
#define OCCL_CTAS_FWD 4       // resident CTAs per SM the forward kernel is compiled for (64 registers)
#endif
#ifndef OCCL_CTAS_GRAD
#define OCCL_CTAS_GRAD 3      // ... and the differentiable kernel (80 registers)
#endif
#define OCCL_WS_BUDGET_MB 8192 // default cap of the per-face scratch (OcclConfig.ws_budget_mb = 0)
#define WDEFER_CAP 128        // per-warp queue of inside hits awaiting their exact depth
#ifndef ROUND_FACES_FWD
#define ROUND_FACES_FWD 256   // faces per round of the tile rasteriser's main phase (one per thread)
#endif
#define ROUND_FACES_GRAD 128  // ... of the differentiable kernel (its records also carry 48 B of vertex tangents)
#define BIG_CAP 16            // 16-word records of the small scratch region behind the depth queues (round tables of the K-overflow paths)
#define CAND_CAP (256 * OCCL_WARPS)  /* 2048 */         // candidate faces of ONE overflowing pixel (12 B each in the selection buffers)
#define TILE_MASK_WORDS 8     // per-env bitmask of non-empty tiles (up to 256 tiles; more: mask unused)
static_assert(OCCL_TILE_STATE_WORDS == 2 * TILE_MASK_WORDS, "OcclOutputs.obs_tile_state: last mask + skip mask per env");
#define OVF_CAP 128           // overflowing (pixel, object) pairs handled per round of the one-pixel fallback
#define HITBUF_CAP (272 * OCCL_WARPS)  /* 2176 */       // hits (12 B each) of one selection round: aliases the face list + depth queue
#define WQ_CAP 64             // (slot, face) pairs queued per warp for the dense evaluation of a round
#define RSLOT_CAP 128         // (pixel, object) slots per selection round (table lives in the small scratch region)
// "Evaluate every hit once" (compile-time tiles of at most REC_MAX_TPX pixels: the dense-mesh and many-object tiles):
// before the main phase the tile counts, per (pixel, object) slot, the faces whose blur box holds the pixel (an upper
// bound U of its hits: a 2-D difference image + prefix sums); slots with U > K get a segment of U entries in the CTA's
// slab of HBM scratch and a flag; the main phase then RECORDS (key, term) of every hit on a flagged slot, and the
// nearest-K rule only has to select.  Slots that did not get a segment fall back to the re-evaluating rounds.
#define REC_MAX_TPX 512
#define REC_SLAB_DENSE 262144      // (key u64, term f32) entries per resident CTA, dense meshes: 3 MB (a 128x4 tile fully
                                   // covered by two overlapping dense objects needs ~270 k: every slot then records)
#define REC_SLAB_SPARSE 32768      // ... other scenes: 384 KB
#define REC_SM_MAX 160             // slabs are indexed by (%smid, resident-CTA slot): B200 has 148 SMs
#define REC_CTAS_PER_SM 4
#define REC_WORDS 16
// observation / occlusion-map stores: written once, read by another kernel (or another GPU) much later
#ifdef OCCL_STREAMING_STORES
#define OCCL_STORE(ptr, v) __stcs((ptr), (v))
#define OCCL_STORE4(ptr, v) __stcs((float4*)(ptr), (v))
#else
#define OCCL_STORE(ptr, v) (*(ptr) = (v))
#define OCCL_STORE4(ptr, v) (*(float4*)(ptr) = (v))
#endif

static thread_local char g_last_err[256] = "";

// ----------------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------------
struct Partial {
  double loss;     // sum occl^2 over the tile
  double objsq;    // sum (sum_i A_i)^2
  double gl[2];    // d loss / d el, d loss / d az
  int ncov[OCCL_MAX_OBJ];
  int nvis[OCCL_MAX_OBJ];
};

// An opaque copy of a shared-memory base address.  ptxas, at a tight register cap, rebuilds every shared-window base from
// SR_CgaCtaId where it is used (S2R + MOV + IADD + LEA, inside the hottest loops); an empty asm makes the value
// unknown to it -- it then lives in a (uniform) register -- and the assume keeps LDS / STS / ATOMS instead of generic accesses.
#define OCCL_OPAQUE_SHARED(ptr)                 \
  do {                                          \
    asm volatile("" : "+l"(ptr));               \
    __builtin_assume(__isShared(ptr));          \
  } while (0)

__device__ __forceinline__ float pix_to_ndc(int i, int S) {
  // PixToNonSquareNdc for a square image: -1 + (2 i + 1) / S           
  return -1.0f + (2.0f * (float)i + 1.0f) / (float)S;
}

// Dual number: value + tangents w.r.t. (elevation, azimuth). The value part uses exactly the
// operations of the oracle / reference so that the pose is bit-identical.
struct Dual {
  float v, d0, d1;
};
__device__ __forceinline__ Dual mk(float v) { return Dual{v, 0.f, 0.f}; }
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return Dual{a.v + b.v, a.d0 + b.d0, a.d1 + b.d1}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return Dual{a.v - b.v, a.d0 - b.d0, a.d1 - b.d1}; }
__device__ __forceinline__ Dual operator-(Dual a) { return Dual{-a.v, -a.d0, -a.d1}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) {
  return Dual{a.v * b.v, a.d0 * b.v + a.v * b.d0, a.d1 * b.v + a.v * b.d1};
}
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
  const float q = a.v / b.v;
  return Dual{q, (a.d0 - q * b.d0) / b.v, (a.d1 - q * b.d1) / b.v};
}
__device__ __forceinline__ Dual dsqrt(Dual a) {
  const float s = sqrtf(a.v);
  const float h = s > 0.f ? 0.5f / s : 0.f;
  return Dual{s, a.d0 * h, a.d1 * h};
}
__device__ __forceinline__ void dnormalize(const Dual v[3], float eps, Dual o[3]) {
  // torch.nn.functional.normalize: v / max(||v||, eps)
  Dual n = dsqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
  Dual d = n.v > eps ? n : mk(eps);
  o[0] = v[0] / d;
  o[1] = v[1] / d;
  o[2] = v[2] / d;
}
__device__ __forceinline__ void dcross(const Dual a[3], const Dual b[3], Dual o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
// trig in double, rounded once to fp32 (bit-identical to the oracle's sin32 / cos32)
__device__ __forceinline__ void sincos32(float a, float* s, float* c) {
  double sd, cd;
  sincos((double)a, &sd, &cd);
  *s = (float)sd;
  *c = (float)cd;
}

// look_at_rotation + T = -R^T C  (pytorch3d renderer/cameras.py; environment.py:367-368)
__device__ void look_at_store(const Dual C[3], float* __restrict__ cam) {
  Dual up[3] = {mk(0.f), mk(1.f), mk(0.f)};
  Dual mC[3] = {mk(0.f) - C[0], mk(0.f) - C[1], mk(0.f) - C[2]};
  Dual x[3], y[3], z[3], t[3];
  dnormalize(mC, 1e-5f, z);
  dcross(up, z, t);
  dnormalize(t, 1e-5f, x);
  dcross(z, x, t);
  dnormalize(t, 1e-5f, y);
  if (fabsf(x[0].v) <= 5e-3f && fabsf(x[1].v) <= 5e-3f && fabsf(x[2].v) <= 5e-3f) {
    dcross(y, z, t);
    dnormalize(t, 1e-5f, x);
  }
  Dual T[3];
  T[0] = -((x[0] * C[0] + x[1] * C[1]) + x[2] * C[2]);
  T[1] = -((y[0] * C[0] + y[1] * C[1]) + y[2] * C[2]);
  T[2] = -((z[0] * C[0] + z[1] * C[1]) + z[2] * C[2]);
  // block 0: values, block 1: d/d_el, block 2: d/d_az ; each [R(9) T(3) C(3) pad]
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    cam[i * 3 + 0] = x[i].v;  cam[16 + i * 3 + 0] = x[i].d0;  cam[32 + i * 3 + 0] = x[i].d1;
    cam[i * 3 + 1] = y[i].v;  cam[16 + i * 3 + 1] = y[i].d0;  cam[32 + i * 3 + 1] = y[i].d1;
    cam[i * 3 + 2] = z[i].v;  cam[16 + i * 3 + 2] = z[i].d0;  cam[32 + i * 3 + 2] = z[i].d1;
    cam[9 + i] = T[i].v;      cam[16 + 9 + i] = T[i].d0;      cam[32 + 9 + i] = T[i].d1;
    cam[12 + i] = C[i].v;     cam[16 + 12 + i] = C[i].d0;     cam[32 + 12 + i] = C[i].d1;
  }
  cam[15] = 0.f; cam[31] = 0.f; cam[47] = 0.f;
}

// ----------------------------------------------------------------------------------------------
// kernel 1a: pose
// ----------------------------------------------------------------------------------------------
// mode 0: step (environment.py:356-365), mode 1: look_at_view_transform (environment.py:308)
__global__ void pose_kernel(int n, int mode, float step_size, const float* __restrict__ action,
                            float* __restrict__ el_p, float* __restrict__ az_p,
                            const float* __restrict__ radius_p, float* __restrict__ cam,
                            float* __restrict__ position, uint32_t* __restrict__ status,
                            const uint8_t* __restrict__ env_mask) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (env_mask && !env_mask[e]) return;
  if (status) status[e] = 0u;
  float el = el_p[e], az = az_p[e];
  const float r = radius_p[e];
  Dual C[3];
  if (mode == 0) {
    float a0 = action[2 * e + 0], a1 = action[2 * e + 1];
    const float nrm = sqrtf(a0 * a0 + a1 * a1);
    if (nrm != 0.f) {
      a0 = a0 / nrm;
      a1 = a1 / nrm;
    }
    el = el + a0 * step_size;
    az = az + a1 * step_size;
    el_p[e] = el;
    az_p[e] = az;
    float se, ce, sa, ca;
    sincos32(el, &se, &ce);
    sincos32(az, &sa, &ca);
    const Dual sin_az{sa, 0.f, ca}, cos_az{ca, 0.f, -sa}, sin_el{se, ce, 0.f}, cos_el{ce, -se, 0.f};
    const Dual rs = mk(r) * sin_az;
    C[0] = rs * cos_el;
    C[1] = rs * sin_el;
    C[2] = mk(r) * cos_az;
  } else {
    float se, ce, sa, ca;
    sincos32(el, &se, &ce);
    sincos32(az, &sa, &ca);
    const Dual sin_az{sa, 0.f, ca}, cos_az{ca, 0.f, -sa}, sin_el{se, ce, 0.f}, cos_el{ce, -se, 0.f};
    const Dual dc = mk(r) * cos_el;
    C[0] = dc * sin_az;
    C[1] = mk(r) * sin_el;
    C[2] = dc * cos_az;
  }
  look_at_store(C, cam + (size_t)e * OCCL_CAM_STRIDE);
  if (position) {
    position[3 * e + 0] = C[0].v;
    position[3 * e + 1] = C[1].v;
    position[3 * e + 2] = C[2].v;
  }
}

__global__ void pose_set_kernel(int n, const float* __restrict__ R, const float* __restrict__ T,
                                const float* __restrict__ C, float* __restrict__ cam) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float* c = cam + (size_t)e * OCCL_CAM_STRIDE;
  for (int i = 0; i < OCCL_CAM_STRIDE; ++i) c[i] = 0.f;
  for (int i = 0; i < 9; ++i) c[i] = R[9 * e + i];
  for (int i = 0; i < 3; ++i) {
    c[9 + i] = T[3 * e + i];
    c[12 + i] = C[3 * e + i];
  }
}

// ----------------------------------------------------------------------------------------------
// kernel 1b: projection
// ----------------------------------------------------------------------------------------------
template <bool GRAD>
__global__ void project_kernel(long long total, int V, const float* __restrict__ cam,
                               const float* __restrict__ verts, long long verts_stride, float s,
                               float z_clip, float4* __restrict__ vproj, float4* __restrict__ vtan,
                               uint32_t* __restrict__ status, const uint8_t* __restrict__ env_mask) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int e = (int)(idx / V);
  if (env_mask && !env_mask[e]) return;
  const int v = (int)(idx - (long long)e * V);
  if (status && v == 0) status[e] = 0u;  // a new transition of env e starts here (face_setup_kernel ORs its flags in later)
  const float* __restrict__ c = cam + (size_t)e * OCCL_CAM_STRIDE;
  const float* __restrict__ p = verts + (size_t)e * verts_stride + (size_t)v * 3;
  const float x = __ldg(p + 0), y = __ldg(p + 1), z = __ldg(p + 2);
  // X_view = X_world R + T   (row-vector convention, left-to-right sums)
  const float xv = ((x * __ldg(c + 0) + y * __ldg(c + 3)) + z * __ldg(c + 6)) + __ldg(c + 9);
  const float yv = ((x * __ldg(c + 1) + y * __ldg(c + 4)) + z * __ldg(c + 7)) + __ldg(c + 10);
  const float zv = ((x * __ldg(c + 2) + y * __ldg(c + 5)) + z * __ldg(c + 8)) + __ldg(c + 11);
  const float xn = (s * xv) / zv;
  const float yn = (s * yv) / zv;
  vproj[idx] = make_float4(xn, yn, zv, 0.f);
  if (GRAD) {
    float t[4];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float* __restrict__ d = c + 16 * (k + 1);
      const float dxv = x * __ldg(d + 0) + y * __ldg(d + 3) + z * __ldg(d + 6) + __ldg(d + 9);
      const float dyv = x * __ldg(d + 1) + y * __ldg(d + 4) + z * __ldg(d + 7) + __ldg(d + 10);
      const float dzv = x * __ldg(d + 2) + y * __ldg(d + 5) + z * __ldg(d + 8) + __ldg(d + 11);
      t[2 * k + 0] = s * (dxv - (xv / zv) * dzv) / zv;
      t[2 * k + 1] = s * (dyv - (yv / zv) * dzv) / zv;
    }
    vtan[idx] = make_float4(t[0], t[1], t[2], t[3]);
  }
}

// ----------------------------------------------------------------------------------------------
// kernels 2-5: face setup / binning and the tile rasteriser
// ----------------------------------------------------------------------------------------------
struct RasterParams {
  int S, n_obj, V, F, K, cull, exact_only;
  int obj_face_start[OCCL_MAX_OBJ + 1];
  int tile_w, tile_h, tiles_x, tiles_y;
  float blur, bbox_r, sigma, inv_sigma, inv_sigma_log2e, z_clip;
  float light[3];
  const float4* vproj;
  const float4* vtan;
  const float* verts;
  long long verts_stride;
  const int* faces;
  long long faces_stride;
  const float* cam;
  Partial* partials;
  // per-env compacted live faces (written by face_setup_kernel)
  uint4* geo;    // [N][F][4]  16 words per live face, see REC_* below
  uint4* rng;    // [N][F]     pixel ranges: soft x, soft y, hard x, hard y  (lo | hi << 16); object id in bits 30..31 of .w
  int* n_live;   // [N]
  const uint32_t* tile_mask;  // [N][TILE_MASK_WORDS]
  const uint32_t* obs_tile_state;  // [N][OCCL_TILE_STATE_WORDS] or null: words 8..15 = tiles whose obs need not be stored
  const float2* shade;        // [N][F]
  int* tile_idx;              // [N][n_tiles][tidx_cap] live-list indices of the faces that touch a tile
  int tidx_cap;
  const int* tile_cnt;        // [N][n_tiles] faces binned to the tile by the setup kernel (nullptr: more than 256 tiles, no binning)
  const int* clip_list;       // [1 + chunk] number of envs with faces cut at z_clip, then their (chunk-local) ids
  const int* env_list;        // [1 + chunk] masked transition: number of flagged envs, then their (chunk-local) ids
  // evaluate-once scratch (nullptr: not available): REC_SM_MAX * REC_CTAS_PER_SM slabs of rec_cap entries
  unsigned long long* rec_keys;
  float* rec_terms;
  unsigned* rec_table;        // [REC_SM_MAX] bitmask of the slab slots in use on each SM (zeroed before the launch)
  int rec_cap;                // entries per slab
  // outputs
  float* obs;
  int obs_planes;  // 4: R, G, B, depth planes (reference layout) ; 2: grey, depth (compact transport layout)
  float* occl;
  float* alphas;
  int* pix_to_face;
  float* bary;
  int* nhits;
  uint32_t* status;
  const uint8_t* env_mask;
};

// Record of a live face (16 words). Words 0..9 and 12..14 are floats.
//   0..8  x0 y0 z0 x1 y1 z1 x2 y2 z2     (x_ndc, y_ndc, z_view)
//   9     area = (float)((double)EdgeFunction(v2,v0,v1) + kEpsilon)
//   10    packed face index (bits 0..27) | object id (bits 28..29) | FAST flag (bit 31)
//   11    (tile list only) hard pixel range, tile-local, 8 bits each: x0 x1 y0 y1
//   12..14  1/l2 of the edges v0v1, v0v2, v1v2  (fast path only)
//   15    (tile list only) soft pixel range, tile-local, 8 bits each
#define REC_FAST 0x80000000u
#define REC_CLIP 0x40000000u   // the face straddles z_clip: words 0..8 hold the UNCUT vertices (see clip_subtris)
#define REC_FIDX_MASK 0x0fffffffu
#define RNG_CLIP 0x80000000u   // same flag in word .z of the range record
#define KEY_CLIP 0x80000000u   // nearest-face key, low word: the hit is on a cut face; bit 30 = second triangle of the cut
#define REC_OBJ_SHIFT 28
#ifndef GROUP_LANES
#define GROUP_LANES 8         // lanes that rasterise one small face; a warp works on 32/GROUP_LANES faces at once
#endif

struct FaceGeo {
  float x0, y0, z0, x1, y1, z1, x2, y2, z2;
  float area;  // (float)((double)EdgeFunction(v2, v0, v1) + kEpsilon)
};

struct PairResult {
  float b0, b1, b2;  // perspective-corrected, unclipped barycentrics
  float dist;        // squared NDC distance to the nearest edge
  float t;           // clamped parameter on the nearest edge
  int edge;          // 0: v0v1, 1: v0v2, 2: v1v2
  bool inside;
};

__device__ __forceinline__ float seg_dist(float px, float py, float ax, float ay, float bx, float by,
                                          float* tt_out) {
  // PointLineDistanceForward
  const float bax = bx - ax, bay = by - ay;
  const float l2 = bax * bax + bay * bay;
  if (l2 <= 1e-8f) {  // float form of (double)l2 <= 1e-8
    const float dx = px - bx, dy = py - by;
    *tt_out = 1.0f;
    return dx * dx + dy * dy;
  }
  const float t = (bax * (px - ax) + bay * (py - ay)) / l2;
  const float tt = fminf(fmaxf(t, 0.0f), 1.0f);
  const float qx = ax + tt * bax, qy = ay + tt * bay;
  const float dx = px - qx, dy = py - qy;
  *tt_out = tt;
  return dx * dx + dy * dy;
}

__device__ __forceinline__ void bary_persp(const FaceGeo& g, float px, float py, float* b0, float* b1,
                                           float* b2) {
  const float e0 = (px - g.x1) * (g.y2 - g.y1) - (py - g.y1) * (g.x2 - g.x1);
  const float e1 = (px - g.x2) * (g.y0 - g.y2) - (py - g.y2) * (g.x0 - g.x2);
  const float e2 = (px - g.x0) * (g.y1 - g.y0) - (py - g.y0) * (g.x1 - g.x0);
  const float w0 = e0 / g.area, w1 = e1 / g.area, w2 = e2 / g.area;
  const float t0 = w0 * g.z1 * g.z2;
  const float t1 = g.z0 * w1 * g.z2;
  const float t2 = g.z0 * g.z1 * w2;
  const float den = fmaxf(t0 + t1 + t2, 1e-8f);
  *b0 = t0 / den;
  *b1 = t1 / den;
  *b2 = t2 / den;
}

// The reference rule, operation for operation.
__device__ __noinline__ PairResult eval_pair(const FaceGeo& g, float px, float py) {
  PairResult r;
  bary_persp(g, px, py, &r.b0, &r.b1, &r.b2);
  r.inside = r.b0 > 0.f && r.b1 > 0.f && r.b2 > 0.f;
  float t01, t02, t12;
  const float d01 = seg_dist(px, py, g.x0, g.y0, g.x1, g.y1, &t01);
  const float d02 = seg_dist(px, py, g.x0, g.y0, g.x2, g.y2, &t02);
  const float d12 = seg_dist(px, py, g.x1, g.y1, g.x2, g.y2, &t12);
  r.dist = fminf(fminf(d01, d02), d12);
  if (d01 <= d02 && d01 <= d12) {
    r.edge = 0;
    r.t = t01;
  } else if (d02 <= d01 && d02 <= d12) {
    r.edge = 1;
    r.t = t02;
  } else {
    r.edge = 2;
    r.t = t12;
  }
  return r;
}

// raw SFU approximations (1-2 ulp): only on the tolerance path (silhouette probabilities)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// IEEE-754 round-to-nearest quotient a / b from y = RN(1/b) (hoisted: b is a per-face constant).
// Two Markstein steps: q <- RN(q + RN(a - q b) y); each remainder is exact (FMA) once q is within an ulp,
// and the second step then rounds correctly.  Requires normal-range operands and quotient, which the
// caller guards (b in [1e-8, 128], |a| in [2^-80, 2^10]); checked against `/` by occl_selftest_div.
__device__ __forceinline__ float div_rn_hoisted(float a, float b, float y) {
  float q = a * y;
  float r = __fmaf_rn(-q, b, a);
  q = __fmaf_rn(r, y, q);
  r = __fmaf_rn(-q, b, a);
  return __fmaf_rn(r, y, q);
}

// clipped-barycentric depth used to order the K nearest soft hits (BarycentricClipForward)
__device__ __forceinline__ float pz_clipped(const FaceGeo& g, float b0, float b1, float b2) {
  float c0 = b0 > 0.f ? b0 : 0.f, c1 = b1 > 0.f ? b1 : 0.f, c2 = b2 > 0.f ? b2 : 0.f;
  const float s = fmaxf(c0 + c1 + c2, 1e-5f);
  // 0 / s is +0 whatever s > 0 is: a clipped (zero) barycentric skips the quotient -- outside hits always have one or two,
  // and a zero numerator sends the IEEE division down its slow-path subroutine (10 % of the config-3 instructions, measured)
  c0 = c0 > 0.f ? c0 / s : 0.f;
  c1 = c1 > 0.f ? c1 / s : 0.f;
  c2 = c2 > 0.f ? c2 / s : 0.f;
  return c0 * g.z0 + c1 * g.z1 + c2 * g.z2;
}

__device__ __forceinline__ float soft_prob(float signed_dist, float sigma) {
  // sigmoid(-dists / sigma)                                          
  const float x = -signed_dist / sigma;
  return 1.0f / (1.0f + expf(-x));
}

// Exact pixel range of an NDC interval [lo, hi]: pixels whose centre c satisfies !(c>hi) && !(c<lo).
// Pixel centres DEcrease with the index; tab[i] = pix_to_ndc(S-1-i, S) (exact reference values).
// For a power-of-two S every centre is the exact value (S - 1 - 2 i) / S, so that the two comparisons reduce to
// i >= (S - 1 - hi S) / 2 and i <= (S - 1 - lo S) / 2, evaluated in double (hi S and the difference are exact there):
// no table look-ups, no correction loops.
__device__ __forceinline__ void ndc_range_to_pixels(const float* __restrict__ tab, float lo, float hi,
                                                    int S, int* i0, int* i1) {
  if ((S & (S - 1)) == 0) {
    const double dS = (double)S;
    double a = ceil(((dS - 1.0) - (double)hi * dS) * 0.5);
    double b = floor(((dS - 1.0) - (double)lo * dS) * 0.5);
    a = fmin(fmax(a, 0.0), dS);          // NaN bounds select the whole range, as the general path does
    b = fmin(fmax(b, -1.0), dS - 1.0);
    *i0 = (hi == hi) ? (int)a : 0;
    *i1 = (lo == lo) ? (int)b : S - 1;
    return;
  }
  const float fS = (float)S;
  float e0 = ceilf(((1.0f - hi) * fS - 1.0f) * 0.5f);   // estimates, exact up to rounding: the loops
  float e1 = floorf(((1.0f - lo) * fS - 1.0f) * 0.5f);  // below settle the last pixel either way
  e0 = fminf(fmaxf(e0, 0.0f), fS);
  e1 = fminf(fmaxf(e1, -1.0f), fS - 1.0f);
  int a = (int)e0, b = (int)e1;
  if (!(e0 == e0)) a = 0;
  if (!(e1 == e1)) b = S - 1;
  while (a < S && tab[a] > hi) ++a;
  while (a > 0 && !(tab[a - 1] > hi)) --a;
  while (b >= 0 && tab[b] < lo) --b;
  while (b < S - 1 && !(tab[b + 1] < lo)) ++b;
  *i0 = a;
  *i1 = b;
}

// Pixel-independent part of CheckPixelInsideFace: the culls and `area`, preceded by the part of
// pytorch3d's clip_faces (renderer/mesh/clip.py) that needs no new geometry: a face whose three vertices are
// all nearer than z_clip is removed; a face that straddles z_clip is cut into one or two new triangles there:
// *straddles is reported, the setup kernel lists the env (clip_list) and the clip-capable instantiation of the tile
// rasteriser rebuilds the cut triangles where it needs them (clip_subtris).
__device__ __forceinline__ bool face_geo(const float4 a, const float4 b, const float4 c, int cull, float z_clip,
                                         FaceGeo* gp, bool* straddles) {
  FaceGeo& g = *gp;
  g.x0 = a.x; g.y0 = a.y; g.z0 = a.z;
  g.x1 = b.x; g.y1 = b.y; g.z1 = b.z;
  g.x2 = c.x; g.y2 = c.y; g.z2 = c.z;
  // face_area = EdgeFunctionForward(v0, v1, v2)
  const float face_area = (g.x0 - g.x1) * (g.y2 - g.y1) - (g.y0 - g.y1) * (g.x2 - g.x1);
  const float zmax = fmaxf(fmaxf(g.z0, g.z1), g.z2);
  const float zmin = fminf(fminf(g.z0, g.z1), g.z2);
  bool skip = zmax < 0.f;
  skip |= zmax < z_clip;
  *straddles = !(zmax < z_clip) && zmin < z_clip;  // the caller cuts the face; the culls below then apply to the pieces
  skip |= (cull && face_area < 0.f);
  skip |= ((double)face_area <= 1e-8 && (double)face_area >= -1e-8);
  skip |= ((double)zmin < 1e-8);
  if (skip) return false;
  // area = EdgeFunctionForward(v2, v0, v1) + kEpsilon   (kEpsilon is a double)
  const float e = (g.x2 - g.x0) * (g.y1 - g.y0) - (g.y2 - g.y0) * (g.x1 - g.x0);
  g.area = (float)((double)e + 1e-8);
  return true;
}

// HardFlatShader lighting of one face: pixel independent, so it is evaluated once per live
// face in the setup kernel.  Returns (ambient + diffuse, specular); colour = x * texel + y.
__device__ __forceinline__ float2 face_lighting(const float* __restrict__ wv, int i0, int i1, int i2,
                                                const float light[3], float camx, float camy, float camz) {
  const float* w0 = wv + 3 * (size_t)i0;
  const float* w1 = wv + 3 * (size_t)i1;
  const float* w2 = wv + 3 * (size_t)i2;
  float v0[3], e1[3], e2[3], ctr[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    v0[k] = __ldg(w0 + k);
    const float q1 = __ldg(w1 + k), q2 = __ldg(w2 + k);
    e1[k] = q1 - v0[k];
    e2[k] = q2 - v0[k];
    ctr[k] = ((v0[k] + q1) + q2) / 3.0f;
  }
  float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
  // (the normalisations multiply by one reciprocal instead of dividing three times, the 64th power is six squarings:
  //  colours are on the tolerance side of the path, compared at 1e-5 relative)
  float nn = sqrtf((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
  nn = 1.0f / (nn > 1e-6f ? nn : 1e-6f);
  n[0] = n[0] * nn; n[1] = n[1] * nn; n[2] = n[2] * nn;
  nn = sqrtf((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
  nn = 1.0f / (nn > 1e-6f ? nn : 1e-6f);
  n[0] = n[0] * nn; n[1] = n[1] * nn; n[2] = n[2] * nn;
  float dir[3] = {light[0] - ctr[0], light[1] - ctr[1], light[2] - ctr[2]};
  float view[3] = {camx - ctr[0], camy - ctr[1], camz - ctr[2]};
  float dn = sqrtf((dir[0] * dir[0] + dir[1] * dir[1]) + dir[2] * dir[2]);
  dn = 1.0f / (dn > 1e-6f ? dn : 1e-6f);
  dir[0] = dir[0] * dn; dir[1] = dir[1] * dn; dir[2] = dir[2] * dn;
  float vn = sqrtf((view[0] * view[0] + view[1] * view[1]) + view[2] * view[2]);
  vn = 1.0f / (vn > 1e-6f ? vn : 1e-6f);
  view[0] = view[0] * vn; view[1] = view[1] * vn; view[2] = view[2] * vn;
  const float cosang = (n[0] * dir[0] + n[1] * dir[1]) + n[2] * dir[2];
  const float diffuse = 0.3f * (cosang > 0.f ? cosang : 0.f);
  const float r0 = -dir[0] + 2.0f * (cosang * n[0]);
  const float r1 = -dir[1] + 2.0f * (cosang * n[1]);
  const float r2 = -dir[2] + 2.0f * (cosang * n[2]);
  float al = (view[0] * r0 + view[1] * r1) + view[2] * r2;
  al = (al > 0.f ? al : 0.f) * (cosang > 0.f ? 1.0f : 0.0f);
  al = al * al; al = al * al; al = al * al; al = al * al; al = al * al; al = al * al;
  return make_float2(0.5f + diffuse, 0.2f * al);
}

// ----------------------------------------------------------------------------------------------
// z-clip: pytorch3d renderer/mesh/clip.py::clip_faces for the frustum MeshRasterizer builds (z_clip_value only)
// ----------------------------------------------------------------------------------------------
// A face with 1 or 2 vertices nearer than z_clip is cut at the plane z = z_clip (in (x_ndc, y_ndc, z_view) space,
// interpolating x_ndc * z, which is proportional to view-space x):
//   one vertex p1 nearer : the quadrilateral in front, as t1 = (p4, p2, p5), t2 = (p5, p2, p3)   (neighbours)
//   two vertices nearer  : the triangle (p1, p4, p5), p1 being the vertex in front
// p2, p3 follow p1 in the face's own order; p4 / p5 = plane intersections of the edges p1p2 / p1p3.
// Same operation order as oracle/oracle.py::clip_faces.  conv = barycentrics of the cut triangle's vertices in the
// uncut face (convert_clipped_rasterization_to_original_faces).
struct SubTris {
  int n;
  FaceGeo g[2];
  bool live[2];
  float conv[2][9];
};

__device__ __forceinline__ float4 clip_cut(const float4 p1, const float4 q, const float zc, float* w_out) {
  const float w = (p1.z - zc) / (p1.z - q.z);
  const float a1 = 1.0f - w;
  const float x = ((p1.x * p1.z) * a1 + (q.x * q.z) * w) / zc;
  const float y = ((p1.y * p1.z) * a1 + (q.y * q.z) * w) / zc;
  *w_out = w;
  return make_float4(x, y, zc, 0.f);
}

__device__ __noinline__ void clip_subtris(const float4 a, const float4 b, const float4 c, const float zc, const int cull,
                                          SubTris* out) {
  const bool c0 = a.z < zc, c1 = b.z < zc, c2 = c.z < zc;
  const int n = (int)c0 + (int)c1 + (int)c2;
  out->n = 0;
  if (n == 0 || n == 3) return;
  const int i = n == 1 ? (c0 ? 0 : (c1 ? 1 : 2)) : (!c0 ? 0 : (!c1 ? 1 : 2));  // the isolated vertex
  const float4 p1 = i == 0 ? a : (i == 1 ? b : c);
  const float4 p2 = i == 0 ? b : (i == 1 ? c : a);
  const float4 p3 = i == 0 ? c : (i == 1 ? a : b);
  const int j = (i + 1) % 3, k = (i + 2) % 3;
  float w2, w3;
  const float4 p4 = clip_cut(p1, p2, zc, &w2);
  const float4 p5 = clip_cut(p1, p3, zc, &w3);
  float e4[3] = {0.f, 0.f, 0.f}, e5[3] = {0.f, 0.f, 0.f}, ei[3] = {0.f, 0.f, 0.f}, ej[3] = {0.f, 0.f, 0.f}, ek[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    ei[t] = t == i ? 1.f : 0.f;
    ej[t] = t == j ? 1.f : 0.f;
    ek[t] = t == k ? 1.f : 0.f;
    e4[t] = t == i ? 1.0f - w2 : (t == j ? w2 : 0.f);
    e5[t] = t == i ? 1.0f - w3 : (t == k ? w3 : 0.f);
  }
  bool unused;
  if (n == 1) {
    out->n = 2;
    out->live[0] = face_geo(p4, p2, p5, cull, 0.f, &out->g[0], &unused);
    out->live[1] = face_geo(p5, p2, p3, cull, 0.f, &out->g[1], &unused);
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      out->conv[0][t] = e4[t]; out->conv[0][3 + t] = ej[t]; out->conv[0][6 + t] = e5[t];
      out->conv[1][t] = e5[t]; out->conv[1][3 + t] = ej[t]; out->conv[1][6 + t] = ek[t];
    }
  } else {
    out->n = 1;
    out->live[0] = face_geo(p1, p4, p5, cull, 0.f, &out->g[0], &unused);
    out->live[1] = false;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      out->conv[0][t] = ei[t]; out->conv[0][3 + t] = e4[t]; out->conv[0][6 + t] = e5[t];
    }
  }
}

// One pixel against a cut face.  Soft (silhouette) and hard (observation) results differ only in the hit rule
// and in the depth; when both triangles of a cut hit, the one with the smaller distance is kept, the second only
// if strictly closer (clipped_faces_neighbor_idx rule of the reference's face loop).
struct ClipPixel {
  bool soft_hit, soft_inside, hard_hit;
  int soft_which, soft_edge;  // cut triangle of the soft hit, its nearest edge (0: v0v1, 1: v0v2, 2: v1v2) ...
  float soft_t;               // ... and the clamped parameter on it (for the gradient)
  float soft_dist, soft_pz;   // squared distance (unsigned), clipped-barycentric depth
  float hard_pz;              // perspective-correct depth of the nearest... of the chosen triangle
  int hard_which;
  float hb0, hb1, hb2;        // barycentrics of the hard hit in the cut triangle
};

__device__ __noinline__ void eval_clip_pixel(const SubTris* st, const float px, const float py, const float blur,
                                             const float bbox_r, ClipPixel* o) {
  o->soft_hit = false; o->hard_hit = false; o->soft_inside = false;
  o->soft_which = 0; o->soft_edge = 0; o->soft_t = 0.f;
  o->soft_dist = 0.f; o->soft_pz = 0.f; o->hard_pz = 0.f; o->hard_which = 0; o->hb0 = o->hb1 = o->hb2 = 0.f;
  float hard_dist = 0.f;
  for (int t = 0; t < st->n; ++t) {
    if (!st->live[t]) continue;
    const FaceGeo& g = st->g[t];
    const float xlo = fminf(fminf(g.x0, g.x1), g.x2), xhi = fmaxf(fmaxf(g.x0, g.x1), g.x2);
    const float ylo = fminf(fminf(g.y0, g.y1), g.y2), yhi = fmaxf(fmaxf(g.y0, g.y1), g.y2);
    if (px > xhi + bbox_r || px < xlo - bbox_r || py > yhi + bbox_r || py < ylo - bbox_r) continue;
    const PairResult r = eval_pair(g, px, py);
    if (r.inside || r.dist < blur) {
      const float pz = pz_clipped(g, r.b0, r.b1, r.b2);
      if (!(pz < 0.f) && (!o->soft_hit || r.dist < o->soft_dist)) {
        o->soft_hit = true; o->soft_inside = r.inside; o->soft_dist = r.dist; o->soft_pz = pz;
        o->soft_which = t; o->soft_edge = r.edge; o->soft_t = r.t;
      }
    }
    if (r.inside && !(px > xhi || px < xlo || py > yhi || py < ylo)) {
      const float pz = r.b0 * g.z0 + r.b1 * g.z1 + r.b2 * g.z2;
      if (!(pz < 0.f) && (!o->hard_hit || r.dist < hard_dist)) {
        o->hard_hit = true; hard_dist = r.dist; o->hard_pz = pz; o->hard_which = t;
        o->hb0 = r.b0; o->hb1 = r.b1; o->hb2 = r.b2;
      }
    }
  }
}

// soft result of one (pixel, cut face) pair from the record's uncut vertices (K-overflow paths)
__device__ __noinline__ bool clip_soft_eval(const uint4* __restrict__ rec, const float px, const float py,
                                            const float z_clip, const int cull, const float blur, const float bbox_r,
                                            bool* inside, float* dist, float* pz) {
  const uint4 q0 = __ldg(rec + 0), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2);
  SubTris st;
  clip_subtris(make_float4(__uint_as_float(q0.x), __uint_as_float(q0.y), __uint_as_float(q0.z), 0.f),
               make_float4(__uint_as_float(q0.w), __uint_as_float(q1.x), __uint_as_float(q1.y), 0.f),
               make_float4(__uint_as_float(q1.z), __uint_as_float(q1.w), __uint_as_float(q2.x), 0.f), z_clip, cull, &st);
  ClipPixel cp;
  eval_clip_pixel(&st, px, py, blur, bbox_r, &cp);
  *inside = cp.soft_inside; *dist = cp.soft_dist; *pz = cp.soft_pz;
  return cp.soft_hit;
}

// Screen-space tangents d(x_ndc, y_ndc)/d(el, az) of the vertices of the cut triangles of face `fidx` (same cases and
// vertex orders as clip_subtris).  An uncut vertex carries the tangent the projection kernel wrote; a plane
// intersection p4 = cut(p1, q), x4 = ((x1 z1)(1 - w) + (xq zq) w) / z_clip with w = (z1 - z_clip) / (z1 - zq), moves
// with x, y AND z of both end points -- pytorch3d gets the same derivative by autograd through clip_faces.
// d z_view / d theta is recomputed from the world vertex and the camera tangent blocks (cut faces are rare).
struct SubTrisTan {
  float4 t[2][3];
};

__device__ __noinline__ void clip_subtris_tan(const RasterParams& p, const int env, const int fidx, const float4 a,
                                              const float4 b, const float4 c, const float zc, SubTrisTan* out) {
  const int* __restrict__ fc = p.faces + (size_t)env * p.faces_stride + 3 * (size_t)fidx;
  const int vi[3] = {__ldg(fc + 0), __ldg(fc + 1), __ldg(fc + 2)};
  const float4* __restrict__ vt = p.vtan + (size_t)env * p.V;
  const float* __restrict__ cam = p.cam + (size_t)env * OCCL_CAM_STRIDE;
  const float4 pv[3] = {a, b, c};
  float4 tv[3];
  float dz[3][2];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    tv[k] = __ldg(vt + vi[k]);
    const float* __restrict__ w = p.verts + (size_t)env * p.verts_stride + 3 * (size_t)vi[k];
    const float x = __ldg(w + 0), y = __ldg(w + 1), z = __ldg(w + 2);
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const float* __restrict__ dc = cam + 16 * (d + 1);
      dz[k][d] = x * __ldg(dc + 2) + y * __ldg(dc + 5) + z * __ldg(dc + 8) + __ldg(dc + 11);
    }
  }
  const bool c0 = a.z < zc, c1 = b.z < zc, c2 = c.z < zc;
  const int n = (int)c0 + (int)c1 + (int)c2;
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int k = 0; k < 3; ++k) out->t[t][k] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n == 0 || n == 3) return;
  const int i = n == 1 ? (c0 ? 0 : (c1 ? 1 : 2)) : (!c0 ? 0 : (!c1 ? 1 : 2));  // the isolated vertex
  const int j = (i + 1) % 3, k = (i + 2) % 3;
  auto cut_tan = [&](const int q) {
    const float4 p1 = pv[i], pq = pv[q];
    const float den = p1.z - pq.z;
    const float w = (p1.z - zc) / den;
    float r[4];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const float dz1 = dz[i][d], dzq = dz[q][d];
      const float dw = (dz1 * den - (p1.z - zc) * (dz1 - dzq)) / (den * den);
      const float dx1 = d == 0 ? tv[i].x : tv[i].z, dy1 = d == 0 ? tv[i].y : tv[i].w;
      const float dxq = d == 0 ? tv[q].x : tv[q].z, dyq = d == 0 ? tv[q].y : tv[q].w;
      r[2 * d + 0] = ((dx1 * p1.z + p1.x * dz1) * (1.0f - w) - (p1.x * p1.z) * dw + (dxq * pq.z + pq.x * dzq) * w + (pq.x * pq.z) * dw) / zc;
      r[2 * d + 1] = ((dy1 * p1.z + p1.y * dz1) * (1.0f - w) - (p1.y * p1.z) * dw + (dyq * pq.z + pq.y * dzq) * w + (pq.y * pq.z) * dw) / zc;
    }
    return make_float4(r[0], r[1], r[2], r[3]);
  };
  const float4 t4 = cut_tan(j), t5 = cut_tan(k);
  if (n == 1) {  // t1 = (p4, p2, p5), t2 = (p5, p2, p3)
    out->t[0][0] = t4; out->t[0][1] = tv[j]; out->t[0][2] = t5;
    out->t[1][0] = t5; out->t[1][1] = tv[j]; out->t[1][2] = tv[k];
  } else {       // (p1, p4, p5)
    out->t[0][0] = tv[i]; out->t[0][1] = t4; out->t[0][2] = t5;
  }
}

// p_k / sigma * d(signed dist)/d(el, az) of a soft hit on a cut face (SURVEY A.7 on the cut triangle)
__device__ __noinline__ void clip_hit_tangent(const RasterParams& p, const SubTris* st, const SubTrisTan* tt,
                                              const ClipPixel* cp, const float px, const float py, float* g0, float* g1) {
  const FaceGeo& g = st->g[cp->soft_which];
  const float4* tq = tt->t[cp->soft_which];
  float ax, ay, bx, by;
  float4 da, db;
  if (cp->soft_edge == 0) { ax = g.x0; ay = g.y0; bx = g.x1; by = g.y1; da = tq[0]; db = tq[1]; }
  else if (cp->soft_edge == 1) { ax = g.x0; ay = g.y0; bx = g.x2; by = g.y2; da = tq[0]; db = tq[2]; }
  else { ax = g.x1; ay = g.y1; bx = g.x2; by = g.y2; da = tq[1]; db = tq[2]; }
  const float t = cp->soft_t;
  const float qx = ax + t * (bx - ax), qy = ay + t * (by - ay);
  const float sgn = cp->soft_inside ? -1.f : 1.f;
  const float gx = sgn * 2.f * (qx - px), gy = sgn * 2.f * (qy - py);
  const float wa = 1.f - t, wb = t;
  const float kk = soft_prob(cp->soft_inside ? -cp->soft_dist : cp->soft_dist, p.sigma) / p.sigma;
  *g0 += kk * (gx * (wa * da.x + wb * db.x) + gy * (wa * da.y + wb * db.y));
  *g1 += kk * (gx * (wa * da.z + wb * db.z) + gy * (wa * da.w + wb * db.w));
}

__device__ __forceinline__ float soft_term(float signed_dist, float inv_sigma_log2e);
// K-overflow round: hits of one cut face on the slots of the round (one lane; rare)
__device__ __noinline__ void clip_round_hits(const uint4* __restrict__ rec, const uint4 rg, const int* s_rslot, const int rn,
                                             unsigned long long* soft_all, unsigned long long* hkey, float* hq,
                                             const float* ndc_x, const float* ndc_y, const int tx0, const int ty0,
                                             const int tile_w, const int tpx, const float z_clip, const int cull,
                                             const float blur, const float bbox_r, const float inv_sigma_log2e) {
  const int fx0 = (int)(rg.x & 0xffffu) - tx0, fx1 = (int)(rg.x >> 16) - tx0;
  const int fy0 = (int)(rg.y & 0xffffu) - ty0, fy1 = (int)(rg.y >> 16) - ty0;
  const int fobj = (int)(rg.w >> 30);
  const unsigned fidx = __ldg(rec + 2).z & REC_FIDX_MASK;
  for (int r = 0; r < rn; ++r) {
    const int slot = s_rslot[r];
    const int obj = slot / tpx, pix = slot - obj * tpx;
    const int ly = pix / tile_w, lx = pix - ly * tile_w;
    if (obj != fobj || lx < fx0 || lx > fx1 || ly < fy0 || ly > fy1) continue;
    bool inside;
    float dist, pz;
    if (!clip_soft_eval(rec, ndc_x[lx], ndc_y[ly], z_clip, cull, blur, bbox_r, &inside, &dist, &pz)) continue;
    const unsigned pos = atomicAdd((unsigned*)(soft_all + slot), 1u);
    if (pos < HITBUF_CAP) {
      hkey[pos] = ((unsigned long long)__float_as_uint(pz) << 32) | (unsigned long long)fidx;
      hq[pos] = soft_term(inside ? -dist : dist, inv_sigma_log2e);
    }
  }
}

// epilogue of a pixel whose nearest face is a cut face: barycentrics in the UNCUT face (the reference converts
// them with the cut triangle's conversion matrix) from the projected vertices of the face
__device__ __noinline__ void clip_hard_bary(const float4* __restrict__ vp, const int i0, const int i1, const int i2,
                                            const float z_clip, const int cull, const int which, const float px,
                                            const float py, float* b0, float* b1, float* b2) {
  const float4 a = __ldg(vp + i0), b = __ldg(vp + i1), c = __ldg(vp + i2);
  SubTris st;
  clip_subtris(a, b, c, z_clip, cull, &st);
  const int t = which < st.n ? which : 0;
  float c0, c1, c2;
  bary_persp(st.g[t], px, py, &c0, &c1, &c2);
  const float* m = st.conv[t];
  *b0 = (c0 * m[0] + c1 * m[3]) + c2 * m[6];
  *b1 = (c0 * m[1] + c1 * m[4]) + c2 * m[7];
  *b2 = (c0 * m[2] + c1 * m[5]) + c2 * m[8];
}

// ----------------------------------------------------------------------------------------------
// kernel 2: per-env face setup + warp-ballot compaction of the live faces
// ----------------------------------------------------------------------------------------------
struct SetupParams {
  int S, V, F, cull, n_obj, grad;
  float z_clip;
  uint32_t* status;
  int obj_face_start[OCCL_MAX_OBJ + 1];
  int tile_w, tile_h, tiles_x, n_tiles;
  float inv_tile_w, inv_tile_h;
  float bbox_r;
  const float4* vproj;
  const int* faces;
  long long faces_stride;
  uint4* geo;
  uint4* rng;
  int* n_live;
  uint32_t* tile_mask;  // [N][TILE_MASK_WORDS] bit t set: some live face's blur box overlaps tile t
  uint32_t* obs_tile_state;  // [N][OCCL_TILE_STATE_WORDS] or null (OcclOutputs.obs_tile_state): incremental delivery of obs
  int* tile_idx;        // [N][n_tiles][tidx_cap] binning: live-list indices of the faces whose blur box overlaps a tile
  int* tile_cnt;        // [N][n_tiles] their number (may exceed tidx_cap: the raster kernel then scans the whole live list)
  int tidx_cap;
  float2* shade;        // [N][F] (ambient + diffuse, specular) of the live faces
  int* clip_list;       // [1 + N] count, then ids, of the envs in which a face was cut at z_clip (zeroed before the launch)
  const float* verts;
  long long verts_stride;
  const float* cam;
  float light[3];
  const uint8_t* env_mask;
};

// BIN: also bin the live faces per tile (dense scenes; a separate instantiation so that the kernel of config 2 is
// not touched: its register allocation is as sensitive as the raster kernel's)
template <bool BIN>
__global__ void __launch_bounds__(SETUP_THREADS, SETUP_CTAS) face_setup_kernel(const SetupParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tab = (float*)smem_raw;  // [S] pixel-centre table
  __shared__ int s_base, s_cut;
  __shared__ uint32_t s_tmask_[TILE_MASK_WORDS];
  __shared__ int s_tcnt_[BIN ? 32 * TILE_MASK_WORDS : 1];  // faces binned per tile (images of up to 256 tiles)
  __shared__ int s_wl_[SETUP_WARPS][64];
  uint32_t* s_tmask = s_tmask_;
  int* s_tcnt = s_tcnt_;
  int (*s_wl)[64] = s_wl_;
  OCCL_OPAQUE_SHARED(tab);
  OCCL_OPAQUE_SHARED(s_tmask);
  OCCL_OPAQUE_SHARED(s_tcnt);
  OCCL_OPAQUE_SHARED(s_wl);
  const int env = blockIdx.x;
  if (p.env_mask && !p.env_mask[env]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = p.S;
  for (int i = tid; i < S; i += SETUP_THREADS) tab[i] = pix_to_ndc(S - 1 - i, S);
  if (BIN)
    for (int i = tid; i < 32 * TILE_MASK_WORDS; i += SETUP_THREADS) s_tcnt[i] = 0;
  if (tid < TILE_MASK_WORDS) s_tmask[tid] = 0u;
  if (tid == 0) { s_base = 0; s_cut = 0; }
  __syncthreads();
  const float4* __restrict__ vp = p.vproj + (size_t)env * p.V;
  const int* __restrict__ faces = p.faces + (size_t)env * p.faces_stride;
  uint4* __restrict__ geo = p.geo + (size_t)env * p.F * 4;
  uint4* __restrict__ rng = p.rng + (size_t)env * p.F;
  // Warps take 64-face chunks in turn.  Pass A (cheap): the culls -- about half the faces of a closed mesh are back
  // faces -- and a ballot compaction of the survivors into the warp's queue; pass B (ranges, record, lighting, binning):
  // dense over the queue, so that its lanes are busy whatever the cull rate.  List slots are reserved with one atomic
  // per pass: the list is in mesh order up to the interleaving of concurrently finishing warps (nothing depends on
  // its order).
  __shared__ int s_next;  // next 64-face chunk (the warps take chunks dynamically: the cull rate varies along the mesh)
  if (tid == 0) s_next = 0;
  __syncthreads();
  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(&s_next, 64);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= p.F) break;
    int nl = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int fa = base + half * 32 + lane;
      bool keep = false;
      if (fa < p.F) {
        const int i0 = __ldg(faces + 3 * fa + 0), i1 = __ldg(faces + 3 * fa + 1), i2 = __ldg(faces + 3 * fa + 2);
        bool straddles = false;
        FaceGeo ga;
        keep = face_geo(__ldg(vp + i0), __ldg(vp + i1), __ldg(vp + i2), p.cull, p.z_clip, &ga, &straddles) || straddles;
      }
      const unsigned kb = __ballot_sync(0xffffffffu, keep);
      if (keep) s_wl[warp][nl + __popc(kb & ((1u << lane) - 1u))] = fa;
      nl += __popc(kb);
    }
    __syncwarp();
   for (int q0 = 0; q0 < nl; q0 += 32) {
    const int f = q0 + lane < nl ? s_wl[warp][q0 + lane] : p.F;
    bool live = false, clipf = false;
    FaceGeo g;
    int sx0 = 0, sx1 = -1, sy0 = 0, sy1 = -1, hx0 = 0, hx1 = -1, hy0 = 0, hy1 = -1;
    if (f < p.F) {
      const int i0 = __ldg(faces + 3 * f + 0), i1 = __ldg(faces + 3 * f + 1), i2 = __ldg(faces + 3 * f + 2);
      bool straddles = false;
      const float4 va = __ldg(vp + i0), vb = __ldg(vp + i1), vc = __ldg(vp + i2);
      live = face_geo(va, vb, vc, p.cull, p.z_clip, &g, &straddles);
      float xlo = fminf(fminf(g.x0, g.x1), g.x2), xhi = fmaxf(fmaxf(g.x0, g.x1), g.x2);
      float ylo = fminf(fminf(g.y0, g.y1), g.y2), yhi = fmaxf(fmaxf(g.y0, g.y1), g.y2);
      if (straddles) {
        // cut at z = z_clip: the record keeps the uncut vertices, its pixel ranges are those of the cut polygon
        atomicOr(p.status + env, OCCL_ST_CLIPPED);
        s_cut = 1;
        SubTris st;
        clip_subtris(va, vb, vc, p.z_clip, p.cull, &st);
        live = false;
        xlo = ylo = 3.0e38f; xhi = yhi = -3.0e38f;
        for (int t = 0; t < st.n; ++t) {
          if (!st.live[t]) continue;
          live = true;
          const FaceGeo& q = st.g[t];
          xlo = fminf(xlo, fminf(fminf(q.x0, q.x1), q.x2)); xhi = fmaxf(xhi, fmaxf(fmaxf(q.x0, q.x1), q.x2));
          ylo = fminf(ylo, fminf(fminf(q.y0, q.y1), q.y2)); yhi = fmaxf(yhi, fmaxf(fmaxf(q.y0, q.y1), q.y2));
        }
        clipf = live;
      }
      if (live) {
        ndc_range_to_pixels(tab, xlo - p.bbox_r, xhi + p.bbox_r, S, &sx0, &sx1);
        ndc_range_to_pixels(tab, ylo - p.bbox_r, yhi + p.bbox_r, S, &sy0, &sy1);
        live = sx0 <= sx1 && sy0 <= sy1;  // faces whose blur box misses every pixel centre
        if (live) {
          ndc_range_to_pixels(tab, xlo - 0.0f, xhi + 0.0f, S, &hx0, &hx1);
          ndc_range_to_pixels(tab, ylo - 0.0f, yhi + 0.0f, S, &hy0, &hy1);
          if (hx0 > hx1 || hy0 > hy1) { hx0 = 0xffff; hx1 = 0; hy0 = 0xffff; hy1 = 0; }
        }
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, live);
    int off = 0;
    if (lane == 0 && bal) off = atomicAdd(&s_base, __popc(bal));
    off = __shfl_sync(0xffffffffu, off, 0);
    if (live) {
      const int slot = off + __popc(bal & ((1u << lane) - 1u));
      // fast-path guards (see eval_fast): magnitudes for which sign(b_i) == sign(e_i) provably and the
      // approximate edge distances stay inside the decision band
      const float bx01 = g.x1 - g.x0, by01 = g.y1 - g.y0, bx02 = g.x2 - g.x0, by02 = g.y2 - g.y0;
      const float bx12 = g.x2 - g.x1, by12 = g.y2 - g.y1;
      const float l01 = bx01 * bx01 + by01 * by01, l02 = bx02 * bx02 + by02 * by02, l12 = bx12 * bx12 + by12 * by12;
      const float zmax = fmaxf(fmaxf(g.z0, g.z1), g.z2), zmin = fminf(fminf(g.z0, g.z1), g.z2);
      const float cmax = fmaxf(fmaxf(fmaxf(fabsf(g.x0), fabsf(g.x1)), fabsf(g.x2)),
                               fmaxf(fmaxf(fabsf(g.y0), fabsf(g.y1)), fabsf(g.y2)));
      const bool fast = !clipf && g.area >= 9.094947e-13f /*2^-40*/ && g.area <= 1024.f && zmin >= 9.765625e-4f && zmax <= 1024.f &&
                        cmax <= 4.0f && l01 > 1e-8f && l02 > 1e-8f && l12 > 1e-8f;
      uint4 q0, q1, q2, q3;
      q0 = make_uint4(__float_as_uint(g.x0), __float_as_uint(g.y0), __float_as_uint(g.z0), __float_as_uint(g.x1));
      q1 = make_uint4(__float_as_uint(g.y1), __float_as_uint(g.z1), __float_as_uint(g.x2), __float_as_uint(g.y2));
      uint32_t obj = 0;
#pragma unroll
      for (int i = 1; i < OCCL_MAX_OBJ; ++i)
        if (i < p.n_obj && f >= p.obj_face_start[i]) obj = i;
      q2 = make_uint4(__float_as_uint(g.z2), __float_as_uint(g.area),
                      (uint32_t)f | (obj << REC_OBJ_SHIFT) | (fast ? REC_FAST : 0u) | (clipf ? REC_CLIP : 0u), 0u);
      if (p.n_tiles <= 32 * TILE_MASK_WORDS) {
        // exact for these small integers: (i + 0.5) / t never lands within 0.5/t of an integer
        const int tx_lo = (int)(((float)sx0 + 0.5f) * p.inv_tile_w), tx_hi = (int)(((float)sx1 + 0.5f) * p.inv_tile_w);
        const int ty_lo = (int)(((float)sy0 + 0.5f) * p.inv_tile_h), ty_hi = (int)(((float)sy1 + 0.5f) * p.inv_tile_h);
        for (int ty = ty_lo; ty <= ty_hi; ++ty)
          for (int tx = tx_lo; tx <= tx_hi; ++tx) {
            const int t = ty * p.tiles_x + tx;
            atomicOr(&s_tmask[t >> 5], 1u << (t & 31));
            // binning: the raster CTA of tile t reads this list instead of range-testing every live face
            if (BIN) {
              const int pos = atomicAdd(&s_tcnt[t], 1);
              if (pos < p.tidx_cap) p.tile_idx[((size_t)env * p.n_tiles + t) * p.tidx_cap + pos] = slot;
            }
          }
      }
      q3 = make_uint4(__float_as_uint(1.0f / l01), __float_as_uint(1.0f / l02), __float_as_uint(1.0f / l12), 0u);
      uint4* o = geo + (size_t)slot * 4;
      o[0] = q0; o[1] = q1; o[2] = q2; o[3] = q3;
      if (hx0 <= hx1) {  // only faces that can own a pixel need a colour
        const float* __restrict__ cam = p.cam + (size_t)env * OCCL_CAM_STRIDE;
        p.shade[(size_t)env * p.F + f] =
            face_lighting(p.verts + (size_t)env * p.verts_stride, __ldg(faces + 3 * f + 0), __ldg(faces + 3 * f + 1),
                          __ldg(faces + 3 * f + 2), p.light, __ldg(cam + 12), __ldg(cam + 13), __ldg(cam + 14));
      }
      rng[slot] = make_uint4((uint32_t)sx0 | ((uint32_t)sx1 << 16), (uint32_t)sy0 | ((uint32_t)sy1 << 16),
                             (uint32_t)hx0 | ((uint32_t)hx1 << 16) | (clipf ? RNG_CLIP : 0u),
                             (uint32_t)hy0 | ((uint32_t)hy1 << 16) | (obj << 30));
    }
   }
   __syncwarp();  // the queue is rewritten by the next chunk
  }
  __syncthreads();
  if (tid == 0 && s_cut) p.clip_list[1 + atomicAdd(p.clip_list, 1)] = env;  // raster_clip_kernel works through this list
  if (tid == 0) p.n_live[env] = s_base;
  if (tid < TILE_MASK_WORDS) {
    const uint32_t cur = s_tmask[tid];
    p.tile_mask[(size_t)env * TILE_MASK_WORDS + tid] = cur;
    if (p.obs_tile_state) {
      // incremental delivery: a tile that was background in the destination and is background now keeps its pixels
      uint32_t* st = p.obs_tile_state + (size_t)env * OCCL_TILE_STATE_WORDS;
      const uint32_t prev = st[tid];
      st[TILE_MASK_WORDS + tid] = ~prev & ~cur;
      st[tid] = cur;
    }
  }
  if (BIN)
    for (int t = tid; t < p.n_tiles; t += SETUP_THREADS) p.tile_cnt[(size_t)env * p.n_tiles + t] = s_tcnt[t];
}

// ----------------------------------------------------------------------------------------------
// kernels 3-5: tile rasteriser
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ int obj_of_face(const RasterParams& p, int f) {
  int o = 0;
#pragma unroll
  for (int i = 1; i < OCCL_MAX_OBJ; ++i)
    if (i < p.n_obj && f >= p.obj_face_start[i]) o = i;
  return o;
}

// soft accumulator of a (pixel, object) slot: two 32-bit words, updated with NATIVE shared-memory atomics
// (ATOMS.ADD / ATOMS.OR; a 64-bit or floating-point shared atomic is a compare-and-swap loop on sm_100a):
//   low word  = L, the sum over the hits of -log2(1 - prob) in fixed point with SOFT_FRAC fractional bits
//               (integer adds commute, so the per-pixel product -- and alpha -- is run-to-run deterministic);
//               once the top-K rule has been applied (SOFT_RESOLVED) the low word holds the float bits of the
//               product itself; during a selection round (SOFT_ROUND) it is the write cursor of the hit list
//   high word = hit count (bits 0..19) | SOFT_RESOLVED (bit 20) | SOFT_ROUND (bit 21) | SOFT_COVERED (bit 22)
//               | count of "strong" hits, factor 1 - prob <= 1/4 (bits 23..31; wraps off the top of the word,
//               which can only lose the strong-hit shortcut, never fake it)
// Terms are clamped to SOFT_TERM_MAX = 26 (a factor below 2^-26 makes alpha == 1.0f on its own) and a slot with
// SOFT_STRONG_NEEDED strong hits is 1.0f whatever L says, so L <= 12 * 26 + 128 * 2 < 2^(32 - SOFT_FRAC) never wraps
// where it is read.
#define SOFT_CNT_MASK 0xfffffu
#define SOFT_RESOLVED 0x00100000u
#define SOFT_ROUND 0x00200000u
#define SOFT_COVERED 0x00400000u
#define SOFT_STRONG_SHIFT 23
#define SOFT_STRONG_ONE (1u << SOFT_STRONG_SHIFT)
#define SOFT_FRAC 22
#define SOFT_TERM_MAX 26.0f
#define SOFT_STRONG_NEEDED 13  // 0.25^13 < 2^-25: that many strong factors make 1 - product == 1.0f exactly
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// -log2(1 - sigmoid(-sd / sigma)) = log2(1 + 2^(-x)), x = sd / (sigma ln 2)            (tolerance side of the path)
//   = max(-x, 0) + log2(1 + u),  u = 2^-|x| in (0, 1].
// log2(1 + u): MUFU.LG2 has an ABSOLUTE error of ~2^-22.6 that does not average out (a pixel just outside a far,
// tiny object collects ~100 hits with u ~ 1e-4 each: alpha ~ 0.01 came out 3e-6 off); for u < 1/4 a degree-5
// polynomial in u (relative error 1.4e-7, unbiased) is used instead, so the error of every weak term is relative.
__device__ __forceinline__ float soft_term(float signed_dist, float inv_sigma_log2e) {
  const float x = signed_dist * inv_sigma_log2e;
  const float a = fabsf(x);
  const float u = ex2_approx(-a);
  float pl = __fmaf_rn(-0.1329696774482727f, u, 0.26198700070381165f);
  pl = __fmaf_rn(pl, u, -0.35743656754493713f);
  pl = __fmaf_rn(pl, u, 0.4807146489620209f);
  pl = __fmaf_rn(pl, u, -0.7213436365127563f);
  pl = __fmaf_rn(pl, u, 1.4426950216293335f);
  const float t = u < 0.25f ? pl * u : lg2_approx(1.0f + u);
  return fminf(x < 0.f ? t + a : t, SOFT_TERM_MAX);
}
__device__ __forceinline__ void soft_accumulate(unsigned long long* slot, float term, bool covered) {
  unsigned* w = (unsigned*)slot;
  atomicAdd(w, __float2uint_rn(term * (float)(1u << SOFT_FRAC)));
  atomicAdd(w + 1, term >= 2.0f ? 1u + SOFT_STRONG_ONE : 1u);
  if (covered) atomicOr(w + 1, SOFT_COVERED);
}
// the same, returning the slot's meta word as it was before this hit (its count = the hit's position in the slot's
// recorded segment, its SOFT_ROUND bit = "this slot records", see REC_*)
__device__ __forceinline__ unsigned soft_accumulate_ret(unsigned long long* slot, float term, bool covered) {
  unsigned* w = (unsigned*)slot;
  atomicAdd(w, __float2uint_rn(term * (float)(1u << SOFT_FRAC)));
  const unsigned old = atomicAdd(w + 1, term >= 2.0f ? 1u + SOFT_STRONG_ONE : 1u);
  if (covered) atomicOr(w + 1, SOFT_COVERED);
  return old;
}
// product of the factors of an unresolved slot from its fixed-point log sum
__device__ __forceinline__ float soft_product(unsigned long long w) {
  const unsigned hi = (unsigned)(w >> 32), lo = (unsigned)(w & 0xffffffffull);
  if (hi & SOFT_RESOLVED) return __uint_as_float(lo);
  if ((hi >> SOFT_STRONG_SHIFT) >= SOFT_STRONG_NEEDED) return 0.0f;
  return ex2_approx(-((float)lo * (1.0f / (float)(1u << SOFT_FRAC))));
}
// The K-overflow paths sum the terms of the hits they keep as 64-bit integers (32 fractional bits: these paths are cold,
// and K hits of a far, tiny object can share one distance, so that 22-bit rounding errors would add up coherently):
// alpha then depends on the SET of kept hits only, not on the order in which the hits arrived -- run-to-run
// deterministic, like the main phase.
__device__ __forceinline__ unsigned long long soft_term_fx(float term) {
  return __float2ull_rn(term * 4294967296.0f);
}
__device__ __forceinline__ float soft_product_fx(unsigned long long fx) {
  return ex2_approx(-((float)fx * (1.0f / 4294967296.0f)));
}
// strong-hit shortcut of the top-K rule: whatever K hits are the nearest, at most (count - strong) of them are weak
__device__ __forceinline__ bool soft_strong_shortcut(unsigned hi, int K) {
  const int cnt = (int)(hi & SOFT_CNT_MASK), strong = (int)(hi >> SOFT_STRONG_SHIFT);
  return K - (cnt - strong) >= SOFT_STRONG_NEEDED;
}

__device__ __forceinline__ unsigned long long pack2f(float lo, float hi) {
  return ((unsigned long long)__float_as_uint(hi) << 32) | (unsigned long long)__float_as_uint(lo);
}
// both tangent sums of a pixel in one 64-bit CAS
__device__ __forceinline__ void grad_accumulate(unsigned long long* slot, float a, float b) {
  unsigned long long old = *slot, assumed;
  do {
    assumed = old;
    const float lo = __uint_as_float((unsigned)(assumed & 0xffffffffull)) + a;
    const float hi = __uint_as_float((unsigned)(assumed >> 32)) + b;
    old = atomicCAS(slot, assumed, pack2f(lo, hi));
  } while (old != assumed);
}

struct TileSmem {
  unsigned long long* hard;  // [tpx]      (z bits << 32) | packed face index ; ~0 = background
  unsigned long long* soft;  // [n_obj][tpx]
  unsigned long long* gacc;  // [n_obj][tpx] two fp32 tangent sums (d/d_el low word, d/d_az high word)  (GRAD)
  float* ndc_x;              // [tile_w]
  float* ndc_y;              // [tile_h]
  uint32_t* list;            // [warps][WBUF_RECS][REC_WORDS]  (aliased by the top-K selection buffers)
  uint32_t* defer;           // [warps][WDEFER_CAP] inside hits awaiting their exact depth
  uint32_t* big;             // [BIG_CAP][REC_WORDS] small scratch region (slot / list tables of the K-overflow paths)
  int* defer_n;
};

// Shared-memory layout of a tile: fixed-size regions first so that, with a compile-time tile, every base
// address is a constant.
__device__ __forceinline__ TileSmem tile_smem_layout(unsigned char* q, const int tile_w, const int tile_h, const int n_obj) {
  TileSmem sm;
  const int tpx = tile_w * tile_h;
  sm.list = (uint32_t*)q;            q += sizeof(uint32_t) * OCCL_WARPS * WBUF_RECS * REC_WORDS;
  sm.defer = (uint32_t*)q;           q += sizeof(uint32_t) * OCCL_WARPS * WDEFER_CAP;
  sm.big = (uint32_t*)q;             q += sizeof(uint32_t) * BIG_CAP * REC_WORDS;
  sm.hard = (unsigned long long*)q;  q += sizeof(unsigned long long) * tpx;
  sm.ndc_x = (float*)q;              q += sizeof(float) * ((tile_w + 1) & ~1);
  sm.ndc_y = (float*)q;              q += sizeof(float) * ((tile_h + 1) & ~1);
  sm.soft = (unsigned long long*)q;  q += sizeof(unsigned long long) * tpx * n_obj;
  sm.gacc = (unsigned long long*)q;
  sm.defer_n = nullptr;
  return sm;
}
// segment table of the evaluate-once path: behind the accumulators (and the tangent sums of the differentiable kernel)
__device__ __forceinline__ unsigned* tile_smem_seg(const TileSmem& sm, const int tpx, const int n_obj, const bool grad) {
  return (unsigned*)(sm.soft + (size_t)tpx * n_obj * (grad ? 2 : 1));
}

__device__ __forceinline__ void load_geo(const uint32_t* __restrict__ rec, FaceGeo* g) {
  g->x0 = __uint_as_float(rec[0]); g->y0 = __uint_as_float(rec[1]); g->z0 = __uint_as_float(rec[2]);
  g->x1 = __uint_as_float(rec[3]); g->y1 = __uint_as_float(rec[4]); g->z1 = __uint_as_float(rec[5]);
  g->x2 = __uint_as_float(rec[6]); g->y2 = __uint_as_float(rec[7]); g->z2 = __uint_as_float(rec[8]);
  g->area = __uint_as_float(rec[9]);
}

// exact depth of an inside hit -> nearest-face key
__device__ __forceinline__ void hard_update(const TileSmem& sm, const FaceGeo& g, int fidx, int pix, float px,
                                            float py, float b0, float b1, float b2, bool have_bary) {
  if (!have_bary) bary_persp(g, px, py, &b0, &b1, &b2);
  const float pz = b0 * g.z0 + b1 * g.z1 + b2 * g.z2;
  if (!(pz < 0.f)) {
    const unsigned long long key =
        ((unsigned long long)__float_as_uint(pz) << 32) | (unsigned long long)(unsigned)fidx;
    atomicMin(sm.hard + pix, key);
  }
}

// A cut face (REC_CLIP) against the pixels of its tile-clipped blur box, by one warp (rare: geometry within
// z_clip = znear/2 of the camera; never on the fast path).
__device__ __noinline__ void raster_clip_record(const uint4* __restrict__ rec, const int cx0, const int cx1, const int cy0,
                                                const int cy1, unsigned long long* soft_all, unsigned long long* hard,
                                                const float* ndc_x, const float* ndc_y, const int tile_w, const int tpx,
                                                const float z_clip, const int cull, const float blur, const float bbox_r,
                                                const float inv_sigma_log2e, const int lane, const RasterParams* gp,
                                                const int env, unsigned long long* gacc_all) {
  const uint4 q0 = __ldg(rec + 0), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2);
  const float4 va = make_float4(__uint_as_float(q0.x), __uint_as_float(q0.y), __uint_as_float(q0.z), 0.f);
  const float4 vb = make_float4(__uint_as_float(q0.w), __uint_as_float(q1.x), __uint_as_float(q1.y), 0.f);
  const float4 vc = make_float4(__uint_as_float(q1.z), __uint_as_float(q1.w), __uint_as_float(q2.x), 0.f);
  SubTris st;
  clip_subtris(va, vb, vc, z_clip, cull, &st);
  const uint32_t w10 = q2.z;
  SubTrisTan tt;
  if (gp) clip_subtris_tan(*gp, env, (int)(w10 & REC_FIDX_MASK), va, vb, vc, z_clip, &tt);  // differentiable step
  unsigned long long* soft = soft_all + (size_t)((w10 >> REC_OBJ_SHIFT) & 3u) * tpx;
  const int w = cx1 - cx0 + 1, n = w * (cy1 - cy0 + 1);
  for (int i = lane; i < n; i += 32) {
    const int ry = i / w, lx = cx0 + (i - ry * w), ly = cy0 + ry;
    const int pix = ly * tile_w + lx;
    ClipPixel cp;
    eval_clip_pixel(&st, ndc_x[lx], ndc_y[ly], blur, bbox_r, &cp);
    if (cp.soft_hit) {
      const float sd = cp.soft_inside ? -cp.soft_dist : cp.soft_dist;
      soft_accumulate(soft + pix, soft_term(sd, inv_sigma_log2e), cp.hard_hit);
      if (gp) {
        float g0 = 0.f, g1 = 0.f;
        clip_hit_tangent(*gp, &st, &tt, &cp, ndc_x[lx], ndc_y[ly], &g0, &g1);
        grad_accumulate(gacc_all + (size_t)((w10 >> REC_OBJ_SHIFT) & 3u) * tpx + pix, g0, g1);
      }
    }
    if (cp.hard_hit) {
      const unsigned low = (w10 & REC_FIDX_MASK) | KEY_CLIP | ((unsigned)cp.hard_which << 30);
      atomicMin(hard + pix, ((unsigned long long)__float_as_uint(cp.hard_pz) << 32) | (unsigned long long)low);
    }
  }
}

// Clip phase of a tile (after the barrier-free phase; only in envs where the setup kernel cut a face): the warps
// look through the faces of the tile for cut ones and rasterise each warp-wide.
__device__ __noinline__ void clip_phase(const uint4* __restrict__ geo, const uint4* __restrict__ rng,
                                        const int* __restrict__ tidx, const int n_cand, const bool use_tidx, const int tx0,
                                        const int ty0, const int tx1, const int ty1, unsigned long long* soft_all,
                                        unsigned long long* hard, const float* ndc_x, const float* ndc_y, const int tile_w,
                                        const int tpx, const float z_clip, const int cull, const float blur,
                                        const float bbox_r, const float inv_sigma_log2e, const RasterParams* gp,
                                        const int env, unsigned long long* gacc_all) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c0 = warp * 32; c0 < n_cand; c0 += OCCL_THREADS) {
    const int ci = c0 + lane;
    bool cut = false;
    int k = 0, cx0 = 0, cx1 = -1, cy0 = 0, cy1 = -1;
    if (ci < n_cand) {
      k = use_tidx ? tidx[ci] : ci;
      const uint4 rg = __ldg(rng + k);
      cx0 = max((int)(rg.x & 0xffffu), tx0);  cx1 = min((int)(rg.x >> 16), tx1);
      cy0 = max((int)(rg.y & 0xffffu), ty0);  cy1 = min((int)(rg.y >> 16), ty1);
      cut = (rg.z & RNG_CLIP) && cx0 <= cx1 && cy0 <= cy1;
    }
    for (unsigned cb = __ballot_sync(0xffffffffu, cut); cb;) {
      const int bl = __ffs(cb) - 1;
      cb &= cb - 1;
      const int kb = __shfl_sync(0xffffffffu, k, bl);
      const int bx0 = __shfl_sync(0xffffffffu, cx0, bl), bx1 = __shfl_sync(0xffffffffu, cx1, bl);
      const int by0 = __shfl_sync(0xffffffffu, cy0, bl), by1 = __shfl_sync(0xffffffffu, cy1, bl);
      raster_clip_record(geo + (size_t)kb * 4, bx0 - tx0, bx1 - tx0, by0 - ty0, by1 - ty0, soft_all, hard, ndc_x, ndc_y,
                         tile_w, tpx, z_clip, cull, blur, bbox_r, inv_sigma_log2e, lane, gp, env, gacc_all);
    }
  }
}

// Round record of a face (tile-local, shared memory): the HOT part is what one (pixel, face) pair of the fast path
// reads -- four 16-byte loads -- the COLD part (depths and `area`) is only needed for the exact depth of inside hits
// and by the reference-order routine.
//   hot  q0: x0 y0 x1 y1
//        q1: x2 y2 | packed face index, object, flags (REC_*) | hard pixel range, tile-local, 8 bits each: x0 x1 y0 y1
//        q2: RN(1/l2) of the edges v0v1, v0v2, v1v2 | soft box: x0 (bits 0..7), y0 (8..15), width (16..31)
//        q3: l2 of the three edges | 1 / width
//   cold z0 z1 z2 area
struct RoundBuf {
  uint4* hot;      // [R][4]
  float4* cold;    // [R]
  float4* tan;     // [R][3] screen-space tangents of the three vertices (GRAD)
  int* pref;       // [R + 1] exclusive prefix sums of the faces' pixel counts
};

__device__ __forceinline__ void round_geo(const RoundBuf& rb, int f, FaceGeo* g) {
  const uint4 a0 = rb.hot[f * 4 + 0], a1 = rb.hot[f * 4 + 1];
  const float4 cz = rb.cold[f];
  g->x0 = __uint_as_float(a0.x); g->y0 = __uint_as_float(a0.y); g->x1 = __uint_as_float(a0.z); g->y1 = __uint_as_float(a0.w);
  g->x2 = __uint_as_float(a1.x); g->y2 = __uint_as_float(a1.y);
  g->z0 = cz.x; g->z1 = cz.y; g->z2 = cz.z; g->area = cz.w;
}

// Rare paths of the pair loop, out of line so that the loop body stays small (the L0 I-cache is ~6 KB):
// the reference-order evaluation of a pair whose guards failed ...
__device__ __noinline__ PairResult eval_pair_round(const RoundBuf rb, int f, float px, float py) {
  FaceGeo g;
  round_geo(rb, f, &g);
  return eval_pair(g, px, py);
}
// ... and the inline exact depth (depth queue full, or the barycentrics are already there)
__device__ __noinline__ void hard_update_cold(unsigned long long* hard, const RoundBuf rb, int f, int pix, float px, float py,
                                              float b0, float b1, float b2, bool have_bary) {
  FaceGeo g;
  round_geo(rb, f, &g);
  if (!have_bary) bary_persp(g, px, py, &b0, &b1, &b2);
  const float pz = b0 * g.z0 + b1 * g.z1 + b2 * g.z2;
  if (!(pz < 0.f)) {
    const unsigned long long key =
        ((unsigned long long)__float_as_uint(pz) << 32) | (unsigned long long)(rb.hot[f * 4 + 1].z & REC_FIDX_MASK);
    atomicMin(hard + pix, key);
  }
}

// Evaluate-once recording of a hit on a flagged slot: its sort key (clipped-barycentric depth, face index) and its
// term -log2(1 - prob) go to the slot's segment of the CTA's scratch slab.  Out of line: the hot loop only tests a bit.
struct RecCtx {
  const unsigned* seg;          // [n_slots] first entry of the slot's segment
  unsigned long long* keys;     // this CTA's slab
  float* terms;
};
__device__ __noinline__ void record_hit(const RoundBuf rb, const int f, const float px, const float py,
                                        unsigned long long* kdst) {
  FaceGeo g;
  round_geo(rb, f, &g);
  float b0, b1, b2;
  bary_persp(g, px, py, &b0, &b1, &b2);
  *kdst = ((unsigned long long)__float_as_uint(pz_clipped(g, b0, b1, b2)) << 32) |
          (unsigned long long)(rb.hot[f * 4 + 1].z & REC_FIDX_MASK);
}

// One (pixel, face) pair: pixel `i` (row-major) of the tile-clipped blur box of round face `f`.
// REC: a hit on a recording slot stores its term at once and returns (pixel | face << 16, entry) in *rq_word / *rq_pos:
// the caller queues these per warp and computes the sort keys (nine divisions each) 32 at a time, every lane busy.
template <bool GRAD, bool REC>
__device__ __forceinline__ void raster_pair(const RasterParams& p, const TileSmem& sm, const RoundBuf& rb, const int tile_w,
                                            const int tpx, const int f, const int i, const RecCtx& rc, unsigned* rq_word,
                                            int* rq_pos) {
  const uint4 a0 = rb.hot[f * 4 + 0], a1 = rb.hot[f * 4 + 1], a2 = rb.hot[f * 4 + 2], a3 = rb.hot[f * 4 + 3];
  const float x0 = __uint_as_float(a0.x), y0 = __uint_as_float(a0.y), x1 = __uint_as_float(a0.z), y1 = __uint_as_float(a0.w);
  const float x2 = __uint_as_float(a1.x), y2 = __uint_as_float(a1.y);
  const uint32_t w10 = a1.z, sbw = a2.w;
  const int w = (int)(sbw >> 16);
  const int ry = (int)(((float)i + 0.5f) * __uint_as_float(a3.w));
  const int lx = (int)(sbw & 0xffu) + (i - ry * w), ly = (int)((sbw >> 8) & 0xffu) + ry;
  const float px = sm.ndc_x[lx], py = sm.ndc_y[ly];
  const bool fast_face = (w10 & REC_FAST) != 0u && !p.exact_only;
  bool inside, have_bary = false;
  float dist, tt, b0 = 0.f, b1 = 0.f, b2 = 0.f;
  int edge;
  bool need_exact = !fast_face;
  if (fast_face) {
    const float bx01 = x1 - x0, by01 = y1 - y0, bx02 = x2 - x0, by02 = y2 - y0, bx12 = x2 - x1, by12 = y2 - y1;
    const float y01 = __uint_as_float(a2.x), y02 = __uint_as_float(a2.y), y12 = __uint_as_float(a2.z);  // RN(1/l2)
    const float l01 = __uint_as_float(a3.x), l02 = __uint_as_float(a3.y), l12 = __uint_as_float(a3.z);
    const float dx0 = px - x0, dy0 = py - y0, dx1 = px - x1, dy1 = py - y1, dx2 = px - x2, dy2 = py - y2;
    // edge functions in the reference's rounding (separate multiply / subtract: this TU has -fmad=false)
    const float e0 = dx1 * by12 - dy1 * bx12;
    const float e1 = dy2 * bx02 - dx2 * by02;
    const float e2 = dx0 * by01 - dy0 * bx01;
    const float ae0 = fabsf(e0), ae1 = fabsf(e1), ae2 = fabsf(e2);
    const float emax = fmaxf(fmaxf(ae0, ae1), ae2), emin = fminf(fminf(ae0, ae1), ae2);
    // sign(b_i) == sign(e_i) when no quotient can underflow (face guards + this spread guard)
    const bool sign_ok = emax >= 9.313226e-10f /*2^-30*/ && emin >= 9.313226e-10f * emax;
    inside = e0 > 0.f && e1 > 0.f && e2 > 0.f;
    // squared distances to the three edges in the reference's operation order (bit-identical to
    // seg_dist for non-degenerate edges, which the FAST flag guarantees): sigmoid(-d/sigma) has slope
    // 1/sigma = 1e4, so even a 1-ulp change of a projected point would move alpha by > 1e-5 relative
    const float n01 = bx01 * dx0 + by01 * dy0, n02 = bx02 * dx0 + by02 * dy0, n12 = bx12 * dx1 + by12 * dy1;
    const float nmin = fminf(fminf(fabsf(n01), fabsf(n02)), fabsf(n12));
    const float nmax = fmaxf(fmaxf(fabsf(n01), fabsf(n02)), fabsf(n12));
    const bool div_ok = nmin >= 8.271806e-25f /*2^-80*/ && nmax <= 1024.f;  // div_rn_hoisted's domain
    const float t01 = fminf(fmaxf(div_rn_hoisted(n01, l01, y01), 0.0f), 1.0f);
    const float t02 = fminf(fmaxf(div_rn_hoisted(n02, l02, y02), 0.0f), 1.0f);
    const float t12 = fminf(fmaxf(div_rn_hoisted(n12, l12, y12), 0.0f), 1.0f);
    const float ux01 = px - (x0 + t01 * bx01), uy01 = py - (y0 + t01 * by01);
    const float ux02 = px - (x0 + t02 * bx02), uy02 = py - (y0 + t02 * by02);
    const float ux12 = px - (x1 + t12 * bx12), uy12 = py - (y1 + t12 * by12);
    const float d01 = ux01 * ux01 + uy01 * uy01;
    const float d02 = ux02 * ux02 + uy02 * uy02;
    const float d12 = ux12 * ux12 + uy12 * uy12;
    dist = fminf(fminf(d01, d02), d12);
    if (d01 <= d02 && d01 <= d12) { edge = 0; tt = t01; }
    else if (d02 <= d12) { edge = 1; tt = t02; }
    else { edge = 2; tt = t12; }
    // shortcuts taken: inside test by signs (no six divisions), divisions by hoisted reciprocal
    need_exact = !(sign_ok && div_ok);
  }
  if (need_exact) {
    const PairResult r = eval_pair_round(rb, f, px, py);
    inside = r.inside; dist = r.dist; tt = r.t; edge = r.edge;
    b0 = r.b0; b1 = r.b1; b2 = r.b2;
    have_bary = true;
  }
  if (!inside && dist >= p.blur) return;
  const int pix = ly * tile_w + lx;
  const float sd = inside ? -dist : dist;
  bool hard_ok = false;
  if (inside) {
    const uint32_t hb = a1.w;
    hard_ok = lx >= (int)(hb & 0xff) && lx <= (int)((hb >> 8) & 0xff) && ly >= (int)((hb >> 16) & 0xff) &&
              ly <= (int)((hb >> 24) & 0xff);
  }
  const int obj = (int)((w10 >> REC_OBJ_SHIFT) & 3u);
  if (REC) {
    const int slot = obj * tpx + pix;
    const float term = soft_term(sd, p.inv_sigma_log2e);
    const unsigned old = soft_accumulate_ret(sm.soft + slot, term, hard_ok);
    if (old & SOFT_ROUND) {
      const unsigned pos = rc.seg[slot] + (old & SOFT_CNT_MASK);
      rc.terms[pos] = term;
      *rq_word = (unsigned)pix | ((unsigned)f << 16);
      *rq_pos = (int)pos;
    }
  } else {
    soft_accumulate(sm.soft + (size_t)obj * tpx + pix, soft_term(sd, p.inv_sigma_log2e), hard_ok);
  }
  if (hard_ok) {
    bool queued = false;
    if (!have_bary) {
      const int d = atomicAdd(sm.defer_n, 1);
      if (d < WDEFER_CAP) {
        sm.defer[d] = (uint32_t)pix | ((uint32_t)f << 16);
        queued = true;
      }
    }
    if (!queued) hard_update_cold(sm.hard, rb, f, pix, px, py, b0, b1, b2, have_bary);
  }
  if (GRAD) {
    // d signed_dist / d theta through the nearest edge, vertices move, pixel fixed
    const int ia = edge == 2 ? 1 : 0, ib = edge == 0 ? 1 : 2;
    const float4 da = rb.tan[f * 3 + ia], db = rb.tan[f * 3 + ib];
    const float ax = ia ? x1 : x0, ay = ia ? y1 : y0, bx = ib == 1 ? x1 : x2, by = ib == 1 ? y1 : y2;
    const float qx = ax + tt * (bx - ax), qy = ay + tt * (by - ay);
    const float sgn = inside ? -1.f : 1.f;
    const float gx = sgn * 2.f * (qx - px), gy = sgn * 2.f * (qy - py);
    const float wa = 1.f - tt, wb = tt;
    const float dsd_el = gx * (wa * da.x + wb * db.x) + gy * (wa * da.y + wb * db.y);
    const float dsd_az = gx * (wa * da.z + wb * db.z) + gy * (wa * da.w + wb * db.w);
    const float prob = rcp_approx(1.0f + ex2_approx(sd * p.inv_sigma_log2e));  // sigmoid(-sd/sigma), ~2 ulp
    const float k = prob * p.inv_sigma;
    grad_accumulate(sm.gacc + (size_t)obj * tpx + pix, k * dsd_el, k * dsd_az);
  }
}

// Tangent terms of one soft hit, re-evaluated from the face index (K-overflow selection of the differentiable
// kernel): adds p_k/sigma * d(signed dist)/d(el, az)  (SURVEY A.7).
__device__ __noinline__ void hit_tangent(const RasterParams& p, int env, int f, float px, float py, float* g0, float* g1) {
  const float4* __restrict__ vp = p.vproj + (size_t)env * p.V;
  const int* __restrict__ faces = p.faces + (size_t)env * p.faces_stride;
  const int i0 = __ldg(faces + 3 * f + 0), i1 = __ldg(faces + 3 * f + 1), i2 = __ldg(faces + 3 * f + 2);
  FaceGeo g;
  bool straddles = false;
  const float4 va = __ldg(vp + i0), vb = __ldg(vp + i1), vc = __ldg(vp + i2);
  face_geo(va, vb, vc, p.cull, p.z_clip, &g, &straddles);
  if (straddles) {  // a face cut at z_clip: the hit is on one of its cut triangles
    SubTris st;
    clip_subtris(va, vb, vc, p.z_clip, p.cull, &st);
    ClipPixel cp;
    eval_clip_pixel(&st, px, py, p.blur, p.bbox_r, &cp);
    if (cp.soft_hit) {
      SubTrisTan tt;
      clip_subtris_tan(p, env, f, va, vb, vc, p.z_clip, &tt);
      clip_hit_tangent(p, &st, &tt, &cp, px, py, g0, g1);
    }
    return;
  }
  const PairResult r = eval_pair(g, px, py);
  const float prob = soft_prob(r.inside ? -r.dist : r.dist, p.sigma);
  const float4* __restrict__ vt = p.vtan + (size_t)env * p.V;
  const float4 ta = __ldg(vt + i0), tb = __ldg(vt + i1), tc = __ldg(vt + i2);
  float ax, ay, bx, by;
  float4 da, db;
  if (r.edge == 0) { ax = g.x0; ay = g.y0; bx = g.x1; by = g.y1; da = ta; db = tb; }
  else if (r.edge == 1) { ax = g.x0; ay = g.y0; bx = g.x2; by = g.y2; da = ta; db = tc; }
  else { ax = g.x1; ay = g.y1; bx = g.x2; by = g.y2; da = tb; db = tc; }
  const float qx = ax + r.t * (bx - ax), qy = ay + r.t * (by - ay);
  const float sgn = r.inside ? -1.f : 1.f;
  const float gx = sgn * 2.f * (qx - px), gy = sgn * 2.f * (qy - py);
  const float wa = 1.f - r.t, wb = r.t;
  const float kk = prob / p.sigma;
  *g0 += kk * (gx * (wa * da.x + wb * db.x) + gy * (wa * da.y + wb * db.y));
  *g1 += kk * (gx * (wa * da.z + wb * db.z) + gy * (wa * da.w + wb * db.w));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) { return __reduce_add_sync(0xffffffffu, v); }

// K-overflow resolution of one tile.  Inlined into the kernel (a non-inlined call measured 5-10 % slower on the
// main phase: ptxas' register allocation of the barrier-free loop is sensitive to what surrounds it); everything
// is recomputed here so that nothing extra stays live across the caller's main phase.
template <bool GRAD, int TW, int TH, bool CLIPF, bool REC>
__device__ __forceinline__ void koverflow_resolve(const RasterParams p, const int env, const int tile, const int n_tidx,
                                                  const RecCtx rc, unsigned char* smem_raw /* the caller's opaque base */) {
  const int tile_w = TW ? TW : p.tile_w, tile_h = TH ? TH : p.tile_h;
  const int tpx = tile_w * tile_h;
  const int tx0 = (tile % p.tiles_x) * tile_w, ty0 = (tile / p.tiles_x) * tile_h;
  const TileSmem sm = tile_smem_layout(smem_raw, tile_w, tile_h, p.n_obj);
  const int n_live = p.n_live[env];
  const int* __restrict__ tidx = p.tile_idx + ((size_t)env * (p.tiles_x * p.tiles_y) + tile) * p.tidx_cap;
  __shared__ double s_red[OCCL_WARPS][3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint4* __restrict__ geo = p.geo + (size_t)env * p.F * 4;
  const uint4* __restrict__ rng = p.rng + (size_t)env * p.F;
  __shared__ int s_ovf_n, s_hit_n;
  __shared__ int s_ovf[OVF_CAP];
  if (tid == 0) { s_ovf_n = 0; s_hit_n = 0; }
  // ---- pixels with more than K hits, hit-list rounds -------------------------------------------
  // The reference keeps the K nearest hits by (pz_clipped, face index).  A round takes the next unresolved
  // (pixel, object) slots -- in slot order, as many as their exactly known hit counts fit HITBUF_CAP -- and
  //  1. marks them (bit 29; the product word becomes the write cursor of the slot's segment),
  //  2. pairs every marked slot with the faces of the tile whose blur box holds its pixel (per-warp queues),
  //     re-evaluates the pairs 32 at a time with the reference-order routine and appends (key, term) of
  //     every hit to the slot's segment,
  //  3. one warp per slot finds the K-th smallest key by bisection over the key bits (early exit once the
  //     K smallest are separated) and sums the (fixed-point) terms below it: product = 2^-sum.
  // Slots with more hits than a whole round holds are left to the one-pixel path below.
  {
    unsigned long long* hkey = (unsigned long long*)sm.list;  // [HITBUF_CAP]
    float* hq = (float*)(hkey + HITBUF_CAP);                    // [HITBUF_CAP]
    static_assert((size_t)HITBUF_CAP * 12 + 4 * OCCL_WARPS * WQ_CAP <= (size_t)4 * OCCL_WARPS * (WBUF_RECS * REC_WORDS + WDEFER_CAP), "hit buffer + pair queues must fit the face list");
    __shared__ int s_rn, s_todo, s_huge, s_rb[4], s_wsum[OCCL_WARPS];
    __shared__ unsigned s_robj;
    int* s_rslot = (int*)sm.big;                                        // [RSLOT_CAP]
    unsigned short* s_roff = (unsigned short*)(s_rslot + RSLOT_CAP);    // [RSLOT_CAP]
    static_assert(RSLOT_CAP * 6 <= 4 * BIG_CAP * REC_WORDS, "round table must fit the small scratch region");
    const int n_slots = tpx * p.n_obj;
    const int per = (n_slots + OCCL_THREADS - 1) / OCCL_THREADS;
    const int s_begin = min(tid * per, n_slots), s_end = min(s_begin + per, n_slots);
    const bool use_tidx = n_tidx <= p.tidx_cap;
    const int n_cand = use_tidx ? n_tidx : n_live;
    // round counters: set here, and again by thread 0 while a round's selections run (everybody has read them by then)
    if (tid == 0) { s_rn = 0; s_todo = 0; s_huge = 0; s_rb[0] = 1 << 30; s_rb[1] = 1 << 30; s_rb[2] = -1; s_rb[3] = -1; s_robj = 0u; }
    {
      bool any = false;
      for (int i = tid; i < n_slots; i += OCCL_THREADS) any |= (int)((unsigned)(sm.soft[i] >> 32) & SOFT_CNT_MASK) > p.K;
      if (!__syncthreads_or(any)) return;  // the common tile: no pixel has more than K hits
    }
    // ---- recorded slots (evaluate-once): the hits are in the slab, a warp per slot only selects --------------------
    if (REC && rc.keys != nullptr) {
      constexpr int LIST_CAP = BIG_CAP * REC_WORDS;
      int* s_list = (int*)sm.big;
      __shared__ int s_ln, s_lnext, s_lmore;
      for (;;) {
        if (tid == 0) { s_ln = 0; s_lnext = 0; s_lmore = 0; }
        __syncthreads();
        for (int i = s_begin; i < s_end; ++i) {
          const unsigned hi = (unsigned)(sm.soft[i] >> 32);
          const int cnt = (int)(hi & SOFT_CNT_MASK);
          if (cnt > p.K && (hi & SOFT_ROUND) && !(hi & SOFT_RESOLVED)) {
            if (soft_strong_shortcut(hi, p.K)) {
              sm.soft[i] = ((unsigned long long)((hi & ~SOFT_ROUND) | SOFT_RESOLVED) << 32);  // product := +0.0f
              if (GRAD) sm.gacc[i] = 0ull;
              atomicOr(p.status + env, OCCL_ST_KOVERFLOW);
            } else {
              const int pos = atomicAdd(&s_ln, 1);
              if (pos < LIST_CAP) s_list[pos] = i; else s_lmore = 1;
            }
          }
        }
        __syncthreads();
        const int ln = min(s_ln, LIST_CAP);
        const bool more = s_lmore != 0;
        if (ln > 0) {
          if (tid == 0) atomicOr(p.status + env, OCCL_ST_KOVERFLOW);
          for (;;) {
            int li = 0;
            if (lane == 0) li = atomicAdd(&s_lnext, 1);
            li = __shfl_sync(0xffffffffu, li, 0);
            if (li >= ln) break;
            const int slot = s_list[li];
            const unsigned hi = (unsigned)(sm.soft[slot] >> 32);
            const int nn = (int)(hi & SOFT_CNT_MASK);
            const unsigned long long* keys = rc.keys + rc.seg[slot];
            const float* __restrict__ terms = rc.terms + rc.seg[slot];
            const int obj = slot / tpx, pix = slot - obj * tpx;
            const int ly = pix / tile_w, lx = pix - ly * tile_w;
            const int want = p.K - 1;           // 0-based rank of the last key kept (nn > K here)
            unsigned long long lsum = 0ull;
            float g0 = 0.f, g1 = 0.f;
            constexpr int KR = 8;               // keys per lane held in registers: slots of up to 256 hits
            if (nn <= 32 * KR) {
              // the slot's keys are read ONCE; the ~12 counting passes of the bisection run on registers
              unsigned long long kr[KR];
              unsigned long long vand = ~0ull, vor = 0ull;
#pragma unroll
              for (int t = 0; t < KR; ++t) {
                const int i = lane + 32 * t;
                kr[t] = i < nn ? keys[i] : ~0ull;   // the padding is larger than every candidate bound
                if (i < nn) { vand &= kr[t]; vor |= kr[t]; }
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                vand &= __shfl_xor_sync(0xffffffffu, vand, o);
                vor |= __shfl_xor_sync(0xffffffffu, vor, o);
              }
              const unsigned long long diff = vand ^ vor;
              unsigned long long bound = ~0ull;   // keys < bound are kept
              if (diff) {
                const int top = 63 - __clzll((long long)diff);
                unsigned long long prefix = top == 63 ? 0ull : (vor & ~((2ull << top) - 1ull));
                bool found = false;
                for (int bit = top; bit >= 0; --bit) {
                  const unsigned long long m = 1ull << bit;
                  if (!(diff & m)) { prefix |= vand & m; continue; }
                  const unsigned long long cand = prefix | m;
                  int c = 0;
#pragma unroll
                  for (int t = 0; t < KR; ++t) c += kr[t] < cand ? 1 : 0;
                  c = __reduce_add_sync(0xffffffffu, c);
                  if (c <= want) prefix = cand;
                  else if (c == want + 1) { bound = cand; found = true; break; }
                }
                if (!found) bound = prefix + 1ull;
              }
#pragma unroll
              for (int t = 0; t < KR; ++t) {
                const int i = lane + 32 * t;
                if (i < nn && kr[t] < bound) {
                  lsum += soft_term_fx(terms[i]);
                  if (GRAD) hit_tangent(p, env, (int)(kr[t] & 0xffffffffull), sm.ndc_x[lx], sm.ndc_y[ly], &g0, &g1);
                }
              }
            } else {
              // more keys than the registers hold: up to 384 are staged in this warp's share of the (idle) round buffers,
              // beyond that the counting passes read the slab through L1
              constexpr int WK = (4 * WBUF_RECS * REC_WORDS) / 8;  // u64 entries per warp
              if (nn <= WK) {
                unsigned long long* wk = (unsigned long long*)sm.list + warp * WK;
                for (int i = lane; i < nn; i += 32) wk[i] = keys[i];
                __syncwarp();
                keys = wk;
              }
              unsigned long long vand = ~0ull, vor = 0ull;
              for (int i = lane; i < nn; i += 32) { const unsigned long long kx = keys[i]; vand &= kx; vor |= kx; }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                vand &= __shfl_xor_sync(0xffffffffu, vand, o);
                vor |= __shfl_xor_sync(0xffffffffu, vor, o);
              }
              const unsigned long long diff = vand ^ vor;
              unsigned long long bound = ~0ull;
              if (diff) {
                const int top = 63 - __clzll((long long)diff);
                unsigned long long prefix = top == 63 ? 0ull : (vor & ~((2ull << top) - 1ull));
                bool found = false;
                for (int bit = top; bit >= 0; --bit) {
                  const unsigned long long m = 1ull << bit;
                  if (!(diff & m)) { prefix |= vand & m; continue; }
                  const unsigned long long cand = prefix | m;
                  int c = 0;
                  for (int i = lane; i < nn; i += 32) c += keys[i] < cand ? 1 : 0;
                  c = __reduce_add_sync(0xffffffffu, c);
                  if (c <= want) prefix = cand;
                  else if (c == want + 1) { bound = cand; found = true; break; }
                }
                if (!found) bound = prefix + 1ull;
              }
              for (int i = lane; i < nn; i += 32) {
                const unsigned long long key = keys[i];
                if (!(key < bound)) continue;
                lsum += soft_term_fx(terms[i]);
                if (GRAD) hit_tangent(p, env, (int)(key & 0xffffffffull), sm.ndc_x[lx], sm.ndc_y[ly], &g0, &g1);
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              lsum += __shfl_down_sync(0xffffffffu, lsum, o);
              if (GRAD) { g0 += __shfl_down_sync(0xffffffffu, g0, o); g1 += __shfl_down_sync(0xffffffffu, g1, o); }
            }
            if (lane == 0) {
              sm.soft[slot] = ((unsigned long long)((hi & ~SOFT_ROUND) | SOFT_RESOLVED) << 32) |
                              (unsigned long long)__float_as_uint(soft_product_fx(lsum));
              if (GRAD) sm.gacc[slot] = pack2f(g0, g1);
            }
            __syncwarp();
          }
        }
        __syncthreads();
        if (!more) break;
      }
    }
    for (;;) {
      int loc = 0, todo = 0, huge = 0;
      for (int i = s_begin; i < s_end; ++i) {
        const unsigned hi = (unsigned)(sm.soft[i] >> 32);
        const int cnt = (int)(hi & SOFT_CNT_MASK);
        if (cnt > p.K && !(hi & SOFT_RESOLVED)) {
          // Whatever K hits are the nearest, at most `weak` of them are weak; if the others number at least
          // SOFT_STRONG_NEEDED their factors (each <= 1/4) already push the product below 2^-25, i.e. the
          // reference's alpha is exactly 1.0f: no selection needed.
          if (soft_strong_shortcut(hi, p.K)) {
            sm.soft[i] = ((unsigned long long)(hi | SOFT_RESOLVED) << 32);  // product := +0.0f
            if (GRAD) sm.gacc[i] = 0ull;
            atomicOr(p.status + env, OCCL_ST_KOVERFLOW);
          } else if (cnt <= HITBUF_CAP) {
            loc += cnt;
            ++todo;
          } else {
            ++huge;
          }
        }
      }
      if (todo) atomicAdd(&s_todo, todo);
      if (huge) atomicAdd(&s_huge, huge);
      int incl = loc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) s_wsum[warp] = incl;
      __syncthreads();
      int off = incl - loc;
      for (int w = 0; w < warp; ++w) off += s_wsum[w];
      for (int i = s_begin; i < s_end && off < HITBUF_CAP; ++i) {
        const unsigned hi = (unsigned)(sm.soft[i] >> 32);
        const int cnt = (int)(hi & SOFT_CNT_MASK);
        if (cnt > p.K && !(hi & SOFT_RESOLVED) && cnt <= HITBUF_CAP) {
          if (off + cnt <= HITBUF_CAP) {
            const int r = atomicAdd(&s_rn, 1);
            if (r < RSLOT_CAP) {
              s_rslot[r] = i;
              s_roff[r] = (unsigned short)off;
              const unsigned h2 = hi | SOFT_ROUND;
              sm.soft[i] = ((unsigned long long)h2 << 32) | (unsigned long long)(unsigned)off;
              const int obj = i / tpx, pix = i - obj * tpx;
              const int ly = pix / tile_w, lx = pix - ly * tile_w;
              atomicMin(&s_rb[0], lx); atomicMin(&s_rb[1], ly); atomicMax(&s_rb[2], lx); atomicMax(&s_rb[3], ly);
              atomicOr(&s_robj, 1u << obj);
            }
          }
          off += cnt;
        }
      }
      __syncthreads();
      const int rn = p.F <= (1 << 25) ? min(s_rn, RSLOT_CAP) : 0;  // queue entries hold 25 bits of face index
      const bool last_round = s_todo == rn;  // nothing left for another round
      const bool none_huge = s_huge == 0;
      if (rn == 0) {
        if (s_todo == 0 && none_huge) return;  // the strong-hit shortcut settled everything
        break;
      }
      if (tid == 0) atomicOr(p.status + env, OCCL_ST_KOVERFLOW);
      // 2. (slot, face) pairs whose blur box holds the slot's pixel: lanes <-> faces of the tile, the round's
      //    slots in turn; the pairs are compacted into a per-warp queue and evaluated 32 at a time with the
      //    reference-order routine (every lane busy), hits are appended to the slot's segment.
      {
        const int bx0 = tx0 + s_rb[0], by0 = ty0 + s_rb[1], bx1 = tx0 + s_rb[2], by1 = ty0 + s_rb[3];
        const unsigned robj = s_robj;
        uint32_t* wq = (uint32_t*)(hq + HITBUF_CAP) + warp * WQ_CAP;
        int qn = 0;  // warp-uniform
        auto drain = [&](const bool flush) {
          while (qn >= 32 || (flush && qn > 0)) {
            const int take = min(qn, 32);
            qn -= take;
            if (lane < take) {
              const uint32_t ent = wq[qn + lane];
              const int slot = s_rslot[ent >> 25];
              const int obj = slot / tpx, pix = slot - obj * tpx;
              const int ly = pix / tile_w, lx = pix - ly * tile_w;
              const uint4* __restrict__ srcg = geo + (size_t)(ent & 0x1ffffffu) * 4;
              const uint4 q0 = __ldg(srcg + 0), q1 = __ldg(srcg + 1), q2 = __ldg(srcg + 2);
              FaceGeo g;
              g.x0 = __uint_as_float(q0.x); g.y0 = __uint_as_float(q0.y); g.z0 = __uint_as_float(q0.z);
              g.x1 = __uint_as_float(q0.w); g.y1 = __uint_as_float(q1.x); g.z1 = __uint_as_float(q1.y);
              g.x2 = __uint_as_float(q1.z); g.y2 = __uint_as_float(q1.w); g.z2 = __uint_as_float(q2.x);
              g.area = __uint_as_float(q2.y);
              const PairResult r = eval_pair(g, sm.ndc_x[lx], sm.ndc_y[ly]);
              if (r.inside || r.dist < p.blur) {
                const float pz = pz_clipped(g, r.b0, r.b1, r.b2);
                const unsigned pos = atomicAdd((unsigned*)(sm.soft + slot), 1u);  // low word = cursor
                if (pos < HITBUF_CAP) {
                  hkey[pos] = ((unsigned long long)__float_as_uint(pz) << 32) | (unsigned long long)(q2.z & REC_FIDX_MASK);
                  hq[pos] = soft_term(r.inside ? -r.dist : r.dist, p.inv_sigma_log2e);
                }
              }
            }
            __syncwarp();
          }
        };
        for (int c0 = warp * 32; c0 < n_cand; c0 += OCCL_THREADS) {
          const int ci = c0 + lane;
          bool ok = false;
          int k = 0;
          uint4 rg = make_uint4(0, 0, 0, 0);
          if (ci < n_cand) {
            k = use_tidx ? tidx[ci] : ci;
            rg = __ldg(rng + k);
            ok = !(bx1 < (int)(rg.x & 0xffffu) || bx0 > (int)(rg.x >> 16) || by1 < (int)(rg.y & 0xffffu) || by0 > (int)(rg.y >> 16)) &&
                 ((robj >> (rg.w >> 30)) & 1u) && !(CLIPF && (rg.z & RNG_CLIP));  // cut faces: separate pass below
          }
          if (!__any_sync(0xffffffffu, ok)) continue;
          const int fx0 = (int)(rg.x & 0xffffu) - tx0, fx1 = (int)(rg.x >> 16) - tx0;
          const int fy0 = (int)(rg.y & 0xffffu) - ty0, fy1 = (int)(rg.y >> 16) - ty0;
          const int fobj = (int)(rg.w >> 30);
          for (int r = 0; r < rn; ++r) {
            const int slot = s_rslot[r];
            const int obj = slot / tpx, pix = slot - obj * tpx;
            const int ly = pix / tile_w, lx = pix - ly * tile_w;
            const bool in = ok && obj == fobj && lx >= fx0 && lx <= fx1 && ly >= fy0 && ly <= fy1;
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (!bal) continue;
            if (in) wq[qn + __popc(bal & ((1u << lane) - 1u))] = ((uint32_t)r << 25) | (uint32_t)k;
            qn += __popc(bal);
            __syncwarp();
            drain(false);
          }
        }
        drain(true);
        // cut faces (z-clip) of the tile, if the env has any: one lane per face, every slot of the round (rare)
        if (CLIPF) {
          for (int ci = tid; ci < n_cand; ci += OCCL_THREADS) {
            const int k = use_tidx ? tidx[ci] : ci;
            const uint4 rg = __ldg(rng + k);
            if ((rg.z & RNG_CLIP) && ((robj >> (rg.w >> 30)) & 1u) &&
                !(bx1 < (int)(rg.x & 0xffffu) || bx0 > (int)(rg.x >> 16) || by1 < (int)(rg.y & 0xffffu) || by0 > (int)(rg.y >> 16)))
              clip_round_hits(geo + (size_t)k * 4, rg, s_rslot, rn, sm.soft, hkey, hq, sm.ndc_x, sm.ndc_y, tx0, ty0, tile_w, tpx,
                              p.z_clip, p.cull, p.blur, p.bbox_r, p.inv_sigma_log2e);
          }
        }
      }
      __syncthreads();
      if (tid == 0) { s_rn = 0; s_todo = 0; s_huge = 0; s_rb[0] = 1 << 30; s_rb[1] = 1 << 30; s_rb[2] = -1; s_rb[3] = -1; s_robj = 0u; }
      // 3. one warp per slot: K-th smallest key, sum of the terms up to it
      for (int r = warp; r < rn; r += OCCL_WARPS) {
        const int slot = s_rslot[r];
        const int off = (int)s_roff[r];
        const unsigned long long w = sm.soft[slot];
        const unsigned hi = (unsigned)(w >> 32);
        const int cnt = (int)(hi & SOFT_CNT_MASK);
        const int got = (int)(unsigned)(w & 0xffffffffull) - off;
        if (got != cnt && lane == 0) atomicOr(p.status + env, OCCL_ST_HITCAP);  // re-evaluation disagrees with the main phase
        const int nn = max(min(got, min(cnt, HITBUF_CAP - off)), 0);
        const unsigned long long* __restrict__ keys = hkey + off;
        unsigned long long vand = ~0ull, vor = 0ull;
        for (int i = lane; i < nn; i += 32) { const unsigned long long kx = keys[i]; vand &= kx; vor |= kx; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          vand &= __shfl_xor_sync(0xffffffffu, vand, o);
          vor |= __shfl_xor_sync(0xffffffffu, vor, o);
        }
        const unsigned long long diff = vand ^ vor;
        const int want = min(p.K, nn) - 1;  // 0-based rank of the last key kept
        unsigned long long bound = ~0ull;   // keys < bound are kept
        if (nn > p.K && diff) {
          const int top = 63 - __clzll((long long)diff);
          unsigned long long prefix = top == 63 ? 0ull : (vor & ~((2ull << top) - 1ull));
          bool found = false;
          for (int bit = top; bit >= 0; --bit) {
            const unsigned long long m = 1ull << bit;
            if (!(diff & m)) { prefix |= vand & m; continue; }
            const unsigned long long cand = prefix | m;
            int c = 0;
            for (int i = lane; i < nn; i += 32) c += keys[i] < cand ? 1 : 0;
            c = __reduce_add_sync(0xffffffffu, c);
            if (c <= want) prefix = cand;
            else if (c == want + 1) { bound = cand; found = true; break; }
          }
          if (!found) bound = prefix + 1ull;
        }
        unsigned long long ls = 0ull;
        float g0 = 0.f, g1 = 0.f;
        const int obj = slot / tpx, pix = slot - obj * tpx;
        const int ly = pix / tile_w, lx = pix - ly * tile_w;
        for (int i = lane; i < nn; i += 32) {
          const unsigned long long key = keys[i];
          if (!(key < bound)) continue;
          ls += soft_term_fx(hq[off + i]);
          if (GRAD) hit_tangent(p, env, (int)(key & 0xffffffffull), sm.ndc_x[lx], sm.ndc_y[ly], &g0, &g1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ls += __shfl_down_sync(0xffffffffu, ls, o);
          if (GRAD) { g0 += __shfl_down_sync(0xffffffffu, g0, o); g1 += __shfl_down_sync(0xffffffffu, g1, o); }
        }
        if (lane == 0) {
          sm.soft[slot] = ((unsigned long long)((hi & ~SOFT_ROUND) | SOFT_RESOLVED) << 32) |
                          (unsigned long long)__float_as_uint(soft_product_fx(ls));
          if (GRAD) sm.gacc[slot] = pack2f(g0, g1);
        }
      }
      __syncthreads();
      if (last_round) {
        if (none_huge) return;
        break;
      }
    }
  }

  // ---- pixels with more hits than a round holds: one pixel at a time, whole CTA -----------------
  // Rounds of at most OVF_CAP pixels; a resolved pixel carries bit 30 of its count word.
  for (bool more = true; more;) {
  for (int i = tid; i < tpx * p.n_obj; i += OCCL_THREADS) {
    const unsigned hi = (unsigned)(sm.soft[i] >> 32);
    if ((int)(hi & SOFT_CNT_MASK) > p.K && !(hi & SOFT_RESOLVED)) {
      // Whatever K hits are the nearest, at most `weak` of them are weak; if the others number at least
      // SOFT_STRONG_NEEDED their factors (each <= 1/4) already push the product below 2^-25, i.e. the
      // reference's alpha is exactly 1.0f: no selection needed.
      if (soft_strong_shortcut(hi, p.K)) {
        sm.soft[i] = ((unsigned long long)(hi | SOFT_RESOLVED) << 32);  // product := +0.0f
        if (GRAD) sm.gacc[i] = 0ull;
        atomicOr(p.status + env, OCCL_ST_KOVERFLOW);
      } else {
        const int s = atomicAdd(&s_ovf_n, 1);
        if (s < OVF_CAP) s_ovf[s] = i;
      }
    }
  }
  __syncthreads();
  int n_ovf = s_ovf_n;
  more = n_ovf > OVF_CAP;
  if (n_ovf > 0) {
    if (tid == 0) atomicOr(p.status + env, OCCL_ST_KOVERFLOW);
    n_ovf = min(n_ovf, OVF_CAP);
    // candidates: the tile's own face list if it fitted, else the env's whole live list
    const bool use_tidx = n_tidx <= p.tidx_cap;
    const int n_cand = use_tidx ? n_tidx : n_live;
    // selection buffers alias the (now idle) face list
    // deterministic order of the overflow list (atomicAdd order is not): sort the small list
    if (tid == 0) {
      for (int a = 1; a < n_ovf; ++a) {
        const int v = s_ovf[a];
        int b = a - 1;
        while (b >= 0 && s_ovf[b] > v) { s_ovf[b + 1] = s_ovf[b]; --b; }
        s_ovf[b + 1] = v;
      }
    }
    __syncthreads();
    // One pixel at a time:
    //  A. the faces of this tile whose blur box holds the pixel are compacted into a candidate list,
    //  B. the candidates are evaluated densely (every thread has work) -> sort key (pz_clipped, face) and factor,
    //  C. the K-th smallest key is found (rank counting for short lists, 8-bit radix select otherwise),
    //  D. the terms at or below that key are summed in fixed point (the differentiable kernel re-evaluates those hits
    //     for their tangent terms).
    for (int oi = 0; oi < n_ovf; ++oi) {
      const int slot = s_ovf[oi];
      if (slot < 0) continue;
      const int obj = slot / tpx;
      const int pix = slot - obj * tpx;
      const int ly = pix / tile_w, lx = pix - ly * tile_w;
      const int xi = tx0 + lx, yi = ty0 + ly;
      const float px = sm.ndc_x[lx], py = sm.ndc_y[ly];
      static_assert((size_t)CAND_CAP * 12 <= (size_t)4 * OCCL_WARPS * (WBUF_RECS * REC_WORDS + WDEFER_CAP), "candidate buffers must fit the selection buffer");
      unsigned long long* ckey = (unsigned long long*)sm.list;  // [CAND_CAP]
      float* cq = (float*)(ckey + CAND_CAP);                    // [CAND_CAP] candidate index, then its term -log2(1 - prob)
      __shared__ unsigned long long s_tau;
      __shared__ int s_hist[256];
      __shared__ int s_sel[2];
      if (tid == 0) { s_hit_n = 0; s_sel[0] = 0; }
      __syncthreads();
      // A
      for (int c0 = 0; c0 < n_cand; c0 += OCCL_THREADS) {
        const int ci = c0 + tid;
        bool ok = false;
        int k = 0;
        if (ci < n_cand) {
          k = use_tidx ? tidx[ci] : ci;
          const uint4 rg = __ldg(rng + k);
          ok = !(xi < (int)(rg.x & 0xffffu) || xi > (int)(rg.x >> 16) || yi < (int)(rg.y & 0xffffu) || yi > (int)(rg.y >> 16)) &&
               (int)(rg.w >> 30) == obj;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        int base0 = 0;
        if (lane == 0 && bal) base0 = atomicAdd(&s_hit_n, __popc(bal));
        base0 = __shfl_sync(0xffffffffu, base0, 0);
        if (ok) {
          const int pos = base0 + __popc(bal & ((1u << lane) - 1u));
          if (pos < CAND_CAP) ((int*)cq)[pos] = k;
        }
      }
      __syncthreads();
      int nc = s_hit_n;
      if (nc > CAND_CAP) {
        if (tid == 0) atomicOr(p.status + env, OCCL_ST_HITCAP);
        nc = CAND_CAP;
      }
      // B
      int my_hits = 0;
      for (int i = tid; i < nc; i += OCCL_THREADS) {
        const int k = ((const int*)cq)[i];
        const uint4* __restrict__ src = geo + (size_t)k * 4;
        const uint4 q0 = __ldg(src + 0), q1 = __ldg(src + 1), q2 = __ldg(src + 2);
        FaceGeo g;
        g.x0 = __uint_as_float(q0.x); g.y0 = __uint_as_float(q0.y); g.z0 = __uint_as_float(q0.z);
        g.x1 = __uint_as_float(q0.w); g.y1 = __uint_as_float(q1.x); g.z1 = __uint_as_float(q1.y);
        g.x2 = __uint_as_float(q1.z); g.y2 = __uint_as_float(q1.w); g.z2 = __uint_as_float(q2.x);
        g.area = __uint_as_float(q2.y);
        bool hit, inside;
        float dist, pz = 0.f;
        if (CLIPF && (q2.z & REC_CLIP)) {
          hit = clip_soft_eval(src, px, py, p.z_clip, p.cull, p.blur, p.bbox_r, &inside, &dist, &pz);
        } else {
          const PairResult r = eval_pair(g, px, py);
          inside = r.inside; dist = r.dist;
          hit = r.inside || r.dist < p.blur;
          if (hit) pz = pz_clipped(g, r.b0, r.b1, r.b2);
        }
        unsigned long long key = ~0ull;
        float q = 0.0f;  // the hit's term -log2(1 - prob)
        if (hit) {
          key = ((unsigned long long)__float_as_uint(pz) << 32) | (unsigned long long)(q2.z & REC_FIDX_MASK);
          q = soft_term(inside ? -dist : dist, p.inv_sigma_log2e);
          ++my_hits;
        }
        ckey[i] = key;
        cq[i] = q;
      }
      if (my_hits) atomicAdd(&s_sel[0], my_hits);
      __syncthreads();
      const int nh = s_sel[0];
      int want = min(p.K, nh) - 1;  // 0-based rank of the last key kept
      // C
      if (nc <= 512) {
        for (int a = tid; a < nc; a += OCCL_THREADS) {
          const unsigned long long ka = ckey[a];
          int rank = 0;
          for (int b = 0; b < nc; ++b) rank += ckey[b] < ka;
          if (rank == want && ka != ~0ull) s_tau = ka;
        }
        __syncthreads();
      } else {
        unsigned long long prefix = 0ull, mask = 0ull;
        for (int shift = 56; shift >= 0; shift -= 8) {
          for (int b = tid; b < 256; b += OCCL_THREADS) s_hist[b] = 0;
          __syncthreads();
          for (int i = tid; i < nc; i += OCCL_THREADS) {
            const unsigned long long key = ckey[i];
            if ((key & mask) == prefix) atomicAdd(&s_hist[(int)((key >> shift) & 255ull)], 1);
          }
          __syncthreads();
          if (warp == 0) {
            int cnt8[8], sum = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) { cnt8[b] = s_hist[lane * 8 + b]; sum += cnt8[b]; }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int t = __shfl_up_sync(0xffffffffu, incl, o);
              if (lane >= o) incl += t;
            }
            const int excl = incl - sum;
            if (want >= excl && want < incl) {
              int r = want - excl, b = 0;
              while (r >= cnt8[b]) { r -= cnt8[b]; ++b; }
              s_sel[0] = lane * 8 + b;
              s_sel[1] = r;
            }
          }
          __syncthreads();
          prefix |= (unsigned long long)s_sel[0] << shift;
          mask |= 255ull << shift;
          want = s_sel[1];
          __syncthreads();
        }
        if (tid == 0) s_tau = prefix;
        __syncthreads();
      }
      const unsigned long long tau = s_tau;
      // D
      unsigned long long ls = 0ull;
      float g0 = 0.f, g1 = 0.f;
      for (int i = tid; i < nc; i += OCCL_THREADS) {
        const unsigned long long key = ckey[i];
        if (key > tau) continue;
        ls += soft_term_fx(cq[i]);
        if (GRAD) hit_tangent(p, env, (int)(key & 0xffffffffull), px, py, &g0, &g1);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ls += __shfl_down_sync(0xffffffffu, ls, o);
        if (GRAD) { g0 += __shfl_down_sync(0xffffffffu, g0, o); g1 += __shfl_down_sync(0xffffffffu, g1, o); }
      }
      if (lane == 0) { s_red[warp][0] = (double)ls; s_red[warp][1] = (double)g0; s_red[warp][2] = (double)g1; }  // (ls < 2^53: exact)
      __syncthreads();
      if (tid == 0) {
        double L = 0.0;
        float G0 = 0.f, G1 = 0.f;
        for (int w = 0; w < OCCL_WARPS; ++w) { L += s_red[w][0]; G0 += (float)s_red[w][1]; G1 += (float)s_red[w][2]; }
        const float P = soft_product_fx((unsigned long long)L);
        const unsigned long long old = sm.soft[slot];
        sm.soft[slot] = (old & 0xffffffff00000000ull) | ((unsigned long long)SOFT_RESOLVED << 32) | (unsigned long long)__float_as_uint(P);
        if (GRAD) sm.gacc[(size_t)obj * tpx + pix] = pack2f(G0, G1);
      }
      __syncthreads();
    }
  }
  if (more) {
    __syncthreads();
    if (tid == 0) s_ovf_n = 0;
    __syncthreads();
  }
  }

}

// TW, TH: compile-time tile shape (0 = take it from the parameters); the fixed 32x32 instantiation turns
// the shared-memory layout and all pixel index arithmetic into constants (register pressure!).
// DBG: the parity / debug outputs (per-object alphas, hit counts, pix_to_face, barycentrics) exist only in this
// instantiation; the production kernel carries no code for them.
// CLIPF: the code for faces cut at z_clip (clip_faces cases 3/4) exists only in this instantiation; it runs in
// raster_clip_kernel, for the (rare) envs in which the setup kernel cut a face.
template <bool GRAD, int TW, int TH, bool DBG, bool CLIPF>
__device__ __forceinline__ void raster_tile(const RasterParams& p, const int env, const int tile) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_tiles = p.tiles_x * p.tiles_y;
  const int tile_w = TW ? TW : p.tile_w, tile_h = TH ? TH : p.tile_h;
  const int tx0 = (tile % p.tiles_x) * tile_w;
  const int ty0 = (tile / p.tiles_x) * tile_h;
  const int tpx = tile_w * tile_h;
  const int S = p.S;

  __shared__ int s_tidx_n;
  __shared__ int s_wdef_n[OCCL_WARPS], s_wsum[OCCL_WARPS];
  __shared__ double s_red[OCCL_WARPS][4];
  __shared__ int s_redi[OCCL_WARPS][2 * OCCL_MAX_OBJ];

  // opaque bases: three to four rebuilt addresses per trip of the pair loop otherwise
  unsigned char* smem_base = smem_raw;
  OCCL_OPAQUE_SHARED(smem_base);
  int* wdef_base = s_wdef_n;
  OCCL_OPAQUE_SHARED(wdef_base);
  TileSmem sm = tile_smem_layout(smem_base, tile_w, tile_h, p.n_obj);
  sm.defer_n = wdef_base + warp;

  // ---- tiles no live face touches: background only ---------------------------------------------
  if (n_tiles <= 32 * TILE_MASK_WORDS &&
      !((__ldg(p.tile_mask + (size_t)env * TILE_MASK_WORDS + (tile >> 5)) >> (tile & 31)) & 1u)) {
    const size_t npix = (size_t)S * S;
    const bool debug_out = DBG && (p.alphas || p.nhits || p.pix_to_face || p.bary);
    // incremental delivery (OcclOutputs.obs_tile_state): the destination already holds this tile's background
    const bool keep_obs = p.obs_tile_state != nullptr &&
        ((__ldg(p.obs_tile_state + (size_t)env * OCCL_TILE_STATE_WORDS + TILE_MASK_WORDS + (tile >> 5)) >> (tile & 31)) & 1u);
    if (!debug_out && (tile_w & 3) == 0 && (S & 3) == 0) {
      // 16-byte stores: four pixels of a row per thread and plane
      const int qw = tile_w >> 2;
      const float inv_qw = 1.0f / (float)qw;
      const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f), neg4 = make_float4(-1.f, -1.f, -1.f, -1.f),
                   zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
      for (int i = tid; i < qw * tile_h; i += OCCL_THREADS) {
        const int ly = (int)(((float)i + 0.5f) * inv_qw), lx = (i - ly * qw) * 4;
        const int xi = tx0 + lx, yi = ty0 + ly;
        if (xi >= S || yi >= S) continue;
        const size_t pix = (size_t)yi * S + xi;
        OCCL_STORE4(p.occl + (size_t)env * npix + pix, zero4);
        if (keep_obs) continue;
        float* o = p.obs + (size_t)env * p.obs_planes * npix + pix;
        OCCL_STORE4(o, one4);
        if (p.obs_planes == 4) {
          OCCL_STORE4(o + npix, one4);
          OCCL_STORE4(o + 2 * npix, one4);
          OCCL_STORE4(o + 3 * npix, neg4);
        } else {
          OCCL_STORE4(o + npix, neg4);
        }
      }
    } else
#pragma unroll 1
    for (int i = tid; i < tpx; i += OCCL_THREADS) {
      const int ly = i / tile_w, lx = i - ly * tile_w;
      const int xi = tx0 + lx, yi = ty0 + ly;
      if (xi >= S || yi >= S) continue;
      const size_t pix = (size_t)yi * S + xi;
      OCCL_STORE(p.occl + (size_t)env * npix + pix, 0.f);
      if (!keep_obs) {
        float* o = p.obs + (size_t)env * p.obs_planes * npix + pix;
        OCCL_STORE(o, 1.0f);
        if (p.obs_planes == 4) { OCCL_STORE(o + npix, 1.0f); OCCL_STORE(o + 2 * npix, 1.0f); OCCL_STORE(o + 3 * npix, -1.0f); } else { OCCL_STORE(o + npix, -1.0f); }
      }
      for (int ob = 0; ob < p.n_obj; ++ob) {
        if (DBG && p.alphas) p.alphas[((size_t)env * p.n_obj + ob) * npix + pix] = 0.f;
        if (DBG && p.nhits) p.nhits[((size_t)env * p.n_obj + ob) * npix + pix] = 0;
      }
      if (DBG && p.pix_to_face) p.pix_to_face[(size_t)env * npix + pix] = -1;
      if (DBG && p.bary) {
        float* bq = p.bary + ((size_t)env * npix + pix) * 3;
        bq[0] = -1.f; bq[1] = -1.f; bq[2] = -1.f;
      }
    }
    if (tid == 0) {
      Partial out;
      out.loss = 0; out.objsq = 0; out.gl[0] = 0; out.gl[1] = 0;
      for (int o = 0; o < OCCL_MAX_OBJ; ++o) { out.ncov[o] = 0; out.nvis[o] = 0; }
      p.partials[(size_t)env * n_tiles + tile] = out;
    }
    return;
  }

  // ---- init accumulators -------------------------------------------------------------------
  for (int i = tid; i < tpx; i += OCCL_THREADS) sm.hard[i] = ~0ull;
  for (int i = tid; i < tpx * p.n_obj; i += OCCL_THREADS) sm.soft[i] = 0ull;
  if (GRAD)
    for (int i = tid; i < tpx * p.n_obj; i += OCCL_THREADS) sm.gacc[i] = 0ull;
  for (int i = tid; i < tile_w; i += OCCL_THREADS) sm.ndc_x[i] = pix_to_ndc(S - 1 - (tx0 + i), S);
  for (int i = tid; i < tile_h; i += OCCL_THREADS) sm.ndc_y[i] = pix_to_ndc(S - 1 - (ty0 + i), S);
  // faces of this tile as binned by the setup kernel; if there is no binning (images of more than 256 tiles) or the
  // tile's list overflowed its capacity, every live face of the env is range-tested here instead
  const int n_bin = p.tile_cnt ? __ldg(p.tile_cnt + (size_t)env * n_tiles + tile) : -1;
  const bool binned = n_bin >= 0 && n_bin <= p.tidx_cap;
  if (tid == 0) s_tidx_n = binned ? n_bin : 0;
  if (tid < OCCL_WARPS) s_wdef_n[tid] = 0;
  __syncthreads();

  const float4* __restrict__ vp = p.vproj + (size_t)env * p.V;
  const int* __restrict__ faces = p.faces + (size_t)env * p.faces_stride;
  const uint4* __restrict__ geo = p.geo + (size_t)env * p.F * 4;
  const uint4* __restrict__ rng = p.rng + (size_t)env * p.F;
  const int n_live = p.n_live[env];
  const int tx1 = tx0 + tile_w - 1, ty1 = ty0 + tile_h - 1;
  int* __restrict__ tidx = p.tile_idx + ((size_t)env * n_tiles + tile) * p.tidx_cap;

  // ---- evaluate-once pre-pass: which slots can exceed K hits, and where their hits will be recorded -----------
  constexpr bool REC = !CLIPF && TW * TH > 0 && TW * TH <= REC_MAX_TPX;
  __shared__ int s_slab;
  RecCtx rc;
  rc.seg = nullptr; rc.keys = nullptr; rc.terms = nullptr;
  if (REC) {
    if (tid == 0) s_slab = -1;
    unsigned* seg = tile_smem_seg(sm, tpx, p.n_obj, GRAD);
    rc.seg = seg;
    const int n_slots = tpx * p.n_obj;
    if (binned && n_bin > p.K && p.rec_keys != nullptr) {  // block-uniform
      const int DW = tile_w + 1, DH = tile_h + 1, DN = DW * DH;
      int* D = (int*)sm.list;  // [n_obj][DH][DW] difference image of the blur boxes (the round buffers are idle)
      // (w + 1)(h + 1) <= w h + w + h + 1 <= 2 REC_MAX_TPX + 2 entries per object: fits the aliased region
      static_assert((size_t)OCCL_MAX_OBJ * (2 * REC_MAX_TPX + 2) * 4 <= (size_t)4 * OCCL_WARPS * WBUF_RECS * REC_WORDS,
                    "difference image must fit the round buffers");
      for (int i = tid; i < p.n_obj * DN; i += OCCL_THREADS) D[i] = 0;
      __syncthreads();
      for (int c = tid; c < n_bin; c += OCCL_THREADS) {
        const uint4 rg = __ldg(rng + __ldg(tidx + c));
        const int x0 = max((int)(rg.x & 0xffffu), tx0) - tx0, x1 = min((int)(rg.x >> 16), tx1) - tx0;
        const int y0 = max((int)(rg.y & 0xffffu), ty0) - ty0, y1 = min((int)(rg.y >> 16), ty1) - ty0;
        if (x0 <= x1 && y0 <= y1) {
          int* Do = D + (int)(rg.w >> 30) * DN;
          atomicAdd(Do + y0 * DW + x0, 1);
          atomicAdd(Do + y0 * DW + x1 + 1, -1);
          atomicAdd(Do + (y1 + 1) * DW + x0, -1);
          atomicAdd(Do + (y1 + 1) * DW + x1 + 1, 1);
        }
      }
      __syncthreads();
      for (int r = warp; r < p.n_obj * DH; r += OCCL_WARPS) {  // prefix sums along x, a warp per (object, row)
        int* row = D + r * DW;
        int carry = 0;
        for (int xb = 0; xb < DW; xb += 32) {
          int v = xb + lane < DW ? row[xb + lane] : 0;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
          }
          v += carry;
          if (xb + lane < DW) row[xb + lane] = v;
          carry = __shfl_sync(0xffffffffu, v, 31);
        }
      }
      __syncthreads();
      for (int c = tid; c < p.n_obj * DW; c += OCCL_THREADS) {  // ... then along y, a thread per (object, column)
        const int o = c / DW, x = c - o * DW;
        int acc = 0;
        for (int y = 0; y < DH; ++y) {
          acc += D[o * DN + y * DW + x];
          D[o * DN + y * DW + x] = acc;
        }
      }
      __syncthreads();
      // segments: slot order, as long as the slab holds them (block scan over the threads' partial sums)
      const int per = (n_slots + OCCL_THREADS - 1) / OCCL_THREADS;
      const int b0 = min(tid * per, n_slots), b1 = min(b0 + per, n_slots);
      int mine = 0;
      for (int i = b0; i < b1; ++i) {
        const int o = i / tpx, pix = i - o * tpx, ly = pix / tile_w, lx = pix - ly * tile_w;
        const int u = D[o * DN + ly * DW + lx];
        mine += u > p.K ? u : 0;
      }
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) s_wsum[warp] = incl;
      __syncthreads();
      int base = incl - mine;
      for (int w = 0; w < warp; ++w) base += s_wsum[w];
      bool any_rec = false;
      for (int i = b0; i < b1; ++i) {
        const int o = i / tpx, pix = i - o * tpx, ly = pix / tile_w, lx = pix - ly * tile_w;
        const int u = D[o * DN + ly * DW + lx];
        if (u > p.K) {
          if (base + u <= p.rec_cap) {
            seg[i] = (unsigned)base;
            sm.soft[i] = (unsigned long long)SOFT_ROUND << 32;  // "this slot records"
            any_rec = true;
          }
          base += u;
        }
      }
      if (__syncthreads_or(any_rec)) {
        if (tid == 0) {  // a slab of scratch: one per resident CTA of this SM
          unsigned smid;
          asm("mov.u32 %0, %%smid;" : "=r"(smid));
          if (smid < REC_SM_MAX)
            for (int b = 0; b < REC_CTAS_PER_SM; ++b)
              if (!(atomicOr(p.rec_table + smid, 1u << b) & (1u << b))) { s_slab = (int)smid * REC_CTAS_PER_SM + b; break; }
        }
        __syncthreads();
        if (s_slab < 0)  // none free (cannot happen with <= REC_CTAS_PER_SM resident CTAs): record nothing
          for (int i = b0; i < b1; ++i) sm.soft[i] = 0ull;
      }
    }
    __syncthreads();
    if (s_slab >= 0) {
      rc.keys = p.rec_keys + (size_t)s_slab * p.rec_cap;
      rc.terms = p.rec_terms + (size_t)s_slab * p.rec_cap;
    }
  }

  // ---- main phase: rounds of up to R faces of the tile ----------------------------------------------------
  // A round (a) loads R face records into shared memory, one per thread, and block-scans their pixel counts
  // (pixels of the tile-clipped blur box), then (b) deals the round's (pixel, face) PAIRS evenly to the warps: warp w
  // takes the pairs [T w / 8, T (w + 1) / 8) of the round's T, lane l the pairs l, l + 32, ... of that range -- every
  // lane has work in every trip whatever the sizes of the faces (a one-pixel sliver and a box side covering the tile
  // are the same to this loop), the warps finish together, consecutive lanes sit on consecutive pixels of (mostly)
  // the same face, so the record loads are broadcasts and the accumulator atomics hit consecutive banks.
  // A pair finds its face by advancing through the prefix sums (a binary search once per round and lane).
  {
    constexpr int R = GRAD ? ROUND_FACES_GRAD : ROUND_FACES_FWD;
    static_assert(R <= OCCL_THREADS && R <= 256, "one face per thread; slot ids are 8 bits in the depth queue");
    static_assert((size_t)R * (64 + 16 + (GRAD ? 48 : 0)) + 4 * (R + 1) <= (size_t)4 * OCCL_WARPS * WBUF_RECS * REC_WORDS,
                  "round buffers must fit the aliased region");
    RoundBuf rb;
    rb.hot = (uint4*)sm.list;
    rb.cold = (float4*)(rb.hot + R * 4);
    rb.tan = (float4*)(rb.cold + R);
    rb.pref = (int*)(rb.tan + (GRAD ? R * 3 : 0));
    // recording queues (REC): 47 two-word entries per warp in the spare end of the aliased region
    constexpr int RQ_CAP = 47;
    static_assert(!REC || (size_t)R * (64 + 16 + (GRAD ? 48 : 0)) + 4 * (R + 1) + 8 * RQ_CAP * OCCL_WARPS <= (size_t)4 * OCCL_WARPS * WBUF_RECS * REC_WORDS,
                  "recording queues must fit behind the round buffers");
    unsigned* rq = (unsigned*)(rb.pref + R + 1) + warp * (2 * RQ_CAP);
    sm.defer = sm.defer + warp * WDEFER_CAP;
    const int n_src = binned ? n_bin : n_live;
    for (int base = 0; base < n_src; base += R) {
      // (a) one candidate per thread
      int npx = 0;
      {
        const int c = base + tid;
        bool keep = false;
        int k = 0;
        uint4 rg = make_uint4(0, 0, 0, 0);
        int cx0 = 0, cx1 = -1, cy0 = 0, cy1 = -1;
        if (tid < R && c < n_src) {
          k = binned ? __ldg(tidx + c) : c;
          rg = __ldg(rng + k);
          cx0 = max((int)(rg.x & 0xffffu), tx0);  cx1 = min((int)(rg.x >> 16), tx1);
          cy0 = max((int)(rg.y & 0xffffu), ty0);  cy1 = min((int)(rg.y >> 16), ty1);
          keep = cx0 <= cx1 && cy0 <= cy1;
        }
        if (!binned) {
          // no list from the setup kernel: remember which live faces touch this tile (the K-overflow and clip passes
          // rescan only those)
          const unsigned bal = __ballot_sync(0xffffffffu, keep);
          int tbase = 0;
          if (lane == 0 && bal) tbase = atomicAdd(&s_tidx_n, __popc(bal));
          tbase = __shfl_sync(0xffffffffu, tbase, 0);
          const int tpos = tbase + __popc(bal & ((1u << lane) - 1u));
          if (keep && tpos < p.tidx_cap) tidx[tpos] = k;
        }
        if (keep) {
          const uint4* __restrict__ src = geo + (size_t)k * 4;
          const uint4 q0 = __ldg(src + 0), q1 = __ldg(src + 1), q2 = __ldg(src + 2), q3 = __ldg(src + 3);
          int hx0 = max((int)(rg.z & 0xffffu), tx0) - tx0, hx1 = min((int)((rg.z >> 16) & 0x7fffu), tx1) - tx0;
          int hy0 = max((int)(rg.w & 0xffffu), ty0) - ty0, hy1 = min((int)((rg.w >> 16) & 0x3fffu), ty1) - ty0;
          if (hx0 > hx1 || hy0 > hy1) { hx0 = 255; hx1 = 0; hy0 = 255; hy1 = 0; }
          const uint32_t hb = (uint32_t)hx0 | ((uint32_t)hx1 << 8) | ((uint32_t)hy0 << 16) | ((uint32_t)hy1 << 24);
          const int w = cx1 - cx0 + 1;
          const uint32_t sbw = (uint32_t)(cx0 - tx0) | ((uint32_t)(cy0 - ty0) << 8) | ((uint32_t)w << 16);
          const float x0 = __uint_as_float(q0.x), y0 = __uint_as_float(q0.y), x1 = __uint_as_float(q0.w), y1 = __uint_as_float(q1.x);
          const float x2 = __uint_as_float(q1.z), y2 = __uint_as_float(q1.w);
          const float bx01 = x1 - x0, by01 = y1 - y0, bx02 = x2 - x0, by02 = y2 - y0, bx12 = x2 - x1, by12 = y2 - y1;
          const float l01 = bx01 * bx01 + by01 * by01, l02 = bx02 * bx02 + by02 * by02, l12 = bx12 * bx12 + by12 * by12;
          rb.hot[tid * 4 + 0] = make_uint4(q0.x, q0.y, q0.w, q1.x);
          rb.hot[tid * 4 + 1] = make_uint4(q1.z, q1.w, q2.z, hb);
          rb.hot[tid * 4 + 2] = make_uint4(q3.x, q3.y, q3.z, sbw);
          rb.hot[tid * 4 + 3] = make_uint4(__float_as_uint(l01), __float_as_uint(l02), __float_as_uint(l12),
                                           __float_as_uint(__fdividef(1.0f, (float)w)));
          rb.cold[tid] = make_float4(__uint_as_float(q0.z), __uint_as_float(q1.y), __uint_as_float(q2.x), __uint_as_float(q2.y));
          if (GRAD) {
            const float4* __restrict__ vt = p.vtan + (size_t)env * p.V;
            const int* __restrict__ fc = faces + 3 * (size_t)(q2.z & REC_FIDX_MASK);
            rb.tan[tid * 3 + 0] = __ldg(vt + __ldg(fc + 0));
            rb.tan[tid * 3 + 1] = __ldg(vt + __ldg(fc + 1));
            rb.tan[tid * 3 + 2] = __ldg(vt + __ldg(fc + 2));
          }
          npx = w * (cy1 - cy0 + 1);
          if (CLIPF && (rg.z & RNG_CLIP)) npx = 0;  // cut face (z-clip): rasterised by the clip phase
        }

      }
      // exclusive block scan of the pixel counts
      int incl = npx;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) s_wsum[warp] = incl;
      __syncthreads();
      int woff = 0, total = 0;
#pragma unroll
      for (int w = 0; w < OCCL_WARPS; ++w) {
        const int v = s_wsum[w];
        if (w < warp) woff += v;
        total += v;
      }
      if (tid < R) rb.pref[tid] = woff + incl - npx;
      if (tid == R - 1 || (R == OCCL_THREADS && tid == OCCL_THREADS - 1)) rb.pref[R] = total;
      __syncthreads();
      // (b) this warp's share of the round's pairs
      const int w_begin = (int)(((long long)total * warp) / OCCL_WARPS), w_end = (int)(((long long)total * (warp + 1)) / OCCL_WARPS);
      int j = w_begin + lane;
      int f = 0;
      {
        int lo = 0, hi = R;  // largest f with pref[f] <= j (empty faces share their successor's prefix: skipped)
#pragma unroll 1
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (rb.pref[mid] <= j) lo = mid; else hi = mid;
        }
        f = lo;
      }
      int fstart = rb.pref[f], fend = rb.pref[f + 1];
      int nrq = 0;  // entries in this warp's recording queue (warp-uniform)
#pragma unroll 1
      for (int jb = w_begin; jb < w_end; jb += 32, j += 32) {
        unsigned rq_word = 0u;
        int rq_pos = -1;
        if (j < w_end) {
          while (j >= fend) { ++f; fstart = fend; fend = rb.pref[f + 1]; }
          raster_pair<GRAD, REC>(p, sm, rb, tile_w, tpx, f, j - fstart, rc, &rq_word, &rq_pos);
        }
        if (REC) {
          // recording queue of this warp (positions from a ballot); 32 keys at a time are computed densely
          const unsigned rqb = __ballot_sync(0xffffffffu, rq_pos >= 0);
          const bool last_trip = jb + 32 >= w_end;
          if (rqb || (last_trip && nrq > 0)) {
            auto drain = [&]() {
              const int take = min(nrq, 32);
              nrq -= take;
              if (lane < take) {
                const unsigned wd = rq[2 * (nrq + lane)], ps = rq[2 * (nrq + lane) + 1];
                const int rpix = (int)(wd & 0xffffu), rly = rpix / tile_w, rlx = rpix - rly * tile_w;
                record_hit(rb, (int)(wd >> 16), sm.ndc_x[rlx], sm.ndc_y[rly], rc.keys + ps);
              }
              __syncwarp();
            };
            if (nrq + __popc(rqb) > RQ_CAP) drain();  // (at most 31 are waiting: this empties the queue)
            if (rq_pos >= 0) {
              const int q = nrq + __popc(rqb & ((1u << lane) - 1u));
              rq[2 * q] = rq_word;
              rq[2 * q + 1] = (unsigned)rq_pos;
            }
            nrq += __popc(rqb);
            __syncwarp();
            while (nrq >= 32 || (last_trip && nrq > 0)) drain();
          }
        }
        __syncwarp();
        // dense exact-depth pass over this warp's queued inside hits, before the queue can overflow
        const int nd = *(volatile int*)sm.defer_n;
        if (nd > WDEFER_CAP - 32 || jb + 32 >= w_end) {
          for (int q = lane; q < min(nd, WDEFER_CAP); q += 32) {
            const uint32_t d = sm.defer[q];
            const int pix = (int)(d & 0xffffu), fq = (int)(d >> 16);
            FaceGeo g;
            round_geo(rb, fq, &g);
            const int ly = pix / tile_w, lx = pix - ly * tile_w;
            hard_update(sm, g, (int)(rb.hot[fq * 4 + 1].z & REC_FIDX_MASK), pix, sm.ndc_x[lx], sm.ndc_y[ly], 0.f, 0.f, 0.f, false);
          }
          __syncwarp();
          if (lane == 0) *sm.defer_n = 0;
          __syncwarp();
        }
      }
      __syncthreads();  // the next round overwrites the records
    }
  }

  // ---- cut faces (z-clip) ------------------------------------------------------------------------
  if (CLIPF) {
    // everything is recomputed from (env, tile): nothing extra may stay live across the main phase
    const int n_t = s_tidx_n;
    const bool use_t = n_t <= p.tidx_cap;
    const int ctx0 = (tile % p.tiles_x) * tile_w, cty0 = (tile / p.tiles_x) * tile_h;
    const TileSmem csm = tile_smem_layout(smem_base, tile_w, tile_h, p.n_obj);
    clip_phase(p.geo + (size_t)env * p.F * 4, p.rng + (size_t)env * p.F,
               p.tile_idx + ((size_t)env * n_tiles + tile) * p.tidx_cap, use_t ? n_t : p.n_live[env], use_t, ctx0, cty0,
               ctx0 + tile_w - 1, cty0 + tile_h - 1, csm.soft, csm.hard, csm.ndc_x, csm.ndc_y, tile_w, tile_w * tile_h, p.z_clip,
               p.cull, p.blur, p.bbox_r, p.inv_sigma_log2e, GRAD ? &p : nullptr, env, csm.gacc);
    __syncthreads();
  }

  // ---- pixels with more than K hits: keep the K nearest by (pz_clipped, face index) -------------
  koverflow_resolve<GRAD, TW, TH, CLIPF, REC>(p, env, tile, s_tidx_n, rc, smem_base);
  if (REC) {
    __syncthreads();
    if (tid == 0 && s_slab >= 0) atomicAnd(p.rec_table + s_slab / REC_CTAS_PER_SM, ~(1u << (s_slab % REC_CTAS_PER_SM)));
  }

  // ---- epilogue: blend, shade, write, reduce ---------------------------------------------------
  const size_t npix = (size_t)S * S;
  double acc_loss = 0.0, acc_obj = 0.0, acc_g0 = 0.0, acc_g1 = 0.0;
  int ncov[OCCL_MAX_OBJ] = {0, 0, 0, 0}, nvis[OCCL_MAX_OBJ] = {0, 0, 0, 0};
  const float inv_tw = 1.0f / (float)tile_w;
#pragma unroll 1  // a fully unrolled epilogue (4 pixels per thread) is 54 KB of code: it evicts the raster loop from the I-cache
  for (int i = tid; i < tpx; i += OCCL_THREADS) {
    const int ly = (int)(((float)i + 0.5f) * inv_tw), lx = i - ly * tile_w;  // exact for i < 2^21
    const int xi = tx0 + lx, yi = ty0 + ly;
    if (xi >= S || yi >= S) continue;
    const size_t pix = (size_t)yi * S + xi;
    const unsigned long long key = sm.hard[i];
    {
      // untouched pixel (the common case): background, nothing to blend or reduce
      unsigned touched = 0u;
#pragma unroll
      for (int o = 0; o < OCCL_MAX_OBJ; ++o)
        if (o < p.n_obj) touched |= (unsigned)(sm.soft[(size_t)o * tpx + i] >> 32);
      if (touched == 0u && key == ~0ull) {
        OCCL_STORE(p.occl + (size_t)env * npix + pix, 0.f);
        float* o = p.obs + (size_t)env * p.obs_planes * npix + pix;
        OCCL_STORE(o, 1.0f);
        if (p.obs_planes == 4) { OCCL_STORE(o + npix, 1.0f); OCCL_STORE(o + 2 * npix, 1.0f); OCCL_STORE(o + 3 * npix, -1.0f); } else { OCCL_STORE(o + npix, -1.0f); }
        for (int ob = 0; ob < p.n_obj; ++ob) {
          if (DBG && p.alphas) p.alphas[((size_t)env * p.n_obj + ob) * npix + pix] = 0.f;
          if (DBG && p.nhits) p.nhits[((size_t)env * p.n_obj + ob) * npix + pix] = 0;
        }
        if (DBG && p.pix_to_face) p.pix_to_face[(size_t)env * npix + pix] = -1;
        if (DBG && p.bary) {
          float* bq = p.bary + ((size_t)env * npix + pix) * 3;
          bq[0] = -1.f; bq[1] = -1.f; bq[2] = -1.f;
        }
        continue;
      }
    }
    float A[OCCL_MAX_OBJ], PR[OCCL_MAX_OBJ];
#pragma unroll
    for (int o = 0; o < OCCL_MAX_OBJ; ++o) {
      A[o] = 0.f;
      PR[o] = 1.f;
      if (o < p.n_obj) {
        const unsigned long long w = sm.soft[(size_t)o * tpx + i];
        PR[o] = soft_product(w);
        A[o] = 1.0f - PR[o];
        const unsigned hi = (unsigned)(w >> 32);
        ncov[o] += (int)((hi / SOFT_COVERED) & 1u);
        if (DBG && p.alphas) p.alphas[((size_t)env * p.n_obj + o) * npix + pix] = A[o];
        if (DBG && p.nhits) p.nhits[((size_t)env * p.n_obj + o) * npix + pix] = (int)(hi & SOFT_CNT_MASK);
      }
    }
    float occl = 0.f, objs = 0.f;
#pragma unroll
    for (int a = 0; a < OCCL_MAX_OBJ; ++a) {
      if (a < p.n_obj) objs = objs + A[a];
#pragma unroll
      for (int b = a + 1; b < OCCL_MAX_OBJ; ++b)
        if (b < p.n_obj) occl = occl + A[a] * A[b];
    }
    OCCL_STORE(p.occl + (size_t)env * npix + pix, occl);
    acc_loss += (double)occl * (double)occl;
    acc_obj += (double)objs * (double)objs;
    if (GRAD) {
      // dA_o = -(1 - A_o) * G_o ; d loss = 2 occl * sum_o (sum_{j != o} A_j) dA_o
      float g0 = 0.f, g1 = 0.f;
#pragma unroll
      for (int o = 0; o < OCCL_MAX_OBJ; ++o) {
        if (o < p.n_obj) {
          const float others = objs - A[o];
          const float c = -PR[o] * others;
          const unsigned long long gw = sm.gacc[(size_t)o * tpx + i];
          g0 += c * __uint_as_float((unsigned)(gw & 0xffffffffull));
          g1 += c * __uint_as_float((unsigned)(gw >> 32));
        }
      }
      acc_g0 += 2.0 * (double)occl * (double)g0;
      acc_g1 += 2.0 * (double)occl * (double)g1;
    }
    // observation: nearest scene face, flat shading, background (1,1,1), depth channel
    float rgb = 1.0f, depth = -1.0f;
    int pf = -1;
    float b0 = -1.f, b1 = -1.f, b2 = -1.f;
    if (key != ~0ull) {
      const unsigned klow = (unsigned)(key & 0xffffffffull);
      pf = (int)(klow & REC_FIDX_MASK);
      depth = __uint_as_float((unsigned)(key >> 32));
      nvis[obj_of_face(p, pf)] += 1;
      const int i0 = __ldg(faces + 3 * pf + 0), i1 = __ldg(faces + 3 * pf + 1), i2 = __ldg(faces + 3 * pf + 2);
      if (CLIPF && (klow & KEY_CLIP)) {
        clip_hard_bary(vp, i0, i1, i2, p.z_clip, p.cull, (int)((klow >> 30) & 1u), sm.ndc_x[lx], sm.ndc_y[ly], &b0, &b1, &b2);
      } else {
      const float4 a = __ldg(vp + i0), b = __ldg(vp + i1), c = __ldg(vp + i2);
      FaceGeo g;
      g.x0 = a.x; g.y0 = a.y; g.z0 = a.z; g.x1 = b.x; g.y1 = b.y; g.z1 = b.z; g.x2 = c.x; g.y2 = c.y; g.z2 = c.z;
      const float e = (g.x2 - g.x0) * (g.y1 - g.y0) - (g.y2 - g.y0) * (g.x1 - g.x0);
      g.area = (float)((double)e + 1e-8);
      bary_persp(g, sm.ndc_x[lx], sm.ndc_y[ly], &b0, &b1, &b2);
      }
      const float2 sh = __ldg(p.shade + (size_t)env * p.F + pf);
      const float texel = (b0 + b1) + b2;
      rgb = sh.x * texel + sh.y;
    }
    float* o = p.obs + (size_t)env * p.obs_planes * npix + pix;
    OCCL_STORE(o, rgb);
    if (p.obs_planes == 4) {
      OCCL_STORE(o + npix, rgb);
      OCCL_STORE(o + 2 * npix, rgb);
      OCCL_STORE(o + 3 * npix, depth);
    } else {
      OCCL_STORE(o + npix, depth);
    }
    if (DBG && p.pix_to_face) p.pix_to_face[(size_t)env * npix + pix] = pf;
    if (DBG && p.bary) {
      float* bq = p.bary + ((size_t)env * npix + pix) * 3;
      bq[0] = b0; bq[1] = b1; bq[2] = b2;
    }
  }
  // block reduction, fixed order -> deterministic given the per-pixel values
  acc_loss = warp_sum(acc_loss);
  acc_obj = warp_sum(acc_obj);
  if (GRAD) { acc_g0 = warp_sum(acc_g0); acc_g1 = warp_sum(acc_g1); }
#pragma unroll
  for (int o = 0; o < OCCL_MAX_OBJ; ++o) { ncov[o] = warp_sum_i(ncov[o]); nvis[o] = warp_sum_i(nvis[o]); }
  if (lane == 0) {
    s_red[warp][0] = acc_loss; s_red[warp][1] = acc_obj; s_red[warp][2] = acc_g0; s_red[warp][3] = acc_g1;
#pragma unroll
    for (int o = 0; o < OCCL_MAX_OBJ; ++o) { s_redi[warp][o] = ncov[o]; s_redi[warp][OCCL_MAX_OBJ + o] = nvis[o]; }
  }
  __syncthreads();
  if (tid == 0) {
    Partial out;
    out.loss = 0; out.objsq = 0; out.gl[0] = 0; out.gl[1] = 0;
    for (int o = 0; o < OCCL_MAX_OBJ; ++o) { out.ncov[o] = 0; out.nvis[o] = 0; }
    for (int w = 0; w < OCCL_WARPS; ++w) {
      out.loss += s_red[w][0]; out.objsq += s_red[w][1]; out.gl[0] += s_red[w][2]; out.gl[1] += s_red[w][3];
      for (int o = 0; o < OCCL_MAX_OBJ; ++o) { out.ncov[o] += s_redi[w][o]; out.nvis[o] += s_redi[w][OCCL_MAX_OBJ + o]; }
    }
    p.partials[(size_t)env * n_tiles + tile] = out;
  }
}


// One CTA per (env, tile).  Envs in which a face was cut at z_clip are left to raster_clip_kernel.
template <bool GRAD, int TW, int TH, bool DBG>
__global__ void __launch_bounds__(OCCL_THREADS, GRAD ? OCCL_CTAS_GRAD : OCCL_CTAS_FWD)
raster_kernel(const RasterParams p) {
  const int n_tiles = p.tiles_x * p.tiles_y;
  const int env = blockIdx.x / n_tiles;
  const int tile = blockIdx.x - env * n_tiles;
  if (p.env_mask && !p.env_mask[env]) return;
  if (*(volatile const uint32_t*)(p.status + env) & OCCL_ST_CLIPPED) return;
  raster_tile<GRAD, TW, TH, DBG, false>(p, env, tile);
}

// Masked transitions (the auto-reset of finished envs inside a step, SubProcVecEnv.py:211-214): usually no env, or a
// handful, is flagged.  One CTA per (env, tile) of the whole batch would launch 65 536 CTAs that look at a flag and
// leave (0.2 ms per config-2 step, measured); instead the flagged envs are listed on the device and a fixed grid of
// CTAs works through the (listed env, tile) items.
__global__ void mask_list_kernel(const int m, const uint8_t* __restrict__ mask, int* __restrict__ list) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < m && mask[e]) list[1 + atomicAdd(list, 1)] = e;
}

template <int TW, int TH, bool DBG>
__global__ void __launch_bounds__(OCCL_THREADS, OCCL_CTAS_FWD) raster_list_kernel(const RasterParams p) {
  const int n_tiles = p.tiles_x * p.tiles_y;
  const int n_env = __ldg(p.env_list);
  const int n_items = n_env * n_tiles;
  // items in tile-major order: a CTA's items (stride gridDim.x) then sweep through the tiles.  Env-major order gave CTA b
  // the SAME tile of every env whenever n_tiles divides the grid (592 = 37 x 16): ten centre tiles here, ten empty corners there.
  for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
    const int tile = it / n_env;
    const int env = __ldg(p.env_list + 1 + (it - tile * n_env));
    if (!(*(volatile const uint32_t*)(p.status + env) & OCCL_ST_CLIPPED)) raster_tile<false, TW, TH, DBG, false>(p, env, tile);
    __syncthreads();
  }
}

// Envs with cut faces (camera within z_clip = znear/2 of the geometry), which the setup kernel has listed: a fixed
// grid of CTAs works through the (listed env, tile) items with the clip-capable instantiation of the tile rasteriser
// (generic tile shape).  With an empty list -- every BASELINE teapot pose -- the few CTAs leave at once.  (One CTA per
// env was the first version: a dense env then ran on 1/592 of the GPU; one CTA per (env, tile) of the whole batch the
// second: 65 536 CTAs that only look at a flag cost 0.12 ms per config-2 step.)
template <bool GRAD>
__global__ void __launch_bounds__(OCCL_THREADS) raster_clip_kernel(const RasterParams p) {
  const int n_tiles = p.tiles_x * p.tiles_y;
  const int n_env = __ldg(p.clip_list);
  const int n_items = n_env * n_tiles;
  for (int it = blockIdx.x; it < n_items; it += gridDim.x) {  // tile-major, as in raster_list_kernel
    const int tile = it / n_env;
    const int env = __ldg(p.clip_list + 1 + (it - tile * n_env));
    raster_tile<GRAD, 0, 0, true, true>(p, env, tile);
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------------------------
// kernel 6: per-env finalisation
// ----------------------------------------------------------------------------------------------
// mode 0: step (environment.py:381-392) ; mode 1: reset (:322-327) ; mode 2: render only
__global__ void finalize_kernel(int n, int mode, int n_tiles, int n_obj, int norm_with_object_size,
                                float done_threshold, float reward_done, float reward_step, float step_size,
                                const Partial* __restrict__ partials, const float* __restrict__ action,
                                float* __restrict__ full_reward, float* __restrict__ object_mass,
                                float* __restrict__ reward, uint8_t* done,
                                float* __restrict__ loss_out, int* __restrict__ n_covered,
                                int* __restrict__ n_visible, float* __restrict__ grad_action,
                                const uint32_t* __restrict__ status, uint32_t* __restrict__ status_or,
                                const uint8_t* env_mask) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (status_or) {
    // running OR of every status word this library has produced (one warp reduction, one atomic when non-zero):
    // the host checks ONE word, when it wants to, instead of scanning (N,) words after every step
    const bool on = e < n && !(env_mask && !env_mask[e]);
    const unsigned st = on ? status[e] : 0u;
    const unsigned all = __reduce_or_sync(0xffffffffu, st);
    if (all && (threadIdx.x & 31) == 0) atomicOr(status_or, all);
  }
  if (e >= n) return;
  if (env_mask && !env_mask[e]) return;
  double loss = 0.0, objsq = 0.0, g0 = 0.0, g1 = 0.0;
  int ncov[OCCL_MAX_OBJ] = {0, 0, 0, 0}, nvis[OCCL_MAX_OBJ] = {0, 0, 0, 0};
  const Partial* __restrict__ pp = partials + (size_t)e * n_tiles;
  for (int t = 0; t < n_tiles; ++t) {
    loss += pp[t].loss; objsq += pp[t].objsq; g0 += pp[t].gl[0]; g1 += pp[t].gl[1];
    for (int o = 0; o < OCCL_MAX_OBJ; ++o) { ncov[o] += pp[t].ncov[o]; nvis[o] += pp[t].nvis[o]; }
  }
  const float lossf = (float)loss;
  if (loss_out) loss_out[e] = lossf;
  for (int o = 0; o < n_obj; ++o) {
    if (n_covered) n_covered[e * n_obj + o] = ncov[o];
    if (n_visible) n_visible[e * n_obj + o] = nvis[o];
  }
  if (mode == 2) return;
  if (mode == 1) {
    full_reward[e] = lossf;                                                  // :323
    object_mass[e] = (norm_with_object_size ? (float)objsq : lossf) + 1.0f;  // :324
    // :327 (no occlusion at reset).  Not written by a masked reset: the mask usually IS the `done` array of the step.
    if (done && !env_mask) done[e] = (uint8_t)(!(lossf > done_threshold));
    return;
  }
  const float mass = object_mass[e];
  float r = full_reward[e] - lossf;  // :382
  full_reward[e] = lossf;            // :384
  const bool fin = lossf < done_threshold;  // :386
  r = r / mass;                             // :387
  r = fin ? r + reward_done : r - reward_step;
  reward[e] = r;
  done[e] = (uint8_t)fin;
  if (grad_action) {
    // d reward / d(el, az) = -(d loss / d(el, az)) / mass ; (el, az) += step * a / |a|
    const float ge = (float)(-g0 / (double)mass), ga = (float)(-g1 / (double)mass);
    const float a0 = action[2 * e + 0], a1 = action[2 * e + 1];
    const float nrm = sqrtf(a0 * a0 + a1 * a1);
    if (nrm != 0.f) {
      const float n0 = a0 / nrm, n1 = a1 / nrm;
      const float dot = ge * n0 + ga * n1;
      grad_action[2 * e + 0] = step_size * (ge - dot * n0) / nrm;
      grad_action[2 * e + 1] = step_size * (ga - dot * n1) / nrm;
    } else {
      grad_action[2 * e + 0] = step_size * ge;
      grad_action[2 * e + 1] = step_size * ga;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// host side: C-ABI
// ----------------------------------------------------------------------------------------------
static int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_last_err, sizeof(g_last_err), "%s: %s", where, cudaGetErrorString(e));
  return OCCL_E_CUDA;
}
#define CK(call, where)                                      \
  do {                                                       \
    cudaError_t _e = (call);                                 \
    if (_e != cudaSuccess) return cuda_fail(_e, where);      \
  } while (0)

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Raise a kernel's dynamic shared-memory limit once per device (not on every launch).
template <void (*KERNEL)(const RasterParams)>
static cudaError_t ensure_dyn_smem(size_t smem) {
  static size_t have[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && have[dev] >= smem) return cudaSuccess;
  e = cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess && dev >= 0 && dev < 64) have[dev] = smem;
  return e;
}

struct WsLayout {
  size_t cam, vproj, vtan, partials, geo, rng, n_live, tile_mask, shade, tile_idx, tile_cnt, clip_list, env_list, total;
  size_t rec_keys, rec_terms, rec_table;  // evaluate-once scratch (0 size when the tile does not record)
  int rec_cap;
  int tidx_cap;
  int n_tiles;
  int chunk;  // envs rasterised per launch: the per-face scratch (geo .. tile_cnt) is sized for this many, not for N
  int sets;   // 1: all envs in one launch; 2: two scratch sets, the chunks alternate between them on two streams
  size_t set_stride;  // bytes from a set-0 scratch buffer to its set-1 twin
};

// the compile-time tiles of at most REC_MAX_TPX pixels record their hits (see REC_*)
static bool rec_capable(const OcclConfig* c) {
  return (c->tile_w == OCCL_TILE2_W && c->tile_h == OCCL_TILE2_H && OCCL_TILE2_W * OCCL_TILE2_H <= REC_MAX_TPX) ||
         (c->tile_w == OCCL_TILE3_W && c->tile_h == OCCL_TILE3_H && OCCL_TILE3_W * OCCL_TILE3_H <= REC_MAX_TPX) ||
         (c->tile_w == OCCL_TILE_W && c->tile_h == OCCL_TILE_H && OCCL_TILE_W * OCCL_TILE_H <= REC_MAX_TPX);
}

static size_t tile_smem_bytes(const OcclConfig* c, int with_grad) {
  const size_t tpx = (size_t)c->tile_w * c->tile_h;
  size_t b = 4 * OCCL_WARPS * WBUF_RECS * REC_WORDS + 4 * OCCL_WARPS * WDEFER_CAP + 4 * BIG_CAP * REC_WORDS + 8 * tpx +
             4 * (size_t)(((c->tile_w + 1) & ~1) + ((c->tile_h + 1) & ~1)) + 8 * tpx * c->n_obj;
  if (with_grad) b += 4 * 2 * tpx * c->n_obj;
  if (rec_capable(c)) b += 4 * tpx * c->n_obj;  // segment table of the evaluate-once path
  return b;
}

__global__ void selftest_div_kernel(unsigned long long n, unsigned long long seed, unsigned long long* mismatches) {
  unsigned long long bad = 0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    // splitmix64 -> two floats: b in [1e-8, 128], |a| in [2^-80, 2^10], random mantissas
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const unsigned ma = (unsigned)z & 0x7fffffu, mb = (unsigned)(z >> 23) & 0x7fffffu;
    const unsigned ea = 127u - 80u + (unsigned)((z >> 46) % 91u);        // 2^-80 .. 2^10
    const unsigned eb = 127u - 27u + (unsigned)((z >> 54) % 34u);        // 2^-27 .. 2^6
    float a = __uint_as_float((ea << 23) | ma), b = __uint_as_float((eb << 23) | mb);
    if (z >> 63) a = -a;
    b = fminf(fmaxf(b, 1.0000001e-8f), 128.f);
    const float y = 1.0f / b;
    if (__float_as_uint(div_rn_hoisted(a, b, y)) != __float_as_uint(a / b)) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

extern "C" int occl_selftest_div(unsigned long long n_samples, unsigned long long seed, unsigned long long* mismatches_dev,
                                 void* stream) {
  if (!mismatches_dev) return OCCL_E_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(mismatches_dev, 0, sizeof(unsigned long long), s) != cudaSuccess) return OCCL_E_CUDA;
  selftest_div_kernel<<<148 * 8, 256, 0, s>>>(n_samples, seed, mismatches_dev);
  if (cudaGetLastError() != cudaSuccess) return OCCL_E_CUDA;
  return OCCL_OK;
}

extern "C" int occl_abi_version(void) { return OCCL_ABI_VERSION; }

extern "C" int occl_ipc_export(const void* ptr, void* handle64_out, size_t* offset_out) {
  if (!ptr || !handle64_out || !offset_out) return OCCL_E_INVALID;
  // cudaIpcGetMemHandle wants the BASE of the cudaMalloc block the pointer lies in (torch's allocator sub-allocates):
  // the driver knows it (cuMemGetAddressRange), resolved at run time so that the library loads without a driver
  typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
  static range_fn fn = nullptr;
  if (!fn) {
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (h) fn = (range_fn)dlsym(h, "cuMemGetAddressRange_v2");
    if (!fn) {
      snprintf(g_last_err, sizeof(g_last_err), "libcuda.so.1 / cuMemGetAddressRange_v2 not available");
      return OCCL_E_CUDA;
    }
  }
  CK(cudaFree(0), "context");  // make sure the runtime's primary context is current for the driver call
  unsigned long long base = 0;
  size_t size = 0;
  if (fn(&base, &size, (unsigned long long)(uintptr_t)ptr) != 0) {
    snprintf(g_last_err, sizeof(g_last_err), "cuMemGetAddressRange failed for %p", ptr);
    return OCCL_E_CUDA;
  }
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base), "cudaIpcGetMemHandle");
  memcpy(handle64_out, &h, sizeof(h));
  *offset_out = (size_t)((unsigned long long)(uintptr_t)ptr - base);
  return OCCL_OK;
}

extern "C" int occl_ipc_open(const void* handle64, void** ptr_out) {
  if (!handle64 || !ptr_out) return OCCL_E_INVALID;
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(&h, handle64, sizeof(h));
  // mapped into the CURRENT device's address space, peer access to the exporting device enabled as needed
  CK(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
  return OCCL_OK;
}

extern "C" int occl_ipc_close(void* ptr) {
  CK(cudaIpcCloseMemHandle(ptr), "cudaIpcCloseMemHandle");
  return OCCL_OK;
}

extern "C" int occl_enable_peer_access(int peer_device) {
  int dev = 0;
  CK(cudaGetDevice(&dev), "cudaGetDevice");
  if (dev == peer_device) return OCCL_OK;
  int can = 0;
  CK(cudaDeviceCanAccessPeer(&can, dev, peer_device), "cudaDeviceCanAccessPeer");
  if (!can) {
    snprintf(g_last_err, sizeof(g_last_err), "device %d cannot access device %d (no NVLink / PCIe peer path)", dev, peer_device);
    return OCCL_E_CUDA;
  }
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    (void)cudaGetLastError();
    return OCCL_OK;
  }
  CK(e, "cudaDeviceEnablePeerAccess");
  return OCCL_OK;
}
extern "C" const char* occl_last_cuda_error(void) { return g_last_err; }

extern "C" int occl_config_resolve(OcclConfig* c, int with_grad) {
  if (!c) return OCCL_E_INVALID;
  if (c->image_size < 1 || c->image_size > 4096) return OCCL_E_INVALID;
  if (c->n_obj < 1 || c->n_obj > OCCL_MAX_OBJ) return OCCL_E_INVALID;
  if (c->n_verts < 1 || c->n_faces < 1) return OCCL_E_INVALID;
  if (c->faces_per_pixel < 1 || c->faces_per_pixel > 128) return OCCL_E_INVALID;
  if (c->obs_planes != 0 && c->obs_planes != 2 && c->obs_planes != 4) return OCCL_E_INVALID;
  if (c->n_faces >= (1 << 28)) return OCCL_E_INVALID;  // packed face index field
  if (c->obj_face_start[0] != 0 || c->obj_face_start[c->n_obj] != c->n_faces) return OCCL_E_INVALID;
  for (int i = 0; i < c->n_obj; ++i)
    if (c->obj_face_start[i + 1] < c->obj_face_start[i]) return OCCL_E_INVALID;
  if (!(c->blur_radius >= 0.f) || !(c->sigma > 0.f)) return OCCL_E_INVALID;
  if (c->tile_w == 0 || c->tile_h == 0) {
    const int S = c->image_size;
    // square tiles split the fewest faces; 32x32 px (1024 px of accumulators) keeps 3 CTAs per SM
    const bool dense = c->n_faces >= OCCL_DENSE_FACES, many = c->n_obj >= 3;
    const int tw = dense ? OCCL_TILE3_W : (many ? OCCL_TILE2_W : OCCL_TILE_W);
    const int th = dense ? OCCL_TILE3_H : (many ? OCCL_TILE2_H : OCCL_TILE_H);
    c->tile_w = S < tw ? S : tw;
    c->tile_h = S < th ? S : th;
  }
  if (c->tile_w < 1 || c->tile_h < 1 || c->tile_w > 256 || c->tile_h > 256) return OCCL_E_INVALID;
  if (tile_smem_bytes(c, with_grad) + 8 * 1024 > 227 * 1024) return OCCL_E_SMEM;
  // the top-K selection buffers alias the face list
  return OCCL_OK;
}

// Per-env buffers (camera, projected vertices, tile partials) are sized for all N envs; the per-FACE scratch of the
// rasteriser (live-face records, pixel ranges, lighting, tile index lists: ~90 B per face plus 4 B per tile-list slot,
// of which only the live / touched part is ever written) is sized for a CHUNK of envs and reused chunk after chunk on
// the stream (face_setup -> raster -> raster_clip per chunk).  The chunk is the largest env count whose scratch fits
// ws_budget_mb (default 8 GiB): 4096 envs x 2 476 faces (config 2) is one chunk of 1.6 GB, 8192 envs x 61 440 faces x
// 256^2 (config 3; 78 GB if sized for N) runs in chunks of ~430 envs.  Measured at config 3 (env-steps/s against the
// budget): 1 GiB 21.9 k, 2 GiB 24.3 k, 4 GiB 26.4 k, 8 GiB 27.5 k, 12 GiB 27.8 k -- the setup kernel is one CTA per env,
// so a chunk should hold several hundred envs to fill the GPU next to the other lane's raster launch.
static int ws_layout(const OcclConfig* c, int n, int with_grad, WsLayout* L) {
  const int tx = (c->image_size + c->tile_w - 1) / c->tile_w;
  const int ty = (c->image_size + c->tile_h - 1) / c->tile_h;
  L->n_tiles = tx * ty;
  L->tidx_cap = c->n_faces < 8192 ? c->n_faces : 8192;
  const size_t per_env = (size_t)c->n_faces * (sizeof(uint4) * 5 + sizeof(float2)) + sizeof(int) * (size_t)L->n_tiles * (L->tidx_cap + 1) +
                         sizeof(int) + sizeof(uint32_t) * TILE_MASK_WORDS;
  const size_t budget = (size_t)(c->ws_budget_mb > 0 ? c->ws_budget_mb : OCCL_WS_BUDGET_MB) << 20;
  size_t chunk = budget / per_env;
  L->sets = 1;
  if (chunk < (size_t)n) {
    // two half-size scratch sets: chunk c + 1 is set up and rasterised (second stream) while the last, heavy tiles of
    // chunk c finish -- a chunked batch then runs like one long launch instead of paying a tail per chunk
    L->sets = 2;
    chunk = budget / 2 / per_env;
    if (chunk < 1) chunk = 1;
    // equal chunks
    const size_t n_chunks = ((size_t)n + chunk - 1) / chunk;
    chunk = ((size_t)n + n_chunks - 1) / n_chunks;
  }
  if (chunk > (size_t)n) chunk = (size_t)n;
  L->chunk = (int)chunk;
  const size_t m = chunk;
  size_t off = 0;
  L->cam = off;      off = align_up(off + sizeof(float) * OCCL_CAM_STRIDE * (size_t)n, 256);
  L->vproj = off;    off = align_up(off + sizeof(float4) * (size_t)n * c->n_verts, 256);
  L->vtan = off;     if (with_grad) off = align_up(off + sizeof(float4) * (size_t)n * c->n_verts, 256);
  L->partials = off; off = align_up(off + sizeof(Partial) * (size_t)n * L->n_tiles, 256);
  L->geo = off;      off = align_up(off + sizeof(uint4) * 4 * m * c->n_faces, 256);
  L->rng = off;      off = align_up(off + sizeof(uint4) * m * c->n_faces, 256);
  L->n_live = off;   off = align_up(off + sizeof(int) * m, 256);
  L->tile_mask = off; off = align_up(off + sizeof(uint32_t) * TILE_MASK_WORDS * m, 256);
  L->shade = off;    off = align_up(off + sizeof(float2) * m * c->n_faces, 256);
  L->tile_idx = off; off = align_up(off + sizeof(int) * m * L->n_tiles * L->tidx_cap, 256);
  L->tile_cnt = off; off = align_up(off + sizeof(int) * m * L->n_tiles, 256);
  L->clip_list = off; off = align_up(off + sizeof(int) * (m + 1), 256);
  L->env_list = off; off = align_up(off + sizeof(int) * (m + 1), 256);
  L->set_stride = off - L->geo;
  if (L->sets == 2) off += L->set_stride;
  L->rec_keys = L->rec_terms = L->rec_table = 0;
  L->rec_cap = 0;
  if (rec_capable(c)) {
    const size_t slabs = (size_t)REC_SM_MAX * REC_CTAS_PER_SM;
    L->rec_table = off; off = align_up(off + sizeof(unsigned) * REC_SM_MAX, 256);
    L->rec_cap = c->n_faces >= OCCL_DENSE_FACES ? REC_SLAB_DENSE : REC_SLAB_SPARSE;
    L->rec_keys = off;  off = align_up(off + sizeof(unsigned long long) * slabs * L->rec_cap, 256);
    L->rec_terms = off; off = align_up(off + sizeof(float) * slabs * L->rec_cap, 256);
  }
  L->total = off;
  return 0;
}

// Two internal streams per device for the chunk pipeline (created once, never destroyed).
static int chunk_streams(cudaStream_t out[2]) {
  static cudaStream_t pool[64][2];
  static bool have[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return OCCL_E_CUDA;
  if (!have[dev]) {
    for (int k = 0; k < 2; ++k)
      if (cudaStreamCreateWithFlags(&pool[dev][k], cudaStreamNonBlocking) != cudaSuccess) return OCCL_E_CUDA;
    have[dev] = true;
  }
  out[0] = pool[dev][0];
  out[1] = pool[dev][1];
  return OCCL_OK;
}

extern "C" size_t occl_workspace_bytes(const OcclConfig* cfg, int n_envs, int with_grad) {
  if (!cfg || n_envs < 1) return 0;
  OcclConfig c = *cfg;
  if (occl_config_resolve(&c, with_grad) != OCCL_OK) return 0;
  WsLayout L;
  ws_layout(&c, n_envs, with_grad, &L);
  return L.total;
}

extern "C" int occl_workspace_offsets(const OcclConfig* cfg, int n_envs, int with_grad, size_t* offsets4) {
  if (!cfg || n_envs < 1 || !offsets4) return OCCL_E_INVALID;
  OcclConfig c = *cfg;
  int rc = occl_config_resolve(&c, with_grad);
  if (rc != OCCL_OK) return rc;
  WsLayout L;
  ws_layout(&c, n_envs, with_grad, &L);
  offsets4[0] = L.cam; offsets4[1] = L.vproj; offsets4[2] = L.vtan; offsets4[3] = L.partials;
  return OCCL_OK;
}

static int launch_pose(const OcclConfig* c, int n, int mode, const float* action, OcclState st, float* cam,
                       float* position, uint32_t* status, const uint8_t* mask, cudaStream_t s) {
  if (!st.elevation || !st.azimuth || !st.radius || !cam) return OCCL_E_INVALID;
  if (mode == 0 && !action) return OCCL_E_INVALID;
  pose_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, mode, c->step_size, action, st.elevation, st.azimuth, st.radius,
                                             cam, position, status, mask);
  CK(cudaGetLastError(), "pose_kernel");
  return OCCL_OK;
}

extern "C" int occl_pose_step(const OcclConfig* cfg, int n, const float* action, OcclState st, float* cam,
                              void* stream) {
  if (!cfg || n < 1) return OCCL_E_INVALID;
  return launch_pose(cfg, n, 0, action, st, cam, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}
extern "C" int occl_pose_lookat(const OcclConfig* cfg, int n, OcclState st, float* cam, void* stream) {
  if (!cfg || n < 1) return OCCL_E_INVALID;
  return launch_pose(cfg, n, 1, nullptr, st, cam, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}
extern "C" int occl_pose_set(int n, const float* R, const float* T, const float* C, float* cam, void* stream) {
  if (n < 1 || !R || !T || !C || !cam) return OCCL_E_INVALID;
  pose_set_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, R, T, C, cam);
  CK(cudaGetLastError(), "pose_set_kernel");
  return OCCL_OK;
}

static int project_impl(const OcclConfig* cfg, int n, const float* cam, OcclScene sc, float* vproj, float* vtan,
                        uint32_t* status, const uint8_t* mask, void* stream) {
  if (!cfg || n < 1 || !cam || !sc.verts || !vproj) return OCCL_E_INVALID;
  const long long total = (long long)n * cfg->n_verts;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (vtan)
    project_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(total, cfg->n_verts, cam, sc.verts, sc.verts_env_stride,
                                                                  cfg->proj_scale, cfg->z_clip, (float4*)vproj,
                                                                  (float4*)vtan, status, mask);
  else
    project_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(total, cfg->n_verts, cam, sc.verts, sc.verts_env_stride,
                                                                   cfg->proj_scale, cfg->z_clip, (float4*)vproj, nullptr,
                                                                   status, mask);
  CK(cudaGetLastError(), "project_kernel");
  return OCCL_OK;
}

extern "C" int occl_project(const OcclConfig* cfg, int n, const float* cam, OcclScene sc, float* vproj, float* vtan,
                            uint32_t* status, void* stream) {
  return project_impl(cfg, n, cam, sc, vproj, vtan, status, nullptr, stream);
}

static int check_outputs(const OcclOutputs& o) {
  if (!o.obs || !o.occl || !o.status) return OCCL_E_INVALID;
  return OCCL_OK;
}

static int raster_impl(const OcclConfig* cfg, int n, OcclScene sc, OcclWorkspace ws, OcclOutputs out, const uint8_t* mask,
                       void* stream) {
  if (!cfg || n < 1 || !sc.verts || !sc.faces || !ws.base) return OCCL_E_INVALID;
  if (check_outputs(out) != OCCL_OK) return OCCL_E_INVALID;
  const int grad = out.grad_action != nullptr;
  OcclConfig c = *cfg;
  int rc = occl_config_resolve(&c, grad);
  if (rc != OCCL_OK) return rc;
  WsLayout L;
  ws_layout(&c, n, grad, &L);
  if (ws.bytes < L.total) return OCCL_E_INVALID;
  unsigned char* base = (unsigned char*)ws.base;
  RasterParams p;
  p.S = c.image_size; p.n_obj = c.n_obj; p.V = c.n_verts; p.F = c.n_faces; p.K = c.faces_per_pixel;
  p.cull = c.cull_backfaces;
  for (int i = 0; i <= OCCL_MAX_OBJ; ++i) p.obj_face_start[i] = i <= c.n_obj ? c.obj_face_start[i] : c.n_faces;
  p.tile_w = c.tile_w; p.tile_h = c.tile_h;
  p.tiles_x = (c.image_size + c.tile_w - 1) / c.tile_w;
  p.tiles_y = (c.image_size + c.tile_h - 1) / c.tile_h;
  p.blur = c.blur_radius; p.bbox_r = sqrtf(c.blur_radius); p.sigma = c.sigma; p.z_clip = c.z_clip;
  p.inv_sigma = 1.0f / c.sigma;
  p.inv_sigma_log2e = (float)(1.4426950408889634 / (double)c.sigma);
  p.exact_only = c.debug_exact;
  p.geo = (uint4*)(base + L.geo); p.rng = (uint4*)(base + L.rng); p.n_live = (int*)(base + L.n_live);
  p.tile_mask = (const uint32_t*)(base + L.tile_mask);
  p.shade = (const float2*)(base + L.shade);
  p.tile_idx = (int*)(base + L.tile_idx); p.tidx_cap = L.tidx_cap;
  // the setup kernel bins the live faces per tile whenever the image has at most 256 tiles (its shared-memory counters);
  // the raster CTAs then read their own list instead of range-testing every live face of the env
  const bool bin = L.n_tiles <= 32 * TILE_MASK_WORDS;
  p.tile_cnt = bin ? (const int*)(base + L.tile_cnt) : nullptr;
  p.light[0] = c.light[0]; p.light[1] = c.light[1]; p.light[2] = c.light[2];
  p.vproj = (const float4*)(base + L.vproj);
  p.vtan = grad ? (const float4*)(base + L.vtan) : nullptr;
  p.verts = sc.verts; p.verts_stride = sc.verts_env_stride;
  p.faces = sc.faces; p.faces_stride = sc.faces_env_stride;
  p.cam = (const float*)(base + L.cam);
  p.partials = (Partial*)(base + L.partials);
  p.rec_keys = nullptr; p.rec_terms = nullptr; p.rec_table = nullptr; p.rec_cap = L.rec_cap;
  if (L.rec_keys) {
    p.rec_keys = (unsigned long long*)(base + L.rec_keys);
    p.rec_terms = (float*)(base + L.rec_terms);
    p.rec_table = (unsigned*)(base + L.rec_table);
    CK(cudaMemsetAsync(p.rec_table, 0, sizeof(unsigned) * REC_SM_MAX, (cudaStream_t)stream), "memset rec table");
  }
  p.obs = out.obs; p.obs_planes = c.obs_planes == 2 ? 2 : 4; p.occl = out.occl; p.alphas = out.alphas; p.pix_to_face = out.pix_to_face; p.bary = out.bary;
  p.nhits = out.nhits; p.status = out.status; p.env_mask = mask;
  // incremental delivery needs the tile mask (images of at most 256 tiles)
  p.obs_tile_state = L.n_tiles <= 32 * TILE_MASK_WORDS ? out.obs_tile_state : nullptr;
  const size_t smem = tile_smem_bytes(&c, grad);
  if ((long long)L.chunk * L.n_tiles > 0x7fffffffLL) return OCCL_E_INVALID;
  SetupParams sp;
  sp.S = c.image_size; sp.V = c.n_verts; sp.F = c.n_faces; sp.cull = c.cull_backfaces; sp.bbox_r = p.bbox_r;
  sp.z_clip = c.z_clip; sp.status = out.status; sp.grad = grad;
  sp.vproj = p.vproj; sp.faces = sc.faces; sp.faces_stride = sc.faces_env_stride;
  sp.geo = p.geo; sp.rng = p.rng; sp.n_live = p.n_live; sp.env_mask = mask;
  sp.tile_mask = (uint32_t*)(base + L.tile_mask);
  sp.tile_idx = p.tile_idx; sp.tile_cnt = bin ? (int*)(base + L.tile_cnt) : nullptr; sp.tidx_cap = L.tidx_cap;
  sp.shade = (float2*)(base + L.shade);
  sp.verts = sc.verts; sp.verts_stride = sc.verts_env_stride; sp.cam = p.cam;
  sp.light[0] = c.light[0]; sp.light[1] = c.light[1]; sp.light[2] = c.light[2];
  sp.n_obj = c.n_obj;
  for (int i = 0; i <= OCCL_MAX_OBJ; ++i) sp.obj_face_start[i] = p.obj_face_start[i];
  sp.tile_w = c.tile_w; sp.tile_h = c.tile_h; sp.tiles_x = p.tiles_x; sp.n_tiles = L.n_tiles;
  sp.inv_tile_w = 1.0f / (float)c.tile_w; sp.inv_tile_h = 1.0f / (float)c.tile_h;
  const bool fixed = c.tile_w == OCCL_TILE_W && c.tile_h == OCCL_TILE_H;
  const bool fixed2 = !fixed && c.tile_w == OCCL_TILE2_W && c.tile_h == OCCL_TILE2_H;
  const bool fixed3 = !fixed && !fixed2 && c.tile_w == OCCL_TILE3_W && c.tile_h == OCCL_TILE3_H;
  const bool dbg = out.alphas || out.pix_to_face || out.bary || out.nhits;
  const size_t npix = (size_t)c.image_size * c.image_size;
  const RasterParams p0 = p;
  // chunk after chunk of envs through the same per-face scratch (see ws_layout): every per-env array is entered at
  // the chunk's first env, so that the kernels index everything by the env's position inside the chunk
  cudaStream_t lanes[2] = {(cudaStream_t)stream, (cudaStream_t)stream};
  cudaEvent_t ev_fork = nullptr;
  if (L.sets == 2) {
    if (chunk_streams(lanes) != OCCL_OK) return cuda_fail(cudaGetLastError(), "chunk streams");
    CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming), "event");
    CK(cudaEventRecord(ev_fork, (cudaStream_t)stream), "fork");
    CK(cudaStreamWaitEvent(lanes[0], ev_fork, 0), "fork");
    CK(cudaStreamWaitEvent(lanes[1], ev_fork, 0), "fork");
    CK(cudaEventDestroy(ev_fork), "event");
  }
  int ci = 0;
  for (int e0 = 0; e0 < n; e0 += L.chunk, ++ci) {
    const int m = n - e0 < L.chunk ? n - e0 : L.chunk;
    const size_t e = (size_t)e0;
    cudaStream_t lane = lanes[ci & 1];
    {
      unsigned char* sb = base + (size_t)(L.sets == 2 ? (ci & 1) : 0) * L.set_stride;
      p.geo = (uint4*)(sb + L.geo); p.rng = (uint4*)(sb + L.rng); p.n_live = (int*)(sb + L.n_live);
      p.tile_mask = (const uint32_t*)(sb + L.tile_mask);
      p.shade = (const float2*)(sb + L.shade);
      p.tile_idx = (int*)(sb + L.tile_idx);
      p.tile_cnt = bin ? (const int*)(sb + L.tile_cnt) : nullptr;
      sp.geo = p.geo; sp.rng = p.rng; sp.n_live = p.n_live; sp.tile_mask = (uint32_t*)(sb + L.tile_mask);
      sp.shade = (float2*)(sb + L.shade); sp.tile_idx = p.tile_idx; sp.tile_cnt = bin ? (int*)(sb + L.tile_cnt) : nullptr;
      sp.clip_list = (int*)(sb + L.clip_list); p.clip_list = sp.clip_list;
      CK(cudaMemsetAsync(sp.clip_list, 0, sizeof(int), lane), "memset clip list");
      p.env_list = (const int*)(sb + L.env_list);
    }
    p.vproj = p0.vproj + e * c.n_verts;
    p.vtan = p0.vtan ? p0.vtan + e * c.n_verts : nullptr;
    p.verts = p0.verts + e * (size_t)p0.verts_stride;
    p.faces = p0.faces + e * (size_t)p0.faces_stride;
    p.cam = p0.cam + e * OCCL_CAM_STRIDE;
    p.partials = p0.partials + e * L.n_tiles;
    p.obs = p0.obs + e * p0.obs_planes * npix;
    p.occl = p0.occl + e * npix;
    p.alphas = p0.alphas ? p0.alphas + e * c.n_obj * npix : nullptr;
    p.pix_to_face = p0.pix_to_face ? p0.pix_to_face + e * npix : nullptr;
    p.bary = p0.bary ? p0.bary + e * npix * 3 : nullptr;
    p.nhits = p0.nhits ? p0.nhits + e * c.n_obj * npix : nullptr;
    p.status = p0.status + e;
    p.obs_tile_state = p0.obs_tile_state ? p0.obs_tile_state + e * OCCL_TILE_STATE_WORDS : nullptr;
    p.env_mask = p0.env_mask ? p0.env_mask + e : nullptr;
    sp.vproj = p.vproj; sp.faces = p.faces; sp.verts = p.verts; sp.cam = p.cam; sp.status = p.status;
    sp.env_mask = p.env_mask;
    sp.obs_tile_state = (uint32_t*)p.obs_tile_state;
    const long long blocks = (long long)m * L.n_tiles;
    const unsigned clip_grid = (unsigned)(blocks < 2 * 148 ? blocks : 2 * 148);
    if (bin) face_setup_kernel<true><<<m, SETUP_THREADS, sizeof(float) * (size_t)c.image_size, lane>>>(sp);
    else face_setup_kernel<false><<<m, SETUP_THREADS, sizeof(float) * (size_t)c.image_size, lane>>>(sp);
    CK(cudaGetLastError(), "face_setup_kernel");
#define OCCL_LAUNCH_RASTER(G, W, H, D)                                                                                   \
  do {                                                                                                                   \
    CK((ensure_dyn_smem<raster_kernel<G, W, H, D>>(smem)), "smem attr");                                                  \
    raster_kernel<G, W, H, D><<<(unsigned)blocks, OCCL_THREADS, smem, lane>>>(p);                                         \
  } while (0)
#define OCCL_LAUNCH_RASTER_G(G)                                                                 \
  do {                                                                                          \
    if (fixed && !dbg) OCCL_LAUNCH_RASTER(G, OCCL_TILE_W, OCCL_TILE_H, false);                  \
    else if (fixed) OCCL_LAUNCH_RASTER(G, OCCL_TILE_W, OCCL_TILE_H, true);                      \
    else if (fixed2 && !dbg) OCCL_LAUNCH_RASTER(G, OCCL_TILE2_W, OCCL_TILE2_H, false);          \
    else if (fixed2) OCCL_LAUNCH_RASTER(G, OCCL_TILE2_W, OCCL_TILE2_H, true);                   \
    else if (fixed3 && !dbg) OCCL_LAUNCH_RASTER(G, OCCL_TILE3_W, OCCL_TILE3_H, false);          \
    else if (fixed3) OCCL_LAUNCH_RASTER(G, OCCL_TILE3_W, OCCL_TILE3_H, true);                   \
    else OCCL_LAUNCH_RASTER(G, 0, 0, true);                                                     \
  } while (0)
    if (p.env_mask && !grad) {
      // masked transition: list the flagged envs, then a fixed grid over (listed env, tile)
      int* el = (int*)p.env_list;
      CK(cudaMemsetAsync(el, 0, sizeof(int), lane), "memset env list");
      mask_list_kernel<<<(m + 255) / 256, 256, 0, lane>>>(m, p.env_mask, el);
      const unsigned lgrid = (unsigned)(blocks < 4 * 148 ? blocks : 4 * 148);
#define OCCL_LAUNCH_LIST(W, H, D)                                                              \
  do {                                                                                         \
    CK((ensure_dyn_smem<raster_list_kernel<W, H, D>>(smem)), "smem attr");                      \
    raster_list_kernel<W, H, D><<<lgrid, OCCL_THREADS, smem, lane>>>(p);                        \
  } while (0)
      if (fixed && !dbg) OCCL_LAUNCH_LIST(OCCL_TILE_W, OCCL_TILE_H, false);
      else if (fixed2 && !dbg) OCCL_LAUNCH_LIST(OCCL_TILE2_W, OCCL_TILE2_H, false);
      else if (fixed3 && !dbg) OCCL_LAUNCH_LIST(OCCL_TILE3_W, OCCL_TILE3_H, false);
      else OCCL_LAUNCH_LIST(0, 0, true);
#undef OCCL_LAUNCH_LIST
    } else if (grad) OCCL_LAUNCH_RASTER_G(true); else OCCL_LAUNCH_RASTER_G(false);
#undef OCCL_LAUNCH_RASTER_G
#undef OCCL_LAUNCH_RASTER
    // envs with faces cut at z_clip (status bit set by the setup kernel): one CTA per env, generic tile
    if (grad) {
      CK(ensure_dyn_smem<raster_clip_kernel<true>>(smem), "smem attr");
      raster_clip_kernel<true><<<clip_grid, OCCL_THREADS, smem, lane>>>(p);
    } else {
      CK(ensure_dyn_smem<raster_clip_kernel<false>>(smem), "smem attr");
      raster_clip_kernel<false><<<clip_grid, OCCL_THREADS, smem, lane>>>(p);
    }
    CK(cudaGetLastError(), "raster_kernel");
  }
  if (L.sets == 2) {
    for (int k = 0; k < 2; ++k) {  // join: the caller's stream continues when both lanes are done
      cudaEvent_t ev_join;
      CK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming), "event");
      CK(cudaEventRecord(ev_join, lanes[k]), "join");
      CK(cudaStreamWaitEvent((cudaStream_t)stream, ev_join, 0), "join");
      CK(cudaEventDestroy(ev_join), "event");
    }
  }
  return OCCL_OK;
}

extern "C" int occl_raster(const OcclConfig* cfg, int n, OcclScene sc, OcclWorkspace ws, OcclOutputs out, void* stream) {
  return raster_impl(cfg, n, sc, ws, out, nullptr, stream);
}

static int finalize_impl(const OcclConfig* cfg, int n, int mode, const float* action, OcclState st, OcclWorkspace ws,
                         OcclOutputs out, const uint8_t* mask, void* stream) {
  if (!cfg || n < 1 || !ws.base || mode < 0 || mode > 2) return OCCL_E_INVALID;
  const int grad = out.grad_action != nullptr;
  OcclConfig c = *cfg;
  int rc = occl_config_resolve(&c, grad);
  if (rc != OCCL_OK) return rc;
  if (mode != 2 && (!st.full_reward || !st.object_mass)) return OCCL_E_INVALID;
  if (mode == 0 && (!out.reward || !out.done)) return OCCL_E_INVALID;
  if (mode == 0 && grad && !action) return OCCL_E_INVALID;
  WsLayout L;
  ws_layout(&c, n, grad, &L);
  if (ws.bytes < L.total) return OCCL_E_INVALID;
  unsigned char* base = (unsigned char*)ws.base;
  finalize_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      n, mode, L.n_tiles, c.n_obj, c.norm_with_object_size, c.done_threshold, c.reward_done, c.reward_step, c.step_size,
      (const Partial*)(base + L.partials), action, st.full_reward, st.object_mass, out.reward, out.done, out.loss,
      out.n_covered, out.n_visible, mode == 0 ? out.grad_action : nullptr, out.status, out.status_or, mask);
  CK(cudaGetLastError(), "finalize_kernel");
  return OCCL_OK;
}

extern "C" int occl_finalize(const OcclConfig* cfg, int n, int mode, const float* action, OcclState st, OcclWorkspace ws,
                             OcclOutputs out, void* stream) {
  return finalize_impl(cfg, n, mode, action, st, ws, out, nullptr, stream);
}

static int run_chain(const OcclConfig* cfg, int n, int mode, const float* action, const float* R, const float* T,
                     const float* C, OcclScene sc, OcclState st, OcclWorkspace ws, OcclOutputs out, const uint8_t* mask,
                     void* stream) {
  if (!cfg || n < 1 || !ws.base) return OCCL_E_INVALID;
  if (check_outputs(out) != OCCL_OK) return OCCL_E_INVALID;
  const int grad = out.grad_action != nullptr;
  OcclConfig c = *cfg;
  int rc = occl_config_resolve(&c, grad);
  if (rc != OCCL_OK) return rc;
  WsLayout L;
  ws_layout(&c, n, grad, &L);
  if (ws.bytes < L.total) return OCCL_E_INVALID;
  unsigned char* base = (unsigned char*)ws.base;
  float* cam = (float*)(base + L.cam);
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == 2) {
    CK(cudaMemsetAsync(out.status, 0, sizeof(uint32_t) * (size_t)n, s), "memset status");
    rc = occl_pose_set(n, R, T, C, cam, stream);
  } else {
    rc = launch_pose(&c, n, mode, action, st, cam, out.position, out.status, mask, s);
  }
  if (rc != OCCL_OK) return rc;
  rc = project_impl(&c, n, cam, sc, (float*)(base + L.vproj), grad ? (float*)(base + L.vtan) : nullptr, out.status, mask,
                    stream);
  if (rc != OCCL_OK) return rc;
  rc = raster_impl(&c, n, sc, ws, out, mask, stream);
  if (rc != OCCL_OK) return rc;
  return finalize_impl(&c, n, mode, action, st, ws, out, mask, stream);
}

extern "C" int occl_step(const OcclConfig* cfg, int n, const float* action, OcclScene sc, OcclState st, OcclWorkspace ws,
                         OcclOutputs out, void* stream) {
  if (!action) return OCCL_E_INVALID;
  return run_chain(cfg, n, 0, action, nullptr, nullptr, nullptr, sc, st, ws, out, nullptr, stream);
}
extern "C" int occl_reset(const OcclConfig* cfg, int n, const uint8_t* env_mask, OcclScene sc, OcclState st,
                          OcclWorkspace ws, OcclOutputs out, void* stream) {
  OcclOutputs o = out;
  o.grad_action = nullptr;
  return run_chain(cfg, n, 1, nullptr, nullptr, nullptr, nullptr, sc, st, ws, o, env_mask, stream);
}
extern "C" int occl_render(const OcclConfig* cfg, int n, const float* R, const float* T, const float* C, OcclScene sc,
                           OcclWorkspace ws, OcclOutputs out, void* stream) {
  if (!R || !T || !C) return OCCL_E_INVALID;
  OcclOutputs o = out;
  o.grad_action = nullptr;
  OcclState st = {nullptr, nullptr, nullptr, nullptr, nullptr};
  return run_chain(cfg, n, 2, nullptr, R, T, C, sc, st, ws, o, nullptr, stream);
}
