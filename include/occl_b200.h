/*
 * occl_b200.h -- C-ABI of the B200-native OcclusionEnv transition (libocclb200.so).
 *
 * The reference (MILAB-IIT-CV/OcclusionEnv) has no FFI seam of its own: the seam is the pair of
 * pytorch3d callables it builds in environment.py:258-284 (silhouette_renderer / phong_renderer,
 * i.e. pytorch3d's _C.rasterize_meshes + SoftSilhouetteShader + HardFlatShader) and the arithmetic
 * around them in OcclusionEnv.reset / .step (environment.py:286-328, 352-396).  Each entry point
 * below names the reference lines it replaces.  INTEGRATION.md shows the ctypes stub a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch tensors); nothing is allocated
 *     or freed inside; entry points are stateless and re-entrant (one host thread per stream);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - return value 0 = OK, negative = OCCL_E_*;
 *   - all floating point is fp32, indices are int32;
 *   - per-env `status` word: OCCL_ST_* bits (never silently wrong: conditions the kernels do not
 *     implement are flagged).
 */
#ifndef OCCL_B200_H_
#define OCCL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCCL_ABI_VERSION 3
#define OCCL_MAX_OBJ 4
#define OCCL_CAM_STRIDE 48 /* floats per env in the camera block, see occl_pose_* */

/* error codes */
#define OCCL_OK 0
#define OCCL_E_INVALID (-1)  /* bad argument / unsupported configuration */
#define OCCL_E_CUDA (-2)     /* a CUDA runtime call or launch failed (see occl_last_cuda_error) */
#define OCCL_E_SMEM (-3)     /* tile does not fit in shared memory */

/* per-env status bits */
#define OCCL_ST_ZCLIP 1u       /* reserved: was "gradient requested through a face cut at z_clip" before the cut
                                  triangles carried tangents; never set */
#define OCCL_ST_CLIPPED 16u    /* informational: some face was cut at z_clip (clip_faces cases 3/4) */
#define OCCL_ST_KOVERFLOW 2u   /* some pixel had more than faces_per_pixel hits (handled: nearest-K rule applied) */
#define OCCL_ST_HITCAP 4u      /* a pixel had more hits than the top-K selection buffer: alpha of that pixel is wrong */
#define OCCL_ST_OVFCAP 8u      /* reserved (overflowing pixels are handled in rounds; never set) */

/* Raster / reward constants: a 1:1 mirror of createRenderers (environment.py:234-284) and of the
 * constants of step() (environment.py:219, 386-392). */
typedef struct OcclConfig {
  int32_t image_size;                       /* S; environment.py:202,250,268 */
  int32_t n_obj;                            /* objects in the scene, 1..OCCL_MAX_OBJ */
  int32_t n_verts;                          /* V of the packed scene mesh */
  int32_t n_faces;                          /* F of the packed scene mesh */
  int32_t obj_face_start[OCCL_MAX_OBJ + 1]; /* face range of object i = [start[i], start[i+1]) */
  int32_t faces_per_pixel;                  /* K = 100; environment.py:252 */
  int32_t cull_backfaces;                   /* 1; environment.py:253,271 */
  int32_t norm_with_object_size;            /* environment.py:208,324 */
  int32_t tile_w, tile_h;                   /* CTA tile in pixels; 0 = automatic */
  float blur_radius;                        /* ln(1/1e-4 - 1) * sigma; environment.py:251 */
  float sigma;                              /* 1e-4; environment.py:242 */
  float proj_scale;                         /* 1/tan(fov/2) of FoVPerspectiveCameras(); :238 */
  float z_clip;                             /* znear/2 = 0.5 (MeshRasterizer z_clip_value) */
  float step_size;                          /* 0.05; environment.py:219 */
  float light[3];                           /* PointLights location (2,2,-2); environment.py:275 */
  float done_threshold;                     /* 0.1; environment.py:386 */
  float reward_done;                        /* +5;  environment.py:390 */
  float reward_step;                        /* -0.2; environment.py:392 (stored as +0.2, subtracted) */
  int32_t debug_exact;                      /* 1: evaluate every (pixel, face) pair with the reference's exact
                                               operation sequence (no guarded fast path); for parity tests */
  int32_t ws_budget_mb;                     /* cap, in MiB, of the rasteriser's per-face scratch inside the workspace
                                               (0 = 8192): the batch is rasterised in chunks of as many envs as fit,
                                               so the workspace does not grow as N x F (occl_workspace_bytes) */
  int32_t obs_planes;                       /* 0 / 4: obs is (N,4,S,S) R, G, B, depth -- the reference's layout
                                               (environment.py:376-378); 2: (N,2,S,S) grey, depth -- the same
                                               information (flat shading of white vertices gives R = G = B) in half
                                               the bytes, for transport to the learner (occlusionenv_b200/dist.py) */
} OcclConfig;

/* Scene mesh in HBM. verts: (V,3) f32 world coordinates, faces: (F,3) i32 into verts.
 * *_env_stride = 0: one mesh shared by the whole batch (teapot configs); otherwise the number of
 * ELEMENTS (floats / ints) between consecutive environments' meshes (per-env meshes, config 3). */
typedef struct OcclScene {
  const float* verts;
  const int32_t* faces;
  int64_t verts_env_stride;
  int64_t faces_env_stride;
} OcclScene;

/* Per-env state of OcclusionEnv (environment.py:302-306,323-324), each (N,) f32. */
typedef struct OcclState {
  float* elevation;
  float* azimuth;
  float* radius;
  float* full_reward;  /* self.fullReward: loss of the previous render */
  float* object_mass;  /* self.objectMass */
} OcclState;

/* Scratch, sized by occl_workspace_bytes(); contents are meaningless between calls. */
typedef struct OcclWorkspace {
  void* base;
  size_t bytes;
} OcclWorkspace;

/* Outputs. Pointers marked [opt] may be NULL. */
typedef struct OcclOutputs {
  float* obs;            /* (N,4,S,S) flat-shaded RGB + depth(-1 bg); environment.py:375-378
                            ((N,2,S,S) grey + depth with OcclConfig.obs_planes = 2); may be PEER memory:
                            the rows are stored by the rasteriser's epilogue wherever this points              */
  float* occl;           /* (N,S,S)  alpha of self.image = sum_{i<j} A_i A_j; environment.py:373   */
  float* reward;         /* (N,)     environment.py:382-392   (not written by occl_reset)          */
  uint8_t* done;         /* (N,)     environment.py:386   (occl_reset without a mask: loss <= threshold;
                                     a masked occl_reset does not write it, so it may alias env_mask)       */
  float* loss;           /* (N,)     new self.fullReward; environment.py:381,384                   */
  float* position;       /* (N,3)    camera centre; info['position'], environment.py:363-365       */
  int32_t* n_covered;    /* (N,n_obj) px hard-covered by object i rendered alone (exact)           */
  int32_t* n_visible;    /* (N,n_obj) px whose nearest scene face belongs to object i (exact)      */
  uint32_t* status;      /* (N,)     OCCL_ST_* bits                                                */
  float* grad_action;    /* [opt] (N,2) d reward / d action; non-NULL selects the differentiable step */
  float* alphas;         /* [opt] (N,n_obj,S,S) per-object soft silhouettes (image_i[...,3])       */
  int32_t* pix_to_face;  /* [opt] (N,S,S)  scene K=1 fragments.pix_to_face (packed face index)     */
  float* bary;           /* [opt] (N,S,S,3) scene K=1 barycentrics (-1 on background)              */
  int32_t* nhits;        /* [opt] (N,n_obj,S,S) soft hits per pixel before the K cut               */
  uint32_t* status_or;   /* [opt] (1,) running OR of every status word written through this struct:
                            ORed into, never cleared, by occl_finalize / occl_step / occl_reset / occl_render
                            (the caller zeroes it when it has looked) -- one word to check instead of (N,) */
  uint32_t* obs_tile_state; /* [opt] (N, OCCL_TILE_STATE_WORDS) INCREMENTAL delivery of `obs`: the state of the
                            destination `obs` points to, owned by the caller next to that destination and filled with
                            0xFF bytes before the first transition into it (and whenever somebody else wrote there).
                            Words 0..7 of an env: bit t = tile t held a face at the last render; the library updates them.
                            A tile that was background at the last render and is background now is NOT stored again
                            (its pixels already are (1,1,1,-1)): the destination ends bit-identical to a full write while a
                            peer destination receives only the tiles that changed.  Ignored (full writes) for images of
                            more than 256 tiles.  NULL = every pixel is written.                                        */
} OcclOutputs;
#define OCCL_TILE_STATE_WORDS 16

int occl_abi_version(void);

/* Human-readable text of the last CUDA error seen by this library on the calling thread. */
const char* occl_last_cuda_error(void);

/* Let kernels launched on the CURRENT device store through pointers into `peer_device`'s memory (NVLink / NVSwitch
 * peer access): needed once per process before OcclOutputs.obs may point into another GPU's gather buffer -- the
 * fused "render + deliver to the learner" path of occlusionenv_b200/dist.py (the cross-GPU torch.stack of
 * SubProcVecEnv.py:219).  Idempotent. */
int occl_enable_peer_access(int peer_device);

/* Map a block another PROCESS exported with cudaIpcGetMemHandle (64-byte handle) into the address space of the
 * calling process's CURRENT device, peer access to the exporting GPU included (cudaIpcMemLazyEnablePeerAccess).
 * The returned pointer may be handed to OcclOutputs.obs of a transition running on this device.  (torch's own
 * CUDA-IPC rebuild maps the block in the EXPORTING device's context of the importing process: its copies work,
 * kernels of another device fault on it -- measured.) */
int occl_ipc_open(const void* handle64, void** ptr_out);
/* The exporting side: the 64-byte CUDA IPC handle of the cudaMalloc block `ptr` lies in, and ptr's offset in it. */
int occl_ipc_export(const void* ptr, void* handle64_out, size_t* offset_out);
int occl_ipc_close(void* ptr);

/* Self-test of the rasteriser's division primitive: counts, over n_samples pseudo-random operand
 * pairs of its guarded domain, the results that are not bit-identical to IEEE `a / b`.
 * mismatches_dev: one device u64 (expected 0). */
int occl_selftest_div(unsigned long long n_samples, unsigned long long seed, unsigned long long* mismatches_dev,
                      void* stream);

/* Fill tile_w/tile_h when they are 0 and validate the configuration. */
int occl_config_resolve(OcclConfig* cfg, int with_grad);

/* Scratch bytes for n_envs environments: camera blocks, projected vertices and tile partials for all of them,
 * plus the per-face scratch of one chunk of envs (see OcclConfig.ws_budget_mb). */
size_t occl_workspace_bytes(const OcclConfig* cfg, int n_envs, int with_grad);

/* Byte offsets inside the workspace of: camera blocks (N,OCCL_CAM_STRIDE) f32, projected vertices
 * (N,V,4) f32, vertex tangents (N,V,4) f32 (with_grad only), tile partials.  For callers that drive
 * the stages separately (occl_pose_* -> occl_project -> occl_raster -> occl_finalize). */
int occl_workspace_offsets(const OcclConfig* cfg, int n_envs, int with_grad, size_t* offsets4);

/* environment.py:356-368: normalise the action, integrate elevation/azimuth in place, camera centre
 * (step convention), look_at_rotation, T = -R^T C.  action (N,2) f32.  Writes the camera block
 * cam (N,OCCL_CAM_STRIDE): R[9] T[3] C[3] pad, then d/d_elevation and d/d_azimuth of the same. */
int occl_pose_step(const OcclConfig* cfg, int n_envs, const float* action, OcclState state, float* cam,
                   void* stream);

/* environment.py:304-308: look_at_view_transform(radius, elevation, azimuth, degrees=False). */
int occl_pose_lookat(const OcclConfig* cfg, int n_envs, OcclState state, float* cam, void* stream);

/* Explicit cameras (what the reference passes as R=, T= to the renderers, environment.py:370-375):
 * R (N,3,3) row-major, T (N,3), C (N,3) camera centres (used by the specular term). */
int occl_pose_set(int n_envs, const float* R, const float* T, const float* C, float* cam, void* stream);

/* MeshRasterizer.transform (world -> view -> NDC, z := view z) for every env and vertex.
 * vproj (N,V,4) f32 = (x_ndc, y_ndc, z_view, 0); vtan [opt] (N,V,4) = d(x,y)/d_el, d(x,y)/d_az.
 * status [opt] (N,): zeroed here -- a transition driven through the split entry points starts with this call
 * (occl_raster then ORs its flags in); pass the same array as OcclOutputs.status. */
int occl_project(const OcclConfig* cfg, int n_envs, const float* cam, OcclScene scene, float* vproj,
                 float* vtan, uint32_t* status, void* stream);

/* pytorch3d rasterize_meshes (silhouette settings, per object) + SoftSilhouetteShader +
 * rasterize_meshes (K=1, scene) + HardFlatShader + hard_rgb_blend + the image products and partial
 * sums of environment.py:373,381.  Consumes vproj/vtan from occl_project in the workspace. */
int occl_raster(const OcclConfig* cfg, int n_envs, OcclScene scene, OcclWorkspace ws, OcclOutputs out,
                void* stream);

/* environment.py:381-392 (mode 0, step) or :322-327 (mode 1, reset): loss, reward, done, state update,
 * and the chain rule to the action for the differentiable step. */
int occl_finalize(const OcclConfig* cfg, int n_envs, int mode, const float* action, OcclState state,
                  OcclWorkspace ws, OcclOutputs out, void* stream);

/* The whole transition: occl_pose_step -> occl_project -> occl_raster -> occl_finalize(step).
 * Replaces OcclusionEnv.step (environment.py:352-396) for N environments. */
int occl_step(const OcclConfig* cfg, int n_envs, const float* action, OcclScene scene, OcclState state,
              OcclWorkspace ws, OcclOutputs out, void* stream);

/* The render half of OcclusionEnv.reset (environment.py:302-328) for N environments whose
 * elevation/azimuth/radius have been written into `state` by the caller.
 * env_mask [opt] (N,) u8: only environments with a non-zero entry are reset, the others keep their
 * state and outputs untouched -- the auto-reset of SimpleVecEnv.step_wait (SubProcVecEnv.py:211-214)
 * without a host round trip (pass the `done` output of occl_step; a masked reset leaves `out.done` alone,
 * so mask and output may be the same array). */
int occl_reset(const OcclConfig* cfg, int n_envs, const uint8_t* env_mask, OcclScene scene, OcclState state,
               OcclWorkspace ws, OcclOutputs out, void* stream);

/* Render from explicit cameras (environment.py:332-336 render(); parity tests): occl_pose_set ->
 * occl_project -> occl_raster.  No state update, no reward. */
int occl_render(const OcclConfig* cfg, int n_envs, const float* R, const float* T, const float* C,
                OcclScene scene, OcclWorkspace ws, OcclOutputs out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OCCL_B200_H_ */
