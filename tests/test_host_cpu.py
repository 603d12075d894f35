"""CPU suite for the host-side logic: spaces, VecEnv contract, meshes, config, C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from occlusionenv_b200 import spaces
from occlusionenv_b200.baseVecEnv import (AlreadySteppingError, CloudpickleWrapper, NotSteppingError, VecEnv,
                                          VecEnvWrapper, tile_images)
from occlusionenv_b200.config import RasterConfig
from occlusionenv_b200.meshes import (default_scene, icosphere, load_obj, load_teapot, pack_scene, procedural_scene)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_box_space():
    b = spaces.Box(0, 1, shape=(4, 8, 8))
    assert b.shape == (4, 8, 8) and b.dtype == np.float32 and b.low.min() == 0 and b.high.max() == 1
    a = spaces.Box(low=-0.1, high=0.1, shape=(2,))
    a.seed(0)
    s = a.sample()
    assert s.shape == (2,) and a.contains(s) and not a.contains(np.array([1.0, 0.0], np.float32))


def test_tile_images():
    imgs = np.arange(5 * 2 * 3 * 1, dtype=np.float32).reshape(5, 2, 3, 1)
    out = tile_images(imgs)
    assert out.shape == (3 * 2, 2 * 3, 1)  # rows = ceil(sqrt(5)) = 3, cols = ceil(5/3) = 2
    assert np.array_equal(out[0:2, 0:3], imgs[0]) and np.array_equal(out[0:2, 3:6], imgs[1])
    assert np.array_equal(out[4:6, 0:3], imgs[4]) and (out[4:6, 3:6] == 0).all()


class _Toy(VecEnv):
    def __init__(self, n):
        super().__init__(n, spaces.Box(0, 1, shape=(1,)), spaces.Box(-1, 1, shape=(2,)))
        self.pending = None
        self.toy_attr = 7

    def reset(self):
        return np.zeros((self.num_envs, 1))

    def step_async(self, actions):
        if self.pending is not None:
            raise AlreadySteppingError()
        self.pending = actions

    def step_wait(self):
        if self.pending is None:
            raise NotSteppingError()
        a, self.pending = self.pending, None
        return np.asarray(a).sum(1, keepdims=True), np.ones(self.num_envs), np.zeros(self.num_envs, bool), [{}] * self.num_envs

    def close(self):
        pass

    def get_attr(self, attr_name, indices=None):
        return [getattr(self, attr_name) for _ in self._get_indices(indices)]

    def set_attr(self, attr_name, value, indices=None):
        setattr(self, attr_name, value)

    def env_method(self, method_name, *a, indices=None, **k):
        return [getattr(self, method_name)(*a, **k) for _ in self._get_indices(indices)]

    def seed(self, seed=None):
        return [seed + i for i in range(self.num_envs)]


class _Wrap(VecEnvWrapper):
    def reset(self):
        return self.venv.reset()

    def step_wait(self):
        return self.venv.step_wait()


def test_vecenv_contract_and_wrapper():
    v = _Toy(3)
    obs, r, d, info = v.step(np.ones((3, 2)))
    assert obs.shape == (3, 1) and (obs == 2).all() and len(info) == 3
    with pytest.raises(NotSteppingError):
        v.step_wait()
    v.step_async(np.ones((3, 2)))
    with pytest.raises(AlreadySteppingError):
        v.step_async(np.ones((3, 2)))
    v.step_wait()
    assert list(v._get_indices(None)) == [0, 1, 2] and v._get_indices(1) == [1]
    w = _Wrap(v)
    assert w.unwrapped is v and w.num_envs == 3 and w.toy_attr == 7
    assert w.get_attr("toy_attr", 0) == [7] and w.seed(5) == [5, 6, 7]
    with pytest.raises(AttributeError):
        w.does_not_exist
    assert w.step(np.zeros((3, 2)))[0].shape == (3, 1)


def test_cloudpickle_wrapper_roundtrip():
    import pickle
    f = CloudpickleWrapper(lambda: 41 + 1)
    g = pickle.loads(pickle.dumps(f))
    assert g.var() == 42


def test_raster_config_constants():
    c = RasterConfig()
    assert c.image_size == 512 and c.faces_per_pixel == 100 and c.sigma == 1e-4
    assert abs(c.blur_radius - 9.2102e-4) < 1e-8 and c.z_clip == 0.5
    assert abs(c.proj_scale - 3 ** 0.5) < 1e-6 and c.light == (2.0, 2.0, -2.0)


def test_meshes(tmp_path):
    v, f = load_teapot()
    assert v.shape == (1292, 3) and f.shape == (2464, 3) and f.min() == 0 and f.max() == 1291
    ref_obj = "/root/reference/data/teapot.obj"
    if os.path.exists(ref_obj):  # only in the build container
        v2, f2 = load_obj(ref_obj)
        assert np.array_equal(v, v2) and np.array_equal(f, f2)
    p = tmp_path / "quad.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1 4//1\nf -4 -3 -2\n")
    qv, qf = load_obj(str(p))
    assert qv.shape == (4, 3) and qf.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2]]
    sc = default_scene("teapot")
    assert sc.n_obj == 2 and sc.verts.shape == (2584, 3) and sc.faces.shape == (4928, 3)
    assert np.allclose(sc.verts[1292:] - sc.verts[:1292], [2, 0, 0])
    ov, of = sc.object(1)
    assert of.min() == 0 and np.array_equal(of, f)
    iv, iff = icosphere(2)
    assert iff.shape == (320, 3) and iv.shape == (162, 3)
    ps = procedural_scene(3, subdiv=2)
    assert ps.n_obj == 3 and ps.faces.shape == (960, 3)
    assert pack_scene([(iv, iff)]).n_obj == 1


def test_cabi_library_exports_every_declared_symbol(cuda_lib):
    hdr = open(os.path.join(ROOT, "include", "occl_b200.h")).read()
    declared = set(re.findall(r"\b(occl_[a-z_0-9]+)\s*\(", hdr))
    from occlusionenv_b200 import _lib
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert raw.occl_abi_version() == _lib.OCCL_ABI_VERSION
    # pure host-side entry points may be called without a GPU
    c = _lib.OcclConfig()
    c.image_size, c.n_obj, c.n_verts, c.n_faces, c.faces_per_pixel = 128, 2, 2584, 4928, 100
    c.obj_face_start[1], c.obj_face_start[2] = 2464, 4928
    c.blur_radius, c.sigma = 9.2e-4, 1e-4
    assert cuda_lib.occl_config_resolve(ctypes.byref(c), 0) == 0 and c.tile_w > 0 and c.tile_h > 0
    assert cuda_lib.occl_workspace_bytes(ctypes.byref(c), 4096, 1) > 4096 * 2584 * 16
    c.n_obj = 9
    assert cuda_lib.occl_config_resolve(ctypes.byref(c), 0) == -1


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from occlusionenv_b200 import _lib
    from occlusionenv_b200.environment import OcclusionEnv
    with pytest.raises(_lib.OcclError):
        OcclusionEnv(img_size=64)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "occlusionenv_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "libocclusion_oracle" not in txt, fn


def test_dataset_generator_file_formats(tmp_path):
    """datasetGenerator.py:99-124: RGB jpg = img_as_ubyte(obs[..., :3]), Occl png = img_as_ubyte(alpha), Depth png =
    depth with background -1 -> 0, times 51, truncated; params.pickle = flat float64 rows [j, el, az, g0, g1]."""
    import pickle

    import cv2

    from occlusionenv_b200 import datasetGenerator as G
    S = 16
    rng = np.random.default_rng(0)
    obs = rng.uniform(0, 1, (4, S, S)).astype(np.float32)
    obs[3] = rng.uniform(2.0, 4.9, (S, S)).astype(np.float32)
    obs[3, :4] = -1.0
    obs[0, 0, 0] = 1.0000001  # shaded colour one ulp above 1: skimage would raise, this clips
    occl = rng.uniform(0, 1, (S, S)).astype(np.float32)
    assert G.img_as_ubyte(np.array([0.0, 0.5, 1.0, -0.2])).tolist() == [0, 128, 255, 0]
    d = G.encode_depth(obs)
    assert (d[:4] == 0).all() and d[5, 5] == int(np.float32(obs[3, 5, 5] * np.float32(51)))
    run = tmp_path / "run_0"
    for sub in ("Depth", "RGB", "Occl"):
        os.makedirs(run / sub)
    G.write_frame(str(run), 3, obs, occl)
    png = cv2.imread(str(run / "Occl" / "3.png"), cv2.IMREAD_UNCHANGED)
    assert png.shape == (S, S) and np.array_equal(png, G.encode_occlusion(occl))
    dep = cv2.imread(str(run / "Depth" / "3.png"), cv2.IMREAD_UNCHANGED)
    assert np.array_equal(dep, d)
    jpg = cv2.imread(str(run / "RGB" / "3.jpg"), cv2.IMREAD_UNCHANGED)
    assert jpg.shape == (S, S, 3)
    rows = np.arange(10, dtype=np.float64).reshape(2, 5)
    G.write_params(str(run), rows)
    back = pickle.load(open(run / "params.pickle", "rb"))
    assert back.shape == (10,) and back.dtype == np.float64 and np.array_equal(back, rows.reshape(-1))


def test_status_bits_match_header():
    """The Python binding's status bits are the header's (include/occl_b200.h)."""
    from occlusionenv_b200 import _lib as L
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "occl_b200.h")).read()
    bits = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define OCCL_ST_(\w+) (\d+)u", hdr)}
    assert bits == {"ZCLIP": L.ST_ZCLIP, "KOVERFLOW": L.ST_KOVERFLOW, "HITCAP": L.ST_HITCAP, "OVFCAP": L.ST_OVFCAP,
                    "CLIPPED": L.ST_CLIPPED}


def test_workspace_is_sized_per_chunk_not_per_batch():
    """the BASELINE config-3 shape (8192 envs x 61 440 faces x 256^2) must fit: the per-face scratch
    is sized for a chunk of envs (OcclConfig.ws_budget_mb), 78 GB if it were sized for N."""
    import ctypes
    from occlusionenv_b200 import _lib as L
    lib = L.load()

    def ws(n, S, V, F, starts, budget=0, grad=0):
        c = L.OcclConfig()
        c.image_size, c.n_obj, c.n_verts, c.n_faces = S, len(starts) - 1, V, F
        for i in range(L.OCCL_MAX_OBJ + 1):
            c.obj_face_start[i] = starts[min(i, len(starts) - 1)]
        c.faces_per_pixel, c.cull_backfaces, c.blur_radius, c.sigma, c.ws_budget_mb = 100, 1, 9.2e-4, 1e-4, budget
        return lib.occl_workspace_bytes(ctypes.byref(c), n, grad)

    c3 = ws(8192, 256, 30726, 61440, [0, 20480, 40960, 61440])
    assert 0 < c3 <= 30 * 2 ** 30, c3
    assert ws(8192, 256, 30726, 61440, [0, 20480, 40960, 61440], grad=1) <= 30 * 2 ** 30
    c2 = ws(4096, 128, 1300, 2476, [0, 2464, 2476])
    assert 0 < c2 < 2 * 2 ** 30
    assert ws(4096, 128, 1300, 2476, [0, 2464, 2476], budget=64) < c2 // 4  # a smaller budget, smaller chunks


def test_peer_view_slices_by_rows():
    from occlusionenv_b200.dist import PeerView
    v = PeerView(0x1000, (8, 2, 4, 4))
    s = v[2:5]
    assert s.shape == (3, 2, 4, 4) and s.data_ptr() == 0x1000 + 2 * 2 * 4 * 4 * 4
    assert v[6:].shape == (2, 2, 4, 4)


def test_struct_fields_match_header_in_order():
    """The ctypes mirrors list the header's struct members in the header's order (a silent mismatch would shift every
    pointer), and the tile-state width of the incremental delivery is the header's."""
    from occlusionenv_b200 import _lib as L
    hdr = open(os.path.join(ROOT, "include", "occl_b200.h")).read()
    for name, cls in (("OcclConfig", L.OcclConfig), ("OcclScene", L.OcclScene), ("OcclState", L.OcclState),
                      ("OcclWorkspace", L.OcclWorkspace), ("OcclOutputs", L.OcclOutputs)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        members = [re.search(r"(\w+)\s*(\[[^\]]*\])?\s*$", part.strip()).group(1)
                   for d in body.split(";") if d.strip() for part in d.split(",")]      # "int32_t tile_w, tile_h;"
        assert members == [f[0] for f in cls._fields_], (name, members)
    assert int(re.search(r"#define OCCL_TILE_STATE_WORDS (\d+)", hdr).group(1)) == L.OCCL_TILE_STATE_WORDS
    assert int(re.search(r"#define OCCL_ABI_VERSION (\d+)", hdr).group(1)) == L.OCCL_ABI_VERSION
