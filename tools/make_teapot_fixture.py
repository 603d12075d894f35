"""Converts the reference's data fixture data/teapot.obj into occlusionenv_b200/data/teapot.npz
(verts f32 (1292,3), faces i32 (2464,3)).  Run once in the build container, where /root/reference
exists; the GPU box only sees the committed .npz."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from occlusionenv_b200.meshes import load_obj  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/teapot.obj"
v, f = load_obj(src)
assert v.shape == (1292, 3) and f.shape == (2464, 3), (v.shape, f.shape)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "occlusionenv_b200", "data", "teapot.npz")
np.savez_compressed(out, verts=v, faces=f)
print("wrote", out, v.shape, f.shape)
