#!/usr/bin/env python
"""Real-data fixture ("the only real-data fixture is the shoe mesh"): the reference's object prior
assets/shoes/1229a2e6e97e_A_basketball_shoes_.obj (configs/custom_shoes.yaml obj_path), read with the minimal OBJ
reader of dynhor_b200/synth.py, centred and scaled exactly like run.py:110-112, stored as float32 vertices and int32
triangles.  Run in the build container (needs /root/reference):

    python tests/golden/make_shoe_mesh.py  ->  tests/golden/shoe_mesh.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dynhor_b200 import synth  # noqa: E402

SRC = "/root/reference/assets/shoes/1229a2e6e97e_A_basketball_shoes_.obj"

if __name__ == "__main__":
    verts, faces = synth.load_obj(SRC, normalize=True)
    assert faces.min() >= 0 and faces.max() < len(verts)
    out = os.path.join(ROOT, "tests", "golden", "shoe_mesh.npz")
    np.savez_compressed(out, verts=verts.astype(np.float32), faces=faces.astype(np.int32))
    print(out, verts.shape, faces.shape, os.path.getsize(out), "bytes;  extent",
          verts.min(0).round(3), verts.max(0).round(3))
