#!/usr/bin/env python
"""Hit statistics of the rasteriser's deferred-depth protocol on bench-shaped frames (host emulation, no GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes  # noqa: E402
import emu_lib  # noqa: E402
from bench import H, W, MESH  # noqa: E402
from dynhor_b200 import synth  # noqa: E402

NAMES = ["entries_pass0", "entries_pass1", "hits", "pretest_rejects", "deferred", "exact", "resolved", "not_deferrable"]
for off in (0, 100, 200):
    seq = synth.make_sequence(1, H, W, mesh=MESH, seed=0, render_fn=None, period=300, frame_offset=off)
    verts, faces = seq["verts"], seq["faces"].astype(np.int32)
    R = emu_lib.rot6d_to_R(seq["rot6d_init"])
    proj, cam = emu_lib.project_pose(verts, R, seq["T_init"], 1.0, seq["K_roi"])
    for order in (0, 1):
        out = np.zeros(8, np.int64)
        emu_lib.lib().emu_raster_stats(emu_lib._p(np.ascontiguousarray(proj[0])), emu_lib._p(faces), len(verts),
                                       len(faces), 512, ctypes.c_float(0.1), ctypes.c_float(100.0), order,
                                       emu_lib._p(out))
        print(off, "order", order, dict(zip(NAMES, out.tolist())))
