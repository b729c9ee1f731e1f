#!/usr/bin/env python
"""Work counters of the backward edge scan on bench-shaped frames (host emulation, no GPU):
items, crossings, tasks, bitmap words, contributing pixels per frame.  python tools/bwd_stats.py [frames]"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emu_lib  # noqa: E402
from bench import H, W, oracle_render_fn  # noqa: E402
import bench as _bench  # noqa: E402
MESH = os.environ.get("DH_STATS_MESH", _bench.MESH)
from dynhor_b200 import synth  # noqa: E402

NAMES = ["front", "items", "span_iters", "crossings", "t_out", "t_out_owner", "words", "words_nz", "pairs", "t_in",
         "in_px", "in_pairs", "max_pairs_task", "n_neg"]


def main():
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    offs = np.linspace(0, 299, nf).astype(int)
    rows = []
    for off in offs:
        seq = synth.make_sequence(1, H, W, mesh=MESH, seed=0, render_fn=oracle_render_fn, period=300,
                                  frame_offset=int(off))
        S, is_ = 256, 512
        verts, faces = seq["verts"], seq["faces"].astype(np.int32)
        R = emu_lib.rot6d_to_R(seq["rot6d_init"])
        proj, cam = emu_lib.project_pose(verts, R, seq["T_init"], 1.0, seq["K_roi"])
        fidx, abits = emu_lib.raster(proj, faces, is_)
        mt = np.where(seq["target_masks"] > 0, 1, np.where(seq["target_masks"] >= 0, 0, -1)).astype(np.int8)
        counts, gpool, pos, neg, rend = emu_lib.loss_epilogue(abits, mt, S, True, np.float32(1e-6))
        out = np.zeros(16, np.int64)
        span_len = np.zeros((2 * len(faces), 6), np.int32)
        emu_lib.lib().emu_backward_stats(emu_lib._p(np.ascontiguousarray(proj[0])), emu_lib._p(faces),
                                         emu_lib._p(np.ascontiguousarray(fidx[0])),
                                         emu_lib._p(np.ascontiguousarray(abits[0])),
                                         emu_lib._p(np.ascontiguousarray(neg[0])), len(verts), len(faces), S, 1,
                                         emu_lib._p(out), emu_lib._p(span_len))
        sl = span_len[:out[1]]
        # items are processed in face order: the given windings of a chunk first, then (per face) ... approximate by order
        nb = len(sl) // 32
        bl = sl[:nb * 32].reshape(nb, 32, 6)
        A = bl.sum(2).max(1).sum()           # one loop over all six spans per lane
        Bv = bl.max(1).sum()                 # six uniform loops
        print("   loop iterations per frame: per-lane-sequential", int(A) + 6 * nb, " uniform-span", int(Bv),
              " ideal", int(sl.sum() / 32))
        rows.append(out[:len(NAMES)])
        print(off, dict(zip(NAMES, out.tolist())))
    m = np.mean(rows, 0)
    print("mean", dict(zip(NAMES, [round(float(x), 1) for x in m])))


if __name__ == "__main__":
    main()
