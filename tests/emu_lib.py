"""ctypes loader + numpy drivers for tests/emu/libdh_emu.so (host build of dh_core.h; TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-s", "-C", _HERE])
        _LIB = ctypes.CDLL(os.path.join(_HERE, "libdh_emu.so"))
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def rot6d_to_R(r6):
    r6 = np.ascontiguousarray(r6, np.float32).reshape(-1, 6)
    R = np.empty((len(r6), 9), np.float32)
    lib().emu_rot6d_to_R(_p(r6), _p(R), len(r6))
    return R.reshape(-1, 3, 3)


def rot6d_backward(r6, G):
    r6 = np.ascontiguousarray(r6, np.float32).reshape(-1, 6)
    G = np.ascontiguousarray(G, np.float64).reshape(-1, 9)
    g6 = np.empty((len(r6), 6), np.float64)
    lib().emu_rot6d_backward(_p(r6), _p(G), _p(g6), len(r6))
    return g6.reshape(-1, 3, 2)


def project_pose(verts_og, R, T, s_abs, K, orig=1.0):
    verts_og = np.ascontiguousarray(verts_og, np.float32)
    R = np.ascontiguousarray(R, np.float32).reshape(-1, 9)
    T = np.ascontiguousarray(T, np.float32).reshape(-1, 3)
    K = np.ascontiguousarray(K, np.float32).reshape(-1, 9)
    B, V = len(R), len(verts_og)
    proj = np.empty((B, V, 4), np.float32)
    cam = np.empty((B, V, 3), np.float32)
    lib().emu_project_pose(_p(verts_og), V, _p(R), _p(T), ctypes.c_float(s_abs), _p(K), ctypes.c_float(orig), B,
                           _p(proj), _p(cam))
    return proj, cam


def project_cam(verts_cam, K, orig=1.0):
    verts_cam = np.ascontiguousarray(verts_cam, np.float32)
    K = np.ascontiguousarray(K, np.float32).reshape(-1, 9)
    B, V = verts_cam.shape[:2]
    proj = np.empty((B, V, 4), np.float32)
    lib().emu_project_cam(_p(verts_cam), V, _p(K), ctypes.c_float(orig), B, _p(proj))
    return proj


def raster(proj, faces, is_, near=0.1, far=100.0):
    proj = np.ascontiguousarray(proj, np.float32)
    faces = np.ascontiguousarray(faces, np.int32)
    B, V = proj.shape[:2]
    fidx = np.empty((B, is_, is_), np.int32)
    abits = np.zeros((B, is_, is_ // 32), np.uint32)
    lib().emu_raster(_p(proj), _p(faces), B, V, len(faces), is_, ctypes.c_float(near), ctypes.c_float(far),
                     _p(fidx), _p(abits))
    return fidx, abits


def loss_epilogue(abits, mask_tri, S, aa, gcoef):
    B = abits.shape[0]
    wprp = (S + 31) // 32
    counts = np.zeros((B, 4), np.int32)
    gpool = np.zeros((B, S, S), np.float32)
    pos = np.zeros((B, S, wprp), np.uint32)
    neg = np.zeros((B, S, wprp), np.uint32)
    rend = np.zeros((B, S, S), np.float32)
    mt = np.ascontiguousarray(mask_tri, np.int8) if mask_tri is not None else None
    lib().emu_loss_epilogue(_p(abits), _p(mt), B, S, int(aa), ctypes.c_float(gcoef), _p(counts), _p(gpool), _p(pos),
                            _p(neg), _p(rend))
    return counts, gpool, pos, neg, rend


def grad_signs(g):
    g = np.ascontiguousarray(g, np.float32)
    n = g.size
    pos = np.zeros(n // 32, np.uint32)
    neg = np.zeros(n // 32, np.uint32)
    lib().emu_grad_signs(_p(g), ctypes.c_longlong(n), _p(pos), _p(neg))
    return pos, neg


def backward(proj, faces, fidx, abits, gpool, pos, neg, S, aa, eps=1e-4, verts_cam=None, K=None, orig=1.0):
    proj = np.ascontiguousarray(proj, np.float32)
    faces = np.ascontiguousarray(faces, np.int32)
    B, V = proj.shape[:2]
    F = len(faces)
    gf = np.zeros((B, 2 * F, 3, 2), np.float32)
    gv = None
    if verts_cam is not None:
        verts_cam = np.ascontiguousarray(verts_cam, np.float32)
        K = np.ascontiguousarray(K, np.float32).reshape(-1, 9)
        gv = np.zeros((B, V, 3), np.float32)
    lib().emu_backward(_p(proj), _p(faces), _p(np.ascontiguousarray(fidx, np.int32)),
                       _p(np.ascontiguousarray(abits, np.uint32)), _p(np.ascontiguousarray(gpool, np.float32)),
                       _p(np.ascontiguousarray(pos, np.uint32)), _p(np.ascontiguousarray(neg, np.uint32)), B, V, F, S,
                       int(aa), ctypes.c_float(eps), _p(gf), _p(verts_cam), _p(K), ctypes.c_float(orig), _p(gv))
    return gf, gv


def smooth_terms(rot6d, trans, scale, moments, V, B_total, lw_smooth, halo_prev=None, halo_next=None):
    rot6d = np.ascontiguousarray(rot6d, np.float32).reshape(-1, 6)
    trans = np.ascontiguousarray(trans, np.float32).reshape(-1, 3)
    B = len(rot6d)
    st = np.zeros((B, 16), np.float64)
    hp = np.ascontiguousarray(halo_prev, np.float32) if halo_prev is not None else None
    hn = np.ascontiguousarray(halo_next, np.float32) if halo_next is not None else None
    mom = np.ascontiguousarray(moments, np.float64)
    lib().emu_smooth_terms(_p(rot6d), _p(trans), _p(hp), _p(hn), ctypes.c_float(scale), _p(mom), V, B, B_total,
                           ctypes.c_double(lw_smooth), _p(st))
    return st


def corr_frames(records, R, T, s_abs, K, S, delta):
    records = np.ascontiguousarray(records, np.float32)
    B, C = records.shape[:2]
    R = np.ascontiguousarray(R, np.float32).reshape(-1, 9)
    T = np.ascontiguousarray(T, np.float32).reshape(-1, 3)
    K = np.ascontiguousarray(K, np.float32).reshape(-1, 9)
    sums = np.zeros((B, 16), np.float32)
    lib().emu_corr_frames(_p(records), B, C, _p(R), _p(T), ctypes.c_float(s_abs), _p(K), ctypes.c_float(S),
                          ctypes.c_float(delta), _p(sums))
    return sums


def adam(p, g, m, v, lr, t):
    lib().emu_adam(_p(p), _p(np.ascontiguousarray(g, np.float32)), _p(m), _p(v), ctypes.c_longlong(p.size),
                   ctypes.c_double(lr), t)


def fx_sum(x):
    """(hi, lo) words and double value of the exact sum of the doubles x (dh_core.h Fx128)."""
    x = np.ascontiguousarray(x, np.float64).reshape(-1)
    out = np.zeros(2, np.uint64)
    val = ctypes.c_double()
    lib().emu_fx_sum(_p(x), len(x), _p(out), ctypes.byref(val))
    return int(out[0]), int(out[1]), float(val.value)


def mesh_moments(verts):
    v = np.asarray(verts, np.float64)
    return np.concatenate([v.sum(0), (v[:, :, None] * v[:, None, :]).sum(0).reshape(-1)])


def full_grads(verts, faces, K_roi, mask_tri, rot6d, trans, S, lw_sil, lw_smooth, scale=1.0, aa=True):
    """The whole fused iteration (forward + backward, no Adam) through the emulated kernel arithmetic.
    Mirrors k_pose_prep / k_project / k_setup_bin / k_raster / k_backward / k_pose_update / k_finalize."""
    B = len(rot6d)
    V = len(verts)
    is_ = 2 * S if aa else S
    R = rot6d_to_R(rot6d)
    s_abs = abs(scale)
    proj, cam = project_pose(verts, R, trans, s_abs, K_roi)
    fidx, abits = raster(proj, faces, is_)
    keep_sum = float((np.asarray(mask_tri) >= 0).sum())
    gcoef = np.float32(np.float32(np.float32(lw_sil) / np.float32(B)) / np.float32(keep_sum))
    counts, gpool, pos, neg, rend = loss_epilogue(abits, mask_tri, S, aa, gcoef)
    gf, _ = backward(proj, faces, fidx, abits, gpool, pos, neg, S, aa)
    # per-face-vertex -> pose gradients (k_backward tail), float64 accumulation
    F = len(faces)
    faces2 = np.concatenate([faces, faces[:, ::-1]], 0)
    gT = np.zeros((B, 3))
    G = np.zeros((B, 3, 3))
    gs = 0.0
    Kr = np.asarray(K_roi, np.float64).reshape(B, 3, 3)
    for b in range(B):
        nz = np.argwhere(np.any(gf[b] != 0, axis=-1))
        for fn, k in nz:
            vid = faces2[fn, k]
            c = cam[b, vid].astype(np.float64)
            gu, gv = gf[b, fn, k].astype(np.float64)
            zc = c[2] + 1e-9
            x_, y_ = c[0] / zc, c[1] / zc
            gu1, gv1 = 2 * gu, -2 * gv
            gx_ = gu1 * Kr[b, 0, 0] + gv1 * Kr[b, 1, 0]
            gy_ = gu1 * Kr[b, 0, 1] + gv1 * Kr[b, 1, 1]
            gc = np.array([gx_ / zc, gy_ / zc, -(gx_ * x_ + gy_ * y_) / zc])
            gT[b] += gc
            G[b] += np.outer(s_abs * verts[vid].astype(np.float64), gc)
            gs += (verts[vid].astype(np.float64) @ R[b].astype(np.float64)) @ gc
    gs *= np.sign(scale)
    st = smooth_terms(rot6d, trans, scale, mesh_moments(verts), V, B, lw_smooth)
    gT += st[:, 0:3]
    G += st[:, 3:12].reshape(B, 3, 3)
    gs += st[:, 12].sum()
    g6 = rot6d_backward(rot6d, G)
    N = (B - 1) * V * 3.0
    out = {
        "loss_smooth_obj": st[:, 13].sum() / N if B > 1 else 0.0,
        "loss_sil_obj": counts[:, 0].sum() / 16.0 / keep_sum / B,
        "iou_object": float(np.mean((counts[:, 1].astype(np.float32) * np.float32(0.25)) /
                                    (counts[:, 2].astype(np.float32) * np.float32(0.25) + np.float32(1e-6)))),
        "grad_rot6d": g6, "grad_trans": gT.reshape(B, 1, 3), "grad_scale": gs,
        "fidx": fidx, "abits": abits, "rend": rend, "proj": proj, "cam": cam, "gpool": gpool, "grad_faces": gf,
    }
    return out


def roi_process(obj_bits, hand_bits, images, S, pad=5.0, expansion=0.3):
    """dh_roi.cu on the host (dh_roi_core.h): boxes, crop mask, tri-state target, image crop."""
    ob = np.ascontiguousarray(obj_bits, np.uint8)
    hb = np.ascontiguousarray(hand_bits, np.uint8) if hand_bits is not None else None
    im = np.ascontiguousarray(images, np.uint8) if images is not None else None
    B, H, W = ob.shape
    bbox = np.zeros((B, 4), np.float32)
    sq = np.zeros((B, 4), np.float32)
    cm = np.zeros((B, S, S), np.uint8)
    tg = np.zeros((B, S, S), np.float32)
    ci = np.zeros((B, 3, S, S), np.float32) if im is not None else None
    st = np.zeros(B, np.int32)
    lib().emu_roi_process(_p(ob), _p(hb), _p(im), B, H, W, S, ctypes.c_float(pad), ctypes.c_float(expansion), _p(bbox),
                          _p(sq), _p(cm), _p(tg), _p(ci), _p(st))
    return {"bbox": bbox, "square_bbox": sq, "crop_mask": cm.astype(bool), "target_crop_mask": tg, "crop_image": ci,
            "status": st}
