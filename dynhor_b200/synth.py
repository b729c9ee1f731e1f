"""Seeded synthetic sequences shaped like the reference's joint-optimisation inputs.

Host-side numpy only.  What `run.py` + stage 1 hand to `joint_optimize` (run.py:155-164,
pose_initializtion.py:460-471) is rebuilt here from a synthetic mesh and trajectory:
  mesh            normalised like run.py:111-112 (centred, max|v| = 0.5)
  camera          run.py:121-122   (focal 1.2*min(H,W), principal point (W//2, H//2))
  ROI box         run.py:37-43     (tight bbox +-5 px, squared, x1.3)  + utils/bbox.py:70-89
  K_roi           utils/camera.py:84-130 then rows 0-1 / REND_SIZE (pose_initializtion.py:275-277,327)
  target masks    tri-state {1 object, 0 background, -1 occluder} (utils/maskutils.py:24-28, jointopt.py:50-53)
Masks are produced by a caller-supplied silhouette renderer (`render_fn`) so that the product never touches
the CPU oracle: bench.py passes the CUDA renderer, the tests pass the oracle.
"""
import math

import numpy as np

REND_SIZE = 256  # ObjTracker/utils/constants.py:2


# ----------------------------------------------------------------------------- meshes
def uv_sphere_mesh(n_rings=50, n_seg=100, seed=0, bump=0.1, axes=(1.0, 0.6, 0.35)):
    """UV sphere with `n_rings` interior rings x `n_seg` segments + 2 poles: V = n_rings*n_seg + 2,
    F = 2*n_rings*n_seg.  Anisotropic, with a seeded low-frequency radial bump to break symmetry; outward
    counter-clockwise winding; centred and scaled to max|v| = 0.5 (run.py:111-112)."""
    rng = np.random.default_rng(seed)
    th = (np.arange(1, n_rings + 1) / (n_rings + 1.0)) * math.pi  # polar angle of interior rings
    ph = (np.arange(n_seg) / float(n_seg)) * 2.0 * math.pi
    T, P = np.meshgrid(th, ph, indexing="ij")
    d = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], -1).reshape(-1, 3)
    d = np.concatenate([np.array([[0.0, 0.0, 1.0]]), d, np.array([[0.0, 0.0, -1.0]])], 0)
    # low-frequency bump: sum of a few random plane waves on the sphere
    r = np.ones(len(d))
    for _ in range(4):
        k = rng.normal(size=3)
        k = k / np.linalg.norm(k) * rng.uniform(1.5, 3.0)
        r += (bump / 4.0) * np.sin(d @ k + rng.uniform(0, 2 * math.pi))
    v = d * r[:, None] * np.asarray(axes)[None, :]
    faces = []
    top, bot = 0, n_rings * n_seg + 1

    def vid(i, j):
        return 1 + i * n_seg + (j % n_seg)

    for j in range(n_seg):
        faces.append((top, vid(0, j), vid(0, j + 1)))
        faces.append((bot, vid(n_rings - 1, j + 1), vid(n_rings - 1, j)))
    for i in range(n_rings - 1):
        for j in range(n_seg):
            a, b, c, e = vid(i, j), vid(i + 1, j), vid(i + 1, j + 1), vid(i, j + 1)
            faces.append((a, b, c))
            faces.append((a, c, e))
    v = v - v.mean(0)
    v = v / np.linalg.norm(v, 2, 1).max() * 0.5
    return v.astype(np.float32), np.asarray(faces, dtype=np.int64)


def icosphere_mesh(subdiv=2, seed=0, axes=(1.0, 0.7, 0.5)):
    """Small test mesh: subdiv 2 -> 162 verts / 320 faces."""
    t = (1.0 + 5 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    v = [np.asarray(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    v = np.asarray(v)
    rng = np.random.default_rng(seed)
    k = rng.normal(size=3)
    v = v * (1.0 + 0.08 * np.sin(2.0 * (v @ k)))[:, None] * np.asarray(axes)[None, :]
    v = v - v.mean(0)
    v = v / np.linalg.norm(v, 2, 1).max() * 0.5
    return v.astype(np.float32), np.asarray(f, dtype=np.int64)


def load_obj(path, normalize=True):
    """Minimal Wavefront reader (v / f lines, triangles), normalised like run.py:110-112."""
    vs, fs = [], []
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith("v "):
                vs.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("f "):
                idx = [int(tok.split("/")[0]) - 1 for tok in line.split()[1:]]
                for k in range(1, len(idx) - 1):
                    fs.append((idx[0], idx[k], idx[k + 1]))
    v = np.asarray(vs, dtype=np.float32)
    if normalize:
        v = v - v.mean(0)
        v = v / np.linalg.norm(v, 2, 1).max() * 1.0 / 2.0
    return v.astype(np.float32), np.asarray(fs, dtype=np.int64)


# ----------------------------------------------------------------------------- poses
def axis_angle_to_matrix(w):
    """Rodrigues, float64.  w [...,3] -> [...,3,3]."""
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w, axis=-1, keepdims=True)
    k = w / np.maximum(th, 1e-30)
    K = np.zeros(w.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    th = th[..., None]
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def gt_trajectory(B, period=None):
    """Ground-truth object poses.  `period` (default B) fixes the motion per frame so that a
    frame-sharded run sees the same sequence as a single-GPU run of the same global length."""
    n = float(B if period is None else period)
    t = np.arange(B, dtype=np.float64)
    w = np.stack([0.8 * np.sin(2 * math.pi * t / n), 0.6 * np.sin(1.4 * math.pi * t / n + 1.0),
                  2 * math.pi * t / n], -1)
    R = axis_angle_to_matrix(w)
    T = np.stack([0.05 * np.sin(4 * math.pi * t / n), 0.05 * np.cos(2 * math.pi * t / n),
                  1.75 + 0.1 * np.sin(2 * math.pi * t / n)], -1)
    return R, T


def perturb_poses(R, T, seed=0, rot_deg=3.0, trans_sigma=0.01):
    rng = np.random.default_rng(seed + 1)
    dw = rng.normal(size=(len(R), 3)) * math.radians(rot_deg)
    Rn = R @ axis_angle_to_matrix(dw)
    Tn = T + rng.normal(size=T.shape) * trans_sigma
    return Rn, Tn


# ----------------------------------------------------------------------------- camera / ROI
def full_frame_K(H, W):
    f = 1.2 * min(H, W)  # run.py:121-122
    return np.array([[f, 0, W // 2], [0, f, H // 2], [0, 0, 1]], dtype=np.float32)


def roi_from_verts(verts_cam, K, H, W, pad=5.0, expansion=0.3):
    """Square ROI (x, y, b, b) around the projected object: run.py:37-43 with the mask's tight bbox replaced
    by the bbox of the projected vertices; utils/bbox.py:70-89 for the squaring."""
    uv = verts_cam @ K.T.astype(np.float64)
    uv = uv[:, :2] / uv[:, 2:3]
    x0, y0 = np.floor(uv.min(0))
    x1, y1 = np.ceil(uv.max(0))
    min_col, max_col = max(x0 - pad, 0.0), min(x1 + pad, float(W))
    min_row, max_row = max(y0 - pad, 0.0), min(y1 + pad, float(H))
    bw, bh = max_col - min_col, max_row - min_row
    cx, cy = min_col + bw / 2.0, min_row + bh / 2.0
    b = max(bw, bh) * (1.0 + expansion)
    return np.array([cx - b / 2.0, cy - b / 2.0, b, b], dtype=np.float32)


def K_crop_resize(K, box_xyxy, size=REND_SIZE):
    """utils/camera.py:84-130 for one box, float32 like the reference."""
    K = np.asarray(K, dtype=np.float32)
    x0, y0, x1, y1 = [np.float32(v) for v in box_xyxy]
    fw = fh = np.float32(size)
    cw, ch = x1 - x0, y1 - y0
    ccj, cci = (x0 + x1) / np.float32(2), (y0 + y1) / np.float32(2)
    cx = K[0, 2] + (cw - 1) / np.float32(2) - ccj
    cy = K[1, 2] + (ch - 1) / np.float32(2) - cci
    center_x, center_y = (cw - 1) / np.float32(2), (ch - 1) / np.float32(2)
    sx, sy = fw / cw, fh / ch
    newK = K.copy()
    newK[0, 0] = sx * K[0, 0]
    newK[1, 1] = sy * K[1, 1]
    newK[0, 2] = (fw - 1) / np.float32(2) + sx * (cx - center_x)
    newK[1, 2] = (fh - 1) / np.float32(2) + sy * (cy - center_y)
    return newK


def make_K_roi(verts, R, T, H, W, size=REND_SIZE):
    """Per-frame normalised ROI intrinsics [B,3,3] (rows 0-1 divided by `size`) and the ROI boxes."""
    K = full_frame_K(H, W)
    Ks, boxes = [], []
    for b in range(len(R)):
        vc = verts.astype(np.float64) @ R[b] + T[b]
        x, y, bb, _ = roi_from_verts(vc, K, H, W)
        Kr = K_crop_resize(K, [x, y, x + bb, y + bb], size)
        Kr[:2] = Kr[:2] / np.float32(size)
        Ks.append(Kr)
        boxes.append([x, y, bb, bb])
    return np.stack(Ks).astype(np.float32), np.asarray(boxes, dtype=np.float32)


# ----------------------------------------------------------------------------- masks
def add_disc_occluder(masks01, seed=0, radius=40.0, frame_offset=0):
    """Tri-state target masks: a seeded disc marked -1 wherever it does not cover the object
    (utils/maskutils.py:24-28).  The disc of a frame depends on (seed, global frame number) only, so a rank that
    builds frames [frame_offset, frame_offset + B) of a longer sequence gets the masks a full build would."""
    B, S, _ = masks01.shape
    yy, xx = np.mgrid[0:S, 0:S]
    out = masks01.astype(np.float32).copy()
    r = radius * S / float(REND_SIZE)
    for b in range(B):
        rng = np.random.default_rng([seed + 2, frame_offset + b])
        cx, cy = rng.uniform(0.2 * S, 0.8 * S, size=2)
        disc = (xx - cx) ** 2 + (yy - cy) ** 2 <= r * r
        out[b][disc & (masks01[b] <= 0)] = -1.0
    return out


def make_sequence(B, H=480, W=640, mesh="uv50x100", seed=0, render_fn=None, size=REND_SIZE, occluder=True,
                  period=None, frame_offset=0, traj_period=None):
    """Build one synthetic joint-optimisation problem.

    render_fn(verts_cam [B,V,3] f32, faces [F,3] i64, K_roi [B,3,3] f32, size) -> [B,size,size] silhouettes
    in [0,1] rendered WITHOUT anti-aliasing; required unless masks are not needed.
    `period`/`frame_offset` let a rank build frames [frame_offset, frame_offset+B) of a longer sequence of
    `period` frames; `traj_period` (default: the sequence length) is the number of frames after which the
    ground-truth motion repeats.
    Returns a dict with the arguments of joint_optimize plus the ground truth."""
    if isinstance(mesh, str):
        if mesh == "uv50x100":
            verts, faces = uv_sphere_mesh(50, 100, seed)
        elif mesh == "uv100x200":
            verts, faces = uv_sphere_mesh(100, 200, seed)
        elif mesh == "ico2":
            verts, faces = icosphere_mesh(2, seed)
        elif mesh == "ico3":
            verts, faces = icosphere_mesh(3, seed)
        else:
            verts, faces = load_obj(mesh)
    else:
        verts, faces = mesh
    n = B + frame_offset if period is None else period
    R_all, T_all = gt_trajectory(n, period=traj_period if traj_period else n)
    Rn_all, Tn_all = perturb_poses(R_all, T_all, seed)
    sl = slice(frame_offset, frame_offset + B)
    R_gt, T_gt, R0, T0 = R_all[sl], T_all[sl], Rn_all[sl], Tn_all[sl]
    K_roi, boxes = make_K_roi(verts, R_gt, T_gt, H, W, size)
    out = {
        "verts": verts, "faces": faces, "K_roi": K_roi, "boxes": boxes,
        "R_gt": R_gt.astype(np.float32), "T_gt": T_gt.astype(np.float32).reshape(B, 1, 3),
        "R_init": R0.astype(np.float32), "T_init": T0.astype(np.float32).reshape(B, 1, 3),
        "rot6d_init": np.ascontiguousarray(R0[:, :, :2]).astype(np.float32),  # geometry.py:38
    }
    if render_fn is not None:
        vc = (verts.astype(np.float64)[None] @ R_gt + T_gt[:, None, :]).astype(np.float32)
        sil = np.asarray(render_fn(vc, faces, K_roi, size))
        m01 = (sil > 0.5).astype(np.float32)
        out["target_masks"] = add_disc_occluder(m01, seed, 40.0, frame_offset) if occluder else m01
    return out


def make_correspondences(seq, C, seed=0, noise_px=0.5, outliers=0.05, size=REND_SIZE, frame_offset=0, device=None):
    """[BUILDER-DEFINED, SURVEY.md 8d]  Synthetic DKM-style dense correspondences for the reprojection term
    (dynhor_b200/corr.py): per frame b, C records (X[3], t[2], w) where X is a surface point of the canonical
    mesh facing the camera in the frame's ground-truth pose, t its ground-truth projection into frame b in ROI
    unit-image coordinates plus N(0, noise_px^2) pixel noise, and w a DKM-like certainty in (0.5, 1]; a fraction
    `outliers` of the targets is replaced by uniform positions.  Every frame draws from its own generator seeded by
    (seed, frame_offset + b): the records of a frame do not depend on which rank builds it.
    Returns float32 [B,C,6] (numpy; a torch tensor on `device` when one is given -- the 50k-record configurations
    are generated on the GPU)."""
    import torch
    dev = torch.device(device) if device is not None else torch.device("cpu")
    verts = torch.from_numpy(seq["verts"].astype(np.float64)).to(dev)
    faces = torch.from_numpy(np.asarray(seq["faces"], np.int64)).to(dev)
    R = torch.from_numpy(seq["R_gt"].astype(np.float64)).to(dev)
    T = torch.from_numpy(seq["T_gt"].astype(np.float64).reshape(-1, 3)).to(dev)
    K = torch.from_numpy(seq["K_roi"].astype(np.float64)).to(dev)
    B = len(R)
    tri = verts[faces]                                           # [F,3,3]
    nrm = torch.linalg.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    cen0 = tri.mean(1)
    out = torch.empty((B, C, 6), dtype=torch.float32, device=dev)
    for b in range(B):
        g = torch.Generator(device=dev).manual_seed((seed + 3) * 1000003 + frame_offset + b)
        nc = nrm @ R[b]                                          # normals in the frame's camera space
        cen = cen0 @ R[b] + T[b]
        vis = torch.nonzero((nc * cen).sum(-1) < 0)[:, 0]        # faces turned towards the camera
        f = vis[torch.randint(0, len(vis), (C,), generator=g, device=dev)]
        uv = torch.rand((C, 2), generator=g, device=dev, dtype=torch.float64)
        flip = uv.sum(1) > 1.0
        uv = torch.where(flip[:, None], 1.0 - uv, uv)
        X = tri[f, 0] + uv[:, :1] * (tri[f, 1] - tri[f, 0]) + uv[:, 1:] * (tri[f, 2] - tri[f, 0])
        c = X @ R[b] + T[b]
        x_, y_ = c[:, 0] / c[:, 2], c[:, 1] / c[:, 2]
        Kb = K[b]
        t = torch.stack([Kb[0, 0] * x_ + Kb[0, 1] * y_ + Kb[0, 2], Kb[1, 0] * x_ + Kb[1, 1] * y_ + Kb[1, 2]], -1)
        t = t + torch.randn((C, 2), generator=g, device=dev, dtype=torch.float64) * (noise_px / float(size))
        bad = torch.rand((C,), generator=g, device=dev) < outliers
        t = torch.where(bad[:, None], torch.rand((C, 2), generator=g, device=dev, dtype=torch.float64), t)
        w = 0.5 + 0.5 * torch.rand((C, 1), generator=g, device=dev, dtype=torch.float64)
        out[b] = torch.cat([X, t, w], -1).float()
    return out if device is not None else out.numpy()


def to_object_parameters(seq):
    """List of per-frame dicts in the layout find_optimal_poses returns (pose_initializtion.py:460-471),
    as torch CPU tensors; the caller moves them to the device like stage 1 does."""
    import torch
    params = []
    B = len(seq["R_init"])
    for b in range(B):
        params.append({
            "rotations": torch.from_numpy(seq["R_init"][b:b + 1].copy()),              # [1,3,3]
            "translations": torch.from_numpy(seq["T_init"][b:b + 1].copy()),           # [1,1,3]
            "K_roi": torch.from_numpy(seq["K_roi"][b:b + 1].copy()).unsqueeze(0),      # [1,1,3,3]
            "target_masks": torch.from_numpy(seq["target_masks"][b:b + 1].copy()),     # [1,S,S]
            "verts": torch.from_numpy(seq["verts"]).unsqueeze(0),
        })
        if "correspondences" in seq:
            params[-1]["correspondences"] = torch.from_numpy(seq["correspondences"][b:b + 1].copy())  # [1,C,6]
    return params


# ----------------------------------------------------------------------------- DINO features
def make_dino_features(N, Fm, P, D, seed=0, noise=0.5, mask_p=0.4, k=10, min_gap=1e-3, device="cpu"):
    """Synthetic patch features shaped like DINOv2 tokens: templates i.i.d. N(0,1), L2-normalised per patch and
    rounded to bf16 precision; each frame = normalise(template[pi(f)] + noise * N(0,1)) so a unique best match
    exists; Bernoulli(mask_p) foreground masks with at least one patch.  Returned as torch fp32 tensors holding
    bf16-representable values (both the oracle and the kernels see the same numbers).  `min_gap` is checked by
    the callers that own an oracle (top-k is only well-posed when consecutive scores differ by more than the
    accumulation noise)."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    perm = torch.randperm(N, generator=g)[:Fm] if Fm <= N else torch.randint(0, N, (Fm,), generator=g)
    gen = torch.Generator(device=device).manual_seed(seed + 1)
    templ = torch.randn(N, P, D, generator=gen, device=device)
    templ = torch.nn.functional.normalize(templ, dim=-1).bfloat16().float()
    frames = templ[perm.to(device)] + noise * torch.nn.functional.normalize(
        torch.randn(Fm, P, D, generator=gen, device=device), dim=-1)
    frames = torch.nn.functional.normalize(frames, dim=-1).bfloat16().float()
    masks = (torch.rand(Fm, P, generator=gen, device=device) < mask_p).float()
    masks[:, 0] = torch.where(masks.sum(1) == 0, torch.ones_like(masks[:, 0]), masks[:, 0])
    return {"templ": templ, "frames": frames, "masks": masks, "match": perm}
