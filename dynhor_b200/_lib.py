"""ctypes binding of libdynhor_b200.so (C ABI declared in include/dynhor_b200.h).

This is the stub a maintainer of the reference would add: plain pointers and sizes, no torch
types cross the boundary.  There is NO CPU fallback: if the shared library is missing, cannot be loaded, or a
call fails, the error is raised to the caller.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdynhor_b200.so")

c_f = ctypes.c_float
c_d = ctypes.c_double
c_i = ctypes.c_int32
c_l = ctypes.c_int64
c_p = ctypes.c_void_p


class DhSil(ctypes.Structure):
    """struct dh_sil (include/dynhor_b200.h)."""
    _fields_ = [
        ("B", c_i), ("V", c_i), ("F", c_i), ("S", c_i), ("aa", c_i),
        ("near_", c_f), ("far_", c_f), ("eps", c_f), ("orig_size", c_f),
        ("faces", c_p), ("K", c_p),
        ("proj", c_p), ("bin_count", c_p), ("bins", c_p), ("fidx", c_p), ("alpha_bits", c_p),
        ("pos_pool", c_p), ("neg_pool", c_p), ("gpool", c_p), ("gmax", c_p), ("owned", c_p), ("negT", c_p), ("row_rng", c_p),
        ("neg_lists", c_p),
    ]


class DhCorr(ctypes.Structure):
    """struct dh_corr (include/dynhor_b200.h) -- builder-defined correspondence term."""
    _fields_ = [
        ("records", c_p), ("C", c_i), ("nslots", c_i), ("delta", c_f), ("pad_", c_f),
        ("w_sum", c_d), ("lw_corr", c_d), ("partials", c_p), ("w_sum_dev", c_p),
    ]


class DhJointOpt(ctypes.Structure):
    """struct dh_jointopt (include/dynhor_b200.h)."""
    _fields_ = [
        ("sil", DhSil),
        ("verts_og", c_p), ("mask_tri", c_p),
        ("rot6d", c_p), ("trans", c_p), ("scale", c_p),
        ("adam_m_rot", c_p), ("adam_v_rot", c_p), ("adam_m_trans", c_p), ("adam_v_trans", c_p),
        ("adam_mv_scale", c_p),
        ("step", c_p), ("hist", c_p), ("max_iters", c_i),
        ("halo_prev", c_p), ("halo_next", c_p),
        ("mailbox", c_p), ("peer_prev", c_p), ("peer_next", c_p),
        ("tick_base", c_i), ("halo_timeout_ms", c_i), ("status", c_p),
        ("rank", c_i), ("world", c_i), ("scale_mode", c_i),
        ("peers", c_p * 16), ("scale_part", c_p),
        ("B_total", c_i),
        ("keep_sum", c_d), ("lw_sil", c_d), ("lw_smooth", c_d), ("lr", c_d),
        ("optimize_scale", c_i),
        ("moments", c_p),
        ("Rmat", c_p), ("smooth_terms", c_p), ("loss_counts", c_p), ("partials", c_p), ("frame_terms", c_p),
        ("nchunks", c_i),
        ("corr", DhCorr),
        ("loss_mode", c_i), ("lw_offscreen", c_d), ("offscreen", c_p), ("frame_coef", c_p), ("iter_ns", c_p),
    ]


# name -> (restype, argtypes); every symbol include/dynhor_b200.h declares
SIGNATURES = {
    "dh_version": (c_i, []),
    "dh_last_error": (ctypes.c_char_p, []),
    "dh_struct_bytes": (c_i, [c_i]),
    "dh_device_info": (c_i, [ctypes.POINTER(c_i)] * 3),
    "dh_tune_set": (c_i, [c_i, c_i]),
    "dh_sil_scratch_bytes": (c_i, [c_i, c_i, c_i, c_i, c_i, ctypes.POINTER(c_l)]),
    "dh_sil_forward": (c_i, [ctypes.POINTER(DhSil), c_p, c_p, c_p]),
    "dh_sil_backward": (c_i, [ctypes.POINTER(DhSil), c_p, c_p, c_p, c_p]),
    "dh_rot6d_to_matrix": (c_i, [c_p, c_p, c_i, c_p]),
    "dh_transform_verts": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_p]),
    "dh_masks_prepare": (c_i, [c_p, c_p, c_p, c_l, c_p]),
    "dh_mesh_moments": (c_i, [c_p, c_i, c_p, c_p]),
    "dh_corr_plan": (c_i, [c_i, c_i, c_i, ctypes.POINTER(c_i)]),
    "dh_corr_eval": (c_i, [c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_i, c_f, c_p, c_i, c_p]),
    "dh_jointopt_scratch_bytes": (c_i, [c_i, c_i, ctypes.POINTER(c_l)]),
    "dh_jointopt_default_chunks": (c_i, [c_i, c_i]),
    "dh_jointopt_run": (c_i, [ctypes.POINTER(DhJointOpt), c_i, c_i, c_p]),
    "dh_jointopt_eval": (c_i, [ctypes.POINTER(DhJointOpt), c_p]),
    "dh_jointopt_grads": (c_i, [ctypes.POINTER(DhJointOpt), c_p, c_p, c_p, c_p]),
    "dh_jointopt_profile": (c_i, [ctypes.POINTER(DhJointOpt), c_i, ctypes.POINTER(c_f), c_p]),
    "dh_jointopt_run_part": (c_i, [ctypes.POINTER(DhJointOpt), c_i, c_p]),
    "dh_bwd_schedule": (c_i, [c_i, ctypes.POINTER(ctypes.c_int32), c_i]),
    "dh_jointopt_release": (c_i, [ctypes.POINTER(DhJointOpt)]),
    "dh_jointopt_probe": (c_i, [ctypes.POINTER(DhJointOpt), c_i, ctypes.POINTER(c_f), c_p]),
    "dh_scale_apply": (c_i, [ctypes.POINTER(DhJointOpt), c_p, c_i, c_p]),
    "dh_dev_alloc": (c_i, [ctypes.POINTER(c_p), c_l]),
    "dh_dev_free": (c_i, [c_p]),
    "dh_memcpy_d2d": (c_i, [c_p, c_p, c_l, c_p]),
    "dh_upload_rows": (c_i, [c_p, ctypes.POINTER(c_p), c_l, c_i, c_p]),
    "dh_ipc_export": (c_i, [c_p, c_p]),
    "dh_ipc_open": (c_i, [c_p, ctypes.POINTER(c_p)]),
    "dh_ipc_close": (c_i, [c_p]),
    "dh_adam_step": (c_i, [c_p, c_p, c_p, c_p, c_l, c_d, c_i, c_p]),
    "dh_dino_workspace_bytes": (c_i, [c_i, c_i, c_l, ctypes.POINTER(c_l)]),
    "dh_dino_plan_info": (c_i, [c_i, c_i, c_l, ctypes.POINTER(c_i)]),
    "dh_dino_topk": (c_i, [c_p, c_p, c_i, c_i, c_l, c_i, c_p, c_p, c_p, c_p, c_l, c_p]),
    "dh_dino_prescale": (c_i, [c_p, c_p, c_i, c_i, c_i, c_p, c_p]),
    "dh_roi_process": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_f, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "dh_roi_process_f32": (c_i, [c_p, c_p, c_i, c_p, c_i, c_i, c_i, c_i, c_f, c_f, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
}

MAILBOX_WORDS = 512      # DH_MAILBOX_WORDS
MAX_RANKS = 16           # DH_MAX_RANKS
SCALE_LOCAL, SCALE_P2P, SCALE_DEFERRED = 0, 1, 2
STATUS_HALO_TIMEOUT = 1
LOSS_JOINT, LOSS_STAGE1 = 0, 1

_LIB = None


class DynhorError(RuntimeError):
    pass


def load():
    """Load the shared library (built in-tree by __graft_entry__.build()); raises if it is not there."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise DynhorError(
                f"{LIB_PATH} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; "
                "g.build()'). dynhor_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header and library out of sync
            fn.restype = res
            fn.argtypes = args
        for which, struct in enumerate((DhSil, DhJointOpt, DhCorr)):
            if lib.dh_struct_bytes(which) != ctypes.sizeof(struct):
                raise DynhorError(f"{struct.__name__}: ctypes layout ({ctypes.sizeof(struct)} B) does not match "
                                  f"the library's ({lib.dh_struct_bytes(which)} B)")
        _LIB = lib
    return _LIB


def check(rc, what=""):
    if rc != 0:
        msg = load().dh_last_error()
        raise DynhorError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
