"""CPU: the C-ABI library loads and exports every symbol include/dynhor_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from helpers import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "dynhor_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dh_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from dynhor_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert lib.dh_version() >= 100


def test_struct_layouts_match_header_sizes():
    from dynhor_b200 import _lib
    # dh_sil: 5 int32 + 4 float + 10 pointers; dh_jointopt embeds it
    assert ctypes.sizeof(_lib.DhSil) == 40 + 15 * 8
    assert ctypes.sizeof(_lib.DhJointOpt) % 8 == 0
    lib = _lib.load()
    for which, struct in enumerate((_lib.DhSil, _lib.DhJointOpt, _lib.DhCorr)):
        assert lib.dh_struct_bytes(which) == ctypes.sizeof(struct)
    out = (ctypes.c_int64 * 13)()
    assert lib.dh_sil_scratch_bytes(2, 10, 20, 64, 1, out) == 0
    assert out[3] == 2 * 128 * 128 * 4 and out[4] == 2 * 128 * 4 * 4
    assert lib.dh_sil_scratch_bytes(0, 10, 20, 64, 1, out) != 0
    assert b"bad arguments" in lib.dh_last_error()


def test_no_cpu_fallback():
    """Product entry points refuse to run without CUDA instead of falling back."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU box")
    from dynhor_b200 import _lib
    from dynhor_b200.jointopt import joint_optimize
    with pytest.raises(_lib.DynhorError):
        joint_optimize([], objvertices=None, objfaces=None)


def test_backward_batch_schedule_covers_every_item_once():
    """dh_bwd_schedule = the kernel's own guided batch schedule (dh_core.h bwd_guided_schedule), a function of the item
    count only: consecutive batches, every item in exactly one of them, never more than 32 items (one per lane),
    at most 95 batches (one row of partial sums each), non-increasing sizes apart from the tail that absorbs crumbs."""
    from dynhor_b200 import _lib
    lib = _lib.load()
    out = (ctypes.c_int32 * 128)()
    for n in list(range(0, 700)) + [1000, 1023, 1024, 1500, 2047, 2048]:
        nb = lib.dh_bwd_schedule(n, out, 128)
        assert 0 <= nb <= 95, (n, nb)
        starts = list(out[:nb + 1])
        assert starts[0] == 0 if nb else True
        assert starts[nb] == n
        sizes = [b - a for a, b in zip(starts, starts[1:])]
        assert all(1 <= s <= 32 for s in sizes), (n, sizes)
        assert sum(sizes) == n
        assert all(a >= b for a, b in zip(sizes[:-2], sizes[1:-1])), (n, sizes)   # shrinking towards the end
    assert lib.dh_bwd_schedule(-1, out, 128) < 0 and lib.dh_bwd_schedule(10, out, 8) < 0
