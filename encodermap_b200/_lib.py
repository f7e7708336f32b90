"""ctypes binding of libemk.so (include/emk.h) and the DLPack hand-off.

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import ctypes as C
from pathlib import Path
from typing import Optional

import torch
from torch.utils.dlpack import to_dlpack

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libemk.so"

EMK_COST_ZERO_OUTPUTS = 1
EMK_COST_NO_GRAD = 2
TILE_ROWS, TILE_COLS = 128, 64
NONE_INDEX = -(2**63)  # INT64_MIN: "None" in a python slice passed over the ABI


class EmkError(RuntimeError):
    """A libemk entry point returned non-zero (negative: argument error, positive: cudaError_t)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libemk error {code}: {message}")
        self.code = code


_lib = None

c_f32p = C.POINTER(C.c_float)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
vp = C.c_void_p
i64 = C.c_int64
dbl = C.c_double
f32 = C.c_float

# every exported symbol of include/emk.h with its argument types (restype int unless noted)
SIGNATURES = {
    "emk_version": ([], C.c_int),
    "emk_last_error": ([], C.c_char_p),
    "emk_build_info": ([], C.c_char_p),
    "emk_probe_fp32": ([C.POINTER(C.c_double)], C.c_int),
    "emk_set_option": ([C.c_char_p, i64], C.c_int),
    "emk_get_option": ([C.c_char_p, c_i64p], C.c_int),
    "emk_triu_pair_count": ([i64], i64),
    "emk_triu_pair_indices": ([i64, c_i32p, c_i32p], C.c_int),
    "emk_backmap_split_counts": ([i64, c_i64p], C.c_int),
    "emk_backmap_split_indices": ([i64, c_i32p, c_i32p, c_i32p, c_i32p], C.c_int),
    "emk_pair_tile_count": ([i64], i64),
    "emk_pair_tile_decode": ([i64, i64, c_i64p, c_i64p], C.c_int),
    "emk_pair_tile_range": ([i64, C.c_int, C.c_int, c_i64p, c_i64p], C.c_int),
    "emk_sigmoid_cost": ([vp, i64, i64, vp, i64, dbl, c_f32p, i64, i64, vp, vp, C.c_uint32, vp], C.c_int),
    "emk_dl_sigmoid_cost": ([vp, vp, dbl, c_f32p, i64, i64, vp, vp, C.c_uint32, vp], C.c_int),
    "emk_sigmoid_cost_host": ([vp, i64, i64, vp, i64, dbl, c_f32p, C.POINTER(dbl), vp], C.c_int),
    "emk_pairwise_dist_periodic": ([vp, i64, i64, dbl, vp, vp], C.c_int),
    "emk_dl_pairwise_dist_periodic": ([vp, dbl, vp, vp], C.c_int),
    "emk_pairwise_dist_periodic_bwd": ([vp, i64, i64, dbl, vp, vp, vp, vp], C.c_int),
    "emk_dl_pairwise_dist_periodic_bwd": ([vp, dbl, vp, vp, vp, vp], C.c_int),
    "emk_pairwise_dist": ([vp, i64, i64, i64, i64, i64, C.c_int, C.c_int, vp, vp], C.c_int),
    "emk_pairwise_dist_bwd": ([vp, i64, i64, i64, i64, i64, C.c_int, C.c_int, vp, vp, vp], C.c_int),
    "emk_dl_pairwise_dist": ([vp, i64, i64, i64, C.c_int, C.c_int, vp, vp], C.c_int),
    "emk_dl_pairwise_dist_bwd": ([vp, i64, i64, i64, C.c_int, C.c_int, vp, vp, vp], C.c_int),
    "emk_cartesian_pair_loss": ([vp, i64, i64, i64, i64, i64, vp, C.c_int, C.c_int, f32, vp, vp, vp, vp], C.c_int),
    "emk_dl_cartesian_pair_loss": ([vp, i64, i64, i64, vp, C.c_int, f32, vp, vp, vp, vp], C.c_int),
    "emk_cartesian_distance_cost": ([vp, i64, i64, i64, i64, i64, vp, i64, c_f32p, i64, i64, vp, vp, C.c_uint32, vp], C.c_int),
    "emk_dl_cartesian_distance_cost": ([vp, i64, i64, i64, vp, c_f32p, i64, i64, vp, vp, C.c_uint32, vp], C.c_int),
    "emk_periodic_distance": ([vp, vp, i64, dbl, vp, vp], C.c_int),
    "emk_periodic_distance_bwd": ([vp, vp, i64, dbl, vp, vp, vp, vp], C.c_int),
    "emk_dl_periodic_distance": ([vp, vp, dbl, vp, vp], C.c_int),
    "emk_dl_periodic_distance_bwd": ([vp, vp, dbl, vp, vp, vp, vp], C.c_int),
    "emk_sigmoid": ([vp, i64, f32, f32, f32, vp, vp], C.c_int),
    "emk_sigmoid_bwd": ([vp, i64, f32, f32, f32, vp, vp, vp], C.c_int),
    "emk_dl_sigmoid": ([vp, f32, f32, f32, vp, vp], C.c_int),
    "emk_dl_sigmoid_bwd": ([vp, f32, f32, f32, vp, vp, vp], C.c_int),
    "emk_periodic_input": ([vp, i64, i64, dbl, vp, vp], C.c_int),
    "emk_periodic_input_bwd": ([vp, i64, i64, dbl, vp, vp, vp], C.c_int),
    "emk_dl_periodic_input": ([vp, dbl, vp, vp], C.c_int),
    "emk_dl_periodic_input_bwd": ([vp, dbl, vp, vp, vp], C.c_int),
    "emk_rotation_matrix": ([vp, vp, i64, vp, vp], C.c_int),
    "emk_dl_rotation_matrix": ([vp, vp, vp, vp], C.c_int),
    "emk_guess_sp2_atoms": ([vp, i64, i64, c_i64p, i64, dbl, dbl, vp, vp], C.c_int),
    "emk_merge_cartesians": ([vp, i64, i64, c_i64p, i64, c_i64p, i64, vp, i64, vp, i64, vp, vp], C.c_int),
    "emk_backbone_amide_atoms": ([vp, i64, i64, c_i64p, i64, c_i64p, i64, dbl, dbl, dbl, dbl, vp, i64, vp], C.c_int),
    "emk_merged_atom_count": ([i64, c_i64p, i64, c_i64p, i64], i64),
    "emk_set_dihedrals": ([vp, i64, i64, c_i32p, c_i32p, c_i32p, c_i32p, i64, vp, i64, vp, vp], C.c_int),
    "emk_sidechain_plan_create": ([i64, c_i32p, C.POINTER(vp)], C.c_int),
    "emk_sidechain_plan_destroy": ([vp], None),
    "emk_sidechain_plan_info": ([vp, c_i64p], C.c_int),
    "emk_sidechain_plan_ops": ([vp, c_i32p], C.c_int),
    "emk_sidechain_backmap": ([vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp], C.c_int),
    "emk_sidechain_backmap_bwd": ([vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp], C.c_int),
    "emk_dl_sidechain_backmap": ([vp, C.POINTER(vp), vp, vp, vp], C.c_int),
    "emk_dl_sidechain_backmap_bwd": ([vp, C.POINTER(vp), vp, vp, C.POINTER(vp), vp], C.c_int),
    "emk_sidechain_saved_size": ([vp], i64),
    "emk_sidechain_pairwise_indices": ([i64, c_i32p, i64, i64, i64, c_i64p], i64),
    "emk_gather_atoms": ([vp, i64, i64, vp, i64, vp, vp], C.c_int),
    "emk_gather_atoms_bwd": ([vp, i64, i64, vp, i64, vp, vp], C.c_int),
    "emk_dl_gather_atoms": ([vp, vp, vp, vp], C.c_int),
    "emk_dl_gather_atoms_bwd": ([vp, vp, vp, vp], C.c_int),
    "emk_column_mean": ([vp, i64, i64, vp, vp], C.c_int),
    "emk_dl_column_mean": ([vp, vp, vp], C.c_int),
    "emk_backmap": ([vp, i64, vp, vp, i64, i64, vp, vp], C.c_int),
    "emk_backmap_bwd": ([vp, i64, vp, vp, vp, i64, i64, vp, vp, vp, vp], C.c_int),
    "emk_dl_backmap": ([vp, vp, vp, vp, vp], C.c_int),
    "emk_dl_backmap_bwd": ([vp, vp, vp, vp, vp, vp, vp, vp], C.c_int),
    "emk_backmap_host": ([vp, vp, vp, i64, i64, vp], C.c_int),
    "emk_chain_in_plane": ([vp, i64, vp, i64, i64, vp, vp], C.c_int),
    "emk_chain_in_plane_bwd": ([vp, i64, vp, vp, i64, i64, vp, vp, vp], C.c_int),
    "emk_dl_chain_in_plane": ([vp, vp, vp, vp], C.c_int),
    "emk_dl_chain_in_plane_bwd": ([vp, vp, vp, vp, vp, vp], C.c_int),
    "emk_dihedrals_to_cartesian": ([vp, vp, i64, i64, i64, C.c_int, vp, vp], C.c_int),
    "emk_dihedrals_to_cartesian_bwd": ([vp, vp, i64, i64, C.c_int, vp, vp], C.c_int),
    "emk_dl_dihedrals_to_cartesian": ([vp, vp, C.c_int, vp, vp], C.c_int),
    "emk_dl_dihedrals_to_cartesian_bwd": ([vp, vp, C.c_int, vp, vp], C.c_int),
    "emk_dihedrals_to_cartesian_chain_bwd": ([vp, i64, vp, vp, i64, i64, C.c_int, vp, vp], C.c_int),
    "emk_dl_dihedrals_to_cartesian_chain_bwd": ([vp, vp, vp, C.c_int, vp, vp], C.c_int),
    "emk_comm_unique_id": ([vp], C.c_int),
    "emk_comm_init": ([C.c_int, C.c_int, vp], C.c_int),
    "emk_comm_info": ([C.POINTER(C.c_int), C.POINTER(C.c_int)], C.c_int),
    "emk_comm_allreduce": ([vp, vp, i64, vp], C.c_int),
    "emk_comm_allgather2": ([vp, vp, i64, vp, vp, i64, vp], C.c_int),
    "emk_comm_reduce_cost_scatter": ([vp, vp, vp, i64, vp], C.c_int),
    "emk_comm_destroy": ([], C.c_int),
}


def lib() -> ctypes.CDLL:
    """Load libemk.so (once).  Raises if it has not been built: there is no other code path."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m encodermap_b200._build` "
                "(needs nvcc; sm_100a).  encodermap_b200 has no CPU or framework fallback."
            )
        handle = ctypes.CDLL(str(LIB_PATH))
        for name, (argtypes, restype) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise EmkError(rc, lib().emk_last_error().decode(errors="replace"))


# ---- DLPack hand-off ---------------------------------------------------------------------------------
_pyapi = ctypes.pythonapi
_pyapi.PyCapsule_GetPointer.restype = ctypes.c_void_p
_pyapi.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]


class DL:
    """Zero-copy view of a torch tensor as `DLManagedTensor*`.

    The capsule is exported with `to_dlpack` and is NOT consumed (not renamed to
    "used_dltensor"), so its destructor releases the export when this object dies; libemk only
    borrows the pointer for the duration of one call."""

    __slots__ = ("capsule", "ptr", "tensor")

    def __init__(self, t: Optional[torch.Tensor]):
        self.tensor = t
        if t is None:
            self.capsule, self.ptr = None, None
        else:
            self.capsule = to_dlpack(t.detach())
            self.ptr = _pyapi.PyCapsule_GetPointer(self.capsule, b"dltensor")

    @property
    def _as_parameter_(self):
        return ctypes.c_void_p(self.ptr)


def stream_of(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def require_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise EmkError(-3, f"{name} is on {t.device}; encodermap_b200 runs on CUDA only (there is no CPU fallback)")
    return t


def f32c(t: torch.Tensor) -> torch.Tensor:
    """float32 + C-contiguous (the reference casts its inputs to float32, autoencoder.py:802)."""
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


def sig_array(sig) -> "ctypes.Array":
    vals = [float(v) for v in sig]
    if len(vals) != 6:
        raise ValueError(f"dist_sig_parameters must have 6 entries, got {len(vals)}")
    return (ctypes.c_float * 6)(*vals)


def set_option(name: str, value: int) -> None:
    check(lib().emk_set_option(name.encode(), int(value)))


def get_option(name: str) -> int:
    v = ctypes.c_int64()
    check(lib().emk_get_option(name.encode(), ctypes.byref(v)))
    return int(v.value)


# ---- host-only helpers (no GPU needed) ---------------------------------------------------------------------
def pair_tile_count(n: int) -> int:
    return int(lib().emk_pair_tile_count(n))


def pair_tile_range(n: int, rank: int, world: int):
    b, e = ctypes.c_int64(), ctypes.c_int64()
    check(lib().emk_pair_tile_range(n, rank, world, ctypes.byref(b), ctypes.byref(e)))
    return int(b.value), int(e.value)


def pair_tile_decode(n: int, tile: int):
    r, c = ctypes.c_int64(), ctypes.c_int64()
    check(lib().emk_pair_tile_decode(n, tile, ctypes.byref(r), ctypes.byref(c)))
    return int(r.value), int(c.value)


def triu_pair_indices(n: int):
    import numpy as np

    cnt = int(lib().emk_triu_pair_count(n))
    i = np.empty(cnt, dtype=np.int32)
    j = np.empty(cnt, dtype=np.int32)
    check(lib().emk_triu_pair_indices(n, i.ctypes.data_as(c_i32p), j.ctypes.data_as(c_i32p)))
    return i, j


def backmap_split_indices(n_atoms: int):
    import numpy as np

    counts = (ctypes.c_int64 * 4)()
    check(lib().emk_backmap_split_counts(n_atoms, counts))
    arrs = [np.empty(int(c), dtype=np.int32) for c in counts]
    check(lib().emk_backmap_split_indices(n_atoms, *[a.ctypes.data_as(c_i32p) for a in arrs]))
    return tuple(arrs)
