"""ctypes binding of ``csrc/libb200grbm.so`` (C ABI declared in ``include/b200grbm.h``).

There is deliberately no fallback: if the shared library has not been built
(``__graft_entry__.build()`` / ``csrc/build.sh``) loading raises, and every compute entry
point returns an error on a machine without an sm_100 device.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libb200grbm.so")

ACCEPT_EXACT = 0
ACCEPT_FAST = 1
ABI_VERSION = 5


class B200Error(RuntimeError):
    """Non-zero return from libb200grbm (message from ``b200grbm_last_error``)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libb200grbm error {code}: {message}")
        self.code = code


class SweepArgs(C.Structure):
    """Mirror of ``b200grbm_sweep_args`` (include/b200grbm.h)."""

    _fields_ = [
        ("struct_size", C.c_uint32),
        ("n", C.c_int32),
        ("n_pad", C.c_int32),
        ("ell_width", C.c_int32),
        ("n_tiles", C.c_int32),
        ("tiles_dev", C.c_void_p),
        ("tile_info_dev", C.c_void_p),
        ("order_dev", C.c_void_p),
        ("chains", C.c_int32),
        ("chains_per_lane", C.c_int32),
        ("threads", C.c_int32),
        ("accept", C.c_int32),
        ("chain_offset", C.c_uint64),
        ("seed", C.c_uint64),
        ("sweep_offset", C.c_uint32),
        ("num_sweeps", C.c_int32),
        ("coef_dev", C.c_void_p),
        ("uniforms_dev", C.c_void_p),
        ("state_in_dev", C.c_void_p),
        ("packed_in_dev", C.c_void_p),
        ("state_out_dev", C.c_void_p),
        ("packed_out_dev", C.c_void_p),
    ]


_i32, _f32, _vp = C.c_int32, C.c_float, C.c_void_p

# name -> argtypes; every symbol include/b200grbm.h declares (tests/test_abi.py checks the
# header against this table and against the built library)
SIGNATURES = {
    "b200grbm_last_error": ([], C.c_char_p),
    "b200grbm_abi_version": ([], _i32),
    "b200grbm_device_info": ([C.POINTER(_i32)] * 4, _i32),
    "b200grbm_set_weights": ([_vp, _vp, _i32, _i32, _f32, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp, _i32, _i32,
                              _vp, _vp, _vp, _vp], _i32),
    "b200grbm_sweep_smem_bytes": ([_i32, _i32, _i32, _i32], C.c_int64),
    "b200grbm_sweep_state_offset": ([_i32], _i32),
    "b200grbm_gibbs_sweeps": ([C.POINTER(SweepArgs), _vp], _i32),
    "b200grbm_last_launch_count": ([], _i32),
    "b200grbm_last_sweep_kernel": ([], _i32),
    "b200grbm_ex2_probe": ([_f32, _f32, C.c_int64, C.POINTER(C.c_double), _vp], _i32),
    "b200grbm_pack_f32": ([_vp, _i32, _i32, _i32, _vp, _i32, _vp, _vp], _i32),
    "b200grbm_pack_i8": ([_vp, _i32, _i32, _i32, _vp, _i32, _vp, _vp], _i32),
    "b200grbm_edge_stats": ([_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp], _i32),
    "b200grbm_energy_forward": ([_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp], _i32),
    "b200grbm_energy_backward": ([_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp], _i32),
    "b200grbm_energy_grad_x": ([_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp], _i32),
    "b200grbm_energy_i8": ([_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp], _i32),
    "b200grbm_energy_packed": ([_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp], _i32),
    "b200grbm_mmd_forward_f32": ([_vp, _i32, _i32, _i32, _i32, _f32, _i32, _f32, _vp, _vp], _i32),
    "b200grbm_mmd_pack_i8": ([_vp, _i32, _i32, _i32, _vp, _vp], _i32),
    "b200grbm_mmd_hist_i8": ([_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp], _i32),
    "b200grbm_mmd_hist_fp4": ([_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp], _i32),
    "b200grbm_pack_fp4_i8": ([_vp, _i32, _i32, _vp, _i32, _vp], _i32),
    "b200grbm_mmd_eval_hist": ([_vp, _i32, _i32, _i32, _i32, _f32, _i32, _f32, _i32, C.c_double, _vp, _vp], _i32),
    "b200grbm_mmd_forward_i8": ([_vp, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _f32, _i32, C.c_double, _vp, _vp, _vp], _i32),
    "b200grbm_spin_extract_f32": ([_vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp, _vp, _i32, _vp, _f32, _vp, _i32, _vp], _i32),
    "b200grbm_spin_extract_i8": ([_vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp, _vp, _i32, _vp, _f32, _vp, _i32, _vp], _i32),
    "b200grbm_transpose_i8": ([_vp, _i32, _i32, _i32, _vp, _i32, _vp], _i32),
    "b200grbm_mmd_coef_i8": ([_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _f32, _vp, _vp, _f32, _f32, _vp, _vp,
                              _i32, _i32, _i32, _vp, _vp, _vp], _i32),
    "b200grbm_mmd_grad_i8": ([_vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp], _i32),
    "b200grbm_gemm_bf16_tn": ([_vp, _vp, _i32, _i32, _i32, _vp, _i32, _vp, _i32, _vp], _i32),
    "b200grbm_mmd_forward_bf16": ([_vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _i32, _f32, _vp, _vp], _i32),
    "b200grbm_mmd_coef_bf16": ([_vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _i32, _f32, _vp, _f32, _f32, _vp, _vp, _i32,
                                _vp], _i32),
    "b200grbm_spin_pack_bits_f32": ([_vp, _i32, _i32, _vp, _i32, _i32, _vp], _i32),
    "b200grbm_spin_pack_bits_i8": ([_vp, _i32, _i32, _vp, _i32, _i32, _vp], _i32),
    "b200grbm_peer_signal": ([_vp, C.c_uint32, _vp], _i32),
    "b200grbm_bits_to_rows": ([C.POINTER(_vp), C.POINTER(_vp), _i32, _i32, _i32, _i32, _i32, _vp, C.c_uint32, _vp], _i32),
    "b200grbm_peer_alloc": ([C.c_int64, C.POINTER(_vp), _vp], _i32),
    "b200grbm_peer_open": ([_vp, C.POINTER(_vp)], _i32),
    "b200grbm_peer_close": ([_vp], _i32),
    "b200grbm_peer_free": ([_vp], _i32),
    "b200grbm_tensor_peak": ([_i32, _i32, C.POINTER(C.c_double), _vp], _i32),
    "b200grbm_mmd_backward_f32": ([_vp, _i32, _i32, _i32, _i32, _f32, _i32, _f32, _vp, _f32, _f32, _vp, _vp, _vp,
                                   _vp], _i32),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the library once; raise loudly if it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or image-generation_b200/csrc/build.sh.  There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (argtypes, restype) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.argtypes = argtypes
        fn.restype = restype
    got = lib.b200grbm_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f"libb200grbm ABI {got} != expected {ABI_VERSION}; rebuild csrc/")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().b200grbm_last_error()
        raise B200Error(rc, msg.decode("utf-8", "replace") if msg else "")


def ptr(t) -> Optional[int]:
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream(device) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def device_info() -> dict:
    lib = load()
    vals = [C.c_int32() for _ in range(4)]
    check(lib.b200grbm_device_info(*[C.byref(v) for v in vals]))
    return dict(sm_count=vals[0].value, cc=(vals[1].value, vals[2].value), smem_optin=vals[3].value)


def ex2_max_rel_error(x_lo: float, x_hi: float, n: int = 1 << 24, device=None) -> float:
    """Largest relative error of the hardware exp2 (``ex2.approx.ftz.f32``) over ``n`` arguments of
    ``[x_lo, x_hi]`` -- the quantity the sweep kernel's acceptance bracket has to cover."""
    import torch

    lib = load()
    out = C.c_double()
    dev = torch.device("cuda" if device is None else device)
    with torch.cuda.device(dev):
        check(lib.b200grbm_ex2_probe(float(x_lo), float(x_hi), int(n), C.byref(out), current_stream(dev)))
    return out.value


def tensor_peak(kind: str = "i8", iters: int = 20000, device=None) -> float:
    """Measured tcgen05 peak of the current device in op/s (2 x MAC): ``kind`` "i8" or "bf16"."""
    import torch

    lib = load()
    out = C.c_double()
    dev = torch.device("cuda" if device is None else device)
    with torch.cuda.device(dev):
        check(lib.b200grbm_tensor_peak(0 if kind == "i8" else 1, int(iters), C.byref(out), current_stream(dev)))
    return out.value
