"""ctypes binding of the C ABI declared in include/femgpu.h (the drop-in boundary).

Loading fails loudly when libfemgpu.so is missing: there is no Python or CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# FEMGPU_LIB selects another build of the same library (the profiling variant, `make prof`)
LIB_PATH = os.environ.get("FEMGPU_LIB") or os.path.join(HERE, "libfemgpu.so")

u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
u64p = C.POINTER(C.c_uint64)
dp = C.POINTER(C.c_double)
fp = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)
H = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/femgpu.h declares
SIGNATURES = {
    "femgpu_create": (C.c_int32, [C.POINTER(H), C.c_double, C.c_double, C.c_uint32, C.c_int32]),
    "femgpu_reset": (C.c_int32, [H, C.c_uint32]),
    "femgpu_destroy": (None, [H]),
    "femgpu_last_error": (C.c_char_p, [H]),
    "femgpu_add_nodes": (C.c_int32, [H, C.c_size_t, u32p, dp, dp, dp]),
    "femgpu_add_truss": (C.c_int32, [H, C.c_size_t, u32p, u32p, u32p, dp, dp, dp]),
    "femgpu_add_beam": (C.c_int32, [H, C.c_size_t, u32p, u32p, u32p] + [dp] * 9),
    "femgpu_add_plate": (C.c_int32, [H, C.c_size_t, u32p, u32p, u32p, u32p, u32p, dp, dp, dp, dp]),
    "femgpu_validate": (C.c_int32, [H, i32p, u32p, i32p]),
    "femgpu_counts": (C.c_int32, [H, u64p, u64p, u64p, u64p]),
    "femgpu_get_numbers": (C.c_int32, [H, C.c_int32, u32p]),
    "femgpu_symbolic": (C.c_int32, [H, i64p, i64p]),
    "femgpu_numeric": (C.c_int32, [H]),
    "femgpu_synchronize": (C.c_int32, [H]),
    "femgpu_assemble": (C.c_int32, [H, i64p, i64p]),
    "femgpu_get_csr": (C.c_int32, [H, i64p, i32p, dp]),
    "femgpu_get_csr_device": (C.c_int32, [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_void_p), i64p, i64p]),
    "femgpu_get_nonzero_coo": (C.c_int32, [H, i64p, i64p, i64p, dp]),
    "femgpu_get_nonzero_csr": (C.c_int32, [H, i64p, i64p, i32p, dp]),
    "femgpu_rotation_elements": (C.c_int32, [H, C.c_int32, C.c_uint32, dp]),
    "femgpu_element_matrix": (C.c_int32, [H, C.c_int32, C.c_uint32, dp]),
    "femgpu_element_slots": (C.c_int32, [H, C.c_int32, C.c_uint32, i64p]),
    "femgpu_add_displacement": (C.c_int32, [H, C.c_size_t, u32p, i32p, dp]),
    "femgpu_add_concentrated_load": (C.c_int32, [H, C.c_size_t, u32p, i32p, dp]),
    "femgpu_add_line_load": (C.c_int32, [H, C.c_size_t, u32p, i32p, dp]),
    "femgpu_add_surface_load": (C.c_int32, [H, C.c_size_t, u32p, i32p, dp]),
    "femgpu_get_forces": (C.c_int32, [H, dp, C.POINTER(C.c_void_p)]),
    "femgpu_separate_sparse": (C.c_int32, [H, i64p, i64p, i64p]),
    "femgpu_separate_direct": (C.c_int32, [H, i64p, i64p, i64p]),
    "femgpu_get_skyline": (C.c_int32, [H, i64p, dp, i64p]),
    "femgpu_get_separated_indexes": (C.c_int32, [H, i64p, i64p]),
    "femgpu_get_separated_csr": (C.c_int32, [H, C.c_int32, i64p, i32p, dp]),
    "femgpu_get_separated_csr_device": (C.c_int32, [H, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                                    C.POINTER(C.c_void_p)]),
    "femgpu_get_separated_dense": (C.c_int32, [H, C.c_int32, dp]),
    "femgpu_separated_rhs": (C.c_int32, [H, dp, C.POINTER(C.c_void_p)]),
    "femgpu_last_separate_ms": (C.c_int32, [H, fp]),
    "femgpu_last_separate_path": (C.c_int32, [H, C.POINTER(C.c_int32)]),
    "femgpu_solve_pcg": (C.c_int32, [H, C.c_int32, C.c_int64, i64p]),
    "femgpu_solve_direct": (C.c_int32, [H]),
    "femgpu_get_ua": (C.c_int32, [H, dp, C.POINTER(C.c_void_p)]),
    "femgpu_set_ua": (C.c_int32, [H, dp]),
    "femgpu_solve_info": (C.c_int32, [H, i64p, dp, fp]),
    "femgpu_global_analysis": (C.c_int32, [H]),
    "femgpu_get_reactions": (C.c_int32, [H, dp, C.POINTER(C.c_void_p)]),
    "femgpu_get_global_result": (C.c_int32, [H, dp, dp]),
    "femgpu_set_displacements": (C.c_int32, [H, dp]),
    "femgpu_element_results": (C.c_int32, [H, C.c_int32, dp, C.POINTER(C.c_void_p)]),
    "femgpu_launch_count": (C.c_int32, [H, C.c_int32, u64p]),
    "femgpu_last_numeric_ms": (C.c_int32, [H, fp]),
    "femgpu_numeric_ms_history": (C.c_int32, [H, C.c_uint32, fp]),
    "femgpu_numeric_kernel_ms": (C.c_int32, [H, C.c_uint32, fp, i32p]),
    "femgpu_device_bytes": (C.c_int32, [H, u64p]),
    "femgpu_stream": (C.c_int32, [H, C.POINTER(C.c_void_p)]),
    "femgpu_fp64_fma_peak": (C.c_int32, [H, dp]),
    "femgpu_dist_unique_id": (C.c_int32, [u8p]),
    "femgpu_dist_init": (C.c_int32, [H, C.c_int32, C.c_int32, u8p]),
    "femgpu_dist_set_ownership": (C.c_int32, [H, C.c_uint32, C.c_uint32]),
    "femgpu_dist_set_node_window": (C.c_int32, [H, C.c_uint32]),
    "femgpu_dist_last_exchange_bytes": (C.c_int32, [H, u64p, u64p]),
    "femgpu_dist_info": (C.c_int32, [H, i32p, u64p]),
}

_lib = None


def load():
    """Load libfemgpu.so and bind every entry point. Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m finite_element_method_b200.build` "
            "(nvcc, sm_100a). femgpu has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
