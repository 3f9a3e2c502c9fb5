"""ctypes binding of include/svof.h.

The same binding drives any shared object implementing the C ABI.  The product
library (CUDA, sm_100a) is located by :func:`load_product`; it raises if the
library is missing -- there is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)

OK = 0
ERR_INVALID_ARG, ERR_BAD_MESH, ERR_BAD_CONFIG, ERR_CUDA, ERR_COMM, ERR_CAPACITY, ERR_STATE, ERR_UNSUPPORTED = \
    -1, -2, -3, -4, -5, -6, -7, -8

PATCH_GENERIC, PATCH_EMPTY, PATCH_PROCESSOR = 0, 1, 2
BC_ZERO_GRADIENT, BC_FIXED_VALUE, BC_INLET_OUTLET = 0, 1, 2

# svof_field
(F_ALPHA, F_ALPHA_PHI, F_DVF, F_INTERFACE_N, F_INTERFACE_D, F_INTERFACE_C, F_INTERFACE_S, F_MIXED_CELLS,
 F_CELL_STATUS, F_FACE_FLATNESS, F_CF, F_SF, F_C, F_V, F_ALPHA_BOUNDARY, F_UN0) = range(16)
# svof_info
(I_N_MIXED, I_MIN_ALPHA_BEFORE, I_MAX_ALPHA_M1_BEFORE, I_MIN_ALPHA_AFTER, I_MAX_ALPHA_M1_AFTER, I_N_BOUND_SWEEPS,
 I_RECONSTRUCTION_TIME, I_ADVECTION_TIME, I_ALPHA_MAPPING_TIME, I_VOLUME, I_GPU_LAUNCHES, I_FLATNESS_MIN,
 I_FLATNESS_MAX, I_FLATNESS_AVG, I_DEVICE_BYTES, I_ERROR_FLAGS, I_DENSE_KERNEL_MS, I_DENSE_KERNEL_LAUNCHES,
 I_N_NEAR, I_H2D_BYTES, I_D2H_BYTES, I_VOLUME_OWNED, I_HALO_BYTES, I_RDF_ITERATIONS, I_SCHEDULE) = range(25)


class SvofPatch(C.Structure):
    _fields_ = [("start", C.c_int32), ("size", C.c_int32), ("kind", C.c_int32), ("nbr_rank", C.c_int32),
                ("alpha_bc", C.c_int32), ("reserved", C.c_int32), ("alpha_value", C.c_double)]


class SvofMesh(C.Structure):
    _fields_ = [("n_points", C.c_int32), ("n_faces", C.c_int32), ("n_internal_faces", C.c_int32),
                ("n_cells", C.c_int32), ("n_patches", C.c_int32), ("reserved", C.c_int32),
                ("points", c_double_p), ("face_offsets", c_int32_p), ("face_points", c_int32_p),
                ("owner", c_int32_p), ("neighbour", c_int32_p), ("patches", C.POINTER(SvofPatch)),
                ("Cf", c_double_p), ("Sf", c_double_p), ("C", c_double_p), ("V", c_double_p)]


class SvofParams(C.Structure):
    _fields_ = [("mixed_cell_tol", C.c_double), ("snap_tol", C.c_double), ("iso_face_tol", C.c_double),
                ("rdf_tol", C.c_double), ("rdf_rel_tol", C.c_double), ("n_alpha_bounds", C.c_int32),
                ("clip", C.c_int32), ("orientation_method", C.c_int32), ("split_warped_face", C.c_int32),
                ("map_alpha_field", C.c_int32), ("write_plic_fields", C.c_int32), ("rdf_iterations", C.c_int32),
                ("mixed_cell_tol_set", C.c_int32), ("alpha_grad_scheme", C.c_int32)]


class SvofComm(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world_size", C.c_int32), ("device", C.c_int32), ("reserved", C.c_int32),
                ("nccl_unique_id", C.c_void_p)]


# every symbol include/svof.h declares: (name, restype, argtypes)
_H = C.c_void_p
SYMBOLS = [
    ("svof_params_default", C.c_int, [C.POINTER(SvofParams)]),
    ("svof_params_set", C.c_int, [C.POINTER(SvofParams), C.c_char_p, C.c_char_p]),
    ("svof_create", C.c_int, [C.POINTER(SvofMesh), C.POINTER(SvofParams), C.POINTER(SvofComm), C.POINTER(_H)]),
    ("svof_destroy", C.c_int, [_H]),
    ("svof_last_error", C.c_char_p, [_H]),
    ("svof_set_alpha", C.c_int, [_H, c_double_p]),
    ("svof_set_phi", C.c_int, [_H, c_double_p]),
    ("svof_set_U", C.c_int, [_H, c_double_p, c_double_p]),
    ("svof_reconstruct", C.c_int, [_H]),
    ("svof_advect", C.c_int, [_H, C.c_double, c_double_p, c_double_p]),
    ("svof_step_device", C.c_int, [_H, C.c_double]),
    ("svof_step_host", C.c_int, [_H, C.c_double, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    ("svof_get_field", C.c_int64, [_H, C.c_int, C.c_void_p, C.c_int64]),
    ("svof_get_info", C.c_int, [_H, C.c_int, c_double_p]),
    ("svof_device_ptr", C.c_int, [_H, C.c_int, C.POINTER(C.c_void_p)]),
    ("svof_device_touch", C.c_int, [_H, C.c_int]),
    ("svof_scatter_alpha_device", C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int64]),
    ("svof_set_phi_device", C.c_int, [_H, C.c_void_p]),
    ("svof_set_U_device", C.c_int, [_H, C.c_void_p, C.c_void_p]),
    ("svof_set_option", C.c_int, [_H, C.c_char_p, C.c_int]),
    ("svof_get_stream", C.c_int, [_H, C.POINTER(C.c_void_p)]),
    ("svof_synchronize", C.c_int, [_H]),
    ("svof_mark", C.c_int, [_H, C.c_int]),
    ("svof_elapsed_ms", C.c_int, [_H, C.c_int, C.c_int, c_double_p]),
    ("svof_last_step_ms", C.c_int, [_H, c_double_p, c_double_p]),
    ("svof_host_alloc", C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    ("svof_host_free", C.c_int, [C.c_void_p]),
    ("svof_cut_faces", C.c_int, [_H, C.c_int32, C.c_int32, c_double_p, c_double_p, c_double_p, c_int32_p,
                                 c_double_p, c_double_p]),
    ("svof_cut_cells", C.c_int, [_H, C.c_int32, c_int32_p, c_double_p, c_double_p, c_int32_p, c_double_p,
                                 c_double_p, c_double_p, c_double_p]),
    ("svof_find_signed_distance", C.c_int, [_H, C.c_int32, c_int32_p, c_double_p, c_double_p, c_int32_p,
                                            c_double_p, c_double_p, c_double_p]),
    ("svof_face_fluxes", C.c_int, [_H, C.c_int32, c_int32_p, c_double_p, c_double_p, c_double_p, C.c_double,
                                   c_double_p, c_double_p]),
    ("svof_plic_surface", C.c_int, [_H, C.c_int64, C.c_int64, c_double_p, c_int32_p, c_int32_p, C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int64)]),
    ("svof_subcell_faces", C.c_int, [_H, C.c_int64, C.c_int64, C.c_int64, c_double_p, c_int32_p, c_int32_p, c_int32_p,
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    # changing meshes
    ("svof_update_points", C.c_int, [_H, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    ("svof_update_mesh", C.c_int, [_H, C.POINTER(SvofMesh)]),
    ("svof_set_interface", C.c_int, [_H, c_double_p, c_double_p]),
    ("svof_map_alpha_field", C.c_int, [_H, C.c_double, C.c_double]),
    ("svof_set_cell_types", C.c_int, [_H, c_int32_p]),
    # decomposed runs
    ("svof_partition_rcb", C.c_int, [C.POINTER(SvofMesh), c_double_p, C.c_int32, c_int32_p]),
    ("svof_decompose", C.c_int, [C.POINTER(SvofMesh), c_int32_p, C.c_int32, C.c_int32, C.POINTER(_H)]),
    ("svof_submesh_mesh", C.c_int, [_H, C.POINTER(SvofMesh)]),
    ("svof_submesh_maps", C.c_int, [_H, C.POINTER(C.c_int32)] + [C.POINTER(c_int32_p)] * 6),
    ("svof_submesh_free", C.c_int, [_H]),
    ("svof_decomp_last_error", C.c_char_p, []),
    ("svof_comm_unique_id", C.c_int, [C.c_void_p]),
    ("svof_halo_setup", C.c_int, [_H, c_int32_p, c_int32_p]),
    ("svof_halo_exchange", C.c_int, [_H]),
    ("svof_halo_setup_faces", C.c_int, [_H, c_int32_p, c_int32_p, c_int32_p]),
    ("svof_halo_exchange_inputs", C.c_int, [_H]),
    ("svof_submesh_face_maps", C.c_int, [_H, C.POINTER(c_int32_p), C.POINTER(c_int32_p)]),
]

PRODUCT_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libsvof_b200.so")


class SvofError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("svof error %d: %s" % (code, msg))
        self.code = code


def load(path):
    """dlopen a library implementing include/svof.h and type its entry points."""
    if not os.path.exists(path):
        raise FileNotFoundError(
            "%s not found: build it first (python -c 'import __graft_entry__ as g; g.build()')" % path)
    # RTLD_LOCAL: the product and the test oracle export the same svof_* names
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    lib._path = path
    return lib


_product = None


def load_product():
    """The CUDA library.  Fails loudly when it has not been built."""
    global _product
    if _product is None:
        _product = load(os.environ.get("SVOF_LIB", PRODUCT_LIB))   # SVOF_LIB: kernel-variant experiments
    return _product


def dptr(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def iptr(a):
    return a.ctypes.data_as(c_int32_p) if a is not None else None


def f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (shape, a.shape))
    return a


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def pinned_array(lib, shape, dtype=np.float64):
    """numpy array over page-locked host memory from svof_host_alloc (lives for the process)."""
    count = int(np.prod(shape))
    n = max(count * np.dtype(dtype).itemsize, 8)
    p = C.c_void_p()
    rc = lib.svof_host_alloc(n, C.byref(p))
    if rc:
        raise SvofError(rc, "svof_host_alloc(%d)" % n)
    buf = (C.c_char * n).from_address(p.value)
    _PINNED_KEEP.append((buf, p))
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


_PINNED_KEEP = []
