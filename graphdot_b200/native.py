"""ctypes binding of libgraphdot_b200.so (C ABI: include/graphdot_b200.h).

The shared library is built in-tree by ``graphdot_b200/csrc/build.py`` (called
from ``__graft_entry__.build()``).  There is no fallback: if the library is
missing, ``load()`` raises, and every product path goes through it.  ctypes
releases the GIL for the duration of each foreign call, so one Python thread
per GPU can drive its own context.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libgraphdot_b200.so')

GDB_OK = 0
JOBS_LIST, JOBS_RECT, JOBS_TRIU = 0, 1, 2
OUT_NONE, OUT_F64, OUT_F32 = 0, 1, 2
NODAL_CODES = {False: 0, True: 1, 'block': 2}


class NativeError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f'libgraphdot_b200 error {code}: {message}')
        self.code = code


class DeviceInfo(C.Structure):
    _fields_ = [('device', C.c_int32), ('sm_count', C.c_int32),
                ('cc_major', C.c_int32), ('cc_minor', C.c_int32),
                ('max_smem_per_block_optin', C.c_int32),
                ('max_smem_per_sm', C.c_int32), ('clock_khz', C.c_int32),
                ('l2_bytes', C.c_int32), ('total_mem', C.c_uint64),
                ('name', C.c_char * 64)]


class FunctorSrc(C.Structure):
    _fields_ = [('theta_decl', C.c_char_p), ('theta_size', C.c_uint32),
                ('expr', C.c_char_p), ('n_jac', C.c_uint32),
                ('jac', C.POINTER(C.c_char_p))]


class ProgramDesc(C.Structure):
    _fields_ = [('node_decl', C.c_char_p), ('node_size', C.c_uint32),
                ('edge_decl', C.c_char_p), ('edge_label_size', C.c_uint32),
                ('edge_label_align', C.c_uint32), ('weighted', C.c_int32),
                ('node_kernel', FunctorSrc), ('edge_kernel', FunctorSrc),
                ('p_start', FunctorSrc),
                ('diagonal', C.c_int32), ('symmetric', C.c_int32),
                ('nodal', C.c_int32), ('lmin', C.c_int32),
                ('eval_gradient', C.c_int32), ('block_size', C.c_int32),
                ('workers_per_thread', C.c_int32),
                ('rows_per_warp', C.c_int32),
                ('slots_per_lane', C.c_int32),
                ('extra_options', C.c_char_p),
                ('cluster_size', C.c_int32), ('cols_per_lane', C.c_int32),
                ('ell_slots', C.c_int32)]


class ProgramInfo(C.Structure):
    _fields_ = [('block_size', C.c_int32), ('num_regs', C.c_int32),
                ('num_regs_small', C.c_int32), ('num_regs_large', C.c_int32),
                ('static_smem', C.c_int32), ('local_bytes', C.c_int32),
                ('max_dynamic_smem', C.c_int32), ('n_jac', C.c_int32),
                ('from_cache', C.c_int32), ('compile_ms', C.c_float)]


class GraphSrc(C.Structure):
    _fields_ = [('n_node', C.c_uint32), ('n_edge', C.c_uint32),
                ('nodes', C.c_void_p), ('edge_i', C.c_void_p),
                ('edge_j', C.c_void_p), ('edge_w', C.c_void_p),
                ('edge_labels', C.c_void_p), ('pool', C.c_void_p),
                ('pool_bytes', C.c_uint32)]


class BatchSrc(C.Structure):
    _fields_ = [('n_graphs', C.c_uint32), ('node_off', C.c_void_p),
                ('edge_off', C.c_void_p), ('pool_off', C.c_void_p),
                ('nodes', C.c_void_p), ('edge_i', C.c_void_p),
                ('edge_j', C.c_void_p), ('edge_w', C.c_void_p),
                ('edge_labels', C.c_void_p), ('pool', C.c_void_p)]


class Layout(C.Structure):
    _fields_ = [('node_size', C.c_uint32), ('edge_label_size', C.c_uint32),
                ('edge_label_align', C.c_uint32), ('weighted', C.c_int32),
                ('n_node_ptr', C.c_uint32), ('node_ptr_offset', C.c_uint32 * 8),
                ('n_edge_ptr', C.c_uint32),
                ('edge_ptr_offset', C.c_uint32 * 8)]


class SolveArgs(C.Structure):
    _fields_ = [('job_mode', C.c_int32), ('jobs', C.c_void_p),
                ('n_jobs', C.c_uint64),
                ('i0', C.c_uint32), ('i1', C.c_uint32),
                ('j0', C.c_uint32), ('j1', C.c_uint32),
                ('starts', C.c_void_p), ('n_starts', C.c_uint32),
                ('q', C.c_float), ('eps', C.c_float),
                ('ftol', C.c_float), ('gtol', C.c_float),
                ('node_theta', C.c_void_p), ('edge_theta', C.c_void_p),
                ('p_theta', C.c_void_p),
                ('gramian', C.c_void_p), ('gradient', C.c_void_p),
                ('nX', C.c_uint32), ('nY', C.c_uint32), ('nJ', C.c_uint32),
                ('row0', C.c_uint32), ('col0', C.c_uint32),
                ('store_diag', C.c_int32), ('normalize', C.c_int32),
                ('upload_graphs', C.c_int32),
                ('stream', C.c_void_p), ('keep_on_device', C.c_int32),
                ('gramian_dev', C.c_void_p), ('gradient_dev', C.c_void_p),
                ('tile', C.c_uint32), ('tile_shrink', C.c_float),
                ('out_dtype', C.c_int32),
                ('out_gram', C.c_void_p), ('out_grad', C.c_void_p),
                ('plane_mask', C.c_void_p), ('async_', C.c_int32),
                ('kernel_ms', C.c_float), ('h2d_ms', C.c_float),
                ('d2h_ms', C.c_float), ('cg_iterations', C.c_uint64),
                ('matvec_products', C.c_uint64),
                ('vector_elements', C.c_uint64), ('h2d_bytes', C.c_uint64),
                ('d2h_bytes', C.c_uint64), ('n_launches', C.c_uint32),
                ('used_small_kernel', C.c_int32), ('grid', C.c_uint32),
                ('smem_bytes', C.c_uint32)]


# every symbol declared in include/graphdot_b200.h: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ('gdb_version', C.c_char_p, []),
    ('gdb_last_error', C.c_char_p, []),
    ('gdb_solver_template', C.c_char_p, []),
    ('gdb_context_create', C.c_int, [C.c_int, C.POINTER(_P)]),
    ('gdb_context_destroy', C.c_int, [_P]),
    ('gdb_context_info', C.c_int, [_P, C.POINTER(DeviceInfo)]),
    ('gdb_context_synchronize', C.c_int, [_P]),
    ('gdb_host_alloc', C.c_int, [C.c_size_t, C.POINTER(_P)]),
    ('gdb_host_free', C.c_int, [_P]),
    ('gdb_host_register', C.c_int, [_P, C.c_size_t]),
    ('gdb_host_unregister', C.c_int, [_P]),
    ('gdb_program_create', C.c_int, [_P, C.POINTER(ProgramDesc),
                                     C.POINTER(_P)]),
    ('gdb_program_info_get', C.c_int, [_P, C.POINTER(ProgramInfo)]),
    ('gdb_program_log', C.c_char_p, [_P]),
    ('gdb_program_source', C.c_char_p, [_P]),
    ('gdb_program_destroy', C.c_int, [_P]),
    ('gdb_render_source', C.c_int, [C.POINTER(ProgramDesc), C.POINTER(_P)]),
    ('gdb_program_compile_only', C.c_int, [C.POINTER(ProgramDesc),
                                           C.POINTER(C.c_uint64)]),
    ('gdb_free', None, [_P]),
    ('gdb_graph_packed_size', C.c_int, [C.POINTER(Layout), C.POINTER(GraphSrc),
                                        C.POINTER(C.c_uint64)]),
    ('gdb_graph_pack', C.c_int, [C.POINTER(Layout), C.POINTER(GraphSrc), _P,
                                 C.c_uint64]),
    ('gdb_graphs_pack_batch', C.c_int, [C.POINTER(Layout), C.POINTER(BatchSrc),
                                        _P, _P, C.c_uint64, C.c_int32]),
    ('gdb_graph_reorder', C.c_int, [C.c_uint32, C.c_uint32, _P, _P, C.c_int32,
                                    _P]),
    ('gdb_graph_count_tiles', C.c_int, [C.c_uint32, C.c_uint32, _P, _P, _P,
                                        C.POINTER(C.c_uint64)]),
    ('gdb_graphset_create', C.c_int, [_P, C.POINTER(Layout), C.c_uint32,
                                      C.POINTER(_P), C.POINTER(C.c_uint64),
                                      C.POINTER(_P)]),
    ('gdb_graphset_upload', C.c_int, [_P]),
    ('gdb_graphset_bytes', C.c_int, [_P, C.POINTER(C.c_uint64)]),
    ('gdb_graphset_destroy', C.c_int, [_P]),
    ('gdb_solve', C.c_int, [_P, _P, _P, C.POINTER(SolveArgs)]),
    ('gdb_last_outputs', C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
]

_lib = None


def load():
    """Load the shared library (once) and declare all prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} is missing: build it with '
            '`python -c "import __graft_entry__ as g; g.build()"` or '
            '`python graphdot_b200/csrc/build.py`. There is no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status):
    if status != GDB_OK:
        raise NativeError(status, load().gdb_last_error().decode())


# Page-locking is slow (cudaHostAlloc of the 96 MB C3 output takes tens of
# milliseconds), and every front-end call asks for fresh output buffers: freed
# blocks are therefore kept in a small size-bucketed pool and handed out again.
_POOL = {}              # capacity -> [ptr, ...]
_POOL_BYTES = [0]
_POOL_LIMIT = 8 << 30   # bytes kept for reuse


def _bucket(nbytes):
    b = 4096
    while b < nbytes:
        b *= 2
    # above 1 MiB: 1/8-octave steps, so that large buffers waste <= 12.5 %
    if b > (1 << 20):
        step = b // 16
        b = -(-nbytes // step) * step
    return b


def pinned_empty(count, dtype):
    """numpy array over page-locked host memory (pooled; released with the
    array)."""
    lib = load()
    dtype = np.dtype(dtype)
    nbytes = max(1, int(count) * dtype.itemsize)
    cap = _bucket(nbytes)
    free = _POOL.get(cap)
    if free:
        ptr = free.pop()
        _POOL_BYTES[0] -= cap
    else:
        p = C.c_void_p()
        check(lib.gdb_host_alloc(cap, C.byref(p)))
        ptr = p.value
    raw = (C.c_ubyte * nbytes).from_address(ptr)
    # the ctypes buffer is the ultimate .base of every view numpy derives from
    # this array, so tying the holder to it keeps the allocation alive exactly
    # as long as any view exists
    raw._gdb_holder = _PinnedHolder(ptr, cap)
    return np.frombuffer(raw, dtype=dtype, count=int(count))


def pinned_pool_clear():
    for cap, ptrs in _POOL.items():
        for ptr in ptrs:
            if _lib is not None:
                _lib.gdb_host_free(C.c_void_p(ptr))
    _POOL.clear()
    _POOL_BYTES[0] = 0


class _PinnedHolder:
    def __init__(self, ptr, cap):
        self.ptr, self.cap = ptr, cap

    def __del__(self):
        try:
            if not self.ptr or _lib is None:
                return
            if _POOL_BYTES[0] + self.cap <= _POOL_LIMIT:
                _POOL.setdefault(self.cap, []).append(self.ptr)
                _POOL_BYTES[0] += self.cap
            else:
                _lib.gdb_host_free(C.c_void_p(self.ptr))
        except Exception:
            pass
