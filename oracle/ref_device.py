"""Launcher for the REFERENCE's own device code (no pycuda).

TEST / BENCH INFRASTRUCTURE.  ``oracle/_ref/<name>/`` holds build products made
in the build container by ``oracle/build_ref_device.py`` from the reference's
unmodified ``template.cu`` + ``graphdot/cpp`` (rendered by the reference's own
code generator, compiled by nvcc for sm_100a) and the reference's own
``OctileGraph`` byte layout of the synthetic graphs.  This module loads the
cubin through the CUDA driver API and launches ``graph_kernel_solver`` exactly
as the reference's ``CUDABackend.__call__`` does (reference
graphdot/kernel/marginalized/_backend_cuda.py:303-367): grid = SMs x 8 blocks,
128 threads, ``shmem_bytes_per_warp`` x 4 dynamic shared memory, one
``pcg_scratch_t`` per block, theta copied into the module's ``__constant__``
symbols, one global job counter.

It is the GPU-side oracle ("results identical to the reference's on the same
inputs") and the "reference back end on 1 GPU" throughput comparator of
BASELINE.json.  torch is used for device memory and events only.
"""
import ctypes as C
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')


def available(name):
    return os.path.exists(os.path.join(REF_DIR, name, 'ref.cubin'))


class _Driver:
    def __init__(self):
        self.lib = C.CDLL('libcuda.so.1')

    def check(self, rc, what):
        if rc != 0:
            msg = C.c_char_p()
            self.lib.cuGetErrorString(rc, C.byref(msg))
            raise RuntimeError(f'{what}: CUDA driver error {rc} '
                               f'{(msg.value or b"").decode()}')


class RefDeviceSolver:
    """One reference configuration (cubin + packed graphs) on ``cuda:device``."""

    def __init__(self, name, device=0, block_per_sm=8, block_size=128):
        import torch
        self.torch = torch
        self.dir = os.path.join(REF_DIR, name)
        self.meta = json.load(open(os.path.join(self.dir, 'meta.json')))
        self.dev = torch.device('cuda', device)
        torch.cuda.set_device(self.dev)
        torch.zeros(1, device=self.dev)          # primary context is current
        self.drv = d = _Driver()
        self.block_size = block_size
        self.n_blocks = (torch.cuda.get_device_properties(self.dev)
                         .multi_processor_count * block_per_sm)
        cubin = open(os.path.join(self.dir, 'ref.cubin'), 'rb').read()
        self.mod = C.c_void_p()
        d.check(d.lib.cuModuleLoadData(C.byref(self.mod), cubin), 'load cubin')
        self.fn = C.c_void_p()
        d.check(d.lib.cuModuleGetFunction(C.byref(self.fn), self.mod,
                                          b'graph_kernel_solver'), 'function')
        ptr, size = self._global('shmem_bytes_per_warp')
        self.shmem = size * block_size // 32
        d.check(d.lib.cuFuncSetAttribute(self.fn, 8, C.c_int(self.shmem)),
                'max dynamic smem')     # CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED
        for sym, hexbytes in self.meta['theta'].items():
            raw = bytes.fromhex(hexbytes)
            if raw:
                p, size = self._global(sym)
                assert size >= len(raw), (sym, size, len(raw))
                d.check(d.lib.cuMemcpyHtoD_v2(C.c_uint64(p), raw,
                                              C.c_size_t(len(raw))), sym)
        self._upload_graphs()

    def _global(self, name):
        ptr, size = C.c_uint64(), C.c_size_t()
        self.drv.check(self.drv.lib.cuModuleGetGlobal_v2(
            C.byref(ptr), C.byref(size), self.mod, name.encode()), name)
        return ptr.value, size.value

    def _dev(self, array):
        t = self.torch.from_numpy(np.ascontiguousarray(array))
        return t.to(self.dev)

    def _upload_graphs(self):
        z = np.load(os.path.join(self.dir, 'graphs.npz'))
        m = self.meta
        node_off, oct_off = z['node_off'], z['oct_off']
        self.n_graphs = len(node_off) - 1
        self.sizes = np.diff(node_off)
        self.d_degree = self._dev(z['degree'].astype(np.float32))
        nodes = np.array(z['nodes'])
        if m.get('node_ptr_offsets'):
            # frozen_array data pointers: pool-relative offsets -> device addresses
            self.d_pool = self._dev(z['pool'])
            rows = nodes.reshape(-1, m['node_size'])
            for off in m['node_ptr_offsets']:
                ptr = rows[:, off:off + 8].copy().view(np.uint64)
                ptr += np.uint64(self.d_pool.data_ptr())
                rows[:, off:off + 8] = ptr.view(np.uint8)
        self.d_nodes = self._dev(nodes)
        self.d_edges = self._dev(z['edges'])
        # octile_t {edge_t* elements; u64 nzmask; u64 nzmask_r; int upper, left}
        oct_dt = np.dtype([('elements', '<u8'), ('nzmask', '<u8'),
                           ('nzmask_r', '<u8'), ('upper', '<i4'),
                           ('left', '<i4')])
        octs = np.zeros(oct_off[-1], dtype=oct_dt)
        octs['elements'] = (self.d_edges.data_ptr()
                            + z['oct_elements'].astype(np.uint64)
                            * np.uint64(m['edge_size']))
        for k in ('nzmask', 'nzmask_r', 'upper', 'left'):
            octs[k] = z['oct_' + k]
        self.d_octs = self._dev(octs.view(np.uint8))
        # graph_t {int n_node, n_octile; float* degree; node_t* node; octile_t* octile}
        g_dt = np.dtype([('n_node', '<i4'), ('n_octile', '<i4'),
                         ('degree', '<u8'), ('node', '<u8'), ('octile', '<u8')])
        gs = np.zeros(self.n_graphs, dtype=g_dt)
        gs['n_node'] = self.sizes
        gs['n_octile'] = np.diff(oct_off)
        gs['degree'] = self.d_degree.data_ptr() + 4 * node_off[:-1].astype(np.uint64)
        gs['node'] = (self.d_nodes.data_ptr()
                      + np.uint64(m['node_size']) * node_off[:-1].astype(np.uint64))
        gs['octile'] = self.d_octs.data_ptr() + 32 * oct_off[:-1].astype(np.uint64)
        self.d_graphs = self._dev(gs.view(np.uint8))

    def solve(self, jobs, q, ftol=1e-8, gtol=1e-6, eps=1e-2, n=None):
        """Symmetric Gram (and Jacobian) over the explicit (i, j) ``jobs`` of
        the first ``n`` graphs; returns (K, dK or None, kernel milliseconds)."""
        torch, d, m = self.torch, self.drv, self.meta
        n = self.n_graphs if n is None else n
        grad = bool(m['traits'].get('eval_gradient'))
        nJ = m['n_jac']
        jobs = np.ascontiguousarray(jobs, dtype=np.uint32).reshape(-1, 2)
        d_jobs = self._dev(jobs)
        d_starts = self._dev(np.arange(n + 1, dtype=np.uint32))
        d_gram = torch.zeros(n * n, dtype=torch.float32, device=self.dev)
        d_grad = torch.zeros(n * n * nJ if grad else 1, dtype=torch.float32,
                             device=self.dev)
        d_counter = torch.zeros(1, dtype=torch.int32, device=self.dev)
        # pcg_scratch_t {float* ptr; size_t stride} per block (reference
        # _backend_cuda.py:93-109, _scratch.py:24-36)
        nmax = int(self.sizes[:n].max()) ** 2 * (2 if grad else 1)
        nmax = (nmax + 15) // 16 * 16
        d_scr = torch.empty(self.n_blocks * 5 * nmax, dtype=torch.float32,
                            device=self.dev)
        scr = np.zeros(self.n_blocks, dtype=[('ptr', '<u8'), ('stride', '<u8')])
        scr['ptr'] = d_scr.data_ptr() + 4 * 5 * nmax * np.arange(
            self.n_blocks, dtype=np.uint64)
        scr['stride'] = nmax
        d_scrd = self._dev(scr.view(np.uint8))
        args = [C.c_uint64(self.d_graphs.data_ptr()),
                C.c_uint64(d_scrd.data_ptr()), C.c_uint64(d_jobs.data_ptr()),
                C.c_uint64(d_starts.data_ptr()), C.c_uint64(d_gram.data_ptr()),
                C.c_uint64(d_grad.data_ptr() if grad else 0),
                C.c_uint64(d_counter.data_ptr()), C.c_uint32(len(jobs)),
                C.c_uint32(n), C.c_uint32(n), C.c_uint32(nJ), C.c_float(q),
                C.c_float(q), C.c_float(eps), C.c_float(ftol), C.c_float(gtol)]
        params = (C.c_void_p * len(args))(*[C.cast(C.pointer(a), C.c_void_p)
                                            for a in args])
        stream = torch.cuda.current_stream()
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        d.check(d.lib.cuLaunchKernel(
            self.fn, self.n_blocks, 1, 1, self.block_size, 1, 1, self.shmem,
            C.c_void_p(stream.cuda_stream), params, None), 'launch')
        ev1.record(stream)
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        K = d_gram.cpu().numpy().reshape(n, n, order='F').astype(float)
        dK = (d_grad.cpu().numpy().reshape(n, n, nJ, order='F').astype(float)
              if grad else None)
        return K, dK, ms


def triu_jobs(n):
    i, j = np.triu_indices(n)
    return np.column_stack([i, j]).astype(np.uint32)
