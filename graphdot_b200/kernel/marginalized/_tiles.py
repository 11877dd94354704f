"""Tiled evaluation of large (normalized) Gram matrices: the pair-job set of a
symmetric Gram is cut into row-block tiles ``rows [i0,i1) x columns [i,n)``
that are independent of each other, so any number of workers -- one per GPU,
threads in one process or one process per GPU -- can pull them from a shared
dynamic queue with no collective on the data path.

The reference has no counterpart (it is single-GPU, reference
_backend_cuda.py:49-52); the per-tile call is the same back-end solve that
``MarginalizedGraphKernel.__call__`` issues, restricted to a tile and with a
tile-sized output (``row0``), followed by the reference's normalization
formulas (reference graphdot/kernel/fix.py:46-73) applied per tile.
"""
import threading

import numpy as np

from ._backend_b200 import B200Backend, PairJobs
from ._kernel import MarginalizedGraphKernel


def row_tiles(n, rows):
    """Row blocks of the upper triangle, largest (top) first."""
    return [(i0, min(i0 + rows, n)) for i0 in range(0, n, rows)]


def tile_pairs(i0, i1, n):
    rows, m = i1 - i0, n - i0
    return rows * m - rows * (rows - 1) // 2


class LocalTileQueue:
    """Dynamic tile queue for worker threads of one process."""

    def __init__(self, tiles):
        self.tiles = list(tiles)
        self._lock = threading.Lock()
        self._next = 0

    def __iter__(self):
        while True:
            with self._lock:
                t = self._next
                self._next += 1
            if t >= len(self.tiles):
                return
            yield self.tiles[t]


class StoreTileQueue:
    """Dynamic tile queue shared by the ranks of a ``torch.distributed`` job
    (one process per GPU): an atomic counter in the rendezvous store hands out
    tile indices, so faster ranks simply take more tiles.  No collective and no
    tensor traffic is involved; ``key`` must be unique per pass."""

    def __init__(self, store, tiles, key):
        self.store, self.tiles, self.key = store, list(tiles), key

    def __iter__(self):
        while True:
            t = self.store.add(self.key, 1) - 1
            if t >= len(self.tiles):
                return
            yield self.tiles[t]


def col_tiles(ny, cols):
    """Column blocks of an X-by-Y job rectangle."""
    return [(j0, min(j0 + cols, ny)) for j0 in range(0, ny, cols)]


class GramTileWorker:
    """Evaluates tiles of a Gram matrix on one device; all graphs are resident
    on the device.

    Symmetric mode (``nx=None``): row-block tiles ``rows [i0,i1) x columns
    [i,n)`` of the Gram matrix of ``graphs`` with itself (``run_tile``).
    Rectangular mode (``nx`` given): ``graphs = X + Y`` and tiles are column
    blocks ``X x Y[j0:j1)`` of the off-diagonal block (``run_cols``), the
    BASELINE configuration C5.

    Tile outputs go either to tile-sized page-locked host arrays that are
    reused from tile to tile, or -- ``run_cols(..., dev=, host=)`` -- into a
    full-size matrix in caller-owned device memory and from there,
    asynchronously, into a full-size host matrix (for instance a shared-memory
    mapping that all per-GPU processes of a node fill: the assembled Gram)."""

    def __init__(self, kernel, graphs, backend=None, eval_gradient=False,
                 max_rows=64, stream=None, nx=None, packed=None):
        self.kernel = kernel
        self.backend = backend or kernel.backend
        self.graphs = list(graphs) if graphs is not None else None
        self.eval_gradient = eval_gradient
        self.stream = stream
        be = self.backend
        self.gs = packed if packed is not None else be.graphset(self.graphs)
        self.n = self.gs.n
        self.nx = nx
        self.ny = None if nx is None else self.n - nx
        T = MarginalizedGraphKernel.traits
        k = kernel
        self.prog = be.program(gs=self.gs, node_kernel=k.node_kernel,
                               edge_kernel=k.edge_kernel, p=k.p,
                               traits=T(eval_gradient=eval_gradient))
        self.prog_diag = be.program(gs=self.gs, node_kernel=k.node_kernel,
                                    edge_kernel=k.edge_kernel, p=k.p,
                                    traits=T(diagonal=True,
                                             eval_gradient=eval_gradient))
        self.nJ = k.n_dims
        if nx is None:
            self.starts = np.arange(self.n + 1, dtype=np.uint32)
            width = self.n
        else:      # rows of X, columns of Y (reference _kernel.py:193-198)
            self.starts = np.concatenate([np.arange(nx), np.arange(
                self.ny + 1)]).astype(np.uint32)
            width = nx
        self.max_rows = max_rows
        self.out = self.dout = None
        if max_rows:
            self.out = be.empty(max_rows * width, np.float32)
            self.dout = (be.empty(max_rows * width * self.nJ, np.float32)
                         if eval_gradient else None)
        self.stats = dict(kernel_ms=0.0, cg_iterations=0, matvec_products=0,
                          vector_elements=0, h2d_bytes=0, d2h_bytes=0,
                          launches=0, pairs=0)

    def _launch(self, prog, jobs, out, dout, nX, nY, starts=None, **kw):
        k = self.kernel
        a = self.backend.launch(self.gs, prog, k.node_kernel, k.edge_kernel,
                                k.p, k.q, k.eps, k.ftol, k.gtol, jobs,
                                self.starts if starts is None else starts,
                                out, dout, nX, nY, self.nJ,
                                stream=self.stream, **kw)
        s = self.stats
        s['kernel_ms'] += a.kernel_ms
        s['cg_iterations'] += a.cg_iterations
        s['matvec_products'] += a.matvec_products
        s['vector_elements'] += a.vector_elements
        s['h2d_bytes'] += a.h2d_bytes
        s['d2h_bytes'] += a.d2h_bytes
        s['launches'] += a.n_launches
        s['pairs'] += len(jobs)
        return a

    def diag(self, upload=False, store=False, fetch=True):
        """Self-similarities (and their Jacobians) of all graphs; ``store``
        keeps them on the device for normalized tiles."""
        n = self.n
        jobs = np.empty(n, dtype=[('i', np.uint32), ('j', np.uint32)])
        jobs['i'] = jobs['j'] = np.arange(n)
        d = dd = None
        if fetch:
            d = self.backend.empty(n, np.float32)
            dd = (self.backend.empty(n * self.nJ, np.float32)
                  if self.eval_gradient else None)
        self._launch(self.prog_diag, jobs, d, dd, n, 1, upload=upload,
                     store_diag=store, keep_on_device=not fetch,
                     starts=np.arange(n + 1, dtype=np.uint32))
        if not fetch:
            return None, None
        d = np.array(d, dtype=float)
        if dd is not None:
            dd = np.array(dd, dtype=float).reshape(n, self.nJ, order='F')
        return d, dd

    def run_tile(self, i0, i1, keep_on_device=False, normalize=False):
        """Symmetric mode: tile ``K[i - i0, j]`` for i in [i0,i1), j in [i,n)
        (zeros left of the diagonal), raw or -- after ``diag(store=True)`` --
        normalized on the device; views into reused pinned buffers."""
        rows = i1 - i0
        assert self.nx is None and rows <= self.max_rows
        out = self.out[:rows * self.n]
        dout = (self.dout[:rows * self.n * self.nJ]
                if self.dout is not None else None)
        self._launch(self.prog, PairJobs.triu(i0, i1, self.n), out, dout,
                     rows, self.n, row0=i0, keep_on_device=keep_on_device,
                     normalize=normalize)
        if keep_on_device:
            return None, None
        K = out.reshape(rows, self.n, order='F')
        dK = (dout.reshape(rows, self.n, self.nJ, order='F')
              if dout is not None else None)
        return K, dK

    def run_cols(self, j0, j1, normalize=False, dev=None, host=None,
                 keep_on_device=False, async_=False):
        """Rectangular mode: the column block ``K[:, j0:j1)`` of X x Y.

        Without ``dev`` the tile lands in the worker's reused pinned buffers
        and is returned as (nx, j1 - j0[, nJ]) views.  With ``dev=(K, dK)``
        -- device addresses of full (nx, ny[, nJ]) Fortran-ordered float32
        matrices -- the tile is written in place there; ``host=(K, dK)``
        (addresses of page-locked matrices of the same shape) makes its
        copy-back follow on the engine's copy stream, overlapping the next
        tile when ``async_`` (finish with ``backend.synchronize()``)."""
        nx, ny, w = self.nx, self.ny, j1 - j0
        jobs = PairJobs.rect(0, nx, nx + j0, nx + j1)
        if dev is None:
            assert w <= self.max_rows
            out = self.out[:w * nx]
            dout = (self.dout[:w * nx * self.nJ]
                    if self.dout is not None else None)
            self._launch(self.prog, jobs, out, dout, nx, w, col0=j0,
                         keep_on_device=keep_on_device, normalize=normalize)
            if keep_on_device:
                return None, None
            return (out.reshape(nx, w, order='F'),
                    dout.reshape(nx, w, self.nJ, order='F')
                    if dout is not None else None)
        hk, hd = host if host is not None else (None, None)
        self._launch(self.prog, jobs, _Addr(hk), _Addr(hd), nx, ny,
                     normalize=normalize, gramian_dev=dev[0],
                     gradient_dev=dev[1], async_=async_)
        return None, None

    @staticmethod
    def normalize_tile(K, dK, i0, d, dd):
        """K_ij / sqrt(K_ii K_jj) and its gradient for one raw tile
        (reference kernel/fix.py:46-73)."""
        rows = K.shape[0]
        sl = d[i0:i0 + rows] ** -0.5
        sr = d ** -0.5
        Kn = sl[:, None] * K * sr[None, :]
        if dK is None:
            return Kn, None
        rl = dd[i0:i0 + rows] / d[i0:i0 + rows, None]
        rr = dd / d[:, None]
        dKn = (sl[:, None, None] * dK * sr[None, :, None]
               - 0.5 * Kn[:, :, None] * (rl[:, None, :] + rr[None, :, :]))
        return Kn, dKn


class _Addr:
    """A raw host address where the back end expects an array
    (``.ctypes.data``); None stays None."""

    def __new__(cls, addr):
        if addr is None:
            return None
        self = object.__new__(cls)
        self.ctypes = self
        self.data = int(addr)
        return self


def gram_tiled(kernel, graphs, devices=(0,), eval_gradient=False,
               normalize=True, tile_rows=64):
    """(Normalized) symmetric Gram matrix -- and Jacobian over ALL
    hyper-parameters -- of ``graphs`` using one worker thread per device that
    pull row-block tiles from a shared queue."""
    n = len(graphs)
    queue = LocalTileQueue(row_tiles(n, tile_rows))
    K = np.zeros((n, n), dtype=np.float32)
    dK = (np.zeros((n, n, kernel.n_dims), dtype=np.float32)
          if eval_gradient else None)
    errors = []

    def work(dev):
        try:
            be = B200Backend(device=dev, block_size=getattr(
                kernel.backend, 'block_size', None))
            w = GramTileWorker(kernel, graphs, be, eval_gradient, tile_rows)
            if normalize:
                w.diag(store=True)
            for i0, i1 in queue:
                Kt, dKt = w.run_tile(i0, i1, normalize=normalize)
                K[i0:i1] = Kt
                if dK is not None:
                    dK[i0:i1] = dKt
        except Exception as e:      # surfaced in the caller's thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(d,)) for d in devices]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    iu = np.triu_indices(n, 1)
    K[(iu[1], iu[0])] = K[iu]
    if dK is not None:
        dK[(iu[1], iu[0])] = dK[iu]
        return K, dK
    return K
