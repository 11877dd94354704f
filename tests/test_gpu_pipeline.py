"""Round-2 paths of the CUDA engine, through the C ABI: pipelined tiled solves
with fused host-side collection (the reference does reshape / active-theta
masking / astype in numpy after the launch, reference
kernel/marginalized/_kernel.py:247-264), caller-owned device outputs,
asynchronous tile copy-back into one host matrix, output-extent validation,
and the BASELINE configurations at their stated sizes (C3 2000-graph output
sampled against the oracle, C4 pairs of 200+ nodes with gradients, a C5
X-by-Y block)."""
import numpy as np
import pytest

from graphdot_b200 import native
from graphdot_b200.kernel.fix import Normalization
from graphdot_b200.kernel.marginalized import MarginalizedGraphKernel
from graphdot_b200.kernel.marginalized._backend_b200 import (B200Backend,
                                                             PairJobs)
from graphdot_b200.kernel.marginalized._tiles import (GramTileWorker,
                                                      col_tiles)
from graphdot_b200.microkernel import (KroneckerDelta, SquareExponential,
                                       TensorProduct)
from graphdot_b200.synthetic import make_config_graphs, make_config_kernel
from oracle import mlgk_oracle as oracle

pytestmark = pytest.mark.gpu
GRAM_RTOL = 1e-5
GRAD_RTOL = 1e-4


@pytest.fixture(scope='module')
def backend():
    return B200Backend()


def rel_err(got, want):
    want = np.asarray(want, float)
    return np.abs(np.asarray(got, float) - want).max() / np.abs(want).max()


def plain_solve(kernel, graphs, nx, symmetric, eval_gradient):
    """One un-tiled launch into float32 buffers, no collection: what a
    reference-style front end gets from ``Backend.__call__``."""
    be = kernel.backend
    n = len(graphs)
    ny = nx if symmetric else n - nx
    T = MarginalizedGraphKernel.traits
    jobs = (PairJobs.triu(0, nx) if symmetric
            else PairJobs.rect(0, nx, nx, n))
    starts = (np.arange(n + 1) if symmetric else np.concatenate(
        [np.arange(nx), np.arange(ny + 1)])).astype(np.uint32)
    K = be.empty(nx * ny, np.float32)
    dK = be.empty(nx * ny * kernel.n_dims, np.float32) if eval_gradient \
        else None
    from graphdot_b200.util import Timer
    be(graphs, kernel.node_kernel, kernel.edge_kernel, kernel.p, kernel.q,
       kernel.eps, kernel.ftol, kernel.gtol, jobs, starts, K, dK, nx, ny,
       kernel.n_dims, T(symmetric=symmetric, eval_gradient=eval_gradient),
       Timer())
    assert be.last['n_launches'] == 1
    K = K.reshape(nx, ny, order='F')
    if dK is not None:
        dK = dK.reshape(nx, ny, kernel.n_dims, order='F')
    return K, dK


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_pipelined_symmetric_solve_is_bit_identical_to_one_launch(backend,
                                                                  dtype):
    """>= 65536 pairs: the front end splits the triangle into row-block
    launches whose column blocks are copied back and converted while the next
    launch runs.  Same bits as one launch + numpy post-processing."""
    G = make_config_graphs('C2', 400)
    kernel = make_config_kernel('C3', backend=backend, dtype=dtype)
    K, dK = kernel(G, eval_gradient=True)
    assert backend.last['n_launches'] > 1
    assert K.dtype == dtype and dK.dtype == dtype
    assert K.shape == (400, 400) and dK.shape == (400, 400, 5)
    K1, dK1 = plain_solve(kernel, G, 400, True, True)
    assert np.array_equal(K, K1.astype(dtype))
    assert np.array_equal(dK, dK1.astype(dtype))
    assert np.array_equal(K, K.T)


def test_pipelined_rectangular_solve_with_fixed_hyperparameters(backend):
    """X x Y with column-block launches; fixed hyper-parameters are dropped
    during collection (reference _kernel.py:249-251)."""
    G = make_config_graphs('C5', 620)
    X, Y = G[:300], G[300:]
    kn = TensorProduct(element=KroneckerDelta(0.5, h_bounds='fixed'),
                       x=SquareExponential(1.0))
    ke = TensorProduct(length=SquareExponential(0.1))
    kernel = MarginalizedGraphKernel(kn, ke, q=0.05, backend=backend)
    assert list(kernel.active_theta_mask) == [True, True, False, True, True]
    K, dK = kernel(X, Y, eval_gradient=True)
    assert backend.last['n_launches'] > 1
    assert K.shape == (300, 320) and dK.shape == (300, 320, 4)
    K1, dK1 = plain_solve(kernel, X + Y, 300, False, True)
    assert np.array_equal(K, K1.astype(float))
    assert np.array_equal(dK, dK1[:, :, [0, 1, 3, 4]].astype(float))
    # float32 with a mask goes through the float32 collection
    k32 = MarginalizedGraphKernel(kn, ke, q=0.05, backend=backend,
                                  dtype=np.float32)
    K32, dK32 = k32(X, Y, eval_gradient=True)
    assert K32.dtype == np.float32 and np.array_equal(K32, K1)
    assert np.array_equal(dK32, dK1[:, :, [0, 1, 3, 4]])


def test_normalized_public_call_pipelined(backend):
    """Normalization(kernel)(G, eval_gradient=True) on enough graphs to
    pipeline: unit diagonal, symmetric, equal to the host formulas."""
    G = make_config_graphs('C2', 380)
    kernel = make_config_kernel('C3', backend=backend)
    K, dK = Normalization(kernel)(G, eval_gradient=True)
    assert backend.last['n_launches'] > 1
    assert np.allclose(np.diag(K), 1.0, atol=2e-7)
    assert np.array_equal(K, K.T)
    Kh, dKh = Normalization(kernel)._host_normalized(G, None, True)
    assert np.allclose(K, Kh, rtol=2e-6, atol=0)
    assert rel_err(dK, dKh) < 1e-5


def test_device_gram_matches_host_result(backend):
    """device_gram writes straight into torch-owned tensors (no clone of the
    engine's buffers, nothing to race with the next solve)."""
    import torch
    G = make_config_graphs('C2', 64)
    kernel = make_config_kernel('C3', backend=backend)
    K, dK = kernel(G, eval_gradient=True)
    for _ in range(2):        # back-to-back solves must not disturb each other
        Kd, dKd = kernel.device_gram(G, eval_gradient=True)
        Kn, dKn = Normalization(kernel).device_gram(G, eval_gradient=True)
    torch.cuda.synchronize()
    assert np.array_equal(Kd.cpu().numpy().astype(float), K)
    assert np.array_equal(dKd.cpu().numpy().astype(float), dK)
    Kh, dKh = Normalization(kernel)(G, eval_gradient=True)
    assert np.array_equal(Kn.cpu().numpy().astype(float), Kh)
    assert np.array_equal(dKn.cpu().numpy().astype(float), dKh)


def test_tile_worker_fills_one_host_matrix_asynchronously(backend):
    """C5-style job: column tiles of X x Y written into a full-size device
    matrix and copied, asynchronously, into ONE page-locked host matrix --
    equal to the public X x Y call."""
    import torch
    G = make_config_graphs('C5', 200)
    nx, ny = 120, 80
    kernel = make_config_kernel('C3', backend=backend)
    w = GramTileWorker(kernel, G, backend, eval_gradient=True, max_rows=0,
                       nx=nx)
    nJ = kernel.n_dims
    Kd = torch.zeros((ny, nx), dtype=torch.float32, device='cuda')
    dKd = torch.zeros((nJ, ny, nx), dtype=torch.float32, device='cuda')
    Kh = native.pinned_empty(nx * ny, np.float32)
    dKh = native.pinned_empty(nx * ny * nJ, np.float32)
    Kh[:] = -1
    dKh[:] = -1
    torch.cuda.synchronize()
    w.diag(store=True, fetch=False)
    for j0, j1 in col_tiles(ny, 24):
        w.run_cols(j0, j1, normalize=True, dev=(Kd.data_ptr(),
                                                dKd.data_ptr()),
                   host=(Kh.ctypes.data, dKh.ctypes.data), async_=True)
    backend.synchronize()
    want_K, want_dK = Normalization(kernel)(G[:nx], G[nx:],
                                            eval_gradient=True)
    got_K = Kh.reshape(nx, ny, order='F')
    got_dK = dKh.reshape(nx, ny, nJ, order='F')
    assert np.array_equal(got_K.astype(float), want_K)
    assert np.array_equal(got_dK.astype(float), want_dK)
    assert np.array_equal(Kd.t().cpu().numpy(), got_K)
    # tile-sized pinned outputs (no device matrix) give the same columns
    w2 = GramTileWorker(kernel, G, backend, eval_gradient=True, max_rows=24,
                        nx=nx)
    Kt, dKt = w2.run_cols(24, 48, normalize=True)
    assert np.array_equal(Kt, got_K[:, 24:48])
    assert np.array_equal(dKt, got_dK[:, 24:48])


def test_outputs_outside_the_buffers_are_rejected(backend):
    """A starts / nX that would make the kernel write out of bounds is an
    error of the ABI call, not a device fault."""
    G = make_config_graphs('C2', 6)
    kernel = make_config_kernel('C2', backend=backend)
    gs = backend.graphset(G)
    T = MarginalizedGraphKernel.traits
    prog = backend.program(gs, kernel.node_kernel, kernel.edge_kernel,
                           kernel.p, T(symmetric=True))
    out = backend.empty(36, np.float32)
    k = kernel
    args = (gs, prog, k.node_kernel, k.edge_kernel, k.p, k.q, k.eps, k.ftol,
            k.gtol, PairJobs.triu(0, 6))
    backend.launch(*args, np.arange(7, dtype=np.uint32), out, None, 6, 6, 5)
    with pytest.raises(native.NativeError, match='outside'):
        backend.launch(*args, np.arange(7, dtype=np.uint32) + 3, out, None,
                       6, 6, 5)
    with pytest.raises(native.NativeError, match='outside'):
        backend.launch(*args, np.arange(7, dtype=np.uint32), out, None, 5, 6,
                       5)
    with pytest.raises(native.NativeError, match='outside'):
        backend.launch(*args, np.arange(7, dtype=np.uint32), out, None, 6, 6,
                       5, row0=2)


def test_c3_full_size_output_sampled_vs_oracle(backend):
    """The 2000-graph normalized Gram + Jacobian that bench.py times, sampled
    against the float64 oracle."""
    G = make_config_graphs('C2', 2000)
    kernel = make_config_kernel('C3', backend=backend)
    K, dK = Normalization(kernel)(G, eval_gradient=True)
    assert K.shape == (2000, 2000) and dK.shape == (2000, 2000, 5)
    assert np.array_equal(K, K.T)
    assert np.allclose(np.diag(K), 1.0, atol=2e-7)
    rng = np.random.default_rng(7)
    kw = dict(knode=kernel.node_kernel, kedge=kernel.edge_kernel, q=kernel.q,
              eval_gradient=True)
    diff = np.zeros(5)
    scale = np.zeros(5)
    for i, j in zip(rng.integers(0, 2000, 24), rng.integers(0, 2000, 24)):
        R, J = oracle.gram([G[i], G[j]], **kw)
        kn = R[0, 1] / np.sqrt(R[0, 0] * R[1, 1])
        dn = (J[0, 1] / np.sqrt(R[0, 0] * R[1, 1])
              - 0.5 * kn * (J[0, 0] / R[0, 0] + J[1, 1] / R[1, 1]))
        assert K[i, j] == pytest.approx(kn, rel=GRAM_RTOL)
        diff = np.maximum(diff, np.abs(dK[i, j] - dn))
        # scale of the terms of the quotient rule (the plane of p cancels to 0)
        scale = np.maximum(scale, np.abs(J[0, 1]) / np.sqrt(R[0, 0] * R[1, 1]))
    assert (diff / scale).max() < GRAD_RTOL


def test_c5_offdiagonal_block_vs_oracle(backend):
    """A block of BASELINE config C5 (seed 5005): X x Y with X and Y disjoint,
    Gram + Jacobian, raw and normalized."""
    G = make_config_graphs('C5', 56)
    X, Y = G[:30], G[30:]
    kernel = make_config_kernel('C3', backend=backend)
    K, dK = kernel(X, Y, eval_gradient=True)
    kw = dict(knode=kernel.node_kernel, kedge=kernel.edge_kernel, q=kernel.q)
    Ko, dKo = oracle.gram(X, Y, eval_gradient=True, **kw)
    assert K.shape == (30, 26)
    assert rel_err(K, Ko) < GRAM_RTOL
    for m in range(5):
        assert rel_err(dK[:, :, m], dKo[:, :, m]) < GRAD_RTOL
    Kn = Normalization(kernel)(X, Y)
    dx = oracle.diag(X, **kw)
    dy = oracle.diag(Y, **kw)
    assert rel_err(Kn, Ko / np.sqrt(np.outer(dx, dy))) < GRAM_RTOL


@pytest.mark.parametrize('pair', [(4, 2), (4, 4), (2, 0)])
def test_c4_full_size_pair_with_gradient_vs_oracle(backend, pair):
    """BASELINE config C4 at its stated size: both graphs have 200-500 nodes
    (203 x 233, 203 x 203 and 233 x 432: N = 41 000 ... 101 000), Convolution
    node kernel over vector features, Gram AND Jacobian against the float64
    oracle (Jacobi-CG to 1e-14 at this size)."""
    G = make_config_graphs('C4', 5)
    a, b = pair
    assert min(len(G[a].nodes), len(G[b].nodes)) >= 200
    kernel = make_config_kernel('C4', backend=backend)
    K, dK = kernel([G[a]], [G[b]], eval_gradient=True)
    assert not backend.last['small_kernel']
    _, ko, go = oracle.solve_pair(G[a], G[b], kernel.node_kernel,
                                  kernel.edge_kernel, kernel.q, kernel.p,
                                  eval_gradient=True)
    assert K[0, 0] == pytest.approx(ko, rel=GRAM_RTOL)
    assert np.allclose(dK[0, 0], go, rtol=GRAD_RTOL,
                       atol=GRAD_RTOL * np.abs(go).max())


def test_c4_symmetric_gram_small_set(backend):
    """Symmetric Gram of two full-size C4 graphs through the public call;
    normalized diagonal is 1."""
    G = make_config_graphs('C4', 5)
    G = [G[4], G[2]]
    kernel = make_config_kernel('C4', backend=backend)
    K = kernel(G)
    Ko = oracle.gram(G, knode=kernel.node_kernel, kedge=kernel.edge_kernel,
                     q=kernel.q)
    assert rel_err(K, Ko) < GRAM_RTOL
    assert np.array_equal(K, K.T)
    Kn = Normalization(kernel)(G)
    assert np.allclose(np.diag(Kn), 1.0, atol=2e-7)


@pytest.mark.parametrize('p_edge, grad', [(0.08, True), (0.3, False),
                                          (0.3, True)])
def test_large_pair_kernel_rare_paths_vs_general_kernel(monkeypatch, p_edge,
                                                        grad):
    """The cluster kernel on inputs unlike C4: weighted molecular-style graphs
    (8-byte edge type) of 70-110 nodes, dense enough (degree up to ~40) that
    rows have more than 8 elements (generic gather loop), columns more than
    16 neighbours (ELL overflow read from global memory) and tile rows touch
    most columns; X-by-Y with different sizes.  Forced by a small
    shared-memory cap; compared with the general kernel and the oracle."""
    from graphdot_b200.synthetic import random_labeled_graph
    rng = np.random.default_rng(11)
    G = [random_labeled_graph(rng, n, p_edge) for n in (71, 96, 110, 83)]
    monkeypatch.setenv('GDB_SMEM_CAP', '120000')
    be = B200Backend()
    kernel = make_config_kernel('C3', backend=be)
    out = kernel(G[:2], G[2:], eval_gradient=grad)
    assert be.last['kernel'] == 'mlgk_solve_large'
    sym = kernel(G, eval_gradient=grad)
    assert be.last['kernel'] == 'mlgk_solve_large'
    monkeypatch.setenv('GDB_FORCE_GENERAL', '1')
    be2 = B200Backend()
    kernel2 = make_config_kernel('C3', backend=be2)
    ref = kernel2(G[:2], G[2:], eval_gradient=grad)
    assert be2.last['kernel'] == 'mlgk_solve'
    sym2 = kernel2(G, eval_gradient=grad)
    if grad:
        assert rel_err(out[0], ref[0]) < 2e-6
        assert rel_err(sym[0], sym2[0]) < 2e-6
        for m in range(5):
            assert rel_err(out[1][:, :, m], ref[1][:, :, m]) < 2e-5
            assert rel_err(sym[1][:, :, m], sym2[1][:, :, m]) < 2e-5
        K = sym[0]
    else:
        assert rel_err(out, ref) < 2e-6
        assert rel_err(sym, sym2) < 2e-6
        K = sym
    assert np.array_equal(K, K.T)
    if p_edge > 0.1:      # the dense oracle is slow on 10^6 edge pairs
        return
    # one pair against the float64 oracle
    _, ko, go = oracle.solve_pair(G[0], G[3], kernel.node_kernel,
                                  kernel.edge_kernel, kernel.q, kernel.p,
                                  eval_gradient=True)
    assert K[0, 3] == pytest.approx(ko, rel=GRAM_RTOL)
    if grad:
        assert np.allclose(sym[1][0, 3], go, rtol=GRAD_RTOL,
                           atol=GRAD_RTOL * np.abs(go).max())


def test_large_pair_kernel_mixed_sizes_and_isolated_nodes(monkeypatch):
    """One graph set with 5 ... 260 nodes (fewer tile rows than CTAs in a
    cluster, sizes that are no multiple of 8 or 32) and a graph with isolated
    nodes (degree 0 -> 1, reference _octilegraph.py:139; rows and columns
    without elements): the cluster kernel equals the general kernel."""
    from graphdot_b200.graph import DataFrame
    from graphdot_b200.synthetic import newman_watts_strogatz
    G = [newman_watts_strogatz(np.random.default_rng(s), n)
         for s, n in ((1, 5), (2, 9), (3, 33), (4, 260), (5, 203))]
    # append three isolated nodes to the 33-node graph
    g = G[2]
    n = len(g.nodes)
    feat = np.empty(n + 3, dtype=object)
    for i in range(n):
        feat[i] = np.asarray(g.nodes['feat'][i])
    for i in range(3):
        feat[n + i] = np.full(8, 0.1 * (i + 1), dtype=np.float32)
    nodes = DataFrame()
    nodes['!i'] = np.arange(n + 3, dtype=np.uint32)
    nodes['feat'] = feat
    G[2] = type(g)(nodes, g.edges, title='isolated')
    be = B200Backend()
    kernel = make_config_kernel('C4', backend=be)
    K, dK = kernel(G, eval_gradient=True)
    assert be.last['kernel'] == 'mlgk_solve_large'
    Kxy = kernel(G[:2], G[2:])
    monkeypatch.setenv('GDB_FORCE_GENERAL', '1')
    be2 = B200Backend()
    kernel2 = make_config_kernel('C4', backend=be2)
    K2, dK2 = kernel2(G, eval_gradient=True)
    assert be2.last['kernel'] == 'mlgk_solve'
    assert np.allclose(K, K2, rtol=5e-6)
    for m in range(dK.shape[2]):
        assert rel_err(dK[:, :, m], dK2[:, :, m]) < 2e-5
    assert np.allclose(Kxy, kernel2(G[:2], G[2:]), rtol=5e-6)
    assert np.array_equal(K, K.T)


def test_large_pair_kernel_module_is_compiled_on_first_need():
    """The cluster kernel is 80 % of the NVRTC time of the template: a program
    that only ever sees small pairs never compiles it (num_regs_large stays 0);
    the first graph set with large pairs compiles and uses it, and the small
    pairs keep their kernel and their results."""
    be = B200Backend()
    kernel = make_config_kernel('C4', backend=be)
    from graphdot_b200.synthetic import newman_watts_strogatz
    small = [newman_watts_strogatz(np.random.default_rng(s), 12 + s)
             for s in range(6)]
    K_small = kernel(small)
    assert be.last['kernel'] == 'mlgk_solve_small'
    progs = list(be._programs.values())
    assert progs and all(be.program_info(p).num_regs_large == 0 for p in progs)
    large = make_config_graphs('C4', 3)
    K_large = kernel(large)
    assert be.last['kernel'] == 'mlgk_solve_large'
    assert any(be.program_info(p).num_regs_large > 0
               for p in be._programs.values())
    assert np.array_equal(kernel(small), K_small)
    assert np.all(np.isfinite(K_large)) and np.array_equal(K_large, K_large.T)


def test_backend_reorder_option_restores_the_tile_footprint():
    """B200Backend(reorder='pbr') relabels large graphs when it packs them:
    shuffled C4 graphs give the Gram matrix of the un-shuffled ones (graph-level
    results do not depend on the labelling) with the shared-memory footprint of
    the natural order; nodal outputs are refused."""
    rng = np.random.default_rng(11)
    natural = make_config_graphs('C4', 4)
    shuffled = [g.permute(rng.permutation(len(g.nodes))) for g in natural]
    be_nat, be_shuf, be_pbr = (B200Backend(), B200Backend(),
                               B200Backend(reorder='pbr'))
    K_nat = make_config_kernel('C4', backend=be_nat)(natural)
    smem_nat = be_nat.last['smem_bytes']
    K_shuf = make_config_kernel('C4', backend=be_shuf)(shuffled)
    smem_shuf = be_shuf.last['smem_bytes']
    kernel = make_config_kernel('C4', backend=be_pbr)
    K_pbr, dK_pbr = kernel(shuffled, eval_gradient=True)
    assert be_pbr.last['kernel'] == 'mlgk_solve_large'
    assert np.allclose(K_shuf, K_nat, rtol=5e-6)
    assert np.allclose(K_pbr, K_nat, rtol=5e-6)
    assert be_pbr.last['smem_bytes'] <= 1.1 * smem_nat < smem_shuf
    _, dK_nat = make_config_kernel('C4', backend=be_nat)(natural,
                                                         eval_gradient=True)
    for m in range(dK_nat.shape[2]):
        assert rel_err(dK_pbr[:, :, m], dK_nat[:, :, m]) < 2e-5
    with pytest.raises(ValueError, match='nodal'):
        kernel(shuffled, nodal=True)
    with pytest.raises(ValueError):
        B200Backend(reorder='metis')


def test_call_memo_is_invalidated_by_an_inplace_permutation():
    """Repeated calls with the same graph objects skip the walk over the
    graphs' caches; relabelling one graph in place (which clears its cache)
    must be seen by the next call."""
    G = make_config_graphs('C2', 6)
    be = B200Backend()
    kernel = make_config_kernel('C2', backend=be)
    R0 = kernel(G, nodal=True)
    assert np.array_equal(kernel(G, nodal=True), R0)       # memo hit
    n = len(G[2].nodes)
    G[2].permute(np.random.default_rng(0).permutation(n), inplace=True)
    R1 = kernel(G, nodal=True)
    R2 = make_config_kernel('C2', backend=B200Backend())(G, nodal=True)
    assert np.array_equal(R1, R2)
    assert not np.array_equal(R1, R0)
    assert np.allclose(kernel(G), make_config_kernel('C2', backend=B200Backend())(G), rtol=1e-6)


@pytest.mark.parametrize('force_general', [False, True])
def test_560_x_530_node_pair_vs_committed_oracle_answer(monkeypatch,
                                                        force_general):
    """560 x 530 nodes (N = 296 800), beyond BASELINE's C4 range: the cluster
    kernel (18 columns per lane) and the general kernel with its global-memory
    arena against the float64 oracle.  The oracle needs 80 s for this pair; its
    answer is committed (tests/golden/large_pair_oracle.json, made by
    make_large_pair_golden.py)."""
    import json
    import os
    import sys
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    sys.path.insert(0, golden)
    from make_large_pair_golden import graphs
    want = json.load(open(os.path.join(golden, 'large_pair_oracle.json')))
    g1, g2 = graphs()
    assert [len(g1.nodes), len(g2.nodes)] == want['sizes']
    if force_general:
        monkeypatch.setenv('GDB_FORCE_GENERAL', '1')
    be = B200Backend()
    kernel = make_config_kernel('C4', backend=be)
    K, dK = kernel([g1], [g2], eval_gradient=True)
    assert be.last['kernel'] == ('mlgk_solve' if force_general
                                 else 'mlgk_solve_large')
    assert K[0, 0] == pytest.approx(want['gram'], rel=GRAM_RTOL)
    go = np.array(want['gradient'])
    assert np.allclose(dK[0, 0], go, rtol=GRAD_RTOL,
                       atol=GRAD_RTOL * np.abs(go).max())


@pytest.mark.parametrize('sizes, which', [((1000, 900), 'mlgk_solve_large'),
                                          ((1500, 1100), 'mlgk_solve')])
def test_largest_pairs_against_the_closed_form(sizes, which):
    """Sizes no oracle solves in seconds: for label-free (Constant) kernels
    K = p^2 n1 n2 / (1 - (1-q)^2) whatever the graphs, so dK/dp = 2 K / p and
    dK/dq = -2 (1-q) K / (1 - (1-q)^2).  1000 x 900 nodes is the top of the
    cluster kernel's range (N = 9e5), 1500 x 1100 (N = 1.65e6) runs in the
    general kernel's global-memory arena."""
    from graphdot_b200.synthetic import newman_watts_strogatz
    g1, g2 = (newman_watts_strogatz(np.random.default_rng(31 + k), n)
              for k, n in enumerate(sizes))
    be = B200Backend()
    kernel = make_config_kernel('C1', backend=be)
    K, dK = kernel([g1], [g2], eval_gradient=True)
    assert be.last['kernel'] == which
    q = kernel.q
    want = sizes[0] * sizes[1] / (1 - (1 - q) ** 2)
    assert K[0, 0] == pytest.approx(want, rel=GRAM_RTOL)
    assert dK.shape == (1, 1, 2)
    assert dK[0, 0, 0] == pytest.approx(2 * want, rel=GRAD_RTOL)
    assert dK[0, 0, 1] == pytest.approx(
        -2 * (1 - q) * want / (1 - (1 - q) ** 2), rel=GRAD_RTOL)


@pytest.mark.parametrize('case', ['large', 'large-mid', 'general-mid'])
def test_large_and_general_kernels_are_bit_reproducible(monkeypatch, case):
    """Identical calls in fresh back ends return identical bits, Jacobian
    included, and the same total number of CG iterations.  (The adjoint solve
    of the cluster kernel once read its right-hand side four elements at a
    time without a barrier after the loop that wrote it one element per
    thread: Jacobians differed in the last bit between runs on mid-size
    graphs.)"""
    from graphdot_b200.synthetic import newman_watts_strogatz
    if case == 'large':
        G = make_config_graphs('C4', 4)
    else:
        G = [newman_watts_strogatz(np.random.default_rng(s), n)
             for s, n in ((1, 41), (2, 56), (3, 64))]
        if case == 'large-mid':
            monkeypatch.setenv('GDB_SMEM_CAP', '40000')
        else:
            monkeypatch.setenv('GDB_FORCE_GENERAL', '1')
    runs = []
    for k in range(4):
        be = B200Backend()
        K, dK = make_config_kernel('C4', backend=be)(G, eval_gradient=True)
        runs.append((K.copy(), dK.copy(), be.last.get('cg_iterations')))
    assert be.last['kernel'] == ('mlgk_solve' if case == 'general-mid'
                                 else 'mlgk_solve_large')
    for K, dK, it in runs[1:]:
        assert np.array_equal(K, runs[0][0])
        assert np.array_equal(dK, runs[0][1])
        assert it == runs[0][2]
