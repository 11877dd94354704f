"""CPU-side checks of the native library: it loads without a GPU driver,
exports every symbol declared in include/graphdot_b200.h, the octile packer
reproduces the graph, and NVRTC compiles the spliced solver for sm_100a for
every fixture kernel (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_graphs, golden_kernels
from graphdot_b200 import native
from graphdot_b200.kernel.marginalized import MarginalizedGraphKernel
from graphdot_b200.kernel.marginalized._backend_b200 import (
    AttributeLayout, B200Backend, state_bytes, struct_decl)
from graphdot_b200.kernel.marginalized.starting_probability import (Adhoc,
                                                                    Uniform)
from graphdot_b200.synthetic import make_config_graphs, make_config_kernel


def test_library_exports_every_declared_symbol():
    lib = native.load()
    header = open(os.path.join(ROOT, 'include', 'graphdot_b200.h')).read()
    declared = set(re.findall(r'\b(gdb_[a-z_0-9]+)\s*\(', header))
    assert len(declared) >= 20
    bound = {name for name, _, _ in native.SYMBOLS}
    assert declared == bound
    for name in declared:
        assert hasattr(lib, name), name
    assert b'graphdot_b200' in lib.gdb_version()
    tpl = lib.gdb_solver_template().decode()
    assert 'mlgk_solve' in tpl and 'gdb_pcg' in tpl


def test_no_device_is_reported_not_hidden():
    import torch
    if torch.cuda.is_available():
        pytest.skip('needs a machine without a GPU')
    ctx = C.c_void_p()
    rc = native.load().gdb_context_create(0, C.byref(ctx))
    assert rc == -2 and native.load().gdb_last_error()
    with pytest.raises(native.NativeError):
        native.pinned_empty(4, np.float32)


def _unpack(blob, edge_size, label_off):
    """Decode a packed blob back into (degree, dense weight/label index)."""
    hdr = np.frombuffer(blob[:80], dtype=np.uint32)
    n, n_oct, nnz, n_tile = hdr[:4].astype(int)
    off_deg, off_node, off_oct, off_trow, off_edge, off_pool, total = \
        hdr[4:11].astype(int)
    assert total == len(blob) and total % 16 == 0
    degree = np.frombuffer(blob[off_deg:off_deg + 4 * n], dtype=np.float32)
    oct_dt = np.dtype([('mask', '<u8'), ('start', '<u4'), ('trow', '<u2'),
                       ('tcol', '<u2')])
    octs = np.frombuffer(blob[off_oct:off_oct + 16 * n_oct], dtype=oct_dt)
    trow = np.frombuffer(blob[off_trow:off_trow + 4 * (n_tile + 1)],
                         dtype=np.uint32)
    entries = {}
    k = 0
    for o, t in enumerate(octs):
        assert t['start'] == k
        assert trow[t['trow']] <= o < trow[t['trow'] + 1]
        for bit in range(64):
            if int(t['mask']) >> bit & 1:
                i, j = 8 * int(t['trow']) + bit // 8, 8 * int(t['tcol']) + bit % 8
                entries[(i, j)] = blob[off_edge + k * edge_size:
                                       off_edge + (k + 1) * edge_size]
                k += 1
    assert k == nnz
    keys = [(int(t['trow']), int(t['tcol'])) for t in octs]
    assert keys == sorted(keys) and len(set(keys)) == len(keys)
    return n, degree, entries


def test_octile_packer_roundtrip():
    be = B200Backend()
    rng = np.random.default_rng(5)
    graphs = make_config_graphs('C2', 6) + make_config_graphs('C1', 3)
    for g in graphs:
        p = be.pack_graph(g)
        assert be.pack_graph(g) is p            # cookie cache
        weighted = '!w' in g.edges
        el = AttributeLayout(g.edges, drop=('!i', '!j', '!w'))
        label_off = 4 if weighted else 0
        edge_size = label_off + el.dtype.itemsize
        n, degree, entries = _unpack(p.blob.tobytes(), edge_size, label_off)
        assert n == len(g.nodes)
        ei, ej = np.asarray(g.edges['!i']), np.asarray(g.edges['!j'])
        w = np.asarray(g.edges['!w']) if weighted else np.ones(len(ei))
        want_deg = np.zeros(n)
        np.add.at(want_deg, ei, w)
        np.add.at(want_deg, ej, w)
        assert np.allclose(degree, want_deg, rtol=1e-6)
        assert len(entries) == 2 * len(ei)
        labels = np.zeros(len(ei), dtype=el.dtype)
        for k, _, _ in el.fields:
            if k in g.edges:
                labels[k] = np.asarray(g.edges[k])
        for k, (i, j) in enumerate(zip(ei, ej)):
            for key in ((i, j), (j, i)):
                raw = entries[key]
                if weighted:
                    assert np.frombuffer(raw[:4], np.float32)[0] == w[k]
                assert raw[label_off:] == labels[k].tobytes()
    # invalidation: unify_datatype / permute(inplace) clear the cookie
    g = graphs[0]
    g.permute(rng.permutation(len(g.nodes)), inplace=True)
    assert be.uuid not in g.cookie


def test_row_index_sections_of_the_blob():
    """elem_meta / row_ptr / row_adj / tile_elem / row_pos / lane_map are consistent with
    the octile-ordered element array (the small-pair kernel indexes through
    them instead of decoding bit masks)."""
    be = B200Backend()
    for g in make_config_graphs('C2', 4) + make_config_graphs('C1', 2):
        blob = be.pack_graph(g).blob.tobytes()
        hdr = np.frombuffer(blob[:80], dtype=np.uint32)
        n, n_oct, nnz, n_tile = hdr[:4].astype(int)
        off = {k: int(v) for k, v in zip(
            ('emeta', 'rowptr', 'rowadj', 'tileelem', 'maxdeg', 'rowpos',
             'lanemap', 'vcols'), hdr[12:20])}
        u32 = lambda o, c: np.frombuffer(blob[o:o + 4 * c], dtype=np.uint32)
        emeta, rowptr = u32(off['emeta'], nnz), u32(off['rowptr'], n + 1)
        rowadj, tileelem = u32(off['rowadj'], nnz), u32(off['tileelem'],
                                                        n_tile + 1)
        rowpos = u32(off['rowpos'], nnz)
        lanemap = u32(off['lanemap'], n)
        rows, cols = emeta & 0xffff, emeta >> 16
        ei, ej = np.asarray(g.edges['!i']), np.asarray(g.edges['!j'])
        want = sorted(set(zip(ei.tolist(), ej.tolist()))
                      | set(zip(ej.tolist(), ei.tolist())))
        assert sorted(zip(rows.tolist(), cols.tolist())) == want
        assert rowptr[0] == 0 and rowptr[-1] == nnz
        deg = np.diff(rowptr.astype(int))
        assert off['maxdeg'] == deg.max()
        # lanes the small-pair kernel wants: chunks of 2 / of 4 neighbour slots
        assert off['vcols'] & 0xffff == np.maximum(1, (deg + 1) // 2).sum()
        assert off['vcols'] >> 16 == np.maximum(1, (deg + 3) // 4).sum()
        # lane map: nodes by decreasing degree (stable) and its inverse
        order, pos = lanemap & 0xffff, lanemap >> 16
        assert np.array_equal(order, np.argsort(-deg, kind='stable'))
        assert np.array_equal(pos[order], np.arange(n))
        for i in range(n):
            adj = rowadj[rowptr[i]:rowptr[i + 1]]
            elem = adj >> 16
            assert np.all(rows[elem] == i)                 # elements of row i
            assert np.array_equal(cols[elem], adj & 0xffff)
            assert np.all(np.diff((adj & 0xffff).astype(int)) > 0)
            assert np.array_equal(rowpos[rowptr[i]:rowptr[i + 1]],
                                  i | (np.arange(deg[i]) << 16))
        for t in range(n_tile):
            sel = rows[tileelem[t]:tileelem[t + 1]]
            assert np.all(sel // 8 == t)
        assert tileelem[-1] == nnz


def test_isolated_node_gets_unit_degree():
    from graphdot_b200 import Graph
    g = Graph({'!i': np.arange(3, dtype=np.uint32)},
              {'!i': np.array([0], np.uint32), '!j': np.array([1], np.uint32)})
    be = B200Backend()
    n, degree, entries = _unpack(be.pack_graph(g).blob.tobytes(), 1, 0)
    assert list(degree) == [1.0, 1.0, 1.0] and len(entries) == 2


def test_variable_length_features_are_pooled(mlgk_golden):
    G = golden_graphs(mlgk_golden['cases']['vario-features'])
    be = B200Backend()
    nl, el, weighted = be._layouts(G[0])
    assert nl.decl == 'frozen_array<int16> rings;'
    assert el.decl == 'frozen_array<int16> spectrum;'
    assert nl.dtype.itemsize == 16 and nl.ptr_offsets == [0] and weighted
    blob = be.pack_graph(G[0]).blob.tobytes()
    hdr = np.frombuffer(blob[:80], dtype=np.uint32)
    off_node, off_pool = int(hdr[5]), int(hdr[9])
    rows = np.frombuffer(blob[off_node:off_node + 3 * 16], dtype=nl.dtype)
    want = [[5, 6], [3], [2, 3, 4]]
    for r, w in zip(rows['rings'], want):
        start = int(r['_data'])        # blob-relative after packing
        assert start >= off_pool and int(r['size']) == len(w)
        got = np.frombuffer(blob[start:start + 2 * len(w)], dtype=np.int16)
        assert list(got) == w


def test_struct_decl_and_state_bytes():
    knode, kedge = golden_kernels('labeled')
    decl = struct_decl(knode.dtype)
    assert decl == ('struct{struct{float32 h;}hybridization;struct{struct{'
                    'float32 length_scale;}k1;struct{float32 c;}k2;}charge;}'
                    'kernel;')
    raw = state_bytes(knode)
    assert np.allclose(np.frombuffer(raw, np.float32), [0.3, 1.0, 0.01])
    from graphdot_b200.microkernel import Product, TensorProduct
    assert state_bytes(Product()) is None
    k = TensorProduct(weight=Product(), label=kedge)
    assert struct_decl(k.dtype).startswith('struct{struct{')
    assert state_bytes(Adhoc(lambda n: 1, 'n.x')) == b'\0'
    assert np.frombuffer(state_bytes(Uniform(2.5)), np.float32)[0] == 2.5


CASES = ['unlabeled', 'labeled', 'weighted', 'vario-features', 'molecular']


TRAITS = [
    dict(symmetric=True), dict(symmetric=True, eval_gradient=True),
    dict(diagonal=True, nodal=True), dict(nodal=True, lmin=1),
    dict(diagonal=True, nodal='block'), dict(diagonal=True, eval_gradient=True,
                                             lmin=1),
    dict(symmetric=True, nodal=True, eval_gradient=True)]
_COMPILED = {}


def _compile_fixture_kernels(golden):
    """NVRTC-compile every (fixture, traits) program once, four at a time on
    host threads (ctypes releases the GIL; each compile runs its kernels'
    modules on threads of its own)."""
    if _COMPILED:
        return _COMPILED
    from concurrent.futures import ThreadPoolExecutor
    lib = native.load()

    def one(job):
        name, t = job
        G = golden_graphs(golden['cases'][name])
        knode, kedge = golden_kernels(name)
        nl, el, weighted = B200Backend._layouts(G[0])
        d, keep, _ = B200Backend._desc(
            nl, el, weighted, knode, kedge, Uniform(1.0),
            MarginalizedGraphKernel.traits(**TRAITS[t]), (96, 1), ())
        size = C.c_uint64()
        rc = lib.gdb_program_compile_only(C.byref(d), C.byref(size))
        return job, (rc, size.value, lib.gdb_last_error().decode())

    jobs = [(name, t) for name in CASES for t in range(len(TRAITS))]
    with ThreadPoolExecutor(4) as pool:
        _COMPILED.update(pool.map(one, jobs))
    return _COMPILED


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('t', range(len(TRAITS)))
def test_nvrtc_compiles_fixture_kernels_for_sm100a(mlgk_golden, name, t):
    rc, size, err = _compile_fixture_kernels(mlgk_golden)[(name, t)]
    assert rc == 0, err
    assert size > 1000


def test_compile_errors_are_reported():
    lib = native.load()
    g = make_config_graphs('C2', 1)[0]
    k = make_config_kernel('C2', backend=B200Backend())
    nl, el, weighted = B200Backend._layouts(g)
    T = MarginalizedGraphKernel.traits
    bad_p = Adhoc(lambda nodes: np.ones(len(nodes)), 'n.no_such_attribute')
    d, keep, _ = B200Backend._desc(nl, el, weighted, k.node_kernel,
                                   k.edge_kernel, bad_p, T(), 32, ())
    assert lib.gdb_program_compile_only(C.byref(d), None) == -3
    assert b'no_such_attribute' in lib.gdb_last_error()
    d, keep, _ = B200Backend._desc(nl, el, weighted, k.node_kernel,
                                   k.edge_kernel, Uniform(1.0), T(), 48, ())
    assert lib.gdb_program_compile_only(C.byref(d), None) == -1
    d, keep, _ = B200Backend._desc(
        nl, el, weighted, k.node_kernel, k.edge_kernel, Uniform(1.0),
        T(nodal=True, eval_gradient=True), (96, 5), ())
    assert lib.gdb_program_compile_only(C.byref(d), None) == -1   # wpt > 4
    for shape, what in (((96, 1, 9), b'rows_per_warp'),
                        ((96, 1, 6, 3), b'slots_per_lane')):
        d, keep, _ = B200Backend._desc(
            nl, el, weighted, k.node_kernel, k.edge_kernel, Uniform(1.0),
            T(symmetric=True), shape, ())
        assert lib.gdb_program_compile_only(C.byref(d), None) == -1
        assert what in lib.gdb_last_error()
    d, keep, _ = B200Backend._desc(
        nl, el, weighted, k.node_kernel, k.edge_kernel, Uniform(1.0),
        T(nodal=True, eval_gradient=True), (96, 1), ())
    assert lib.gdb_program_compile_only(C.byref(d), None) == 0    # nodal Jacobian


def test_pair_jobs_descriptor_matches_explicit_lists():
    from graphdot_b200.kernel.marginalized._backend_b200 import PairJobs
    t = np.asarray(PairJobs.triu(0, 7))
    i, j = np.triu_indices(7)
    assert len(PairJobs.triu(0, 7)) == 28
    assert np.array_equal(t['i'], i) and np.array_equal(t['j'], j)
    t = np.asarray(PairJobs.triu(2, 5, 9))      # rows 2..4, columns i..8
    want = [(a, b) for a in range(2, 5) for b in range(a, 9)]
    assert len(PairJobs.triu(2, 5, 9)) == len(want)
    assert [tuple(x) for x in t.tolist()] == want
    r = np.asarray(PairJobs.rect(1, 3, 4, 7))
    assert [tuple(x) for x in r.tolist()] == \
        [(a, b) for a in range(1, 3) for b in range(4, 7)]
    assert B200Backend.array(PairJobs.rect(0, 1, 0, 1)).mode == 1


def test_heterogeneous_graphs_raise_type_error():
    be = B200Backend()
    kernel = make_config_kernel('C2', backend=be)
    a = make_config_graphs('C2', 1)[0]
    b = make_config_graphs('C1', 1)[0]
    with pytest.raises(TypeError):
        kernel([a, b])


def test_launch_shape_follows_the_graph_set():
    """_pick_block: ~6 rows per warp for the small-pair kernel, two neighbour
    slots per lane for sparse (molecular) graphs and four for dense ones; the
    shape macros reach the rendered source."""
    be = B200Backend()
    assert be._pick_block(np.array([16, 24]), True, 2.2) == (128, 1, 6, 2)
    assert be._pick_block(np.array([10, 20]), False, 5.5) == (128, 1, 5, 4)
    assert be._pick_block(np.array([40]), False, 2.0) == (224, 2, 6, 2)
    assert be._pick_block(np.array([40]), True, 2.0)[0] in (128, 256)   # general kernel
    assert B200Backend(block_size=96)._pick_block(
        np.array([24]), True, 2.2) == (96, 1, 8, 2)
    assert B200Backend(slots_per_lane=4)._pick_block(
        np.array([24]), True, 2.2)[3] == 4
    from graphdot_b200.kernel.marginalized._backend_b200 import preset_sources
    src = preset_sources()['c3_molecular_grad']
    for line in ('#define GDB_BLOCK 128', '#define GDB_RPW 6',
                 '#define GDB_ADJ 2', '#define GDB_MIN_BLOCKS_SMALL 5'):
        assert line in src


def test_rcm_reordering_shrinks_the_octile_footprint():
    """reference graph/reorder/rcm.py: a ring lattice with shuffled labels
    packs into far fewer octiles after reverse Cuthill-McKee; the packed header
    agrees with the host-side tile count; solver-relevant content (degrees,
    element count) is unchanged."""
    from graphdot_b200.reorder import octile_count, rcm
    from graphdot_b200.synthetic import newman_watts_strogatz
    rng = np.random.default_rng(7)
    g = newman_watts_strogatz(rng, 120, k=4, p=0.0)
    shuffled = g.permute(rng.permutation(120))
    perm = rcm(shuffled)
    assert sorted(perm.tolist()) == list(range(120))
    ordered = shuffled.permute(perm)
    be = B200Backend()

    def header(graph):
        return be.pack_graph(graph).blob[:16].view(np.int32)

    h_shuf, h_ord = header(shuffled), header(ordered)
    assert h_shuf[1] == octile_count(shuffled)
    assert h_ord[1] == octile_count(ordered) == octile_count(shuffled, perm)
    assert h_ord[1] < 0.5 * h_shuf[1]
    assert h_ord[0] == h_shuf[0] == 120 and h_ord[2] == h_shuf[2]


@pytest.mark.skipif(not os.path.exists('/usr/local/cuda/bin/nvcc'),
                    reason='needs nvcc')
def test_large_pair_kernel_fits_its_register_budget(tmp_path):
    """The cluster kernel runs two CTAs of 256 threads per SM: ptxas may use
    128 registers and must not spill (at 512 threads / 64 registers it spilled
    316 B per thread and ran 7 % slower).  Compiled
    for sm_100a from the rendered C4 source, large-pair module only."""
    import subprocess
    from graphdot_b200.kernel.marginalized._backend_b200 import preset_sources
    src = tmp_path / 'c4.cu'
    src.write_text(preset_sources()['c4_convolution'])
    out = subprocess.run(
        ['/usr/local/cuda/bin/nvcc', '-gencode',
         'arch=compute_100a,code=sm_100a', '-std=c++17', '-O3',
         '--use_fast_math', '-lineinfo', '-DGDB_BUILD_MASK=4', '-Xptxas',
         '-v', '-cubin', str(src), '-o', str(tmp_path / 'c4.cubin')],
        capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    log = out.stdout + out.stderr
    assert 'mlgk_solve_large' in log
    assert '0 bytes spill stores, 0 bytes spill loads' in log
    regs = [int(n) for n in re.findall(r'Used (\d+) registers', log)]
    assert regs and max(regs) <= 128
