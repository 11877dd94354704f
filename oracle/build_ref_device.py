#!/usr/bin/env python
"""Build the REFERENCE's own device code for the benchmark configurations.

TEST / BENCH INFRASTRUCTURE (build container only: needs /root/reference).

What it does, entirely with the reference's own code (imported in place under
tests/golden/_refshim.py, nothing copied into the repository):

* builds the reference ``OctileGraph`` byte layout of every synthetic graph
  (reference graphdot/kernel/marginalized/_octilegraph.py:37-177);
* renders the reference's CUDA source for the configuration with the
  reference's own code generator (``CUDABackend.gencode_kernel`` /
  ``gencode_probability`` / ``Template.render``, reference
  _backend_cuda.py:157-228, :282-293) into a temporary file;
* compiles it with nvcc for sm_100a with the reference's flags
  (reference _backend_cuda.py:122-129) against /root/reference/graphdot/cpp;
* writes ONLY build products to ``oracle/_ref/<name>/``: the cubin, the packed
  graph arrays (npz) and a json with struct sizes and hyper-parameter bytes.

``oracle/_ref`` is git-ignored and travels to the GPU box, where
``oracle/ref_device.py`` launches the cubin through the CUDA driver API
(no pycuda).  Run:  python oracle/build_ref_device.py
"""
import ctypes
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
sys.path.insert(0, ROOT)
import _refshim  # noqa: E402

OUT = os.path.join(HERE, '_ref')
REF = _refshim.REFERENCE_ROOT


class _Mem(bytearray):
    """Stand-in for a pycuda managed allocation: ``int(mem)`` is its address
    (the reference calls ``int(array.base)``, _octilegraph.py:172-189)."""

    def __int__(self):
        return ctypes.addressof(ctypes.c_char.from_buffer(self))


def _managed(shape, dtype, order='C', mem_flags=0):
    dt = np.dtype(dtype)
    n = int(np.prod(shape))
    mem = _Mem(max(1, n * dt.itemsize))
    return np.ndarray(shape, dtype=dt, buffer=mem, order=order)


def install_shim():
    _refshim.install()
    drv = sys.modules['pycuda.driver']
    drv.managed_empty = _managed
    drv.managed_zeros = lambda s, d, o='C', mem_flags=0: _zero(_managed(s, d, o))
    drv.managed_empty_like = lambda a, mem_flags=0: _managed(a.shape, a.dtype)


def _zero(a):
    a[...] = 0
    return a


def to_ref_graph(g):
    """graphdot_b200 Graph -> reference Graph with identical columns."""
    from graphdot import Graph as RefGraph
    from graphdot.minipandas import DataFrame as RefFrame

    def frame(df):
        out = RefFrame()
        for k in df.columns:
            out[k] = np.asarray(df[k])
        return out
    return RefGraph(frame(g.nodes), frame(g.edges), title=g.title)


def build(name, config, n_graphs, traits_kw):
    from graphdot.codegen import Template
    from graphdot.codegen.cpptool import decltype
    from graphdot.kernel.marginalized import MarginalizedGraphKernel as RefMGK
    from graphdot.kernel.marginalized._backend_cuda import CUDABackend
    from graphdot.kernel.marginalized._octilegraph import OctileGraph
    from graphdot.kernel.marginalized.starting_probability import Uniform
    from graphdot import microkernel as mk
    from graphdot_b200.synthetic import make_config_graphs

    if config in ('C2', 'C3', 'C5'):
        knode = mk.TensorProduct(element=mk.KroneckerDelta(0.5),
                                 x=mk.SquareExponential(1.0))
        kedge = mk.TensorProduct(length=mk.SquareExponential(0.1))
    elif config == 'C1':
        knode, kedge = mk.Constant(1.0), mk.Constant(1.0)
    elif config == 'C4':
        knode = mk.TensorProduct(feat=mk.Convolution(mk.SquareExponential(1.0)))
        kedge = mk.TensorProduct(length=mk.SquareExponential(0.2))
    else:
        raise KeyError(config)
    p = Uniform(1.0)
    traits = RefMGK.traits(**traits_kw)

    graphs = [to_ref_graph(g) for g in make_config_graphs(config, n_graphs)]
    ogs = [OctileGraph(g) for g in graphs]
    og0 = ogs[0]
    weighted, node_t, edge_t = og0.weighted, og0.node_t, og0.edge_t
    if weighted:                      # reference _backend_cuda.py:274-276
        kedge_dev = mk.TensorProduct(weight=mk.Product(), label=kedge)
    else:
        kedge_dev = kedge

    # ---- render with the reference's code generator ----------------------
    tpl = Template(os.path.join(REF, 'graphdot/kernel/marginalized/template.cu'))
    with tpl.context(traits=traits) as t:
        source = t.render(
            node_kernel=CUDABackend.gencode_kernel(knode, 'node_kernel'),
            edge_kernel=CUDABackend.gencode_kernel(kedge_dev, 'edge_kernel'),
            p_start=CUDABackend.gencode_probability(p, 'p_start'),
            node_t=decltype(node_t), edge_t=decltype(edge_t))
    out = os.path.join(OUT, name)
    os.makedirs(out, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        cu = os.path.join(tmp, 'ref.cu')
        open(cu, 'w').write(source)
        cmd = ['nvcc', '-std=c++14', '-O3', '--use_fast_math',
               '--expt-relaxed-constexpr', '--maxrregcount=64', '-lineinfo',
               '-Xptxas', '-v', '-gencode', 'arch=compute_100a,code=sm_100a',
               '-I', os.path.join(REF, 'graphdot/cpp'), '-cubin', cu, '-o',
               os.path.join(out, 'ref.cubin')]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode:
            raise RuntimeError(res.stderr)
        ptxas = [ln for ln in res.stderr.splitlines() if 'registers' in ln]

    # ---- graphs in the reference's device layout (pointers -> offsets) -----
    node_off, oct_off, edge_off = [0], [0], [0]
    degree, nodes, edges = [], [], []
    # variable-length node features (reference _octilegraph.py:45-88): node_t
    # holds frozen_array{ptr, size}; the host pointers of the build container
    # become offsets into one saved feature pool that the launcher relocates
    ptr_fields = [(k, node_t.fields[k][1]) for k in node_t.names
                  if node_t.fields[k][0].names == ('ptr', 'size')]
    pool, pool_at = [], 0
    octs = {k: [] for k in ('elements', 'nzmask', 'nzmask_r', 'upper', 'left')}
    for og in ogs:
        assert og.node_t == node_t and og.edge_t == edge_t
        degree.append(np.asarray(og.degree))
        aos = np.array(og.nodes_aos)            # private copy
        for key, _ in ptr_fields:
            ptrs = aos[key]['ptr'].astype(np.int64)
            sizes = aos[key]['size'].astype(np.int64)
            inner = np.dtype(key.rsplit('::', 1)[1])
            base = int(ptrs.min())
            nbytes = int((ptrs - base).max()
                         + sizes[np.argmax(ptrs)] * inner.itemsize)
            raw = np.frombuffer(ctypes.string_at(base, nbytes), np.uint8)
            pad = (-pool_at) % 16
            pool.append(np.zeros(pad, np.uint8))
            pool_at += pad
            aos[key]['ptr'] = ptrs - base + pool_at
            pool.append(raw)
            pool_at += nbytes
        nodes.append(aos.view(np.uint8).reshape(-1))
        edges.append(np.asarray(og.edges_aos).view(np.uint8).reshape(-1))
        base = int(og.edges_aos.base)
        o = og.octiles
        octs['elements'].append((o['elements'].astype(np.int64) - base)
                                // edge_t.itemsize + edge_off[-1])
        for k in ('nzmask', 'nzmask_r', 'upper', 'left'):
            octs[k].append(np.asarray(o[k]))
        node_off.append(node_off[-1] + og.n_node)
        oct_off.append(oct_off[-1] + og.n_octile)
        edge_off.append(edge_off[-1] + len(og.edges_aos))
    np.savez(os.path.join(out, 'graphs.npz'),
             node_off=np.array(node_off), oct_off=np.array(oct_off),
             degree=np.concatenate(degree), nodes=np.concatenate(nodes),
             edges=np.concatenate(edges),
             pool=(np.concatenate(pool) if pool else np.zeros(0, np.uint8)),
             **{'oct_' + k: np.concatenate(v) for k, v in octs.items()})

    def state(obj):
        dt = np.dtype(obj.dtype)
        return np.array([obj.state], dtype=dt).tobytes().hex() if dt.itemsize \
            else ''
    meta = dict(
        name=name, config=config, n_graphs=n_graphs, traits=traits_kw,
        weighted=bool(weighted), node_size=node_t.itemsize,
        edge_size=edge_t.itemsize,
        node_ptr_offsets=[int(off + node_t.fields[k][0].fields['ptr'][1])
                          for k, off in ptr_fields],
        n_jac=len(p.theta) + 1 + sum(1 for _ in _flat(knode.theta))
        + sum(1 for _ in _flat(kedge.theta)),
        theta={'node_kernel': state(knode), 'edge_kernel': state(kedge_dev),
               'p_start': state(p)},
        nvcc=' '.join(cmd[:-4]), ptxas=ptxas,
        provenance='reference graphdot 0.8.1: template.cu + graphdot/cpp '
                   'rendered by the reference code generator, OctileGraph '
                   'layout by the reference packer')
    json.dump(meta, open(os.path.join(out, 'meta.json'), 'w'), indent=1)
    print(name, ptxas)


def _flat(t):
    for x in t:
        if isinstance(x, (list, tuple)):
            yield from _flat(x)
        else:
            yield x


def main():
    install_shim()
    n = int(os.environ.get('GDB_REF_GRAPHS', 2000))
    build('c2_gram', 'C2', n, dict(symmetric=True))
    build('c3_grad', 'C2', n, dict(symmetric=True, eval_gradient=True))
    build('c1_gram', 'C1', 100, dict(symmetric=True))
    build('c4_gram', 'C4', int(os.environ.get('GDB_REF_C4_GRAPHS', 60)),
          dict(symmetric=True))


if __name__ == '__main__':
    main()
