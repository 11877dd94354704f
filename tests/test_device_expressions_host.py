"""The DEVICE expressions of the microkernels (``gen_expr``: the C++ text that
NVRTC splices into the solver) evaluated on the host against the values of
the reference's own microkernels (tests/golden/microkernel_reference.json).

``test_microkernel.py`` pins ``__call__`` (host evaluation) to those values; the
oracle uses ``__call__``.  This test closes the loop for the device side
without a GPU: the generated functor source -- hyper-parameter struct
(``struct_decl``), value expression and Jacobian expressions, exactly the
strings ``B200Backend`` hands to ``gdb_program_create`` -- is compiled by g++
together with csrc/mlgk_prelude.cuh (``__device__`` & co. defined away) and
run on the golden samples.  A bug shared by ``__call__`` and ``gen_expr`` can
no longer pass both sides.  float32 arithmetic: 2e-5 relative."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from graphdot_b200.kernel.marginalized._backend_b200 import (_Functor,
                                                             state_bytes)
from graphdot_b200.microkernel import (  # noqa: F401  (eval namespace)
    Additive, Composite, Constant, Convolution, DotProduct, KroneckerDelta,
    Normalize, Product, RationalQuadratic, SquareExponential, TensorProduct)

inf = np.inf
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HEAD = r'''
#include <cmath>
#include <cstdio>
#include <cstring>
#define __host__
#define __device__
#define __forceinline__ inline
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
// CUDA fast-math intrinsics the expressions may name (glibc declares the same
// identifiers, hence macros after <cmath>)
static inline float gdb_host_div(float a, float b) { return a / b; }
#define __powf powf
#define __logf logf
#define __expf expf
#define __fdividef gdb_host_div
#include "mlgk_prelude.cuh"
'''


def _literal(v):
    return f'{v}' if isinstance(v, int) else f'{float(v)!r}f'


def _arg_decl(name, v, defs):
    """C++ definition of one sample argument (x and y share the struct type
    ``arg_t``, as the nodes / edge labels of a graph pair do)."""
    if isinstance(v, dict):
        if name == 'x':
            fields = ''.join(f'{"int" if isinstance(x, int) else "float"} {k};'
                             for k, x in v.items())
            defs.append(f'struct arg_t {{ {fields} }};')
        init = ', '.join(_literal(x) for x in v.values())
        defs.append(f'arg_t {name} = {{ {init} }};')
    elif isinstance(v, list):
        init = ', '.join(f'{float(x)!r}f' for x in v)
        defs.append(f'float {name}_d[] = {{ {init} }}; '
                    f'frozen_array<float> {name} = {{ {name}_d, {len(v)} }};')
    else:
        defs.append(f'float {name} = {float(v)!r}f;')


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
def test_device_expressions_reproduce_reference_values(microkernel_golden,
                                                       tmp_path):
    body, expect = [], []
    for n, item in enumerate(microkernel_golden['items']):
        k = eval(item['expr'])
        f = _Functor(k, ('x1', 'x2'))
        theta = state_bytes(k) or b''
        body.append(f'struct k{n}_theta_t {{ {f.theta_decl} }};')
        body.append(
            f'struct k{n}_t : k{n}_theta_t {{\n'
            f'  template<class X> float operator()(X const &x1, X const &x2) '
            f'const {{ return ({f.expr}); }}\n'
            f'  template<class X> void jacobian(X const &x1, X const &x2, '
            f'float *j) const {{\n'
            + ''.join(f'    j[{i}] = ({e});\n' for i, e in enumerate(f.jac))
            + '  }\n};')
        init = ', '.join(str(b) for b in theta) or '0'
        body.append(f'static const unsigned char k{n}_bytes[] = {{ {init} }};')
        body.append(f'static_assert(sizeof(k{n}_theta_t) == {max(1, len(theta))}'
                    f' || {len(theta)} == 0, "theta layout");')
        body.append(f'void run{n}() {{\n  k{n}_t k;\n'
                    f'  memcpy((void *)&k, k{n}_bytes, {len(theta)});')
        for m, s in enumerate(item['samples']):
            defs = []
            _arg_decl('x', s['x'], defs)
            _arg_decl('y', s['y'], defs)
            nj = len(f.jac)
            body.append('  {\n    ' + '\n    '.join(defs) + f'''
    float j[{max(1, nj)}] = {{0}};
    k.jacobian(x, y, j);
    printf("{n} {m} %.9g", (double)k(x, y));
    for (int i = 0; i < {nj}; ++i) printf(" %.9g", (double)j[i]);
    printf("\\n");
  }}''')
            expect.append((n, m, item, s, nj))
        body.append('}')
    main = 'int main() {\n' + ''.join(
        f'  run{n}();\n' for n in range(len(microkernel_golden['items']))
    ) + '  return 0;\n}\n'
    src = tmp_path / 'exprs.cpp'
    src.write_text(HEAD + '\n'.join(body) + '\n' + main)
    exe = tmp_path / 'exprs'
    subprocess.run(['g++', '-std=c++17', '-O1', '-w', '-I',
                    os.path.join(ROOT, 'graphdot_b200', 'csrc'), str(src),
                    '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True,
                         text=True).stdout.strip().splitlines()
    assert len(out) == len(expect)
    checked_jac = 0
    for line, (n, m, item, s, nj) in zip(out, expect):
        tok = line.split()
        assert (int(tok[0]), int(tok[1])) == (n, m)
        got = [float(t) for t in tok[2:]]
        assert got[0] == pytest.approx(s['f'], rel=2e-5, abs=1e-7), item['expr']
        assert len(got) == 1 + nj == 1 + len(item['theta'])
        if len(s['jac']) != nj:
            continue    # reference Add.__call__ bug, see test_microkernel.py
        scale = max(1.0, max((abs(v) for v in s['jac']), default=0.0))
        assert np.allclose(got[1:], s['jac'], rtol=5e-5, atol=2e-6 * scale), \
            (item['expr'], s, got)
        checked_jac += nj
    assert checked_jac > 40
