#!/usr/bin/env python
"""Build-container check (needs /root/reference): the B200 back end accepts
the REFERENCE's own Graph / microkernel / starting-probability objects.

For every fixture of reference test/kernel/marginalized/test_kernel.py:129-170
the reference's microkernels' gen_expr() strings, dtypes and states are fed to
B200Backend (packing + NVRTC compile for sm_100a, no GPU needed), proving that
the spliced-expression dialect and struct layouts are compatible
(INTEGRATION.md section 1).  TEST INFRASTRUCTURE; cannot run on the GPU box.
"""
import ctypes as C
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refshim  # noqa: E402

_refshim.install()
spec = importlib.util.spec_from_file_location(
    'ref_test_kernel', os.path.join(
        _refshim.REFERENCE_ROOT, 'test/kernel/marginalized/test_kernel.py'))
ref_test = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_test)

from graphdot.kernel.marginalized import MarginalizedGraphKernel as RefMGK  # noqa: E402
from graphdot.kernel.marginalized.starting_probability import Uniform as RefUniform  # noqa: E402
from graphdot_b200 import native  # noqa: E402
from graphdot_b200.kernel.marginalized._backend_b200 import (B200Backend,  # noqa: E402
                                                             state_bytes)

lib = native.load()
be = B200Backend()
ok = True
for name, case in ref_test.case_dict.items():
    G = case['graphs']
    packed = [be.pack_graph(g) for g in G]
    assert packed[0].key == packed[1].key
    nl, el, weighted = be._layouts(G[0])
    for traits in (RefMGK.traits(symmetric=True),
                   RefMGK.traits(symmetric=True, eval_gradient=True),
                   RefMGK.traits(diagonal=True, nodal=True, lmin=1)):
        d, keep, _ = be._desc(nl, el, weighted, case['knode'], case['kedge'],
                              RefUniform(1.0), traits, 64, ())
        size = C.c_uint64()
        rc = lib.gdb_program_compile_only(C.byref(d), C.byref(size))
        status = 'ok' if rc == 0 else lib.gdb_last_error().decode()[:400]
        ok &= rc == 0
        print(f'{name:16s} {tuple(traits)} cubin {size.value:7d} B  {status}')
    for k in (case['knode'], case['kedge'], RefUniform(2.0)):
        b = state_bytes(k)
        print(f'    theta bytes {0 if b is None else len(b):3d}  '
              f'{type(k).__name__}')
sys.exit(0 if ok else 1)
