#!/usr/bin/env python
"""Stage the reference's Python package for the GPU box (build container only).

Copies ``/root/reference/graphdot`` (Python files only) and the reference's
hot-path test module (for its fixtures, ``case_dict``) into the git-ignored
``baseline/_ref/``, which travels to the GPU box with the repository snapshot.
``tests/test_gpu_reference_frontend.py`` then drives the REFERENCE's own
``MarginalizedGraphKernel`` / microkernels / ``Graph`` objects through
``backend=B200Backend()`` on a real B200 (the plug point of reference
graphdot/kernel/marginalized/_backend_factory.py:6-8).  TEST INFRASTRUCTURE:
nothing here is imported by the product, nothing is committed to git.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = '/root/reference'
DST = os.path.join(ROOT, 'baseline', '_ref')


def stage():
    if not os.path.isdir(os.path.join(SRC, 'graphdot')):
        return False
    pkg = os.path.join(DST, 'graphdot')
    if os.path.isdir(pkg):
        shutil.rmtree(pkg)
    shutil.copytree(os.path.join(SRC, 'graphdot'), pkg,
                    ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    tdir = os.path.join(DST, 'test', 'kernel', 'marginalized')
    os.makedirs(tdir, exist_ok=True)
    shutil.copy(os.path.join(SRC, 'test/kernel/marginalized/test_kernel.py'),
                tdir)
    return True


if __name__ == '__main__':
    ok = stage()
    print('staged' if ok else 'no /root/reference here', DST)
    sys.exit(0 if ok else 1)
