#!/usr/bin/env python
"""Golden vectors for the GPR caller from the REFERENCE's own
``GaussianProcessRegressor`` (reference model/gaussian_process/gpr.py), run in
the build container under ``_refshim``:

    python tests/golden/make_gpr_golden.py

TEST INFRASTRUCTURE.  Writes gpr_reference.json next to this file (committed);
cannot run on the GPU box (no /root/reference there)."""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402

_refshim.install()

from graphdot.model.gaussian_process import GaussianProcessRegressor  # noqa: E402
from gpr_cases import RBF, THETAS, data  # noqa: E402

warnings.simplefilter('ignore')
X, y, y_masked, Z = data()
out = {'cases': []}
for s, L in THETAS:
    for reg in ('+', '*'):
        for normalize_y in (False, True):
            gpr = GaussianProcessRegressor(RBF(s, L), alpha=1e-4,
                                           normalize_y=normalize_y,
                                           regularization=reg)
            gpr.fit(X, y_masked)
            lml, dlml = gpr.log_marginal_likelihood(eval_gradient=True)
            sq, dsq = gpr.squared_loocv_error(eval_gradient=True)
            mean, std = gpr.predict(Z, return_std=True)
            _, cov = gpr.predict(Z, return_cov=True)
            loo, loo_std = gpr.predict_loocv(X, y_masked, return_std=True)
            out['cases'].append(dict(
                s=s, L=L, regularization=reg, normalize_y=normalize_y,
                lml=float(lml), dlml=np.asarray(dlml).tolist(),
                sqloocv=float(sq), dsqloocv=np.asarray(dsq).tolist(),
                mean=mean.tolist(), std=std.tolist(), cov=cov.tolist(),
                loo=loo.tolist(), loo_std=loo_std.tolist()))
# a singular Gram matrix (duplicate inputs, alpha = 0): pseudoinverse path
gpr = GaussianProcessRegressor(RBF(1.0, 1.0), alpha=0, beta=1e-8)
Xd = np.array([0.0, 0.0, 1.0, 1.0, 2.0])
yd = np.array([0.1, 0.3, 1.0, 1.2, -0.5])
gpr.fit(Xd, yd)
out['singular'] = dict(X=Xd.tolist(), y=yd.tolist(),
                       mean=gpr.predict(np.array([0.0, 0.5, 1.0, 2.0])).tolist(),
                       lml=float(gpr.log_marginal_likelihood()))
# hyper-parameter optimisation from a fixed start
gpr = GaussianProcessRegressor(RBF(1.0, 1.0), alpha=1e-4, optimizer=True)
gpr.fit(X, y, tol=1e-8)
out['optimized'] = dict(theta=np.asarray(gpr.kernel.theta).tolist(),
                        lml=float(gpr.log_marginal_likelihood()))
json.dump(out, open(os.path.join(HERE, 'gpr_reference.json'), 'w'), indent=1)
print('wrote', len(out['cases']), 'cases; optimized theta', out['optimized'])
