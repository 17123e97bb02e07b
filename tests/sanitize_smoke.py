"""One small launch of every kernel (for compute-sanitizer): python tests/sanitize_smoke.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import harmonic_b200 as hm  # noqa: E402
from harmonic_b200 import _lib  # noqa: E402
from tests import util  # noqa: E402

cases = [
    ("realnvp tc", util.make_realnvp(64, seed=1, standardize=True, ws=0.3), _lib.HB_PATH_TCGEN05),
    ("realnvp generic 64", util.make_realnvp(64, seed=1, ws=0.3), _lib.HB_PATH_SIMT),
    ("realnvp small 5", util.make_realnvp(5, 6, 2, seed=1, standardize=True, ws=0.3), _lib.HB_PATH_SIMT),
    ("realnvp generic 5", util.make_realnvp(5, seed=1, ws=0.3), _lib.HB_PATH_SIMT_GENERIC),
    ("rqspline fast", util.make_rqspline(2, 8, 8, (64, 64), seed=1, ws=0.3), _lib.HB_PATH_AUTO),
    ("rqspline generic", util.make_rqspline(5, 2, 6, (16,), seed=1, ws=0.3), _lib.HB_PATH_AUTO),
]
for name, m, path in cases:
    m.kernel_path = path
    ch = util.gaussian_chains(m.ndim, 7, 333, seed=0, ragged=True)
    ev = hm.Evidence(ch.nchains, m)
    ev.add_chains(ch)
    ref = util.oracle_evidence(m, ch)
    util.assert_evidence_close(ev, ref)
    p = m.predict(ch.samples[:777])
    util.check_predict(m, ch.samples[:777], p)
    eps = np.random.default_rng(2).standard_normal((301, m.ndim))
    util.check_sample(m, eps, m.sample(301, eps=eps))
    print("ok", name, ev.ln_evidence_inv)
ev = hm.Evidence(5, cases[0][1])
ev.add_lnpred(np.random.default_rng(0).standard_normal(5000), np.random.default_rng(1).standard_normal(5000),
              [0, 10, 1000, 1000, 4000, 5000])
print("ok precomputed", ev.ln_evidence_inv)
