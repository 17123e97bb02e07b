import sys, numpy as np
sys.path.insert(0, '.')
from tests import util
from harmonic_b200 import _lib
m = util.make_realnvp(64, seed=5, ws=0.5)
m.kernel_path = _lib.HB_PATH_TCGEN05
x = np.random.default_rng(6).standard_normal((1000, 64)) * 1.2
got = m.predict(x)
want = util.oracle_predict(m, x)
err = np.abs(got - want) / np.maximum(np.abs(want), 1)
print("tc max rel err", err.max(), "median", np.median(err))
print(got[:5], want[:5])
