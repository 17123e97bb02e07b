import sys
sys.path.insert(0, ".")
import numpy as np, torch
from tests import util
from harmonic_b200 import _lib
m = util.make_realnvp(64, seed=5, standardize=True, ws=0.3)
m.kernel_path = _lib.HB_PATH_TCGEN05
rng = np.random.default_rng(1)
x = rng.standard_normal((200000, 64)).astype(np.float32)
xd = torch.from_numpy(x).cuda()
a = m.predict(xd).cpu().numpy()
for i in range(5):
    b = m.predict(xd).cpu().numpy()
    print("device rerun mismatches", int((a != b).sum()))
h = m.predict(x)
print("host vs device mismatches", int((a != h).sum()))
