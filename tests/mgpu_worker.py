"""torchrun worker: chain-sharded add_chains over NCCL must reproduce the single-GPU statistics."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import harmonic_b200 as hm  # noqa: E402
from tests import util  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for kind in ("realnvp64", "rqspline"):
        m = util.make_realnvp(64, seed=3, standardize=True, ws=0.3) if kind == "realnvp64" else \
            util.make_rqspline(3, 3, 8, (32, 32), seed=3, ws=0.3)
        m.device = local
        ch = util.gaussian_chains(m.ndim, 4 * world + 1, 600, seed=5, ragged=True)  # same on every rank
        bounds = np.linspace(0, ch.nchains, world + 1).astype(int)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        mine = ch.get_sub_chains(list(range(lo, hi)))
        ev = hm.Evidence(ch.nchains, m)
        for rep in range(2):  # two streaming calls
            ev.add_chains(mine, chain_offset=lo, comm=True)
        single = hm.Evidence(ch.nchains, m)
        for rep in range(2):
            single.add_chains(ch)
        for k in ("ln_evidence_inv", "ln_evidence_inv_var", "ln_evidence_inv_var_var", "n_eff", "lnargmax",
                  "lnargmin", "lnpredictmax", "lnprobmin"):
            a, b = float(getattr(ev, k)), float(getattr(single, k))
            if not np.isclose(a, b, rtol=1e-9, atol=1e-9):
                ok = False
                print("rank", rank, kind, k, a, b)
        if not np.allclose(ev.nsamples_per_chain, single.nsamples_per_chain):
            ok = False
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU_OK" if t.item() == 1.0 else "MGPU_FAIL")
    dist.destroy_process_group()
    return 0 if t.item() == 1.0 else 1


if __name__ == "__main__":
    sys.exit(main())
