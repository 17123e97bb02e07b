"""N > 1 host logic on CPU (gloo, world_size 2): chain sharding arithmetic and the single collective
(one all-gather of per-chain partials) -- checked against the oracle's merge rule."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def merge_pairs(parts):
    """Reference merge of (m, s, n, nnan) rows (what hb_accum_merge does on the device)."""
    out = np.zeros_like(parts[0])
    out[:, 0] = -np.inf
    for p in parts:
        for c in range(p.shape[0]):
            m0, s0 = out[c, 0], out[c, 1]
            m1, s1 = p[c, 0], p[c, 1]
            if s1 > 0:
                if s0 > 0:
                    M = max(m0, m1)
                    out[c, 1] = s0 * np.exp(m0 - M) + s1 * np.exp(m1 - M)
                    out[c, 0] = M
                else:
                    out[c, 0], out[c, 1] = m1, s1
            out[c, 2] += p[c, 2]
            out[c, 3] += p[c, 3]
    return out


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from harmonic_b200.evidence import _allgather

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nchains, per = 6, 50
    rng = np.random.default_rng(11)
    lnarg = rng.standard_normal(nchains * per) * 3.0  # same data on every rank
    lo, hi = rank * nchains // world, (rank + 1) * nchains // world
    part = np.zeros((nchains, 4))
    part[:, 0] = -np.inf
    for c in range(lo, hi):  # this rank folds only ITS chains
        seg = lnarg[c * per:(c + 1) * per]
        part[c] = [seg.max(), np.exp(seg - seg.max()).sum(), per, 0]
    glob = np.array([lnarg[lo * per:hi * per].sum(), (hi - lo) * per, 1, -1, 2, -2, 3, -3], dtype=np.float64)
    rows = _allgather(np.concatenate([part.reshape(-1), glob]), True)
    merged = merge_pairs([r[:-8].reshape(nchains, 4) for r in rows])
    L = merged[:, 0] + np.log(merged[:, 1])
    want = np.array([np.log(np.exp(lnarg[c * per:(c + 1) * per]).sum()) for c in range(nchains)])
    ok = bool(np.allclose(L, want, rtol=1e-12) and merged[:, 2].sum() == nchains * per
              and rows.shape == (world, nchains * 4 + 8) and np.isclose(rows[:, -8].sum(), lnarg.sum()))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_and_merge_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]
