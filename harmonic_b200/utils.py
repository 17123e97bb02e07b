"""Chain utilities on the path between ``Chains`` and ``Evidence.add_chains``.

``split_data`` mirrors harmonic/utils.py:129-200 (first chains -> training set, the rest -> test set, same
argument checks) but as index arithmetic: both halves are zero-copy views of the parent's arrays, so it also works
on device-resident chains (``Chains.from_arrays`` with CUDA tensors) without moving a byte.  The plotting helpers and
the cross-validation index helpers of the reference module are outside the accelerated path.
"""
from .chains import Chains


def split_data(chains, training_proportion: float = 0.5):
    """(chains_train, chains_test): the first ``int(nchains * training_proportion)`` chains and the remaining ones.

    Raises ValueError if ``training_proportion`` is not strictly between 0 and 1 or a side would be empty
    (harmonic/utils.py:163-172)."""
    if training_proportion <= 0.0 or training_proportion >= 1.0:
        raise ValueError("training_proportion must be strictly between 0 and 1.")
    nchains_train = int(chains.nchains * training_proportion)
    nchains_test = chains.nchains - nchains_train
    if nchains_train < 1:
        raise ValueError("nchains for training set must strictly greater than 0.")
    if nchains_test < 1:
        raise ValueError("nchains for test set must strictly greater than 0.")
    start = [int(s) for s in chains.start_indices]
    cut = start[nchains_train]
    train = Chains.from_arrays(chains.samples[start[0]:cut], chains.ln_posterior[start[0]:cut],
                               [s - start[0] for s in start[:nchains_train + 1]])
    test = Chains.from_arrays(chains.samples[cut:start[-1]], chains.ln_posterior[cut:start[-1]],
                              [s - cut for s in start[nchains_train:]])
    return train, test
