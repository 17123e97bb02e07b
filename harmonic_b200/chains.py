"""Sample container: the input layout of the hot path.

Mirrors ``harmonic.chains.Chains`` (harmonic/chains.py:7-482): ``samples`` (N, ndim) row-major,
``ln_posterior`` (N,), ``start_indices`` (nchains+1), chains contiguous and possibly of unequal
length.  Differences, all additive:
* ``Chains.from_arrays`` wraps existing arrays WITHOUT copying -- including float32 and
  device-resident tensors (anything with ``__cuda_array_interface__``), which lets
  ``Evidence.add_chains`` run with no host<->device traffic;
* ``add_chains_2d_list`` slices by the given indexes (the reference multiplies the chain number by
  the current chain's length, harmonic/chains.py:180-192, which is only right for equal lengths).
"""
import copy

import numpy as np


class Chains:
    def __init__(self, ndim: int):
        if ndim < 1:
            raise ValueError("ndim must be greater than 0")
        self.nchains = 0
        self.start_indices = [0]
        self.ndim = ndim
        self.nsamples = 0
        self.samples = np.empty((0, self.ndim))
        self.ln_posterior = np.empty((0))
        self._start_np = None  # int64 image of start_indices (start_array)

    def start_array(self):
        """``start_indices`` as a contiguous int64 array (what the C ABI takes), cached while the list keeps
        its length (every method of this class that edits the list drops the cache)."""
        a = self._start_np
        if a is None or a.shape[0] != len(self.start_indices):
            a = self._start_np = np.ascontiguousarray(self.start_indices, dtype=np.int64)
        return a

    # ---------------------------------------------------------------- zero-copy constructor
    @classmethod
    def from_arrays(cls, samples, ln_posterior, start_indices):
        """Wrap (N, ndim) samples, (N,) ln_posterior and nchains+1 start indices as they are
        (numpy float32/float64 or CUDA tensors).  No copy, no dtype change."""
        ndim = int(samples.shape[1])
        ch = cls(ndim)
        sa = np.array(start_indices, dtype=np.int64).reshape(-1)  # private copy
        if sa.shape[0] < 1 or sa[0] != 0 or np.any(np.diff(sa) < 0):
            raise ValueError("start_indices must start at 0 and be non-decreasing")
        start = sa.tolist()
        if start[-1] != int(samples.shape[0]) or int(ln_posterior.shape[0]) != int(samples.shape[0]):
            raise ValueError("Length of sample and ln_posterior arrays do not match")
        ch.samples = samples
        ch.ln_posterior = ln_posterior
        ch.start_indices = start
        ch._start_np = sa
        ch.nchains = len(start) - 1
        ch.nsamples = start[-1]
        return ch

    def _is_host(self):
        return isinstance(self.samples, np.ndarray)

    # ---------------------------------------------------------------- harmonic/chains.py:35-71
    def add_chain(self, samples, ln_posterior):
        nsamples_in, ndim_in = samples.shape[0], samples.shape[1]
        if ndim_in != self.ndim:
            raise ValueError("ndim of new chain does not match previous chains")
        if nsamples_in != ln_posterior.shape[0]:
            raise ValueError("Length of sample and ln_posterior arrays do not match")
        if not self._is_host():
            raise ValueError("cannot append to a device-resident Chains; build it with from_arrays")
        self.samples = np.concatenate((self.samples, samples))
        self.ln_posterior = np.concatenate((self.ln_posterior, ln_posterior))
        self.nsamples += nsamples_in
        self.start_indices.append(self.nsamples)
        self.nchains += 1
        self._start_np = None

    def _add_many(self, samples, ln_posterior, bounds):
        """Append several chains with ONE concatenation (the reference re-concatenates per chain)."""
        if not self._is_host():
            raise ValueError("cannot append to a device-resident Chains; build it with from_arrays")
        self.samples = np.concatenate((self.samples, samples[bounds[0]:bounds[-1]]))
        self.ln_posterior = np.concatenate((self.ln_posterior, ln_posterior[bounds[0]:bounds[-1]]))
        base = self.nsamples - bounds[0]
        for b in bounds[1:]:
            self.start_indices.append(int(base + b))
        self.nsamples = self.start_indices[-1]
        self.nchains += len(bounds) - 1
        self._start_np = None

    # ---------------------------------------------------------------- harmonic/chains.py:73-130
    def add_chains_2d(self, samples, ln_posterior, nchains_in: int):
        if (samples.shape[0] % nchains_in) != 0:
            raise ValueError("The number of samples is not a multiple of the number of chains")
        if samples.shape[1] != self.ndim:
            raise ValueError("ndim of new chain does not match previous chains")
        if samples.shape[0] != ln_posterior.shape[0]:
            raise ValueError("Length of sample and ln_posterior arrays do not match")
        per = samples.shape[0] // nchains_in
        self._add_many(samples, ln_posterior, [i * per for i in range(nchains_in + 1)])

    # ---------------------------------------------------------------- harmonic/chains.py:132-194
    def add_chains_2d_list(self, samples, ln_posterior, nchains_in: int, chain_indexes):
        if samples.shape[1] != self.ndim:
            raise ValueError("ndim of new chain does not match previous chains")
        if len(chain_indexes) != nchains_in + 1:
            raise ValueError("Length of index list is not nchains_in + 1")
        if samples.shape[0] != ln_posterior.shape[0]:
            raise ValueError("Length of sample and ln_posterior arrays do not match")
        self._add_many(samples, ln_posterior, [int(i) for i in chain_indexes])

    # ---------------------------------------------------------------- harmonic/chains.py:196-236
    def add_chains_3d(self, samples, ln_posterior):
        nchains_in, nsamples_in, ndim_in = samples.shape
        if ndim_in != self.ndim:
            raise ValueError("ndim of new chain does not match previous chains")
        if nchains_in != ln_posterior.shape[0] or nsamples_in != ln_posterior.shape[1]:
            raise ValueError("Length of sample and ln_posterior arrays do not match")
        self._add_many(
            samples.reshape((nchains_in * nsamples_in, ndim_in)),
            ln_posterior.reshape((nchains_in * nsamples_in,)),
            [i * nsamples_in for i in range(nchains_in + 1)],
        )

    # ---------------------------------------------------------------- harmonic/chains.py:238-306
    def get_chain_indices(self, i: int):
        if i < 0:
            raise ValueError("Chain number must be positive")
        if i >= self.nchains:
            raise ValueError("Chain number is greater than nchains-1")
        return self.start_indices[i], self.start_indices[i + 1]

    def get_sub_chains(self, chains_wanted):
        out = Chains(self.ndim)
        for i in chains_wanted:
            if i < 0 or i >= self.nchains:
                raise ValueError("chains_wanted contains index out of bounds")
            a, b = self.start_indices[i], self.start_indices[i + 1]
            out.add_chain(self.samples[a:b, :], self.ln_posterior[a:b])
        return out

    # ---------------------------------------------------------------- harmonic/chains.py:308-370
    def add(self, other):
        if self.ndim != other.ndim:
            raise ValueError("ndim of other Chain object does not match this Chain object")
        self._add_many(other.samples, other.ln_posterior, list(other.start_indices))

    def shallowcopy(self):
        return copy.copy(self)

    def deepcopy(self):
        return copy.deepcopy(self)

    def nsamples_per_chain(self):
        return np.diff(np.asarray(self.start_indices))

    # ---------------------------------------------------------------- harmonic/chains.py:372-417
    def remove_burnin(self, nburn: int = 100):
        per = self.nsamples_per_chain()
        if np.any(per <= nburn):
            raise ValueError("nburn must be less than number of samples in chain")
        keep = np.concatenate([np.arange(a + nburn, b) for a, b in
                               zip(self.start_indices[:-1], self.start_indices[1:])])
        self.samples = self.samples[keep]
        self.ln_posterior = self.ln_posterior[keep]
        self.start_indices = [0] + [int(v) for v in np.cumsum(per - nburn)]
        self.nsamples = self.start_indices[-1]
        self._start_np = None

    # ---------------------------------------------------------------- harmonic/chains.py:419-482
    def split_into_blocks(self, nblocks: int = 100):
        """Split every chain into blocks (new chains), block counts proportional to chain length;
        the last block of a chain absorbs the remainder so no sample is dropped."""
        if nblocks <= self.nchains:
            raise ValueError("nblocks must be greater then number of chains")
        per = self.nsamples_per_chain()
        nb = np.round(nblocks * (per / self.nsamples)).astype(int)
        nb[nb == 0] = 1
        off = nblocks - int(nb.sum())
        if off != 0:
            k = int(np.argmax(nb))
            nb[k] += off
            if nb[k] < 1:
                raise ValueError("Adjusted block number for chain less than 1.")
        new_start = [0]
        for c in range(self.nchains):
            a, b = self.start_indices[c], self.start_indices[c + 1]
            step = (b - a) // int(nb[c])
            new_start.extend(int(a + k * step) for k in range(1, int(nb[c])))
            new_start.append(int(b))
        self.start_indices = new_start
        self.nchains = nblocks
        self._start_np = None
