"""CPU oracle for the harmonic hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``harmonic_b200``) never imports it and has no CPU fallback.

Pinning status
--------------
* ``oracle.evidence`` (accumulate + finalise half: ``harmonic/evidence.py:125-189,
  215-358,438-497``) is PINNED: it reproduces the reference's own golden numbers
  (``tests/test_evidence.py:95-144,177-181,199-203,283-329,432-516``) and is checked
  against the *unmodified* reference ``evidence.py`` executed in the build container
  through a numpy stand-in for ``jax.numpy`` (``tests/golden/make_golden.py``).
* ``oracle.flows`` (flow half: ``harmonic/flows.py`` wiring + the tfp-jax / distrax
  arithmetic it calls) is **parity unpinned**: tfp / distrax / jax are not installed
  in this image and are unpinned, un-vendored dependencies of the reference
  (``requirements/requirements-core.txt:13-17``).  The restatement follows the
  published algorithms (see ``oracle/flows.py`` header) and is property-checked
  (densities integrate to one at several temperatures, NaN propagation), but no
  per-sample golden values from the real reference exist.
"""
