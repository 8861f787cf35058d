"""CPU oracle for the careless ELBO-gradient + Adam step.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy + torch float64 autograd) of the reference
algorithm in rs-station/careless v0.5.4.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only
as the checker or as the timed CPU baseline -- never as the product path.  The product
(``careless_b200``) fails loudly if its CUDA library is missing; it never routes here.

Pinning status
--------------
The reference cannot run in the build container: tensorflow 2.18, tensorflow-probability
0.25, tf_keras, reciprocalspaceship and gemmi are all absent and there is no network.
The arithmetic of the path lives in those third-party packages (pins:
``/root/reference/pyproject.toml:14-22``); their published algorithms are restated here
and every call site is cited ``file:line`` relative to ``/root/reference``.

* PINNED by the reference's own closed-form known-answer tests (see
  ``tests/test_oracle_kat.py``): Wilson pdfs (``tests/models/priors/test_wilson.py:13-29``),
  truncated-normal 4th moment vs ``scipy.stats.truncnorm``
  (``tests/models/merging/test_truncated_normal.py:29-42``), likelihood definitions vs
  ``scipy.stats`` (``tests/models/likelihoods/test_mono.py:12-51``), the Laue convolution
  identity (``tests/models/likelihoods/test_laue.py:11-36``), Rice / folded-normal
  log-densities vs ``scipy.stats.rice`` / ``foldnorm``.
* PARITY UNPINNED: the ELBO value, per-parameter gradients and the Adam trajectory.  The
  reference's tests only assert finiteness for these
  (``tests/models/merging/test_variational_mono.py:72-75``), so there is no stored vector
  to pin them to.  They are instead cross-checked internally: autograd vs finite
  differences, and the truncated-normal sampler gradient vs the implicit-function
  derivative of the inverse CDF.
"""
