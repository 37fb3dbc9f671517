"""CPU oracle for the caribou_hi forward-model likelihood hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  ``caribou_hi_b200`` (the product) never
imports anything from here; the product path fails loudly when its CUDA library is
missing instead of falling back to this code.

What it is
----------
A literal CPU restatement (NumPy fp64 for values, torch-CPU fp64 + autograd for
gradients) of the arithmetic in the reference's ``caribou_hi/physics.py`` and of the
six ``add_priors``/``add_likelihood`` compositions, each function citing the
reference file:line it follows.  The reference itself is a PyTensor/PyMC program;
``pytensor``, ``pymc`` and ``bayes_spec`` are NOT installable in this image (no
network, not in the wheelhouse), so the reference cannot be imported or run, here or
on the GPU box.

Pinning status
--------------
* physics helper functions and the spin temperature are PINNED against the
  reference's own known-answer tests (``tests/test_physics.py``) and against every
  number printed by ``docs/source/notebooks/optimization.ipynb`` cell 5
  (``tests/test_oracle_kat.py``).
* The pseudo-Voigt profile, radiative transfer, logp, gradients and predicted spectra
  are NOT pinned by any value in the reference tree (its tests check shape / no-NaN
  only): for those outputs **parity is unpinned** -- the oracle
  is a line-by-line restatement cross-checked by two independent derivations
  (torch autograd vs. the analytic reverse pass, and NumPy vs. torch), nothing more.
* Third-party arithmetic restated from its published definition:
  PyMC ``Normal.logp`` (pymc>=5, notebook ran 5.22.0) and bayes_spec's polynomial
  baseline (bayes_spec>=1.7.8) -- see ``oracle/likelihood.py``.
"""
