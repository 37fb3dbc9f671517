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
network, not in the wheelhouse), so the reference *program* cannot be run, here or
on the GPU box -- but its *source* can: ``tests/golden/make_golden_ref.py`` imports the
unmodified ``/root/reference/caribou_hi`` package in the build container on eager
torch-float64 stand-ins for those three dependencies (``tests/golden/ref_shim/``) and
commits what the reference's own expressions evaluate to.

Pinning status
--------------
* PINNED to the reference's own source (``tests/golden/ref_*.npz``, ``refphys.npz``,
  ``ref_free_rvs.json``; ``tests/test_golden.py``): all 13 functions of ``physics.py``
  (values; torch-autograd VJPs of spin temperature, pseudo-Voigt profile and radiative
  transfer), and for all six model classes x {Gaussian, pseudo-Voigt}: every
  ``pm.Deterministic`` of ``add_priors``/``add_deterministics``, the predicted spectra
  (``mu`` of the observed Normals), the summed likelihood, its torch-autograd gradient
  with respect to every free RV and baseline coefficient, and the name / distribution
  family / parameters / shape of every free RV.  The oracle reproduces them to 1e-12
  (values) / 1e-11 (gradients); the CUDA path to the 1e-10 bar.
* PINNED against the reference's own known-answer tests (``tests/test_physics.py``) and
  every number printed by ``docs/source/notebooks/optimization.ipynb`` cell 5
  (``tests/test_oracle_kat.py``).
* NOT pinned (third-party behaviour the stand-ins restate rather than execute, so
  **parity is unpinned for these pieces only**): PyMC's prior log-densities, default
  transforms and Jacobians (``oracle/priors_ref.py``), PyMC's ``Normal.logp`` closed
  form, and bayes_spec's baseline normalisation.
* Third-party arithmetic restated from its published definition:
  PyMC ``Normal.logp`` (pymc>=5, notebook ran 5.22.0) and bayes_spec's polynomial
  baseline (bayes_spec>=1.7.8) -- see ``oracle/likelihood.py``.
"""
