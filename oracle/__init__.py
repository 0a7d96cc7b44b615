"""CPU oracle for the PINN residual-and-gradient hot path (TEST INFRASTRUCTURE ONLY).

This package restates, in float64 on the CPU, the algorithm of the reference
(slitvinov/PINN_for_quantum_wavefunction_surfaces):

* ``ref_autograd``  - the reference's own way of computing the loss: model forward
  followed by a Laplacian built from nested ``torch.autograd.grad`` calls
  (poc/main.py:82-120, 247-303, 341-355 and train.py:8-10, 41-57).  This is the
  "port" timed as the CPU baseline / reference arm by ``bench.py``.
* ``closed_form``   - the closed-form (forward-mode Taylor) statement of the same
  quantities, which is the specification of the CUDA
  kernels, plus a hand-written reverse sweep in numpy.
* ``layout``        - the packed parameter order.

Parity is PINNED: ``tests/golden/*.npz`` hold outputs of the real reference code
(loaded from /root/reference by ``tests/golden/make_golden.py``) and
``tests/test_oracle.py`` checks this oracle against them.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import this package.  The product package
``pinn_for_quantum_wavefunction_surfaces_b200`` never does; it fails loudly when its
CUDA library is missing instead of falling back to anything in here.
"""
