"""oracle/ -- CPU restatement of the reference's two hot paths.  TEST INFRASTRUCTURE ONLY.

This package is the *checker*, never the product:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
  ``--impl reference`` legs may import, link or execute anything under ``oracle/``;
* ``stochvolmodels_b200`` never imports it and has no CPU fallback -- the product path raises
  when the CUDA library is missing.

Parity status: **pinned**.  Every function here is checked in ``tests/test_oracle_golden.py``
against fixtures under ``tests/golden/*.npz`` that were produced by running the unmodified
reference (ArturSepp/StochVolModels 2.2.0, ``/root/reference``) in the build container with
``tests/golden/make_golden.py`` (numpy 2.3.5 / scipy 1.18.1 / numba 0.65.0).

Third-party arithmetic on the path that is *not* under ``/root/reference``:
``scipy.integrate.solve_ivp(method='RK45')`` (scipy 1.18.1, ``scipy/integrate/_ivp/rk.py``,
``common.py``) -- restated in :func:`oracle.mgf.rk45_grid`; anchored on the reference's call
site ``pricers/logsv/affine_expansion.py:300-301`` and on the golden ``a_t1`` / ``log_mgf``
arrays (which are the outputs of that call site).

Modules
-------
``oracle.mc``   MC steppers (LogSV, Heston), time grid, payoffs, chain loops, Philox4x32-10 +
                Box-Muller restatement of the device RNG (numpy, exact uint32 arithmetic).
``oracle.mgf``  transform grid, legacy Simpson weights, affine-expansion ODE terms, SciPy-RK45
                controller clone, LogSV / Heston log-MGF grids, Fourier vanilla sums, chain loops.
``oracle/csrc`` plain-C (OpenMP) port of the MC chain pricer with the same Philox stream, used as
                the timed CPU baseline and as a large-N cross-check of the fp64 GPU path.
"""
