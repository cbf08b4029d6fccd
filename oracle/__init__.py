"""CPU oracle for the LanczosNet spectral-convolution forward path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lanczosnetwork_b200/`` imports this
package.  The only callers are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and there only
as the checker / the timed CPU baseline, never as the product.

The oracle is a functional restatement (plain torch-CPU / numpy, dtype
parametrised so an fp64 run gives the rounding budget) of the reference
algorithm.  Every function cites the reference file:line it follows.  It is
pinned against the reference itself: ``tests/golden/make_golden.py`` imports
the real reference classes from ``/root/reference`` in the build container and
writes their inputs/outputs to ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays those through this oracle.
"""
from . import graph_prep, lanczos_oracle, segment_oracle  # noqa: F401
