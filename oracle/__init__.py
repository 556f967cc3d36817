"""CPU oracle for the composed-score sampler hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The
product path (``composable_diffusion_models_b200``) never imports this package
and fails loudly when its CUDA library is missing.

What it is: a plain fp32 (optionally fp64) CPU restatement, in functional
torch ops, of the reference's algorithms for the path SURVEY.md section 8 names.
Each function cites the reference file:line it follows.

Parity pin: the reference (mo-rsa24/composable_diffusion_models) ships NO
tests, golden vectors or known-answer values, so the
oracle is pinned the other way the task allows: ``oracle/make_golden.py``
imports the *unmodified* reference modules from ``/root/reference`` (they run
on CPU here), feeds them seeded synthetic weights and injected noise, and
commits the reference's own outputs under ``tests/golden/``.
``tests/test_oracle_golden.py`` then checks every oracle function against
those fixtures.  The reference itself cannot travel to the GPU box, the
fixtures and this restatement do.
"""
