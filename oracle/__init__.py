"""CPU oracle for the SocialMediaModel hot path -- TEST INFRASTRUCTURE ONLY.

This package is a Float64 NumPy restatement of the reference's algorithm
(LuzieH/SocialMediaModel: src/abm.jl, src/pde.jl, src/setting.jl,
src/ensemble.jl).  It is the *checker* for the CUDA path; it is never the
product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
Nothing under ``socialmediamodel_b200/`` imports it, and the product raises
if the CUDA library is missing instead of falling back here.

PARITY UNPINNED: the reference ships no golden vectors (test/runtests.jl:1-4
and src/tests.jl:35-41 are assertion-free smoke tests; data/ is empty) and
Julia is not installed in this image, so the reference itself cannot be run
to produce fixtures.  The oracle is instead pinned to the closed-form
known-answer tests that follow from the reference source (SURVEY.md section
4, invariants 1-6) -- see tests/test_oracle_kats.py -- and every function
cites the reference file:line it restates.
"""
