"""CPU oracle for the cVAE-ensemble hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``multi_modal_normative_modeling_b200``) may import
this package.  The only permitted importers are ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py``, and there only as the checker or the timed CPU baseline -- never as
the thing shipped.

Every function cites the reference file:line (relative to the upstream repository
soz223/multi_modal_normative_modeling) whose behaviour it restates.

Parity pin: the reference has no tests of its own.  The
oracle is pinned instead against outputs of the reference's ``cVAE.py`` imported
in the build container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``),
against the stored ``deviation/**`` CSV identities, and against the sklearn /
numpy known-answer vectors of SURVEY.md A.4.
"""
