"""CPU oracle for the NGU episodic-novelty hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the CPU
arm that is timed beside the GPU path.  ``ngu_b200`` never imports it.

Parity pinning: the reference (minoring/NGU) ships no tests, golden vectors or
known-answer fixtures for this path, so the oracle is
pinned against outputs of the unmodified reference module executed in the
build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).
"""
