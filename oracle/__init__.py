"""CPU oracle for the FO1 hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it,
and only as the checker / the timed CPU baseline.  The product path
(``vlm-fo1_b200`` + ``libfo1.so``) never imports this package and fails loudly
if the CUDA library is missing.

Each module restates one stage of the reference (om-ai-lab/VLM-FO1 @ e6bef8d)
in plain torch-CPU / numpy fp32 and cites the reference file:line it follows.
Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 8c), so every restatement is pinned against outputs of the
reference's *own modules* executed in the build container
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``); ``tests/test_oracle_*.py``
replays those fixtures without needing ``/root/reference``.
"""
