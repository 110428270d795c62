"""TEST INFRASTRUCTURE ONLY.

CPU/fp32 restatement of the reference's grounding path (ekazakos/grove).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under
``grove_b200/`` imports it: the product path is the CUDA library and fails
loudly without it.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md §8c), so the oracle is pinned against *outputs of the reference's
own modules run in the build container* (``oracle/make_golden.py`` imports
``/root/reference`` and writes ``tests/golden/*.npz``); ``tests/test_oracle_golden.py``
replays them on every run.
"""
