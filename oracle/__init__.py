"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

A CPU restatement (numpy / torch-CPU / plain C) of the reference's algorithm for the
graph-embedding train + full-sort evaluation hot path of
chenzheng5555/tag-aware-recommendation.  It is the checker the CUDA path is
compared against; it is never the thing shipped or measured as the product.

Who may import / execute anything under ``oracle/``:
  * ``tests/``                              (parity checks),
  * ``__graft_entry__.smoke()``             (one small check on cuda:0),
  * ``bench.py``'s ``cpu_baseline`` leg and ``bench.py --impl reference`` (CPU timing arm).
The product package ``tag-aware-recommendation_b200/`` must not import it; the product fails
loudly when ``libtagrec_b200.so`` is missing instead of falling back to this code.

Pinning.  The reference has no tests, golden vectors or fixtures of its own
(SURVEY.md §4), so this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced by
importing the unmodified reference in the build container (``tests/golden/make_golden.py``,
committed, writes ``tests/golden/*.npz``).  ``tests/test_oracle_vs_golden.py`` checks every
function here against those fixtures (bit-exact for CSR structure/values, sampled triples,
top-K ids; <=1e-6 for fp32 tensors; <=1e-12 for metrics).  The arithmetic itself lives in
third-party packages the reference does not pin (torch / scipy / numpy / scikit-learn; the
effective pins are this image's versions: torch 2.11.0, numpy 2.3.5, scipy 1.18.1,
scikit-learn 1.9.0).

Each function cites the reference file:line it restates (paths relative to the reference root).
"""
