"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT.

CPU restatement of the reference's columnar hot path (marsupialtail/quokka @ 1caf62e):
scan+filter+project -> hash partition -> hash join / group-by -> top-k, and the sorted as-of join.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package.  The product (`quokka_b200/`) never does: it fails loudly when the
CUDA library is missing instead of falling back to anything in here.

Parity pinning: the reference has no native code and cannot be imported in
the build container (polars / duckdb / ray / redis / sqlglot are absent), so the oracle is pinned
against the reference's own fixtures instead:
  * apps/graph_api/tutorials/a.csv x b.csv          -> inner-join row count 10 118 (lesson2.1.py:64-68)
  * apps/time-series/test_trade2.csv x test_quote2.csv -> backward as-of by symbol (asof_join.py:6-18)
  * pyquokka/sql_utils.py:313-325,389-395           -> aggregate decomposition docstring examples
and cross-checked against pandas (merge / merge_asof) and pyarrow Acero (group_by) which are the
engines of the same family the reference delegates to.  See tests/test_oracle_golden.py.

What stays "parity unpinned": the reference could not be run (here or on the GPU box) and asserts nothing
about executor outputs beyond the fixtures above, so the Q1 / Q3 / Q5 numerics on the synthetic TPC-H-shaped
data are pinned to this oracle + pandas / Acero agreement only, not to an execution of the reference.
"""
