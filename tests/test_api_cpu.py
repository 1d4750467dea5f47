"""Host logic of the operator API on CPU: planner, expression compiler, edge functions, executor protocol
and driver, with the kernels replaced by tests/cpu_shim.py (numpy oracle).  The same cases run against the
real kernels in tests/test_gpu_api.py."""
import pytest

import api_cases as A
import cpu_shim


@pytest.fixture
def qc(monkeypatch):
    cpu_shim.install(monkeypatch)
    from quokka_b200.df import QuokkaContext
    return QuokkaContext()


def test_q1_sql(qc): A.case_q1_sql(qc)
def test_q1_dict_api(qc): A.case_q1_dict_api(qc)
def test_q3(qc): A.case_q3(qc)
def test_q5(qc): A.case_q5(qc)
def test_join_kinds(qc, golden_dir): A.case_join_kinds(qc, golden_dir)
@pytest.mark.parametrize("tag", ["0", "2"])
def test_asof(qc, golden_dir, tag): A.case_asof(qc, golden_dir, tag)
def test_asof_reference_result(qc, golden_dir): A.case_asof_reference_result(qc, golden_dir)
def test_windows(qc, golden_dir): A.case_windows(qc, golden_dir)
def test_parquet_q1(qc, tmp_path): A.case_parquet_q1(qc, tmp_path)
def test_parquet_device(qc, tmp_path): A.case_parquet_device(qc, tmp_path)
def test_misc_ops(qc): A.case_misc_ops(qc)
def test_scalar_aggs(qc): A.case_scalar_aggs(qc)
def test_string_key_join(qc): A.case_string_key_join(qc)
def test_agg_types(qc): A.case_agg_types(qc)
def test_q6_and_semi_anti(qc): A.case_q6_and_semi_anti(qc)
def test_q10_q18(qc): A.case_q10_q18(qc)
def test_q14_q17_q19(qc): A.case_q14_q17_q19(qc)
def test_q4_q12(qc): A.case_q4_q12(qc)
def test_q7_q8(qc): A.case_q7_q8(qc)
def test_q9_q11_q13(qc): A.case_q9_q11_q13(qc)
def test_q15_q16_q20_q22(qc): A.case_q15_q16_q20_q22(qc)
def test_q2_q21(qc): A.case_q2_q21(qc)
def test_string_funcs_and_nulls(qc): A.case_string_funcs_and_nulls(qc)
def test_case_like_extract(qc): A.case_case_like_extract(qc)
def test_csv_q1(qc, tmp_path): A.case_csv_q1(qc, tmp_path)
def test_custom_host_executor(qc): A.case_custom_host_executor(qc)
def test_union_clip_transform(qc, tmp_path): A.case_union_clip_transform(qc, tmp_path)
def test_asof_parquet(qc, golden_dir, tmp_path): A.case_asof_parquet(qc, golden_dir, tmp_path)
def test_count_distinct_and_writer(qc, tmp_path): A.case_count_distinct_and_writer(qc, tmp_path)
def test_executor_protocol(qc, golden_dir): A.case_executor_protocol(qc, golden_dir)


def test_context_needs_cuda_without_shim():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from quokka_b200 import _lib
    from quokka_b200.df import QuokkaContext
    with pytest.raises(_lib.QkError, match="no CPU execution path"):
        QuokkaContext()


def test_q3_plan_uses_transitive_semi_join_reduction(qc):
    """Q3's plan: customer is staged before orders, its Bloom filter is applied on the orders scan (o_custkey comes
    from the build side of lineitem x orders), and the filter of the reduced orders is applied on the lineitem scan."""
    qc.set_config("broadcast_rows", 100)          # at SF-0.01 every build side would otherwise be broadcast
    A.case_q3(qc)
    g = qc.last_graph
    blooms = {ti.bloom_key: (a.id, ti) for a in g.actors.values() for _, _, ti in a.targets if ti.bloom_key is not None}
    assert set(blooms) == {"o_custkey", "l_orderkey"}
    assert all(ti.bloom is not None for _, ti in blooms.values())
    stages = {a.id: a.stage for a in g.actors.values() if a.kind == "input"}
    by_first_col = {}
    for a in g.actors.values():
        if a.kind == "input":
            by_first_col[a.obj.table.column_names[0][:2]] = a.stage
    assert by_first_col["c_"] < by_first_col["o_"] < by_first_col["l_"]       # customer, then orders, then lineitem
    # switching the push-down off gives the reference's two-stage shape and the same answer
    qc.set_config("bloom_pushdown", False)
    A.case_q3(qc)
    g2 = qc.last_graph
    keys = {ti.bloom_key for a in g2.actors.values() for _, _, ti in a.targets if ti.bloom_key is not None}
    assert keys == {"l_orderkey", "o_custkey"}
    qc.set_config("bloom_join", False)
    A.case_q3(qc)
    assert not any(ti.bloom_key for a in qc.last_graph.actors.values() for _, _, ti in a.targets)


def test_explain_physical_plan_of_q3(qc, capsys):
    qc.set_config("broadcast_rows", 100)
    li, od, cu, *_ = A.tables()
    d = qc.from_arrow(li).join(qc.from_arrow(od), left_on="l_orderkey", right_on="o_orderkey")
    d = qc.from_arrow(cu).join(d, left_on="c_custkey", right_on="o_custkey")
    d = d.filter_sql("c_mktsegment = 'BUILDING' and o_orderdate < date '1995-03-15' and l_shipdate > date '1995-03-15'")
    g = d.groupby(["l_orderkey", "o_orderdate", "o_shippriority"]).agg_sql("sum(l_extendedprice * (1 - l_discount)) as revenue")
    g.top_k(["revenue", "o_orderdate"], 10, descending=[True, False]).explain(mode="physical")
    out = capsys.readouterr().out
    print(out)
    assert "stage -2" in out and "stage -1" in out and "stage 0" in out
    assert "where (c_mktsegment = 'BUILDING')" in out and "where (l_shipdate > date '1995-03-15')" in out
    assert "bloom(o_custkey in build keys" in out and "bloom(l_orderkey in build keys" in out
    assert "PartialAgg" in out and "SQLAggExecutor" in out and "ConcatThenSQLExecutor [single channel]" in out
