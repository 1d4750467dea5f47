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
def test_parquet_q1(qc, tmp_path): A.case_parquet_q1(qc, tmp_path)
def test_misc_ops(qc): A.case_misc_ops(qc)
def test_executor_protocol(qc, golden_dir): A.case_executor_protocol(qc, golden_dir)


def test_context_needs_cuda_without_shim():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from quokka_b200 import _lib
    from quokka_b200.df import QuokkaContext
    with pytest.raises(_lib.QkError, match="no CPU execution path"):
        QuokkaContext()
