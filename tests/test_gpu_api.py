"""The operator API against the real sm_100a kernels (same cases as tests/test_api_cpu.py)."""
import pytest

import api_cases as A

pytestmark = pytest.mark.gpu


@pytest.fixture
def qc():
    from quokka_b200.df import QuokkaContext
    return QuokkaContext()


def test_q1_sql(qc):
    A.case_q1_sql(qc)
    # the partial aggregate of Q1 must have gone through the fused TMA kernel, not the interpreter
    from quokka_b200 import ops
    assert ops.launch_count() > 0


def test_q1_dict_api(qc): A.case_q1_dict_api(qc)
def test_q3(qc): A.case_q3(qc)
def test_q5(qc): A.case_q5(qc)
def test_join_kinds(qc, golden_dir): A.case_join_kinds(qc, golden_dir)
@pytest.mark.parametrize("tag", ["0", "1", "2"])
def test_asof(qc, golden_dir, tag): A.case_asof(qc, golden_dir, tag)
def test_parquet_q1(qc, tmp_path): A.case_parquet_q1(qc, tmp_path)
def test_misc_ops(qc): A.case_misc_ops(qc)
def test_executor_protocol(qc, golden_dir): A.case_executor_protocol(qc, golden_dir)
