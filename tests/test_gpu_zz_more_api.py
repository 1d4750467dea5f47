"""More DataStream programs against the real kernels (collected last: these cases were added after the
round's last GPU session and have so far only run against tests/cpu_shim.py and on 2-rank gloo)."""
import pytest

import api_cases as A

pytestmark = pytest.mark.gpu


@pytest.fixture
def qc():
    from quokka_b200.df import QuokkaContext
    return QuokkaContext()


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
