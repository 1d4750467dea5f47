"""Device Parquet decode (qk_parquet_decode) against pyarrow's reading of the same files.  Collected last: the
decoder was written after the round's last GPU session -- its host walker and the per-value decode function have run
in the CPU container (tests/test_parquet_decode.py), the CUDA kernel around them has not run on hardware yet."""
import pyarrow.parquet as pq
import pytest
import torch

import api_cases as A
import parquet_cases as P

pytestmark = pytest.mark.gpu
CUDA = torch.device("cuda", 0)


@pytest.fixture
def qc():
    from quokka_b200.df import QuokkaContext
    return QuokkaContext()


@pytest.mark.parametrize("version,dict_on,page", P.LINEITEM_SHAPES)
def test_lineitem_shapes(tmp_path, version, dict_on, page): P.case_lineitem_shapes(tmp_path, CUDA, version, dict_on, page)
def test_required_columns_and_dictionary_fallback(tmp_path): P.case_required_and_fallback(tmp_path, CUDA)
def test_strings_share_codes_across_row_groups(tmp_path): P.case_strings_share_codes(tmp_path, CUDA)
def test_bit_widths(tmp_path): P.case_bit_widths(tmp_path, CUDA)
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_files(tmp_path, seed): P.case_random_files(tmp_path, CUDA, seed)
def test_spark_layout(tmp_path): P.case_spark_layout(tmp_path, CUDA)
def test_decimals(tmp_path):
    P.case_decimals(tmp_path, CUDA)
    P.case_decimals(tmp_path, CUDA, "snappy")
def test_parquet_device_api(qc, tmp_path): A.case_parquet_device(qc, tmp_path)


# ---- SNAPPY pages: qk_parquet_inflate (warp per page) + qk_parquet_page_runs (thread per page) + qk_parquet_decode
@pytest.mark.parametrize("version,dict_on,page", P.LINEITEM_SHAPES)
def test_snappy_lineitem_shapes(tmp_path, version, dict_on, page): P.case_lineitem_shapes(tmp_path, CUDA, version, dict_on, page, compression="snappy")
def test_snappy_fallback_strings_widths(tmp_path):
    P.case_required_and_fallback(tmp_path, CUDA, "snappy")
    P.case_strings_share_codes(tmp_path, CUDA, "snappy")
    P.case_bit_widths(tmp_path, CUDA, "snappy")
def test_snappy_streams(tmp_path): P.case_snappy_streams(tmp_path, CUDA)


# ---- ZSTD pages: lane 0 of a warp runs the sequential frame decoder (csrc/zstd_core.h) with the warp's workspace slot
@pytest.mark.parametrize("version,dict_on,page", P.LINEITEM_SHAPES)
def test_zstd_lineitem_shapes(tmp_path, version, dict_on, page): P.case_lineitem_shapes(tmp_path, CUDA, version, dict_on, page, compression="zstd")
def test_zstd_fallback_strings_widths(tmp_path):
    P.case_required_and_fallback(tmp_path, CUDA, "zstd")
    P.case_strings_share_codes(tmp_path, CUDA, "zstd")
    P.case_bit_widths(tmp_path, CUDA, "zstd")
def test_zstd_streams(tmp_path): P.case_snappy_streams(tmp_path, CUDA, "zstd")


# ---- GZIP pages: lane 0 runs the DEFLATE decoder (csrc/deflate_core.h)
@pytest.mark.parametrize("version,dict_on,page", P.LINEITEM_SHAPES[:2])
def test_gzip_lineitem_shapes(tmp_path, version, dict_on, page): P.case_lineitem_shapes(tmp_path, CUDA, version, dict_on, page, compression="gzip")
def test_gzip_streams(tmp_path): P.case_snappy_streams(tmp_path, CUDA, "gzip")


def test_sf1_lineitem_q1_columns(tmp_path):
    """6 M rows x the seven Q1 columns, Polars-style row groups of 100 000 (apps/convert.py:5-19): decoded columns are
    bit-identical to the generator's."""
    from oracle import tpch_gen as G
    from quokka_b200 import synth
    names = ["l_shipdate", "l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax"]
    li = G.gen_lineitem(1, columns=names)
    for compression in (None, "snappy", "zstd", "gzip"):
        _check_sf1(tmp_path, li, names, compression)


def _check_sf1(tmp_path, li, names, compression):
    from oracle import tpch_gen as G
    from quokka_b200 import synth
    path = str(tmp_path / "sf1.parquet")
    pq.write_table(G.to_arrow(li), path, compression=compression, row_group_size=100_000)
    d = P.read(path, CUDA, names)
    for n in names:
        exp = synth.column(n, 1)
        got = d[n].data
        if n in ("l_returnflag", "l_linestatus"):           # string column: compare through the dictionaries
            lut = torch.tensor([synth.DICTIONARIES[n].index(v) for v in d[n].dictionary], device=CUDA)
            got = lut[got.long()].to(exp.dtype)
        assert got.dtype == exp.dtype and torch.equal(got, exp), n
