"""Synthetic TPC-H-shaped / SIP-shaped tables generated directly in HBM (qk_synth_column).

Same counter-based hash as the CPU-side generator the tests and the CPU baseline use, so the device
columns are bit-identical to it (checked in tests/test_gpu_kernels.py).  Used by bench.py to hold SF-100
(600 037 902 lineitem rows, 22.8 GB of Q1 columns) resident without a host copy."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

T_ORDERS, T_LINEITEM, T_CUSTOMER, T_SUPPLIER, T_TRADES, T_QUOTES = 1, 2, 3, 4, 8, 9

# name -> (table, column id in qk_synth_column, torch dtype)
COLUMNS = {
    "o_orderkey": (T_ORDERS, 0, torch.int64), "o_custkey": (T_ORDERS, 1, torch.int64),
    "o_orderdate": (T_ORDERS, 2, torch.int32), "o_shippriority": (T_ORDERS, 3, torch.int32),
    "l_orderkey": (T_LINEITEM, 0, torch.int64), "l_partkey": (T_LINEITEM, 1, torch.int64),
    "l_suppkey": (T_LINEITEM, 2, torch.int64), "l_linenumber": (T_LINEITEM, 3, torch.int32),
    "l_quantity": (T_LINEITEM, 4, torch.float64), "l_extendedprice": (T_LINEITEM, 5, torch.float64),
    "l_discount": (T_LINEITEM, 6, torch.float64), "l_tax": (T_LINEITEM, 7, torch.float64),
    "l_returnflag": (T_LINEITEM, 8, torch.uint8), "l_linestatus": (T_LINEITEM, 9, torch.uint8),
    "l_shipdate": (T_LINEITEM, 10, torch.int32), "l_commitdate": (T_LINEITEM, 11, torch.int32),
    "l_receiptdate": (T_LINEITEM, 12, torch.int32),
    "c_custkey": (T_CUSTOMER, 0, torch.int64), "c_nationkey": (T_CUSTOMER, 1, torch.int64),
    "c_mktsegment": (T_CUSTOMER, 2, torch.uint8),
    "s_suppkey": (T_SUPPLIER, 0, torch.int64), "s_nationkey": (T_SUPPLIER, 1, torch.int64),
}
TICK_COLUMNS = {
    T_TRADES: {"time": (0, torch.int64), "symbol": (1, torch.int32), "size": (2, torch.float32), "price": (3, torch.float32)},
    T_QUOTES: {"time": (0, torch.int64), "symbol": (1, torch.int32), "bid": (2, torch.float32), "ask": (3, torch.float32),
               "bsize": (4, torch.float32), "asize": (5, torch.float32)},
}
DICTIONARIES = {"l_returnflag": ["A", "N", "R"], "l_linestatus": ["F", "O"],
                "c_mktsegment": ["AUTOMOBILE", "BUILDING", "FURNITURE", "HOUSEHOLD", "MACHINERY"]}
DATE_COLUMNS = {"o_orderdate", "l_shipdate", "l_commitdate", "l_receiptdate"}
# the two fixed dimension tables of TPC-H (spec 4.2.3): 25 nations in 5 regions
NATIONS = ["ALGERIA", "ARGENTINA", "BRAZIL", "CANADA", "EGYPT", "ETHIOPIA", "FRANCE", "GERMANY", "INDIA", "INDONESIA", "IRAN",
           "IRAQ", "JAPAN", "JORDAN", "KENYA", "MOROCCO", "MOZAMBIQUE", "PERU", "CHINA", "ROMANIA", "SAUDI ARABIA", "VIETNAM",
           "RUSSIA", "UNITED KINGDOM", "UNITED STATES"]
NATION_REGION = [0, 1, 1, 1, 4, 0, 3, 3, 2, 2, 4, 4, 2, 4, 0, 0, 0, 1, 2, 3, 4, 2, 3, 3, 1]
REGIONS = ["AFRICA", "AMERICA", "ASIA", "EUROPE", "MIDDLE EAST"]


def nation_table():
    import pyarrow as pa
    return pa.table({"n_nationkey": pa.array(range(25), pa.int64()), "n_name": pa.array(NATIONS),
                     "n_regionkey": pa.array(NATION_REGION, pa.int64())})


def region_table():
    import pyarrow as pa
    return pa.table({"r_regionkey": pa.array(range(5), pa.int64()), "r_name": pa.array(REGIONS)})
_TQ = {torch.uint8: L.QK_U8, torch.int32: L.QK_I32, torch.int64: L.QK_I64, torch.float32: L.QK_F32, torch.float64: L.QK_F64}


def sizes(sf: float) -> dict:
    """Row counts at scale factor sf (TPC-H cardinalities; lineitem exact at SF-1 and SF-100)."""
    n_lineitem = 6_001_215 if sf == 1 else 600_037_902 if sf == 100 else int(round(6_000_000 * sf))
    return dict(orders=int(round(1_500_000 * sf)), lineitem=n_lineitem, customer=int(round(150_000 * sf)),
                supplier=max(1, int(round(10_000 * sf))), part=max(1, int(round(200_000 * sf))))


def _sizes_arr(sf: float, n_symbols: int = 1, gap: int = 1):
    s = sizes(sf)
    return (C.c_int64 * 6)(s["orders"], s["customer"], s["supplier"], s["part"], n_symbols, gap)


def column(name: str, sf: float, lo: int = 0, hi: int | None = None, device="cuda") -> torch.Tensor:
    table, cid, dt = COLUMNS[name]
    total = sizes(sf)[{T_ORDERS: "orders", T_LINEITEM: "lineitem", T_CUSTOMER: "customer", T_SUPPLIER: "supplier"}[table]]
    hi = total if hi is None else hi
    out = torch.empty(hi - lo, dtype=dt, device=device)
    L.check(L.lib().qk_synth_column(table, cid, _sizes_arr(sf), lo, hi - lo, out.data_ptr(), _TQ[dt],
                                    torch.cuda.current_stream().cuda_stream), "qk_synth_column")
    return out


def table(names, sf: float, lo: int = 0, hi: int | None = None, device="cuda") -> dict:
    return {n: column(n, sf, lo, hi, device) for n in names}


def ticks(table_id: int, n: int, n_symbols: int, lo: int = 0, hi: int | None = None, gap: int = 1000,
          columns=None, device="cuda") -> dict:
    hi = n if hi is None else hi
    out = {}
    arr = (C.c_int64 * 6)(1, 1, 1, 1, n_symbols, gap)
    for name, (cid, dt) in TICK_COLUMNS[table_id].items():
        if columns is not None and name not in columns:
            continue
        t = torch.empty(hi - lo, dtype=dt, device=device)
        L.check(L.lib().qk_synth_column(table_id, cid, arr, lo, hi - lo, t.data_ptr(), _TQ[dt],
                                        torch.cuda.current_stream().cuda_stream), "qk_synth_column")
        out[name] = t
    return out
