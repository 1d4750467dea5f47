"""Synthetic TPC-H-shaped tables, numpy side (TEST INFRASTRUCTURE -- see oracle/__init__.py).

No dbgen is available offline, so both arms (CPU oracle and CUDA product) generate the same
TPC-H-*shaped* data from a counter-based hash: every value is a pure function of
(table, column, row index), so any row range can be produced independently on the host (here) and on
the device (quokka_b200/csrc/synth.cu) and the two are BIT-IDENTICAL, fp64 measures included (they
are integer cents divided by 100.0 once, IEEE-exact on both sides).

Shapes follow SURVEY.md Appendix B / TPC-H spec 4.2: sparse order keys (8 of every 32), 1..7 lines
per order (fixed 7-order pattern, mean 4), custkeys never a multiple of 3, quantities 1..50,
discount 0.00..0.10, tax 0.00..0.08, order dates 1992-01-01..1998-08-02, ship = order + 1..121 days.
Column types are what apps/convert.py (reference) would infer: int64 keys, float64 measures,
date32 dates, dictionary codes (uint8) for the single-character flags / segments.
"""
from __future__ import annotations

import numpy as np

U64 = np.uint64
MASK32 = U64(0xFFFFFFFF)
GOLDEN = U64(0x9E3779B97F4A7C15)
SEED_BASE = 0x5EED0000

# table ids
T_ORDERS, T_LINEITEM, T_CUSTOMER, T_SUPPLIER, T_PART = 1, 2, 3, 4, 5
# column ids used as hash streams
(C_CUSTKEY, C_ORDERDATE, C_SUPPKEY, C_PARTKEY, C_QUANTITY, C_DISCOUNT, C_TAX, C_SHIPDELTA,
 C_COMMITDELTA, C_RECEIPTDELTA, C_RETFLAG, C_NATION, C_SEGMENT) = range(1, 14)
# host-only streams (columns the CUDA generator does not produce: used by API tests through from_arrow only)
C_SHIPMODE, C_SHIPINSTRUCT, C_BRAND, C_TYPE, C_SIZE, C_CONTAINER, C_PRIORITY = range(14, 21)
(C_PNAME, C_SCOMMENT, C_OCOMMENT, C_ACCTBAL, C_PHONE, C_AVAILQTY, C_SUPPLYCOST, C_ORDERSTATUS) = range(21, 29)
T_PARTSUPP = 6

DAY_1992_01_01 = 8035
ORDERDATE_SPAN = 2406          # 1992-01-01 .. 1998-08-02 inclusive
DAY_1995_06_17 = 9298
DAY_1998_09_02 = 10471         # date '1998-12-01' - interval '90' day  (Q1 cutoff)
DAY_1995_03_15 = 9204          # Q3 cutoff
DAY_1994_01_01 = 8766          # Q5 range start
DAY_1995_01_01 = 9131          # Q5 range end (exclusive)

RETURNFLAG_DICT = ["A", "N", "R"]
LINESTATUS_DICT = ["F", "O"]
SEGMENT_DICT = ["AUTOMOBILE", "BUILDING", "FURNITURE", "HOUSEHOLD", "MACHINERY"]
NATIONS = ["ALGERIA", "ARGENTINA", "BRAZIL", "CANADA", "EGYPT", "ETHIOPIA", "FRANCE", "GERMANY",
           "INDIA", "INDONESIA", "IRAN", "IRAQ", "JAPAN", "JORDAN", "KENYA", "MOROCCO", "MOZAMBIQUE",
           "PERU", "CHINA", "ROMANIA", "SAUDI ARABIA", "VIETNAM", "RUSSIA", "UNITED KINGDOM",
           "UNITED STATES"]
NATION_REGION = [0, 1, 1, 1, 4, 0, 3, 3, 2, 2, 4, 4, 2, 4, 0, 0, 0, 1, 2, 3, 4, 2, 3, 3, 1]
REGIONS = ["AFRICA", "AMERICA", "ASIA", "EUROPE", "MIDDLE EAST"]
PRIORITY_DICT = ["1-URGENT", "2-HIGH", "3-MEDIUM", "4-NOT SPECIFIED", "5-LOW"]
SHIPMODE_DICT = ["REG AIR", "AIR", "RAIL", "SHIP", "TRUCK", "MAIL", "FOB"]
SHIPINSTRUCT_DICT = ["DELIVER IN PERSON", "COLLECT COD", "NONE", "TAKE BACK RETURN"]
BRAND_DICT = [f"Brand#{m}{n}" for m in range(1, 6) for n in range(1, 6)]
TYPE_DICT = [f"{a} {b} {c}" for a in ("STANDARD", "SMALL", "MEDIUM", "LARGE", "ECONOMY", "PROMO")
             for b in ("ANODIZED", "BURNISHED", "PLATED", "POLISHED", "BRUSHED") for c in ("TIN", "NICKEL", "BRASS", "STEEL", "COPPER")]
CONTAINER_DICT = [f"{a} {b}" for a in ("SM", "LG", "MED", "JUMBO", "WRAP") for b in ("CASE", "BOX", "BAG", "JAR", "PKG", "PACK", "CAN", "DRUM")]

# host-only string columns for the rest of tpch.py's programs (Q2, Q9, Q11, Q13, Q15, Q16, Q20-Q22): small value lists in the
# spirit of TPC-H 4.2.2.13 (P_NAME = colour words) and 4.2.3 (comments carrying the phrases the queries look for)
COLOURS = ["almond", "blue", "chocolate", "forest", "green", "ivory", "khaki", "lemon", "maroon", "navy", "olive", "plum"]
PNAME_DICT = [f"{a} {b}" for a in COLOURS for b in COLOURS if a != b]                      # 132 two-colour names
SCOMMENT_DICT = ["carefully final deposits", "blithely even Customer accounts", "quick Customer slow Complaints nag", "ironic packages",
                 "Customer Recommends regular ideas", "furiously Customer bold Complaints", "pending requests haggle", "silent foxes"]
OCOMMENT_DICT = ["final deposits sleep", "special pending requests", "requests above the special ideas", "special packages about the requests",
                 "even instructions", "bold special theodolites", "quickly regular requests", "express accounts special furious requests nag",
                 "ironic pinto beans", "unusual asymptotes"]
ORDERSTATUS_DICT = ["F", "O", "P"]

# lines per order inside a block of 7 consecutive orders (sum 28 -> mean 4 lines/order)
LINES_PATTERN = [4, 1, 7, 3, 5, 2, 6]
PAT_ORDER = np.repeat(np.arange(7), LINES_PATTERN).astype(np.int64)          # 28 entries
PAT_LINE = np.concatenate([np.arange(c) for c in LINES_PATTERN]).astype(np.int64) + 1


def sizes(sf: float):
    """Row counts at scale factor `sf` (TPC-H cardinalities, lineitem exact at SF-1/SF-100)."""
    n_orders = int(round(1_500_000 * sf))
    n_customer = int(round(150_000 * sf))
    n_supplier = max(1, int(round(10_000 * sf)))
    n_part = max(1, int(round(200_000 * sf)))
    if sf == 1:
        n_lineitem = 6_001_215
    elif sf == 100:
        n_lineitem = 600_037_902
    else:
        n_lineitem = int(round(6_000_000 * sf))
    return dict(orders=n_orders, lineitem=n_lineitem, customer=n_customer,
                supplier=n_supplier, part=n_part)


def mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finalizer on uint64 arrays (wraps mod 2^64)."""
    x = x.astype(U64, copy=True)
    x ^= x >> U64(30)
    x *= U64(0xBF58476D1CE4E5B9)
    x ^= x >> U64(27)
    x *= U64(0x94D049BB133111EB)
    x ^= x >> U64(31)
    return x


def stream_seed(table: int, col: int) -> np.uint64:
    return mix64(np.array([SEED_BASE + table * 256 + col], dtype=U64))[0]


def hash_u64(table: int, col: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        return mix64((idx.astype(U64) + U64(1)) * GOLDEN + stream_seed(table, col))


def uniform(table: int, col: int, idx: np.ndarray, n: int) -> np.ndarray:
    """Integer uniform in [0, n) by multiply-shift of the top 32 hash bits (n < 2^32)."""
    h = hash_u64(table, col, idx)
    return (((h >> U64(32)) * U64(n)) >> U64(32)).astype(np.int64)


# ---------------------------------------------------------------- orders
def order_key(oidx: np.ndarray) -> np.ndarray:
    oidx = oidx.astype(np.int64)
    return (oidx // 8) * 32 + (oidx % 8) + 1


def order_date(oidx: np.ndarray) -> np.ndarray:
    return (DAY_1992_01_01 + uniform(T_ORDERS, C_ORDERDATE, oidx, ORDERDATE_SPAN)).astype(np.int32)


def order_custkey(oidx: np.ndarray, n_customer: int) -> np.ndarray:
    n_valid = n_customer - n_customer // 3
    u = uniform(T_ORDERS, C_CUSTKEY, oidx, n_valid)
    return u + u // 2 + 1          # 1,2,4,5,7,8,... never a multiple of 3


def gen_orders(sf: float, lo: int = 0, hi: int | None = None, columns=None) -> dict:
    sz = sizes(sf)
    hi = sz["orders"] if hi is None else hi
    j = np.arange(lo, hi, dtype=np.int64)
    cols = {
        "o_orderkey": lambda: order_key(j),
        "o_custkey": lambda: order_custkey(j, sz["customer"]),
        "o_orderdate": lambda: order_date(j),
        "o_shippriority": lambda: np.zeros(len(j), dtype=np.int32),
    }
    extra = {"o_orderpriority": lambda: uniform(T_ORDERS, C_PRIORITY, j, 5).astype(np.uint8),      # host only, on request
             "o_comment": lambda: uniform(T_ORDERS, C_OCOMMENT, j, len(OCOMMENT_DICT)).astype(np.uint8),
             "o_orderstatus": lambda: uniform(T_ORDERS, C_ORDERSTATUS, j, 3).astype(np.uint8)}
    return {c: (cols[c] if c in cols else extra[c])() for c in (columns or cols)}


# ---------------------------------------------------------------- lineitem
def line_order_index(i: np.ndarray, n_orders: int) -> np.ndarray:
    i = i.astype(np.int64)
    return ((i // 28) * 7 + PAT_ORDER[i % 28]) % n_orders


def gen_lineitem(sf: float, lo: int = 0, hi: int | None = None, columns=None) -> dict:
    sz = sizes(sf)
    hi = sz["lineitem"] if hi is None else hi
    i = np.arange(lo, hi, dtype=np.int64)
    oidx = line_order_index(i, sz["orders"])
    cache: dict = {}

    def memo(name, fn):
        if name not in cache:
            cache[name] = fn()
        return cache[name]

    def odate():
        return memo("odate", lambda: order_date(oidx).astype(np.int64))

    def shipdate():
        return memo("ship", lambda: odate() + 1 + uniform(T_LINEITEM, C_SHIPDELTA, i, 121))

    def receiptdate():
        return memo("rcpt", lambda: shipdate() + 1 + uniform(T_LINEITEM, C_RECEIPTDELTA, i, 30))

    def quantity():
        return memo("qty", lambda: 1 + uniform(T_LINEITEM, C_QUANTITY, i, 50))

    def partkey():
        return memo("pk", lambda: 1 + uniform(T_LINEITEM, C_PARTKEY, i, sz["part"]))

    def extprice():
        pk = partkey()
        retail_cents = 90000 + ((pk // 10) % 20001) + 100 * (pk % 1000)
        return (quantity() * retail_cents).astype(np.float64) / 100.0

    def returnflag():
        ra = uniform(T_LINEITEM, C_RETFLAG, i, 2)          # 0 -> 'A', 1 -> 'R'
        code = np.where(receiptdate() <= DAY_1995_06_17, ra * 2, 1)
        return code.astype(np.uint8)

    cols = {
        "l_orderkey": lambda: order_key(oidx),
        "l_partkey": partkey,
        "l_suppkey": lambda: 1 + uniform(T_LINEITEM, C_SUPPKEY, i, sz["supplier"]),
        "l_linenumber": lambda: PAT_LINE[i % 28].astype(np.int32),
        "l_quantity": lambda: quantity().astype(np.float64),
        "l_extendedprice": extprice,
        "l_discount": lambda: uniform(T_LINEITEM, C_DISCOUNT, i, 11).astype(np.float64) / 100.0,
        "l_tax": lambda: uniform(T_LINEITEM, C_TAX, i, 9).astype(np.float64) / 100.0,
        "l_returnflag": returnflag,
        "l_linestatus": lambda: (shipdate() > DAY_1995_06_17).astype(np.uint8),
        "l_shipdate": lambda: shipdate().astype(np.int32),
        "l_commitdate": lambda: (odate() + 30 + uniform(T_LINEITEM, C_COMMITDELTA, i, 61)).astype(np.int32),
        "l_receiptdate": lambda: receiptdate().astype(np.int32),
    }
    extra = {      # only on request (TPC-H spec 4.2.3 value lists); not part of the default column set nor of synth.cu
        "l_shipmode": lambda: uniform(T_LINEITEM, C_SHIPMODE, i, 7).astype(np.uint8),
        "l_shipinstruct": lambda: uniform(T_LINEITEM, C_SHIPINSTRUCT, i, 4).astype(np.uint8),
    }
    return {c: (cols[c] if c in cols else extra[c])() for c in (columns or cols)}


# ---------------------------------------------------------------- customer / supplier / dims
def gen_customer(sf: float, lo: int = 0, hi: int | None = None, columns=None) -> dict:
    sz = sizes(sf)
    hi = sz["customer"] if hi is None else hi
    j = np.arange(lo, hi, dtype=np.int64)
    cols = {
        "c_custkey": lambda: j + 1,
        "c_nationkey": lambda: uniform(T_CUSTOMER, C_NATION, j, 25),
        "c_mktsegment": lambda: uniform(T_CUSTOMER, C_SEGMENT, j, 5).astype(np.uint8),
    }

    def phone():           # country code = nation key + 10 (TPC-H 4.2.2.9), then three hashed groups
        nk, h = uniform(T_CUSTOMER, C_NATION, j, 25), uniform(T_CUSTOMER, C_PHONE, j, 10**9)
        return np.array([f"{a + 10}-{b // 10**6:03d}-{b // 1000 % 1000:03d}-{b % 10000:04d}" for a, b in zip(nk.tolist(), h.tolist())], dtype=object)
    extra = {"c_acctbal": lambda: (uniform(T_CUSTOMER, C_ACCTBAL, j, 1_100_000).astype(np.float64) - 99_999.0) / 100.0,   # -999.99 .. 9999.99
             "c_phone": phone}
    return {c: (cols[c] if c in cols else extra[c])() for c in (columns or cols)}


def gen_supplier(sf: float, lo: int = 0, hi: int | None = None, columns=None) -> dict:
    sz = sizes(sf)
    hi = sz["supplier"] if hi is None else hi
    j = np.arange(lo, hi, dtype=np.int64)
    cols = {
        "s_suppkey": lambda: j + 1,
        "s_nationkey": lambda: uniform(T_SUPPLIER, C_NATION, j, 25),
    }
    extra = {"s_name": lambda: np.array([f"Supplier#{k + 1:09d}" for k in j.tolist()], dtype=object),
             "s_acctbal": lambda: (uniform(T_SUPPLIER, C_ACCTBAL, j, 1_100_000).astype(np.float64) - 99_999.0) / 100.0,
             "s_comment": lambda: uniform(T_SUPPLIER, C_SCOMMENT, j, len(SCOMMENT_DICT)).astype(np.uint8)}
    return {c: (cols[c] if c in cols else extra[c])() for c in (columns or cols)}


def gen_part(sf: float, lo: int = 0, hi: int | None = None, columns=None) -> dict:
    """part (host only): the retail price is the formula lineitem's extended price is built from."""
    sz = sizes(sf)
    hi = sz["part"] if hi is None else hi
    j = np.arange(lo, hi, dtype=np.int64)
    pk = j + 1
    cols = {
        "p_partkey": lambda: pk,
        "p_brand": lambda: uniform(T_PART, C_BRAND, j, 25).astype(np.uint8),
        "p_type": lambda: uniform(T_PART, C_TYPE, j, 150).astype(np.uint8),
        "p_size": lambda: (1 + uniform(T_PART, C_SIZE, j, 50)).astype(np.int32),
        "p_container": lambda: uniform(T_PART, C_CONTAINER, j, 40).astype(np.uint8),
        "p_retailprice": lambda: (90000 + ((pk // 10) % 20001) + 100 * (pk % 1000)).astype(np.float64) / 100.0,
    }
    extra = {"p_name": lambda: uniform(T_PART, C_PNAME, j, len(PNAME_DICT)).astype(np.uint8)}
    return {c: (cols[c] if c in cols else extra[c])() for c in (columns or cols)}


def gen_partsupp(sf: float, columns=None) -> dict:
    """partsupp (host only): four suppliers per part, TPC-H 4.2.3's key formula."""
    sz = sizes(sf)
    S = sz["supplier"]
    j = np.arange(4 * sz["part"], dtype=np.int64)
    pk, i = j // 4 + 1, j % 4
    cols = {
        "ps_partkey": lambda: pk,
        "ps_suppkey": lambda: (pk + i * (S // 4 + (pk - 1) // S)) % S + 1,
        "ps_availqty": lambda: (1 + uniform(T_PARTSUPP, C_AVAILQTY, j, 9999)).astype(np.int32),
        "ps_supplycost": lambda: (100 + uniform(T_PARTSUPP, C_SUPPLYCOST, j, 99_901)).astype(np.float64) / 100.0,
    }
    return {c: cols[c]() for c in (columns or cols)}


def gen_nation() -> dict:
    return {"n_nationkey": np.arange(25, dtype=np.int64),
            "n_name": np.array(NATIONS, dtype=object),
            "n_regionkey": np.array(NATION_REGION, dtype=np.int64)}


def gen_region() -> dict:
    return {"r_regionkey": np.arange(5, dtype=np.int64), "r_name": np.array(REGIONS, dtype=object)}


DICTIONARIES = {"l_returnflag": RETURNFLAG_DICT, "l_linestatus": LINESTATUS_DICT,
                "c_mktsegment": SEGMENT_DICT, "o_orderpriority": PRIORITY_DICT, "l_shipmode": SHIPMODE_DICT, "l_shipinstruct": SHIPINSTRUCT_DICT,
                "p_brand": BRAND_DICT, "p_type": TYPE_DICT, "p_container": CONTAINER_DICT,
                "p_name": PNAME_DICT, "s_comment": SCOMMENT_DICT, "o_comment": OCOMMENT_DICT, "o_orderstatus": ORDERSTATUS_DICT}


def to_arrow(cols: dict):
    """numpy column dict -> pyarrow.Table with the reference's Parquet types
    (date32 dates, dictionary<uint8,string> flags)."""
    import pyarrow as pa
    arrays, names = [], []
    for name, v in cols.items():
        if name in DICTIONARIES:
            arr = pa.DictionaryArray.from_arrays(pa.array(v, type=pa.uint8()),
                                                 pa.array(DICTIONARIES[name], type=pa.string()))
        elif name.endswith("date"):
            arr = pa.array(v.astype(np.int32), type=pa.int32()).cast(pa.date32())
        elif v.dtype == object:
            arr = pa.array(list(v), type=pa.string())
        else:
            arr = pa.array(v)
        arrays.append(arr)
        names.append(name)
    return pa.table(arrays, names=names)


# ---------------------------------------------------------------- SIP-shaped trades / quotes
ASOF_SEED = 0xA50F
T_TRADES, T_QUOTES = 8, 9
C_TIME, C_SYMBOL, C_PAYLOAD0 = 20, 21, 22


def gen_ticks(table: int, n: int, n_symbols: int, lo: int = 0, hi: int | None = None,
              mean_gap_ns: int = 1000) -> dict:
    """Globally time-sorted tick stream: time[i] = i*gap + jitter (jitter < gap keeps it sorted,
    duplicates possible between neighbouring rows are avoided), symbol skewed (square-law)."""
    hi = n if hi is None else hi
    i = np.arange(lo, hi, dtype=np.int64)
    t = i * mean_gap_ns + uniform(table, C_TIME, i, mean_gap_ns)
    u = uniform(table, C_SYMBOL, i, 1 << 24).astype(np.float64) / float(1 << 24)
    sym = np.minimum((u * u * n_symbols).astype(np.int32), n_symbols - 1)
    out = {"time": t, "symbol": sym.astype(np.int32)}
    if table == T_TRADES:
        out["size"] = (uniform(table, C_PAYLOAD0, i, 10000).astype(np.float32)) / np.float32(100.0)
        out["price"] = (uniform(table, C_PAYLOAD0 + 1, i, 100000).astype(np.float32)) / np.float32(100.0)
    else:
        out["bid"] = (uniform(table, C_PAYLOAD0, i, 100000).astype(np.float32)) / np.float32(100.0)
        out["ask"] = (uniform(table, C_PAYLOAD0 + 1, i, 100000).astype(np.float32)) / np.float32(100.0)
        out["bsize"] = (uniform(table, C_PAYLOAD0 + 2, i, 1000).astype(np.float32)) / np.float32(10.0)
        out["asize"] = (uniform(table, C_PAYLOAD0 + 3, i, 1000).astype(np.float32)) / np.float32(10.0)
    return out
