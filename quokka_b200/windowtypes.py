"""Window and trigger descriptors -- same names, constructor arguments and meaning as pyquokka/windowtypes.py:6-145.
`aggregation_dict` maps a new column name to a SQL aggregate over the stream's columns, e.g.
{"avg_bid": "AVG(bid)", "max_disc_price": "MAX(price * (1 - discount))", "n": "count(*)"} (windowtypes.py:14-20); the
reference turns them into Polars expressions / DuckDB window SQL, here they are parsed by expr.py and evaluated by the
window kernels (csrc/window.cu)."""
from __future__ import annotations

import datetime

from . import expr as E

_OPS = {"sum", "min", "max", "count", "avg"}


class Window:
    def __init__(self, order_by, partition_by, aggregation_dict=None) -> None:
        assert order_by is not None, "order_by is not set"
        self.order_by = order_by
        assert partition_by is not None, "partition_by is not set, currently does not support unpartitioned windows"
        self.partition_by = partition_by
        self.aggregation_dict = dict(aggregation_dict or {})

    def add_aggregation(self, new_col, sql_agg) -> None:
        assert new_col not in self.aggregation_dict, "new_col already exists in aggregation_dict"
        self.aggregation_dict[new_col] = sql_agg

    def parsed(self):
        """[(new column, op, argument Node | None)]: one aggregate call per entry."""
        out = []
        for name, sql in self.aggregation_dict.items():
            n = E.parse(sql)
            if n.kind != "agg" or n.value not in _OPS:
                raise E.ExprError(f"window aggregation {sql!r} must be one SUM / AVG / MIN / MAX / COUNT call")
            arg = None if (not n.args or n.args[0].kind == "star") else n.args[0]
            if arg is None and n.value != "count":
                raise E.ExprError(f"{sql!r}: only COUNT takes *")
            out.append((name, n.value, arg))
        return out

    def get_required_cols(self):
        need = set()
        for _, _, arg in self.parsed():
            if arg is not None:
                need |= arg.columns()
        return need

    def get_new_cols(self):
        return list(self.aggregation_dict.keys())

    @staticmethod
    def ticks(val, unit: str | None = None) -> int:
        """A window length as an integer in the time column's own unit (ints as they are; timedeltas in `unit`)."""
        if isinstance(val, bool) or not isinstance(val, (int, datetime.timedelta)):
            raise Exception("Unsupported value type, only int and datetime.timedelta are supported for now for window hops and sizes")
        if isinstance(val, int):
            return val
        per_s = {"s": 1, "ms": 10 ** 3, "us": 10 ** 6, "ns": 10 ** 9}[unit or "us"]
        return int(round(val.total_seconds() * per_s))


class HoppingWindow(Window):
    def __init__(self, order_by, partition_by, hop, size, aggregation_dict=None) -> None:
        super().__init__(order_by, partition_by, aggregation_dict)
        self.hop = hop
        self.size = size


class TumblingWindow(HoppingWindow):
    def __init__(self, order_by, partition_by, size, aggregation_dict=None) -> None:
        super().__init__(order_by, partition_by, size, size, aggregation_dict)


class SlidingWindow(Window):
    def __init__(self, order_by, partition_by, size_before, aggregation_dict=None) -> None:
        super().__init__(order_by, partition_by, aggregation_dict)      # size_after is not supported by the reference either
        self.size_before = size_before


class SessionWindow(Window):
    def __init__(self, order_by, partition_by, timeout, aggregation_dict=None) -> None:
        super().__init__(order_by, partition_by, aggregation_dict)
        self.timeout = timeout


class Trigger:
    def __init__(self) -> None:
        pass


class OnEventTrigger(Trigger):
    def __init__(self) -> None:
        super().__init__()


class OnCompletionTrigger(Trigger):
    """Triggers on completion of the window (the delay is accepted and, as in the reference, not used)."""

    def __init__(self, delay=None) -> None:
        super().__init__()
        self.delay = delay
