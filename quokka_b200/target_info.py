"""TargetInfo and the partitioner descriptors -- same names, constructor arguments and meaning as
pyquokka/target_info.py:4-71 (the reference stores a sqlglot predicate; here it is an expr.Node or a
SQL string, compiled to a libqk program when the edge runs)."""
from __future__ import annotations

from . import expr as E


class Partitioner:
    """How the rows of a produced batch are assigned to the consumer's channels.  `describe()` is what
    explain() prints; the work itself is done by edge.apply_partitioner (qk_partition_plan + scatter)."""
    kind = "abstract"

    def describe(self) -> str:
        return self.kind

    def __str__(self) -> str:
        return self.describe()


class PassThroughPartitioner(Partitioner):
    kind = "pass_thru"            # the batch stays on the producing rank's channel


class BroadcastPartitioner(Partitioner):
    kind = "broadcast"            # every consumer channel receives the whole batch


class HashPartitioner(Partitioner):
    kind = "hash"

    def __init__(self, key: str) -> None:
        if not isinstance(key, str) or not key:
            raise TypeError("HashPartitioner needs a column name")
        self.key = key

    def describe(self) -> str:
        return self.key


class RangePartitioner(Partitioner):
    kind = "range"

    def __init__(self, key: str, total_range: int) -> None:
        if type(total_range) is not int:
            raise TypeError("total_range must be an int (the cardinality estimate of the key)")
        self.key, self.total_range = key, total_range

    def describe(self) -> str:
        return f"range partitioner on {self.key}, range estimate {self.total_range}"


class FunctionPartitioner(Partitioner):
    kind = "custom partitioner"

    def __init__(self, func) -> None:
        if not callable(func):
            raise TypeError("FunctionPartitioner needs a callable (table, source_channel, n) -> {channel: table}")
        self.func = func


class TargetInfo:
    """partitioner, predicate, projection (set of column names | None), batch_funcs (list of callables
    DeviceTable -> DeviceTable | None) -- pyquokka/target_info.py:4-30."""

    def __init__(self, partitioner, predicate, projection, batch_funcs: list, edge_ops=None, stable=False) -> None:
        from .edge import EdgeOps
        self.partitioner = partitioner
        self.predicate = predicate
        self.projection = projection
        self.batch_funcs = list(batch_funcs or [])
        self.stable = stable            # keep row order through the filter (ordered streams)
        self.bloom_key = None           # probe edge of a shuffled join: key column for the semi-join reduction
        self.bloom_source = None        # actor id of the join whose build keys make the filter
        self.bloom = None               # ops.Bloom, filled by the runtime once that build side is complete
        self.edge_ops = edge_ops.copy() if edge_ops is not None else EdgeOps()
        if predicate is not None:
            p = E.parse(predicate) if isinstance(predicate, str) else predicate
            self._pending_pred = p
        else:
            self._pending_pred = None
        self.lowered = False

    def and_predicate(self, predicate) -> None:
        p = E.parse(predicate) if isinstance(predicate, str) else predicate
        self._pending_pred = p if self._pending_pred is None else E.binop("and", self._pending_pred, p)

    def predicate_required_columns(self) -> set:
        return set() if self._pending_pred is None else self._pending_pred.columns()

    def append_batch_func(self, f) -> None:
        self.batch_funcs.append(f)

    def bind(self, raw_names):
        """Folds the predicate given at construction into the edge ops once the producer schema is known."""
        if self._pending_pred is not None:
            self.edge_ops.filter(self._pending_pred, raw_names)
            self._pending_pred = None

    def __str__(self) -> str:
        return ("partitioner: " + str(self.partitioner) + "\n\t  predicate: " +
                (self.edge_ops.pred.sql() if self.edge_ops.pred is not None else "TRUE") +
                "\n\t  projection: " + str(self.projection) + "\n\t batch_funcs: " + str(self.batch_funcs))
