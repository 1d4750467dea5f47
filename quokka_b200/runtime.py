"""In-process, push-based driver: the part of the reference runtime that sits between operators --
TaskGraph construction (pyquokka/quokka_runtime.py:118-394), the two worker loops
(IOTaskManager.execute / ExecTaskManager.execute, pyquokka/core.py:484-965) and the shuffle
(TaskManager.push + the Flight mailbox, core.py:276-376, flight.py:44-264) -- re-thought for one process per
GPU (SPMD over torch.distributed):

  * every rank builds the same graph; an actor has one channel per rank (or a single channel on rank 0);
  * a produced batch is pushed through each outgoing edge (partition_fn), exchanged with an all-to-all of
    the partitioned column buffers (NCCL over NVLink; gloo on CPU for tests) and handed to the consumer's
    `execute` immediately -- partitions are consumed as they arrive, there is no global barrier between
    operators, only the stage rule "every build input before the first probe batch" (df.py:1558-1568);
  * empty partitions still take part in the exchange (the reference's `__empty__` sentinel,
    core.py:333-335) so all ranks issue the same sequence of collectives.

Fault tolerance (lineage, HBQ spill, Redis tables, the coordinator) is out of scope: SURVEY.md section 8.
"""
from __future__ import annotations

import copy
import os

import pyarrow as pa
import torch
import torch.distributed as dist

from . import _lib as L
from .columns import DeviceColumn, DeviceTable, as_device_table, concat_tables, unify_dictionaries
from .edge import partition_fn
from .placement_strategy import CustomChannelsStrategy, SingleChannelStrategy
from .target_info import PassThroughPartitioner, TargetInfo


def _default_device():
    from . import columns
    return columns.default_device()


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


# ------------------------------------------------------------------ the shuffle
_DT = {"torch.uint8": torch.uint8, "torch.int32": torch.int32, "torch.int64": torch.int64,
       "torch.float32": torch.float32, "torch.float64": torch.float64}


class Exchange:
    """All-to-all of partitioned column buffers.  Counts + schema travel in one small object all-gather,
    then ONE all_to_all_single per column (variable splits) moves the payload device-to-device."""

    def __init__(self, device):
        self.device = device
        self.bytes_sent = 0
        self.calls = 0

    def __call__(self, parts: dict, n_target: int, single_owner: int | None = None) -> list:
        """parts: {target_channel: DeviceTable}.  Target channel c lives on rank c (or on `single_owner`
        when the consumer has one channel).  Returns the tables received by this rank (one per source)."""
        w = world_size()
        if w == 1:
            return [p for _, p in sorted(parts.items()) if p is not None and len(p) > 0]
        me = rank()
        owner = (lambda ch: single_owner) if single_owner is not None else (lambda ch: ch)
        by_rank = {}
        for ch, p in parts.items():
            if p is not None and len(p) > 0:
                by_rank.setdefault(owner(ch), []).append(p)
        sends = {r: concat_tables(ps) for r, ps in by_rank.items()}
        any_t = next(iter(sends.values()), None)
        header = {"counts": [len(sends[r]) if r in sends else 0 for r in range(w)],
                  "schema": None if any_t is None else [(n, str(c.data.dtype), c.dictionary, c.arrow_type, c.valid is not None)
                                                        for n, c in any_t.columns.items()]}
        headers = [None] * w
        dist.all_gather_object(headers, header)
        self.calls += 1
        schema = next((h["schema"] for h in headers if h["schema"] is not None), None)
        if schema is None:
            return []
        recv_counts = [headers[s]["counts"][me] for s in range(w)]
        send_counts = header["counts"]
        # dictionaries: all ranks re-code onto the sorted union so codes mean the same everywhere
        out_cols = {}
        for i, (name, dt, _, atype, has_valid) in enumerate(schema):
            dtype = _DT[dt]
            dicts = [h["schema"][i][2] for h in headers if h["schema"] is not None]
            union = None
            if any(d is not None for d in dicts):
                union = sorted(set().union(*[set(d) for d in dicts if d is not None]))
            pieces = []
            for r in range(w):
                if r in sends:
                    c = sends[r][name]
                    if union is not None and c.dictionary != union:
                        c = unify_dictionaries([DeviceColumn(torch.zeros(0, dtype=c.data.dtype, device=c.data.device), union, c.arrow_type), c])[1][1]
                    pieces.append(c.data)
            send = torch.cat(pieces) if pieces else torch.zeros(0, dtype=dtype, device=self.device)
            recv = torch.empty(sum(recv_counts), dtype=dtype, device=self.device)
            dist.all_to_all_single(recv, send.contiguous(), recv_counts, send_counts)
            self.bytes_sent += (sum(send_counts) - send_counts[me]) * send.element_size()
            valid = None
            if has_valid:
                vp = [sends[r][name].valid for r in range(w) if r in sends]
                vs = torch.cat(vp) if vp else torch.zeros(0, dtype=torch.uint8, device=self.device)
                valid = torch.empty(sum(recv_counts), dtype=torch.uint8, device=self.device)
                dist.all_to_all_single(valid, vs.contiguous(), recv_counts, send_counts)
            out_cols[name] = (recv, union, atype, valid)
        tables, lo = [], 0
        for s in range(w):
            hi = lo + recv_counts[s]
            if hi > lo:
                tables.append(DeviceTable({n: DeviceColumn(d[lo:hi], u, a, None if v is None else v[lo:hi])
                                           for n, (d, u, a, v) in out_cols.items()}))
            lo = hi
        return tables


# ------------------------------------------------------------------ the graph
class _Actor:
    def __init__(self, aid, kind, obj, stage, single):
        self.id, self.kind, self.obj, self.stage, self.single = aid, kind, obj, stage, single
        self.targets = []          # (target actor id, stream_id, TargetInfo)
        self.sources = {}          # stream_id -> source actor id
        self.blocking = False
        self.done = False
        self.instance = None       # this rank's executor instance (one per (actor, channel): core.py:526-527)
        self.results = []
        self.ordered = False


class TaskGraph:
    """Same construction calls as the reference's TaskGraph (quokka_runtime.py:118,314,370,383,394)."""

    def __init__(self, context=None) -> None:
        self.context = context
        self.actors: dict[int, _Actor] = {}
        self.current_actor = 0
        self.device = getattr(context, "device", None) or _default_device()
        self.exchange = Exchange(self.device)
        self.profile = bool(os.environ.get("QK_PROFILE"))

    # -- construction
    def new_input_reader_node(self, reader, stage=0, placement_strategy=None):
        a = _Actor(self.current_actor, "input", reader, stage, False)
        a.ordered = hasattr(reader, "sorted_by")
        self.actors[a.id] = a
        self.current_actor += 1
        return a.id

    def _new_exec(self, streams, functionObject, stage, placement_strategy, source_target_info, blocking):
        single = isinstance(placement_strategy, SingleChannelStrategy)
        a = _Actor(self.current_actor, "exec", functionObject, stage, single)
        a.blocking = blocking
        for stream_id, src in streams.items():
            ti = source_target_info.get(stream_id) if source_target_info else None
            if ti is None:
                ti = TargetInfo(PassThroughPartitioner(), None, None, [])
            a.sources[stream_id] = src
            self.actors[src].targets.append((a.id, stream_id, ti))
        self.actors[a.id] = a
        self.current_actor += 1
        return a.id

    def new_non_blocking_node(self, streams, functionObject, stage=0, placement_strategy=CustomChannelsStrategy(1),
                              source_target_info={}, assume_sorted={}):
        return self._new_exec(streams, functionObject, stage, placement_strategy, source_target_info, False)

    def new_blocking_node(self, streams, functionObject, stage=0, placement_strategy=CustomChannelsStrategy(1),
                          source_target_info={}, transform_fn=None, assume_sorted={}):
        return self._new_exec(streams, functionObject, stage, placement_strategy, source_target_info, True)

    def create(self):
        for a in self.actors.values():
            if a.kind == "exec":
                a.instance = copy.deepcopy(a.obj)
            else:
                a.obj.device = self.device
                if self.context is not None and getattr(a.obj, "dictionaries", None) is None:
                    a.obj.dictionaries = self.context.dictionaries
        return self

    # -- execution
    def _owns(self, actor: _Actor) -> bool:
        return (not actor.single) or rank() == 0

    def _push(self, actor: _Actor, table):
        for tgt_id, stream_id, ti in actor.targets:
            tgt = self.actors[tgt_id]
            n = 1 if tgt.single else world_size()
            if table is not None and len(table.columns) > 0:
                ti.bind(table.column_names)
                parts = partition_fn(ti, table, rank(), n)
            else:
                parts = {}
            received = self.exchange(parts, n, single_owner=0 if tgt.single else None)
            out = None
            if self._owns(tgt) and received:
                out = tgt.instance.execute(received, stream_id, rank())
                out = as_device_table(out) if out is not None else None
            self._emit(tgt, out)

    def _emit(self, actor: _Actor, out):
        if actor.blocking:
            if out is not None and len(out) > 0:
                actor.results.append(out)
        else:
            self._push(actor, out)

    def _finish(self, actor: _Actor):
        actor.done = True
        for tgt_id, _, _ in actor.targets:
            tgt = self.actors[tgt_id]
            if tgt.done or not all(self.actors[s].done for s in tgt.sources.values()):
                continue
            out = None
            if self._owns(tgt):
                out = tgt.instance.done(rank())
                out = as_device_table(out) if out is not None else None
            self._emit(tgt, out)
            self._finish(tgt)

    def run(self):
        w, me = world_size(), rank()
        inputs = sorted((a for a in self.actors.values() if a.kind == "input"), key=lambda a: (a.stage, a.id))
        for a in inputs:
            state = a.obj.get_own_state(w)
            if a.ordered and w > 1:
                # ordered stream: batches enter in global order, one source channel after the other
                for ch in sorted(state):
                    for lineage in state[ch]:
                        batch = a.obj.execute(ch, lineage)[1] if ch == me else None
                        self._push(a, as_device_table(batch, self.device) if batch is not None else None)
            else:
                mine = state.get(me, [])
                rounds = len(mine)
                if w > 1:
                    t = torch.tensor([rounds], device=self.device, dtype=torch.int64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    rounds = int(t.item())
                for i in range(rounds):
                    batch = a.obj.execute(me, mine[i])[1] if i < len(mine) else None
                    self._push(a, as_device_table(batch, self.device) if batch is not None else None)
            self._finish(a)
        return self

    def results(self, actor_id) -> list:
        return self.actors[actor_id].results


def gather_to_all(tables: list, device) -> list:
    """collect(): every rank receives every rank's sink batches (quokka_dataset.py:107-117: the result is
    the unordered concatenation of all sink batches)."""
    w = world_size()
    local = concat_tables(tables) if tables else None
    if w == 1:
        return [local] if local is not None and len(local) > 0 else []
    ex = Exchange(device)
    out = []
    # broadcast partitioner semantics: send my batch to every rank
    parts = {r: local for r in range(w)} if local is not None and len(local) > 0 else {}
    return ex(parts, w)
