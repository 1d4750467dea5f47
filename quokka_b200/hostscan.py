"""Host-resident (pinned) Arrow-layout columns streamed through the fused scan->filter->aggregate
kernel: the device half of what the reference does per input batch in IOTaskManager.execute ->
push -> partition_fn (pyquokka/core.py:866-965, :152-195) when the source is host memory.

Chunks are copied host->device on two copy streams into two staging buffers while the previous chunk is
being aggregated, so the PCIe copy (the bound of this path: 38 B/row over ~50 GB/s) overlaps the kernel."""
from __future__ import annotations

import torch

from . import _lib as L
from . import expr as E
from . import ops


class HostQ1Stream:
    def __init__(self, names, host_cols, pred_sql, agg_sqls, group_names, group_card, device, chunk_rows=1 << 24,
                 dictionaries=None):
        self.names = list(names)
        self.host = list(host_cols)
        for h in self.host:
            if h.is_cuda or not h.is_pinned():
                raise L.QkError("HostQ1Stream expects pinned host tensors")
        self.n = self.host[0].numel()
        self.device = device
        self.chunk = int(min(chunk_rows, max(self.n, 1)))
        sch = {c: E.ColumnInfo(i, ops.qk_dtype(t), (dictionaries or {}).get(c)) for i, (c, t) in enumerate(zip(self.names, self.host))}
        self.pred = E.compile_expr(E.parse(pred_sql), sch) if pred_sql else None
        self.aggs = [E.compile_expr(E.parse(a), sch) for a in agg_sqls]
        self.gcols = [sch[g].slot for g in group_names]
        self.state = ops.DenseAggState(group_card, [L.AGG_SUM] * len(self.aggs), device)
        self.staging = [[torch.empty(self.chunk, dtype=h.dtype, device=device) for h in self.host] for _ in range(2)]
        self.copy_streams = [torch.cuda.Stream(device=device) for _ in range(2)]
        self.compute = torch.cuda.Stream(device=device)
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.out_acc = torch.empty_like(self.state.acc, device="cpu").pin_memory()
        self.out_cnt = torch.empty_like(self.state.cnt, device="cpu").pin_memory()

    def run(self):
        st = self.state
        with torch.cuda.stream(self.compute):
            st.acc.zero_(); st.cnt.zero_()
        for b in range(2):
            self.consumed[b].record(self.compute)
        for i, lo in enumerate(range(0, self.n, self.chunk)):
            hi = min(lo + self.chunk, self.n)
            b = i & 1
            cs = self.copy_streams[b]
            cs.wait_event(self.consumed[b])                      # staging buffer b is free again
            with torch.cuda.stream(cs):
                for h, d in zip(self.host, self.staging[b]):
                    d[:hi - lo].copy_(h[lo:hi], non_blocking=True)
                self.copied[b].record(cs)
            self.compute.wait_event(self.copied[b])
            with torch.cuda.stream(self.compute):
                st.update([d[:hi - lo] for d in self.staging[b]], self.pred, self.gcols, self.aggs)
                self.consumed[b].record(self.compute)
        with torch.cuda.stream(self.compute):
            self.out_acc.copy_(st.acc, non_blocking=True)
            self.out_cnt.copy_(st.cnt, non_blocking=True)
        self.compute.synchronize()
        return {"acc": self.out_acc, "cnt": self.out_cnt,
                "bytes": self.out_acc.numel() * 8 + self.out_cnt.numel() * 8}
