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
import pickle
import threading
import zlib

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops
from .columns import DeviceColumn, DeviceTable, as_device_table, concat_tables, unify_dictionaries
from .edge import Parts, partition_fn
from .placement_strategy import CustomChannelsStrategy, SingleChannelStrategy
from .target_info import BroadcastPartitioner, PassThroughPartitioner, TargetInfo


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


def _flat(t: torch.Tensor) -> torch.Tensor:
    """A dense 1-D buffer (a one-row slice of a 2-D state keeps its row stride and reports is_contiguous())."""
    if t.dim() == 1 and (t.numel() == 0 or t.stride(0) == 1):
        return t
    out = torch.empty(t.numel(), dtype=t.dtype, device=t.device)      # fresh storage: unit stride guaranteed
    out.copy_(t.reshape(-1))
    return out


class PeerMailbox:
    """A receive buffer per rank, mapped into every other rank's address space (CUDA IPC / fabric handles
    through torch symmetric memory), so that the partition kernel of a producer can store rows directly into
    the consumer's HBM over NVLink: compute (partition-scatter) and collective (all-to-all) in ONE kernel."""

    def __init__(self, device, nbytes: int):
        import torch.distributed._symmetric_memory as symm
        self.nbytes = int(nbytes)
        self.buf = symm.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]

    def barrier(self):
        self.hdl.barrier(channel=0)


class NativeLink:
    """The process-wide symmetric allocation behind the peer-memory shuffle (csrc/exchange.cu): `nchannels` channels,
    each a control block + a mailbox.  Channel 0 belongs to the driver thread, channel 1 + i to pipeline lane i, so
    that chunks in flight on different CUDA streams never share flags or mailbox space and every channel sees the
    same sequence of exchanges on every rank."""

    CTRL = L.XCHG_CTRL_BYTES

    def __init__(self, device, nchannels: int, mailbox_bytes: int):
        mailbox_bytes = (int(mailbox_bytes) + 4095) // 4096 * 4096
        head = (nchannels * self.CTRL + 65535) // 65536 * 65536
        self.box = PeerMailbox(device, head + nchannels * mailbox_bytes)
        self.box.buf[:head].zero_()
        torch.cuda.synchronize(device)
        self.box.barrier()                                   # every control block is zero before anybody posts
        torch.cuda.synchronize(device)
        w, me = world_size(), rank()
        tmo = int(float(os.environ.get("QK_XCHG_TIMEOUT_S", "120")) * 1000)     # a rank may lag by a first-touch allocation of tens of GB
        self.channels = [ops.XchgChannel(w, me, [p + i * self.CTRL for p in self.box.ptrs],
                                         [p + head + i * mailbox_bytes for p in self.box.ptrs], mailbox_bytes, device, tmo)
                         for i in range(nchannels)]
        self.mailbox_bytes = mailbox_bytes


_lane_streams = {}                   # device index -> the lanes' CUDA streams (process-wide: their allocator pools stay warm)


def lane_streams(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _lane_streams:
        _lane_streams[key] = [torch.cuda.Stream(device) for _ in range(N_LANES)]
    return _lane_streams[key]


_link = {"obj": None, "failed": False}
_lane = threading.local()            # .index: pipeline lane of the calling thread (absent = the driver thread)
N_LANES = max(1, int(os.environ.get("QK_LANES", "2")))


def lane_channel() -> int:
    return 1 + getattr(_lane, "index", -1)


def native_link(device):
    """The peer-memory link, or None (no CUDA / single rank / gloo / QK_P2P=0 / symmetric memory unavailable).  Creating
    it is a collective; all ranks agree on the outcome (a rank-local failure makes every rank fall back to NCCL)."""
    if _link["failed"] or device.type != "cuda" or world_size() == 1 or os.environ.get("QK_P2P", "1") == "0":
        return None
    if dist.get_backend() != "nccl":
        return None
    if _link["obj"] is None:
        ok, err = 1, None
        try:
            obj = NativeLink(device, 1 + N_LANES, int(float(os.environ.get("QK_MAILBOX_MB", "2048")) * (1 << 20)))
        except Exception as e:                                   # no P2P / fabric support on this box, or out of memory
            ok, err, obj = 0, e, None
        flag = torch.tensor([ok], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            _link["failed"] = True
            if rank() == 0:
                print(f"[quokka_b200] peer-memory shuffle unavailable ({type(err).__name__ if err else 'a peer failed'}: {err}); using NCCL all-to-all", flush=True)
            return None
        _link["obj"] = obj
    return _link["obj"]


_mailbox = {"obj": None, "failed": False}


def peer_mailbox(device, nbytes: int):
    """Round 1's mailbox for qk_scatter_peer, kept for the collective (NCCL) path's opt-in QK_P2P_LEGACY=1."""
    if os.environ.get("QK_P2P_LEGACY", "0") != "1" or _mailbox["failed"] or device.type != "cuda" or world_size() == 1:
        return None
    if dist.get_backend() != "nccl":
        return None
    if _mailbox["obj"] is None:
        ok = 1
        try:
            _mailbox["obj"] = PeerMailbox(device, nbytes)
        except Exception:
            ok = 0
        flag = torch.tensor([ok], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)              # a rank-local failure must not leave the others in rendezvous
        if int(flag.item()) == 0:
            _mailbox["failed"], _mailbox["obj"] = True, None
            return None
    return _mailbox["obj"]


class Exchange:
    """All-to-all of partitioned column buffers (the reference's push -> Flight do_put / do_get,
    core.py:276-376, flight.py:44-264).

    Per call: one small all-gather of the per-destination row counts (+ schema checksums), then one
    variable-split all-to-all per column (grouped ncclSend / ncclRecv) that moves every column slice straight
    out of the partition kernel's output -- rows of one destination are already contiguous there, so nothing
    is packed or concatenated.  Column names / dtypes / dictionaries
    travel (pickled) only when an edge is first used or a dictionary changes."""

    def __init__(self, device, mailbox_bytes: int = 6 << 30):
        self.device = device
        self.bytes_sent = 0
        self.calls = 0
        self.peer_calls = 0
        self.schemas = {}          # edge key -> [(name, dtype str, dictionary, arrow type, has_valid)]
        self.recv_totals = {}      # edge key -> rows every rank has RECEIVED on that edge so far (known to all ranks from the counts)
        self.mailbox_bytes = mailbox_bytes

    def _peer_path(self, parts, allmeta, schema, w, me):
        """Hash-partitioned rows go straight into the receivers' mailboxes (qk_scatter_peer): every rank knows
        the whole counts matrix, hence where its rows start inside each receiver's columns.  Returns None (all
        ranks alike) when the payload does not fit the mailbox."""
        counts = allmeta[:, :w]                                   # counts[s][d]
        n_recv = [int(counts[:, d].sum()) for d in range(w)]
        maxrecv = max(n_recv)
        src = next(r for r in range(w) if int(allmeta[r, w + 2]) > 0)
        ncols = int(allmeta[src, w + 2])
        col_w = [int(allmeta[src, w + 3 + i]) & 255 for i in range(ncols)]
        if any(int(allmeta[r, w + 3 + i]) & 256 for r in range(w) for i in range(ncols)):
            return None                                           # validity masks: NCCL path
        bases, off = [], 0
        for wd in col_w:
            bases.append(off)
            off += (maxrecv * wd + 255) // 256 * 256
        mb = peer_mailbox(self.device, self.mailbox_bytes)
        if off > mb.nbytes:
            return None
        table, dest, doffs = parts.pending if parts is not None and parts.pending is not None else (None, None, None)
        mb.barrier()                                              # every rank is done reading the previous contents
        if table is not None and len(table) > 0:
            cols_ = []
            for i in range(ncols):
                c = table[schema[i][0]]
                union = schema[i][2]
                if union is not None and c.dictionary != union:
                    c = unify_dictionaries([DeviceColumn(torch.zeros(0, dtype=c.data.dtype, device=c.data.device), union, c.arrow_type), c])[1][1]
                cols_.append(_flat(c.data))
            row_off = [int(counts[:me, d].sum()) for d in range(w)]
            ptrs = [[mb.ptrs[d] + bases[i] for i in range(ncols)] for d in range(w)]
            ops.scatter_peer(cols_, dest, doffs, ptrs, row_off)
            self.bytes_sent += sum((int(counts[me].sum()) - int(counts[me, me])) * wd for wd in col_w)
        mb.barrier()                                              # every row addressed to me has landed
        self.peer_calls += 1
        if n_recv[me] == 0 or schema is None:
            return []
        out = {}
        for i, (name, dt, union, atype, _) in enumerate(schema):
            view = mb.buf[bases[i]: bases[i] + n_recv[me] * col_w[i]].view(_DT[dt])
            out[name] = DeviceColumn(view.clone(), union, atype)   # the mailbox is reused by the next exchange
        return [DeviceTable(out)]

    @staticmethod
    def _schema_of(t: DeviceTable):
        return [(n, str(c.data.dtype), c.dictionary, c.arrow_type, c.valid is not None) for n, c in t.columns.items()]

    # ------------------------------------------------------------------ the peer-memory path (csrc/exchange.cu)
    W_COUNTS, W_HASH, W_CACHED, W_NCOLS, W_WIDTHS, MAXC = 0, 16, 17, 18, 20, 24

    def allgather_words(self, words: list) -> list:
        """[[words of rank 0], [words of rank 1], ...]: a handful of integers from every rank.  On the peer-memory link
        this is one meta round (one tiny kernel + one host wait); otherwise an all-gather."""
        w = world_size()
        if w == 1:
            return [list(words)]
        link = native_link(self.device)
        if link is not None:
            m = link.channels[lane_channel()].meta([0] * 16 + [int(x) for x in words])
            return [[int(x) for x in m[r, 16:16 + len(words)]] for r in range(w)]
        t = torch.tensor([int(x) for x in words], dtype=torch.int64, device=self.device)
        out = torch.empty(w * len(words), dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(out, t)
        return out.cpu().view(w, len(words)).tolist()

    def _share_objects(self, ch, obj):
        """Every rank's `obj` (pickled) to every rank over the link: used only when schemas disagree."""
        w, me = ch.world, ch.rank
        blob = pickle.dumps(obj)
        buf = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(self.device)
        m = ch.meta([len(blob)] * w)
        sizes = [int(m[s_, 0]) for s_ in range(w)]
        if sum(sizes) > ch.mailbox_bytes:
            raise L.QkError("exchange: schema negotiation does not fit the mailbox")
        off = [sum(sizes[:s_]) for s_ in range(w)]
        ch.push([buf], [0] * w, [len(blob)] * w, [[off[me]] for _ in range(w)])
        out = torch.empty(sum(sizes), dtype=torch.uint8, device=self.device)
        ch.recv([0], [out])
        host = out.cpu().numpy().tobytes()
        return [pickle.loads(host[off[s_]:off[s_] + sizes[s_]]) for s_ in range(w)]

    def _native_call(self, link, parts, single_owner, edge_key) -> list:
        """One exchange over the link.  Host round trips: ONE (the meta matrix); collectives: none."""
        ch = link.channels[lane_channel()]
        w, me = ch.world, ch.rank
        # ---- what do I send?  mode "scatter": an unmaterialised hash partition (dest / device offsets);
        #      mode "ranges": a table + one contiguous row range per destination rank
        table, dest, doffs, lo, hi = None, None, None, [0] * w, [0] * w
        if isinstance(parts, Parts):
            if single_owner is None and parts.pending is not None and parts.doffs is not None and parts.doffs.numel() == w + 1:
                table, dest, doffs = parts.pending
                if len(table) == 0:
                    table = dest = doffs = None
            else:
                t = parts.table
                if t is not None and len(t) > 0:
                    table = t
                    if single_owner is not None:
                        hi[single_owner] = len(t)
                    else:
                        offs = parts.offsets
                        for d in range(min(w, len(offs) - 1)):
                            lo[d], hi[d] = offs[d], offs[d + 1]
        else:
            owner = (lambda c_: single_owner) if single_owner is not None else (lambda c_: c_)
            by_rank = {}
            for c_, p_ in sorted(parts.items()):
                if p_ is not None and len(p_) > 0:
                    by_rank.setdefault(owner(c_), []).append(p_)
            flat = [p_ for r in sorted(by_rank) for p_ in by_rank[r]]
            if flat and all(p_ is flat[0] for p_ in flat) and all(len(v) == 1 for v in by_rank.values()):
                table = flat[0]                                   # broadcast / single owner: the same rows to every destination
                for r in by_rank:
                    hi[r] = len(table)
            elif flat:
                table = concat_tables(flat)
                pos = 0
                for r in sorted(by_rank):
                    n_r = sum(len(p_) for p_ in by_rank[r])
                    lo[r], hi[r] = pos, pos + n_r
                    pos += n_r
        self.calls += 1
        # ---- meta round: counts (from the device in scatter mode), schema checksums, column widths
        cached = self.schemas.get(edge_key) if edge_key is not None else None
        mine = self._schema_of(table) if table is not None else None
        if mine is not None and len(mine) > self.MAXC:
            raise L.QkError(f"exchange: more than {self.MAXC} columns on one edge")
        words = [0] * ops.XchgChannel.META
        for d in range(w):
            words[d] = hi[d] - lo[d]
        words[self.W_HASH] = (zlib.crc32(repr(mine).encode()) | 1) if mine is not None else 0
        words[self.W_CACHED] = (zlib.crc32(repr(cached).encode()) | 1) if cached is not None else 0
        words[self.W_NCOLS] = len(mine) if mine is not None else 0
        if mine is not None:
            for i, (_, dt, _, _, hv) in enumerate(mine):
                words[self.W_WIDTHS + i] = _DT[dt].itemsize | (256 if hv else 0)
        m = ch.meta(words, doffs if dest is not None else None)
        counts = m[:, :w]                                               # counts[s][d]
        if edge_key is not None:
            tot = self.recv_totals.setdefault(edge_key, [0] * w)
            for d in range(w):
                tot[d] += int(counts[:, d].sum())
        if dest is not None:
            offs = [0]
            for d in range(w):
                offs.append(offs[-1] + int(counts[me, d]))
            parts.offsets = offs                                        # my own counts came back with everybody's
        if int(counts.sum()) == 0:
            return []
        # ---- schema agreement (objects travel only when the checksums disagree)
        hashes = set(int(x) for x in m[:, self.W_HASH] if x != 0)
        negotiate = len(hashes) != 1
        if not negotiate:
            common = next(iter(hashes))
            for r in range(w):
                if int(m[r, self.W_HASH]) == 0 and int(counts[:, r].sum()) > 0 and int(m[r, self.W_CACHED]) != common:
                    negotiate = True
        if negotiate:
            known = [h for h in self._share_objects(ch, mine) if h is not None]
            schema = []
            for i, (name, dt, _, atype, _) in enumerate(known[0]):
                dicts = [h[i][2] for h in known]
                union = sorted(set().union(*[set(d_) for d_ in dicts if d_ is not None])) if any(d_ is not None for d_ in dicts) else None
                schema.append((name, dt, union, atype, any(h[i][4] for h in known)))
            ch.meta([0])        # the objects used this channel's epoch and mailbox: the payload gets a fresh epoch (and the barrier)
        else:
            schema = mine if mine is not None else (cached if words[self.W_CACHED] == next(iter(hashes)) else None)
        if edge_key is not None and schema is not None:
            self.schemas[edge_key] = schema
        src = next(r for r in range(w) if int(m[r, self.W_NCOLS]) > 0)
        ncols = int(m[src, self.W_NCOLS])
        col_w = [int(m[src, self.W_WIDTHS + i]) & 255 for i in range(ncols)]
        col_valid = [any(int(m[r, self.W_WIDTHS + i]) & 256 for r in range(w)) for i in range(ncols)]
        widths = col_w + [1 for v in col_valid if v]                     # validity masks travel as extra uint8 columns
        n_recv = [int(counts[:, d].sum()) for d in range(w)]
        bases = []                                                       # bases[d][c]: column c in rank d's mailbox
        for d in range(w):
            row, off = [], 0
            for wd in widths:
                row.append(off)
                off += (n_recv[d] * wd + 255) // 256 * 256
            bases.append(row)
            if off > ch.mailbox_bytes:
                if edge_key is not None:                       # the rounds below count their own rows
                    tot = self.recv_totals[edge_key]
                    for d2 in range(w):
                        tot[d2] -= int(counts[:, d2].sum())
                return self._native_oversize(link, parts, table, lo, hi, single_owner, edge_key, off, ch.mailbox_bytes)
        # ---- payload
        send = []
        if table is not None:
            valids = []
            for i in range(ncols):
                c = table[schema[i][0]]
                union = schema[i][2]
                if union is not None and c.dictionary != union:
                    c = unify_dictionaries([DeviceColumn(torch.zeros(0, dtype=c.data.dtype, device=c.data.device), union, c.arrow_type), c])[1][1]
                d_ = _flat(c.data)
                send.append(d_.view(torch.uint8) if d_.dtype == torch.bool else d_)
                if col_valid[i]:
                    valids.append(c.valid if c.valid is not None else torch.ones(len(c), dtype=torch.uint8, device=self.device))
            send += valids
        row_off = [int(counts[:me, d].sum()) for d in range(w)]
        dst_off = [[bases[d][c] + row_off[d] * widths[c] for c in range(len(send))] for d in range(w)]
        if dest is not None and table is not None:
            ch.push_scatter(send, dest, doffs, dst_off)
        else:
            ch.push(send, lo, hi, dst_off)
        self.peer_calls += 1
        self.bytes_sent += (int(counts[me].sum()) - int(counts[me, me])) * sum(widths)
        if n_recv[me] == 0 or schema is None:
            ch.recv([], [])
            return []
        outs = [torch.empty(n_recv[me], dtype=_DT[schema[i][1]], device=self.device) for i in range(ncols)]
        outs += [torch.empty(n_recv[me], dtype=torch.uint8, device=self.device) for v in col_valid if v]
        ch.recv(bases[me], outs)
        vmap, k = {}, ncols
        for i, v in enumerate(col_valid):
            if v:
                vmap[i] = outs[k]
                k += 1
        tables, r0 = [], 0
        for s_ in range(w):
            r1 = r0 + int(counts[s_, me])
            if r1 > r0:
                tables.append(DeviceTable({schema[i][0]: DeviceColumn(outs[i][r0:r1], schema[i][2], schema[i][3],
                                                                      None if i not in vmap else vmap[i][r0:r1])
                                           for i in range(ncols)}))
                tables[-1].src_rank = s_         # ordered consumers put the ranks' time ranges back in order with this
            r0 = r1
        return tables

    def _native_oversize(self, link, parts, table, lo, hi, single_owner, edge_key, need, have) -> list:
        """Somebody's share does not fit its mailbox (every rank sees that in the same meta matrix): the rows go in R
        rounds, each a complete exchange of a slice of every sender's rows."""
        R = -(-int(need) // int(have)) * 2
        out = []
        w = world_size()
        if isinstance(parts, Parts) and parts.pending is not None and single_owner is None:
            # an unmaterialised hash partition: every round is the fused scatter push of a slice of the INPUT rows; the
            # slice's own plan comes from the destinations the full plan already assigned (no local scatter, no copy)
            t, dest, doffs = parts.pending
            n = len(t)
            inner = doffs[1:w].contiguous()
            for r in range(R):
                a, b = n * r // R, n * (r + 1) // R
                if b > a:
                    which = torch.bucketize(dest[a:b].long(), inner, right=True).to(torch.int32)
                    d_r, o_r = ops.partition_plan(which, w, L.PART_CODE)
                    sub = Parts(None, None, (t.slice(a, b), d_r, o_r))
                else:
                    sub = {}
                out += self._native_call(link, sub, None, edge_key)
            return out
        for r in range(R):
            sub = {}
            if table is not None:
                for d in range(w):
                    n_d = hi[d] - lo[d]
                    a, b = lo[d] + n_d * r // R, lo[d] + n_d * (r + 1) // R
                    if b > a:
                        sub[d] = table.slice(a, b)
            out += self._native_call(link, sub, None, edge_key)
        return out

    def __call__(self, parts, n_target: int, single_owner: int | None = None, edge_key=None) -> list:
        """parts: {target_channel: DeviceTable} or edge.Parts.  Target channel c lives on rank c (or on
        `single_owner` when the consumer has a single channel).  Returns the tables received by this rank,
        one per source rank that sent rows."""
        w = world_size()
        if w > 1:
            link = native_link(self.device)
            if link is not None:
                return self._native_call(link, parts, single_owner, edge_key)
        if isinstance(parts, Parts):
            if w == 1:
                return parts.tables()
            peer_flag = 1 if (single_owner is None and parts.pending is not None and len(parts.offsets) == w + 1) else 0
            table, counts = (parts.pending[0] if peer_flag else parts.table), [0] * w
            if table is not None and len(table) > 0:
                if single_owner is not None:
                    counts[single_owner] = len(table)
                else:
                    for ch in range(len(parts.offsets) - 1):
                        counts[ch] += parts.offsets[ch + 1] - parts.offsets[ch]
            else:
                table = None
        else:
            if w == 1:
                return [p for _, p in sorted(parts.items()) if p is not None and len(p) > 0]
            # nothing to send: this rank can follow whichever path the others take
            peer_flag = 1 if (single_owner is None and not any(p is not None and len(p) > 0 for p in parts.values())) else 0
            owner = (lambda ch: single_owner) if single_owner is not None else (lambda ch: ch)
            by_rank = {}
            for ch, p in sorted(parts.items()):
                if p is not None and len(p) > 0:
                    by_rank.setdefault(owner(ch), []).append(p)
            counts = [sum(len(p) for p in by_rank.get(r, [])) for r in range(w)]
            ordered = [p for r in range(w) for p in by_rank.get(r, [])]
            table = concat_tables(ordered) if ordered else None
        me = rank()
        self.calls += 1
        # Schema agreement without pickling: every rank publishes a checksum of its table's schema (0 = no
        # table), of the schema it has cached for this edge, and the column widths.  Names / dtypes /
        # dictionaries are exchanged as objects only when the checksums disagree (a dictionary grew) or when a
        # rank that will RECEIVE rows has neither a table nor a matching cached schema.  Payload moves as bytes,
        # so a rank that neither sends nor receives can take part in the collectives knowing only the widths.
        MAXC = 24
        cached = self.schemas.get(edge_key) if edge_key is not None else None
        mine = self._schema_of(table) if table is not None else None
        h_mine = (zlib.crc32(repr(mine).encode()) | 1) if mine is not None else 0
        h_cached = (zlib.crc32(repr(cached).encode()) | 1) if cached is not None else 0
        widths = [0] * MAXC
        if mine is not None:
            if len(mine) > MAXC:
                raise L.QkError(f"exchange: more than {MAXC} columns on one edge")
            for i, (_, dt, _, _, hv) in enumerate(mine):
                widths[i] = _DT[dt].itemsize | (256 if hv else 0)
        meta = torch.tensor(counts + [h_mine, h_cached, len(mine) if mine is not None else 0] + widths + [peer_flag],
                            dtype=torch.int64, device=self.device)
        M = w + 3 + MAXC + 1
        allmeta = torch.empty(w * M, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(allmeta, meta)
        allmeta = allmeta.cpu().view(w, M)
        if int(allmeta[:, :w].sum()) == 0:
            return []
        hashes = set(int(x) for x in allmeta[:, w].tolist() if x != 0)
        negotiate = len(hashes) != 1
        if not negotiate:
            common = next(iter(hashes))
            for r in range(w):
                if int(allmeta[r, w]) == 0 and int(allmeta[:, r].sum()) > 0 and int(allmeta[r, w + 1]) != common:
                    negotiate = True
        if negotiate:
            headers = [None] * w
            dist.all_gather_object(headers, mine)
            known = [h for h in headers if h is not None]
            schema = []
            for i, (name, dt, _, atype, has_valid) in enumerate(known[0]):
                dicts = [h[i][2] for h in known]
                union = sorted(set().union(*[set(d) for d in dicts if d is not None])) if any(d is not None for d in dicts) else None
                schema.append((name, dt, union, atype, any(h[i][4] for h in known)))
        else:
            schema = mine if mine is not None else (cached if h_cached == next(iter(hashes)) else None)
        if edge_key is not None and schema is not None:
            self.schemas[edge_key] = schema
        if int(allmeta[:, M - 1].sum()) == w and peer_mailbox(self.device, self.mailbox_bytes) is not None:
            got = self._peer_path(parts if isinstance(parts, Parts) else None, allmeta, schema, w, me)
            if got is not None:
                return got
        if isinstance(parts, Parts) and table is not None and parts.pending is not None:
            table = parts.table                                   # NCCL path: group the rows locally first
        src = next(r for r in range(w) if int(allmeta[r, w + 2]) > 0)
        ncols = int(allmeta[src, w + 2])
        col_w = [int(allmeta[src, w + 3 + i]) & 255 for i in range(ncols)]
        col_valid = [any(int(allmeta[r, w + 3 + i]) & 256 for r in range(w)) for i in range(ncols)]
        recv_counts = [int(allmeta[s_, me]) for s_ in range(w)]
        send_counts = [int(c) for c in counts]
        n_recv, n_send = sum(recv_counts), sum(send_counts)
        received = []
        grouped = os.environ.get("QK_EXCHANGE") == "grouped"        # opt-in: ONE grouped send/recv call for all columns
        pending_ops, sent_keep = [], []
        for i in range(ncols):
            send_b = torch.zeros(0, dtype=torch.uint8, device=self.device)
            valid_b = torch.zeros(0, dtype=torch.uint8, device=self.device)
            if table is not None:
                name = schema[i][0]
                c = table[name]
                union = schema[i][2]
                if union is not None and c.dictionary != union:
                    c = unify_dictionaries([DeviceColumn(torch.zeros(0, dtype=c.data.dtype, device=c.data.device), union, c.arrow_type), c])[1][1]
                d = _flat(c.data)
                send_b = d.view(torch.uint8)
                if col_valid[i]:
                    valid_b = c.valid if c.valid is not None else torch.ones(len(c), dtype=torch.uint8, device=self.device)
            # one variable-split all-to-all per column (grouped ncclSend/ncclRecv inside NCCL), straight out of the
            # partition kernel's output: rows of one destination are contiguous there
            wd = col_w[i]
            recv_b = torch.empty(n_recv * wd, dtype=torch.uint8, device=self.device)
            valid_r = torch.empty(n_recv, dtype=torch.uint8, device=self.device) if col_valid[i] else None
            if grouped:
                # every rank knows the whole counts matrix, so both ends of each transfer agree on which ones are empty
                for buf_s, buf_r, unit in ((send_b, recv_b, wd),) + (((valid_b, valid_r, 1),) if col_valid[i] else ()):
                    so = ro = 0
                    for r in range(w):
                        ns, nr = send_counts[r] * unit, recv_counts[r] * unit
                        if r == me:
                            if ns:
                                buf_r[ro:ro + nr].copy_(buf_s[so:so + ns])
                        else:
                            if ns:
                                pending_ops.append(dist.P2POp(dist.isend, buf_s[so:so + ns], r))
                            if nr:
                                pending_ops.append(dist.P2POp(dist.irecv, buf_r[ro:ro + nr], r))
                        so, ro = so + ns, ro + nr
                    sent_keep.append(buf_s)
            else:
                dist.all_to_all_single(recv_b, send_b, [c_ * wd for c_ in recv_counts], [c_ * wd for c_ in send_counts])
                if col_valid[i]:
                    dist.all_to_all_single(valid_r, valid_b, recv_counts, send_counts)
            self.bytes_sent += (n_send - send_counts[me]) * wd
            received.append((recv_b, valid_r))
        if pending_ops:
            for req in dist.batch_isend_irecv(pending_ops):
                req.wait()
        del sent_keep
        if n_recv == 0 or schema is None:
            return []
        roff = [0]
        for c_ in recv_counts:
            roff.append(roff[-1] + c_)
        out_cols = {}
        for (name, dt, union, atype, _), (recv_b, valid_r) in zip(schema, received):
            out_cols[name] = (recv_b.view(_DT[dt]), union, atype, valid_r)
        tables = []
        for s_ in range(w):
            lo, hi = roff[s_], roff[s_ + 1]
            if hi > lo:
                tables.append(DeviceTable({n: DeviceColumn(d[lo:hi], u, a, None if v is None else v[lo:hi])
                                           for n, (d, u, a, v) in out_cols.items()}))
                tables[-1].src_rank = s_
        return tables


# ------------------------------------------------------------------ the graph
def _takes_device_tables(executor) -> bool:
    """The Executor protocol hands `execute` a list of pyarrow.Table (pyquokka/executors/base_executor.py:26-32,
    core.py:624-632).  The executors of this package take the device-resident batches directly; any other class (a
    user's plug-in, written against the reference) gets Arrow tables on the host unless it opts in with
    `device_tables = True`.  Either kind may return a DeviceTable, a pyarrow.Table, a pandas frame or None."""
    flag = getattr(executor, "device_tables", None)
    return bool(flag) if flag is not None else type(executor).__module__.startswith("quokka_b200.")


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
        self.lock = threading.Lock()   # one execute() at a time per executor instance (core.py:493), whatever the lane
        self.tail = None               # CUDA event after the last execute(): the next caller's stream waits for it


class TaskGraph:
    """Same construction calls as the reference's TaskGraph (quokka_runtime.py:118,314,370,383,394)."""

    def __init__(self, context=None) -> None:
        self.context = context
        self.actors: dict[int, _Actor] = {}
        self.current_actor = 0
        self.device = getattr(context, "device", None) or _default_device()
        self.exchange = Exchange(self.device)
        self.profile = bool(os.environ.get("QK_PROFILE"))      # like the reference's PROFILE flag (core.py:20-30)
        self.timings = {}
        self._bind_lock = threading.Lock()
        self._lane_streams = None
        self.lanes_used = 0

    def _timed(self, label, fn, *a, **kw):
        if not self.profile:
            return fn(*a, **kw)
        import time
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(*a, **kw)
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        self.timings[label] = self.timings.get(label, 0.0) + (time.perf_counter() - t0)
        return out

    def report(self):
        return {k: round(v * 1e3, 3) for k, v in sorted(self.timings.items(), key=lambda kv: -kv[1])}

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
                with self._bind_lock:
                    ti.bind(table.column_names)
                parts = self._timed(f"edge {actor.id}->{tgt_id} partition_fn", partition_fn, ti, table, rank(), n)
            else:
                parts = {}
            if n > 1 and isinstance(ti.partitioner, PassThroughPartitioner) and isinstance(parts, dict):
                # channel c feeds channel c: nothing crosses ranks, and every rank knows that from the plan alone, so
                # the whole exchange (metadata all-gather, host sync, all-to-all) is skipped -- by all ranks alike
                received = [p for _, p in sorted(parts.items()) if p is not None and len(p) > 0]
            else:
                received = self._timed(f"edge {actor.id}->{tgt_id} exchange", self.exchange, parts, n,
                                       single_owner=0 if tgt.single else None, edge_key=(actor.id, tgt_id, stream_id))
            out = None
            if self._owns(tgt) and received:
                if not _takes_device_tables(tgt.instance):      # a user's Executor: the reference protocol, list[pyarrow.Table]
                    received = [b.to_arrow() for b in received]
                with tgt.lock:
                    cuda = self.device.type == "cuda"
                    if cuda and tgt.tail is not None:           # state last touched on another lane's stream
                        torch.cuda.current_stream().wait_event(tgt.tail)
                    out = self._timed(f"actor {tgt_id} {type(tgt.instance).__name__}.execute[{stream_id}]", tgt.instance.execute, received, stream_id, rank())
                    out = as_device_table(out) if out is not None else None
                    if cuda:
                        tgt.tail = torch.cuda.Event()
                        tgt.tail.record()
            # An executor whose execute() never returns rows for this stream (the build side of a join, an aggregate that
            # answers in done(): `silent_streams` on the class, the same on every rank) sends nothing downstream: no rank
            # starts the (empty) exchanges a None would otherwise cascade through -- one meta round trip each at > 1 rank
            quiet = _is_silent(tgt.instance, stream_id)
            if quiet and out is not None and len(out) > 0:
                raise L.QkError(f"{type(tgt.instance).__name__}.execute returned rows on stream {stream_id}, which it declares silent")
            self._emit(tgt, out, quiet)

    def _emit(self, actor: _Actor, out, quiet: bool = False):
        if actor.blocking:
            if out is not None and len(out) > 0:
                actor.results.append(out)
        elif not quiet:
            self._push(actor, out)

    def _publish_bloom(self, actor: _Actor):
        """`actor` just delivered the last build batch of every join it feeds on stream 1: where the planner
        asked for it, build the Bloom filter of each channel's build keys, all-gather the filters and hand them
        to the probe edge, so the probe-side scan drops non-joining rows BEFORE they are partitioned and sent."""
        for tgt_id, stream_id, edge in actor.targets:
            tgt = self.actors[tgt_id]
            if stream_id != 1 or not hasattr(tgt.instance, "make_bloom") or tgt.single:
                continue
            replicated = isinstance(edge.partitioner, BroadcastPartitioner)     # every rank holds ALL build keys
            sinks = [ti for a in self.actors.values() for _, _, ti in a.targets
                     if ti.bloom_key is not None and ti.bloom_source == tgt_id]
            if not sinks:
                continue
            w, me = world_size(), rank()
            n_local = tgt.instance.build_rows()
            no_filter = 0 if tgt.instance.bloom_ok() else 1      # string keys: codes of unrelated dictionaries
            if w > 1:
                ekey = (actor.id, tgt_id, stream_id)
                tot, sch = self.exchange.recv_totals.get(ekey), self.exchange.schemas.get(ekey)
                if tot is not None and (sch is not None or sum(tot) == 0) and native_link(self.device) is not None:
                    # every rank already knows how many build rows every rank holds (the exchanges' count matrices) and
                    # whether the key is a string column (the agreed schema): no extra agreement round
                    n_local = max(tot)
                    right = getattr(tgt.instance, "right_on", None)
                    # (string keys: codes of unrelated dictionaries; fp64 keys are matched on their bit pattern, not filtered)
                    no_filter = 1 if (sch is not None and any(name == right and (dic is not None or "float" in dt) for name, dt, dic, _, _ in sch)) else 0
                else:
                    rows = self.exchange.allgather_words([n_local, no_filter])
                    n_local, no_filter = max(r[0] for r in rows), max(r[1] for r in rows)
            if no_filter:
                continue
            words = ops.Bloom.words_for(-(-n_local // w) if replicated else n_local)
            local = self._timed(f"actor {tgt_id} bloom build", tgt.instance.make_bloom, words, w)
            bits = local.bits
            if w > 1 and not replicated:            # a replicated build side yields the complete filter on every rank
                mine = bits[me * words:(me + 1) * words].contiguous()
                if native_link(self.device) is not None:     # my slice to every rank over the peer-memory link
                    t = DeviceTable({"bits": DeviceColumn(mine)})
                    got = self.exchange({r: t for r in range(w)}, w, edge_key=("bloom", tgt_id))
                    bits = torch.cat([g["bits"].data for g in got]) if len(got) > 1 else got[0]["bits"].data
                else:
                    allbits = torch.empty(w * words, dtype=bits.dtype, device=self.device)
                    dist.all_gather_into_tensor(allbits, mine)
                    bits = allbits
            for ti in sinks:
                ti.bloom = ops.Bloom(bits, words, w)

    def _finish(self, actor: _Actor):
        actor.done = True
        self._publish_bloom(actor)
        for tgt_id, _, _ in actor.targets:
            tgt = self.actors[tgt_id]
            if tgt.done or not all(self.actors[s].done for s in tgt.sources.values()):
                continue
            out = None
            if self._owns(tgt):
                out = self._timed(f"actor {tgt_id} {type(tgt.instance).__name__}.done", tgt.instance.done, rank())
                out = as_device_table(out) if out is not None else None
            # An executor whose done() never emits (joins, pass-through sinks: `emits_on_done = False` on the class, the
            # same on every rank) sends nothing downstream: no rank starts the exchange for it.
            if out is not None or getattr(tgt.instance, "emits_on_done", True):
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
                fixed = getattr(a.obj, "fixed_rounds", None)
                if fixed is not None:
                    rounds = fixed                              # the same on every rank by construction: no agreement round
                elif w > 1:
                    rounds = max(r[0] for r in self.exchange.allgather_words([rounds]))
                if not self._run_laned(a, mine, rounds):
                    for i in range(rounds):
                        batch = a.obj.execute(me, mine[i])[1] if i < len(mine) else None
                        self._push(a, as_device_table(batch, self.device) if batch is not None else None)
            self._finish(a)
        return self

    def _run_laned(self, a: _Actor, mine: list, rounds: int) -> bool:
        """Chunk-pipelined execution of one reader: chunk i runs depth-first through the graph (scan -> partition ->
        exchange -> the consumer's execute -> ... ) on lane i % L, each lane a host thread with its own CUDA stream and
        its own exchange channel.  While one lane waits for the one host round trip of an exchange (or for a row count),
        the other has already enqueued the scan of the next chunk: exchange(i) overlaps scan(i+1) and build / probe(i-1),
        and partitions are consumed as they arrive (the reference pushes batches as produced, core.py:654,946).
        Executors see one execute() at a time (per-actor lock + event chain).  Needs: a reader whose execute() is
        re-entrant (`concurrent`), an unordered stream, and -- across ranks -- the peer-memory link (channels make the
        order of exchanges per lane the same on every rank; library collectives from several threads would not be)."""
        lanes = min(N_LANES, rounds)
        if lanes < 2 or self.device.type != "cuda" or not getattr(a.obj, "concurrent", False) or a.ordered:
            return False
        if world_size() > 1 and native_link(self.device) is None:
            return False
        if self._lane_streams is None:
            self._lane_streams = lane_streams(self.device)
        main = torch.cuda.current_stream()
        me = rank()
        errors = []

        def work(li):
            try:
                torch.cuda.set_device(self.device)
                _lane.index = li
                st = self._lane_streams[li]
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    for i in range(li, rounds, lanes):
                        batch = a.obj.execute(me, mine[i])[1] if i < len(mine) else None
                        self._push(a, as_device_table(batch, self.device) if batch is not None else None)
            except BaseException as e:                          # surfaces on the driver thread
                errors.append(e)

        threads = [threading.Thread(target=work, args=(li,), name=f"qk-lane-{li}") for li in range(lanes)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for st in self._lane_streams[:lanes]:
            main.wait_stream(st)
        if errors:
            raise errors[0]
        self.lanes_used = max(self.lanes_used, lanes)
        return True

    def results(self, actor_id) -> list:
        return self.actors[actor_id].results

    def describe(self) -> str:
        """The physical plan: actors in execution order with their stage, and every edge with its partitioner,
        predicate, folded batch functions and semi-join (Bloom) reduction."""
        lines = []
        for a in sorted(self.actors.values(), key=lambda a: (a.stage if a.kind == "input" else 1 << 30, a.id)):
            what = type(a.obj).__name__ + (" [single channel]" if a.single else "") + (" [sink]" if a.blocking else "")
            lines.append(f"actor {a.id}: {what}" + (f"  stage {a.stage}" if a.kind == "input" else ""))
            for tgt, sid, ti in a.targets:
                e = ti.edge_ops
                bits = [f"partitioner={ti.partitioner}"]
                if e.pred is not None:
                    bits.append(f"where {e.pred.sql()}")
                if e.defs is not None:
                    bits.append("cols=" + ",".join(n if d.kind == "col" and d.value == n else f"{n}:={d.sql()}" for n, d in e.defs.items()))
                if ti.batch_funcs:
                    bits.append("batch_funcs=" + ",".join(type(f).__name__ if not callable(f) or hasattr(f, "keys") else getattr(f, "__name__", "fn") for f in ti.batch_funcs))
                if ti.bloom_key is not None:
                    bits.append(f"bloom({ti.bloom_key} in build keys of actor {ti.bloom_source})")
                lines.append(f"    -> actor {tgt} stream {sid}: " + "; ".join(bits))
        return "\n".join(lines)


def _is_silent(instance, stream_id) -> bool:
    s = getattr(instance, "silent_streams", ())
    return s == "all" or stream_id in s


def gather_to_all(tables: list, device) -> list:
    """collect(): every rank receives every rank's sink batches (quokka_dataset.py:107-117: the result is
    the unordered concatenation of all sink batches)."""
    w = world_size()
    local = concat_tables(tables) if tables else None
    if w == 1:
        return [local] if local is not None and len(local) > 0 else []
    ex = Exchange(device)
    # broadcast partitioner semantics: send my batch to every rank
    parts = {r: local for r in range(w)} if local is not None and len(local) > 0 else {}
    return ex(parts, w)
