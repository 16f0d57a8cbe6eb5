"""Per-GPU executor of one piece: device-resident batch state, prefill, CUDA-graph decode
bursts, wavefront micro-batch groups.  One ``GpuRunner`` per process/GPU; rank 0 also owns
the token buffers the last piece's sampler writes into over NVLink.

Burst protocol (all ranks execute the same plan):
    prefill(new sequences)  : eager launches, flags for the activation hop, host barrier after
    decode(n steps)         : for step: for group g: replay graph[g]   -- device-side flags only
                              (piece i works on group g+1 while piece i+1 works on group g)
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .. import ops
from ..models.config import ModelConfig, balanced_split, piece_units, split_layers
from ..models.native import BatchMeta, Handoff, NativePiece
from ..models.weights import load_or_init
from ..parallel.mesh import MeshComm
from .kv import PAGE


@dataclass
class SeqInit:
    """Everything a rank needs to start a sequence (part of the replicated plan)."""
    slot: int
    prompt: List[int]
    pages: List[int]
    temperature: float = 0.7
    top_p: float = 0.95
    repetition_penalty: float = 1.15
    seed: int = 0


class GpuRunner:
    def __init__(self, cfg: ModelConfig, model: str = "", rank: int = 0, world: int = 1,
                 device: Optional[torch.device] = None, max_batch: int = 32, groups: int = 1,
                 max_seq_len: int = 4096, max_prefill_tokens: int = 2048, num_pages: int = 0, seed: int = 0,
                 hist_len: int = 4096, control_group=None, use_graphs: bool = True, quant: str = "bf16"):
        assert max_batch % groups == 0, "max_batch must be divisible by the number of micro-batch groups"
        self.cfg, self.rank, self.world = cfg, rank, world
        self.device = torch.device(device if device is not None else f"cuda:{rank}")
        torch.cuda.set_device(self.device)
        self.max_batch, self.groups, self.gb = max_batch, groups, max_batch // groups
        self.max_seq_len = max_seq_len
        self.max_pages_per_seq = (max_seq_len + PAGE - 1) // PAGE
        self.max_prefill_tokens = max_prefill_tokens
        self.hist_len = hist_len
        self.use_graphs = use_graphs
        # piece boundaries in half-layer units (attention block | MLP block), balanced for the wavefront: the last
        # piece also streams the lm_head and runs the sampler.  B2B_UNIT_BOUNDS="0,3,8" overrides the search (tests).
        env_bounds = os.environ.get("B2B_UNIT_BOUNDS", "")
        bounds = [int(v) for v in env_bounds.split(",")] if env_bounds and world > 1 else None
        self.unit_ranges = piece_units(cfg, world, bounds)
        assert len(self.unit_ranges) == world, f"{cfg.name}: cannot split into {world} pieces ({self.unit_ranges})"
        self.units = self.unit_ranges[rank]
        self.layers = list(range(self.units[0] // 2, (self.units[1] + 1) // 2))
        self.first, self.last = rank == 0, rank == world - 1
        if num_pages <= 0:
            num_pages = 1 + max_batch * self.max_pages_per_seq
        self.num_pages = num_pages
        tensors = load_or_init(model, cfg, self.layers, self.first, self.last, device=self.device,
                               dtype=torch.bfloat16, seed=seed)
        max_tokens = max(max_prefill_tokens, self.gb)
        self.piece = NativePiece(cfg, self.layers, self.first, self.last, tensors, self.device, max_tokens,
                                 max(self.gb, 64), num_pages, quant=quant, units=self.units)
        del tensors
        self.mesh = MeshComm(rank, world, self.device, cfg.hidden_size, max_tokens, groups, self.gb, hist_len,
                             control_group)
        dev, i32 = self.device, torch.int32
        B = max_batch
        # ---- device-resident batch state (static addresses -> CUDA graphs)
        if world > 1 and self.first:
            self.tokens = self.mesh.local_view("tok", (B,), "i32")
            self.history = self.mesh.local_view("hist", (B, hist_len), "i32")
        else:
            self.tokens = torch.zeros(B, device=dev, dtype=i32)
            self.history = torch.zeros((B, hist_len), device=dev, dtype=i32)
        self.positions = torch.zeros(B, device=dev, dtype=i32)
        self.kv_len = torch.zeros(B, device=dev, dtype=i32)
        self.q_len = torch.zeros(B, device=dev, dtype=i32)
        self.slots = torch.full((B,), -1, device=dev, dtype=i32)
        self.block_table = torch.zeros((B, self.max_pages_per_seq), device=dev, dtype=i32)
        self.q_start = torch.arange(self.gb, device=dev, dtype=i32)
        self.temperature = torch.zeros(B, device=dev)
        self.top_p = torch.ones(B, device=dev)
        self.rep_pen = torch.ones(B, device=dev)
        self.seeds = torch.zeros(B, device=dev, dtype=i32)
        self.hist_pos = torch.zeros(B, device=dev, dtype=i32)
        self.step_ctr = torch.zeros(1, device=dev, dtype=i32)
        self.seen = torch.zeros((B, (cfg.vocab_size + 31) // 32), device=dev, dtype=i32)
        self.tok_local = torch.zeros(B, device=dev, dtype=i32)     # last rank's own copy of sampled ids
        self.graphs: Dict[int, torch.cuda.CUDAGraph] = {}
        self._pf: Dict[int, dict] = {}             # prefill graphs per token bucket
        self.stream = torch.cuda.Stream(device=dev)
        self.kernel_launches = 0
        self.h2d_bytes = 0
        self.decode_splits = 1
        if max_seq_len > 1024:
            self.decode_splits = max(1, min(self.piece.max_splits, ops.NUM_SMS // max(1, self.gb * cfg.n_kv_heads)))
        ops.native().init_kernels(self.device.index)
        if world == 1:
            self.warmup()
            self._ensure_graphs()

    # ------------------------------------------------------------------ helpers
    def _grp(self, t: torch.Tensor, g: int) -> torch.Tensor:
        return t[g * self.gb:(g + 1) * self.gb]

    def launches_per_decode_step(self) -> int:
        """Native kernel launches one rank issues per group decode step (for reporting)."""
        c, n = self.cfg, len(self.layers)
        if c.norm == "ln":
            per = 9
        elif c.post_norms:
            per = 7
        else:
            per = 5
        total = 1 + n * per                      # decode_advance + layers
        total -= (3 if self.piece.head_skip_attn else 0) + (2 if self.piece.tail_skip_mlp else 0)   # half-layer ends
        if self.first:
            total += 1                           # embed
        if self.last:
            total += 2 + (0 if c.norm == "rms" else 1)   # lm_head GEMM + sampler (+ final LN)
            if not self.first:
                total += 1                       # input-slot release
        return total

    # ------------------------------------------------------------------ prefill
    def prefill(self, seqs: Sequence[SeqInit]) -> None:
        """Run the prompts of ``seqs`` through this rank's piece (chunked), install their decode
        state, and (last rank) sample their first token into rank 0's token buffer."""
        if not seqs:
            return
        dev, i32 = self.device, torch.int32
        with torch.cuda.stream(self.stream):
            # per-sequence state of all admitted sequences in a handful of batched copies (one tiny launch per
            # field and sequence cost ~25 ms for 256 admissions)
            S = len(seqs)
            bt_host = torch.zeros((S, self.max_pages_per_seq), dtype=i32)
            for i, s in enumerate(seqs):
                bt_host[i, :len(s.pages)] = torch.tensor(s.pages, dtype=i32)
            fl_host = torch.tensor([[s.temperature, s.top_p, s.repetition_penalty] for s in seqs], dtype=torch.float32)
            seed_host = torch.tensor([int(s.seed) & 0x7FFFFFFF for s in seqs], dtype=i32)
            rows = torch.tensor([s.slot for s in seqs], dtype=torch.int64).to(dev, non_blocking=True)
            fl = fl_host.to(dev, non_blocking=True)
            self.block_table.index_copy_(0, rows, bt_host.to(dev, non_blocking=True))
            self.temperature.index_copy_(0, rows, fl[:, 0].contiguous())
            self.top_p.index_copy_(0, rows, fl[:, 1].contiguous())
            self.rep_pen.index_copy_(0, rows, fl[:, 2].contiguous())
            self.seeds.index_copy_(0, rows, seed_host.to(dev, non_blocking=True))
            self.hist_pos.index_fill_(0, rows, 0)
            self.seen.index_fill_(0, rows, 0)
            self.h2d_bytes += bt_host.numel() * 4 + fl_host.numel() * 4 + S * (4 + 8)
            if self.last:
                flat = [t for s in seqs for t in s.prompt]
                self.h2d_bytes += 8 * len(flat)
                owner = [s.slot for s in seqs for _ in s.prompt]
                ops.mark_seen(torch.tensor(flat, dtype=i32).to(dev, non_blocking=True),
                              torch.tensor(owner, dtype=i32).to(dev, non_blocking=True), self.seen, self.cfg.vocab_size)
            if self._prefill_graph_ok(seqs):
                self._prefill_single_graph(seqs[0])
                self.stream.synchronize()
                return
            # pack prompts into chunks of <= max_prefill_tokens tokens
            work: List[Tuple[SeqInit, int, int]] = []          # (seq, start, end)
            for s in seqs:
                L = len(s.prompt)
                for c0 in range(0, L, self.max_prefill_tokens):
                    work.append((s, c0, min(L, c0 + self.max_prefill_tokens)))
            batch: List[Tuple[SeqInit, int, int]] = []
            used = 0
            chunks: List[List[Tuple[SeqInit, int, int]]] = []
            for w in work:
                n = w[2] - w[1]
                same_seq = any(x[0].slot == w[0].slot for x in batch)
                if batch and (used + n > self.max_prefill_tokens or same_seq or len(batch) >= self.piece.max_seqs):
                    chunks.append(batch)
                    batch, used = [], 0
                batch.append(w)
                used += n
            if batch:
                chunks.append(batch)
            # Chunks travel through the pieces as a wavefront: piece i works on chunk c+1 while piece i+1 works on
            # chunk c.  The handoff slot is flow-controlled on the device (release flag / consumer ack), so no host
            # barrier is needed between chunks (B2B_PREFILL_BARRIER=1 restores the serialised behaviour).
            serial = os.environ.get("B2B_PREFILL_BARRIER", "0") == "1"
            for ci, chunk in enumerate(chunks):
                self._prefill_chunk(chunk)
                if self.world > 1 and serial:
                    self.stream.synchronize()
                    self.mesh.barrier()
            lens = torch.tensor([len(s.prompt) for s in seqs], dtype=i32).to(dev, non_blocking=True)
            self.positions.index_copy_(0, rows, lens - 1)
            self.kv_len.index_copy_(0, rows, lens)
            self.q_len.index_fill_(0, rows, 1)
        self.stream.synchronize()
        self.mesh.barrier()

    # ---- graph-captured prefill for a single short prompt (the common chat / TTFT case) ----------
    PF_BUCKETS = (16, 32, 64)

    def _prefill_graph_ok(self, seqs) -> bool:
        # single short prompt (the chat / TTFT case): one H2D copy + one graph replay; B2B_PREFILL_GRAPH=0 disables
        return (os.environ.get("B2B_PREFILL_GRAPH", "1") == "1" and self.use_graphs and self.world == 1
                and len(seqs) == 1 and 0 < len(seqs[0].prompt) <= self.PF_BUCKETS[-1]
                and self.max_prefill_tokens >= self.PF_BUCKETS[-1])

    def _pf_state(self, tb: int):
        """Static staging + captured graph for prompts padded to ``tb`` tokens.  One pinned host
        buffer / one H2D copy carries ids, positions, slots and the scalars."""
        st = self._pf.get(tb)
        if st is not None:
            return st
        dev, i32 = self.device, torch.int32
        mp = self.max_pages_per_seq
        n = 3 * tb + 8 + mp
        host = torch.zeros(n, dtype=i32).pin_memory()
        stage = torch.zeros(n, device=dev, dtype=i32)
        ids, pos, slots = stage[0:tb], stage[tb:2 * tb], stage[2 * tb:3 * tb]
        qlen, kvlen, row32, last32 = (stage[3 * tb + i:3 * tb + i + 1] for i in range(4))
        bt = stage[3 * tb + 8:3 * tb + 8 + mp].view(1, mp)              # this sequence's block-table row
        qstart = torch.zeros(1, device=dev, dtype=i32)
        last64 = torch.zeros(1, device=dev, dtype=torch.int64)
        # the captured kernels address these tensors by raw pointer: they must outlive this function (qstart / last64
        # used to be locals of the capture closure -> freed, reallocated, and the replays read garbage offsets)
        st = {"host": host, "stage": stage, "graph": None, "keep": (qstart, last64)}

        def body():
            # no torch gather/scatter ops in here: every per-sequence input is either staged by the single
            # H2D copy or addressed through the device-side row offset `row32`
            last64.copy_(last32)
            meta = BatchMeta(ids=ids, positions=pos, slots=slots, q_start=qstart, q_len=qlen, kv_len=kvlen,
                             block_table=bt, n_tokens=tb, n_seqs=1, max_q=tb, last_idx=last64)
            out = self.piece.forward(meta)
            ops.sample(out, self.tokens, seen=self.seen, temperature=self.temperature, top_p=self.top_p,
                       rep_penalty=self.rep_pen, seeds=self.seeds, step=self.step_ctr, vocab=self.cfg.vocab_size,
                       softcap=self.cfg.final_softcap, history=self.history.data_ptr(), hist_pos=self.hist_pos,
                       hist_stride=self.hist_len, row_base=row32.data_ptr())
            ops.native().set_decode_state(self.positions, self.kv_len, self.q_len, row32.data_ptr(), kvlen.data_ptr())

        saved = (self.seen.clone(), self.tokens.clone(), self.history[:, 0].clone(), self.hist_pos.clone(),
                 self.positions.clone(), self.kv_len.clone(), self.q_len.clone(), self.step_ctr.clone())
        with torch.cuda.stream(self.stream):
            slots.fill_(-1)                         # warm-up run writes no KV
            qlen.fill_(1); kvlen.fill_(1)
            body()                                   # eager warm-up (allocations, descriptor cache)
            self.stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.stream, capture_error_mode="thread_local"):
                body()
        self.stream.synchronize()
        for dst, src in zip((self.seen, self.tokens, self.history[:, 0], self.hist_pos, self.positions, self.kv_len,
                             self.q_len, self.step_ctr), saved):
            dst.copy_(src)
        torch.cuda.synchronize(self.device)
        st["graph"] = g
        self._pf[tb] = st
        return st

    def _prefill_single_graph(self, s: SeqInit) -> None:
        L = len(s.prompt)
        tb = next(b for b in self.PF_BUCKETS if L <= b)
        st = self._pf_state(tb)
        h = st["host"]
        h.zero_()
        h[0:L] = torch.tensor(s.prompt, dtype=torch.int32)
        h[tb:tb + L] = torch.arange(L, dtype=torch.int32)
        h[2 * tb:3 * tb] = -1
        h[2 * tb:2 * tb + L] = torch.tensor([s.pages[p // PAGE] * PAGE + p % PAGE for p in range(L)], dtype=torch.int32)
        h[3 * tb + 0], h[3 * tb + 1], h[3 * tb + 2], h[3 * tb + 3] = L, L, s.slot, L - 1
        h[3 * tb + 8:3 * tb + 8 + len(s.pages)] = torch.tensor(s.pages, dtype=torch.int32)
        self.h2d_bytes += h.numel() * 4
        with torch.cuda.stream(self.stream):
            st["stage"].copy_(h, non_blocking=True)
            st["graph"].replay()
        self.kernel_launches += self.launches_per_decode_step()

    def _prefill_chunk(self, chunk) -> None:
        dev, i32 = self.device, torch.int32
        ids, pos, slots, q_start, q_len, kv_len, rows, last_idx, finals = [], [], [], [], [], [], [], [], []
        off = 0
        for s, c0, c1 in chunk:
            n = c1 - c0
            ids += s.prompt[c0:c1]
            pos += list(range(c0, c1))
            slots += [s.pages[p // PAGE] * PAGE + p % PAGE for p in range(c0, c1)]
            q_start.append(off)
            q_len.append(n)
            kv_len.append(c1)
            rows.append(s.slot)
            last_idx.append(off + n - 1)
            finals.append(c1 == len(s.prompt))
            off += n
        T, S = off, len(chunk)

        def t(v, dt=i32):
            h = torch.tensor(v, dtype=dt).pin_memory()      # host inputs are staged in pinned memory
            self.h2d_bytes += h.numel() * h.element_size()
            return h.to(dev, non_blocking=True)

        row_idx = t(rows, torch.int64)
        meta = BatchMeta(ids=t(ids), positions=t(pos), slots=t(slots), q_start=t(q_start), q_len=t(q_len),
                         kv_len=t(kv_len), block_table=self.block_table.index_select(0, row_idx).contiguous(),
                         n_tokens=T, n_seqs=S, max_q=max(q_len), last_idx=t(last_idx, torch.int64))
        hand = self.mesh.handoff(0)
        if self.first:
            hand.in_flag = hand.in_epoch = 0          # prompt ids come from the host, not from the token ring
        x_in = None
        if not self.first:
            x_in = self.mesh.C.tensor_from_ptr(hand.in_x, [T, self.cfg.hidden_size], "bf16", dev.index)
        out = self.piece.forward(meta, x_in=x_in, hand=hand if self.world > 1 else None)
        if self.last:
            # sample the first generated token of every sequence whose prompt is complete
            tok_tmp = torch.zeros(S, device=dev, dtype=i32)
            hp = torch.zeros(S, device=dev, dtype=i32)
            sel = lambda x: x.index_select(0, row_idx).contiguous()
            seen_sel = sel(self.seen)
            ops.sample(out, tok_tmp, seen=seen_sel, temperature=sel(self.temperature), top_p=sel(self.top_p),
                       rep_penalty=sel(self.rep_pen), seeds=sel(self.seeds), step=self.step_ctr,
                       vocab=self.cfg.vocab_size, softcap=self.cfg.final_softcap)
            fin_i = [i for i, f in enumerate(finals) if f]          # host-side selection: no device sync per chunk
            if fin_i:
                fin = t(fin_i, torch.int64)
                fin_rows = t([rows[i] for i in fin_i], torch.int64)
                self.seen.index_copy_(0, fin_rows, seen_sel.index_select(0, fin))
                self._publish_first_tokens(fin_rows, tok_tmp.index_select(0, fin))

    def _publish_first_tokens(self, rows: torch.Tensor, toks: torch.Tensor) -> None:
        """token ring + history[.., 0] on rank 0 (peer stores when world > 1)."""
        if self.world == 1:
            self.tokens.index_copy_(0, rows, toks)
            self.history[:, 0].index_copy_(0, rows, toks)
            self.hist_pos.index_fill_(0, rows, 1)
            return
        C = self.mesh.C
        tok_remote = C.tensor_from_ptr(self.mesh.remote_first["tok"], [self.max_batch], "i32", self.device.index)
        hist_remote = C.tensor_from_ptr(self.mesh.remote_first["hist"], [self.max_batch, self.hist_len], "i32",
                                        self.device.index)
        tok_remote.index_copy_(0, rows, toks)
        hist_remote[:, 0].index_copy_(0, rows, toks)
        self.hist_pos.index_fill_(0, rows, 1)

    # ------------------------------------------------------------------- decode
    def _decode_group(self, g: int) -> None:
        """One decode step of micro-batch group g on this rank (graph-capturable)."""
        gb = self.gb
        pos, kvl, slots, ql = (self._grp(x, g) for x in (self.positions, self.kv_len, self.slots, self.q_len))
        bt = self._grp(self.block_table, g)
        ops.native().decode_advance(pos, kvl, slots, ql, bt)
        meta = BatchMeta(ids=self._grp(self.tokens, g), positions=pos, slots=slots, q_start=self.q_start, q_len=ql,
                         kv_len=kvl, block_table=bt, n_tokens=gb, n_seqs=gb, max_q=1, splits=self.decode_splits)
        hand = self.mesh.handoff(g)
        x_in = None
        if not self.first:
            x_in = self.mesh.C.tensor_from_ptr(hand.in_x, [gb, self.cfg.hidden_size], "bf16", self.device.index)
        out = self.piece.forward(meta, x_in=x_in, hand=hand if self.world > 1 else None)
        if self.last:
            hist_ptr = (self.mesh.history_ptr(g) if self.world > 1 else self._grp(self.history, g).data_ptr())
            ops.sample(out, self._grp(self.tokens, g) if self.world == 1 else self._grp(self.tok_local, g),
                       seen=self._grp(self.seen, g), temperature=self._grp(self.temperature, g),
                       top_p=self._grp(self.top_p, g), rep_penalty=self._grp(self.rep_pen, g),
                       seeds=self._grp(self.seeds, g), step=self.step_ctr, vocab=self.cfg.vocab_size,
                       softcap=self.cfg.final_softcap, history=hist_ptr, hist_pos=self._grp(self.hist_pos, g),
                       hist_stride=self.hist_len, peer_tokens=hand.out_x if self.world > 1 else 0,
                       signal_flag=hand.out_flag if self.world > 1 else 0,
                       signal_epoch=hand.out_epoch if self.world > 1 else 0,
                       done_counter=(self.mesh._flag(self.mesh.local, g, 5) if self.world > 1 else 0))
            if g == self.groups - 1:
                self.step_ctr.add_(1)

    def _ensure_graphs(self) -> None:
        if self.graphs or not self.use_graphs:
            return
        if self.world > 1:
            # graphs that spin on peer flags cannot be warmed up in isolation; capture without a dry run
            pass
        with torch.cuda.stream(self.stream):
            for g in range(self.groups):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=self.stream, capture_error_mode="thread_local"):
                    self._decode_group(g)
                self.graphs[g] = graph
        self.stream.synchronize()

    def warmup(self) -> None:
        """Eagerly run every kernel variant once (sets smem attributes, fills the TMA descriptor
        cache) with all rows inactive, so nothing is written to the KV cache. Single rank only;
        multi-rank warm-up happens through a real burst."""
        if self.world > 1:
            return
        saved = self.q_len.clone()
        self.q_len.zero_()
        pos, kvl = self.positions.clone(), self.kv_len.clone()
        with torch.cuda.stream(self.stream):
            for g in range(self.groups):
                self._decode_group(g)
        self.stream.synchronize()
        self.q_len.copy_(saved)
        self.positions.copy_(pos)
        self.kv_len.copy_(kvl)
        self.hist_pos.zero_()
        self.step_ctr.zero_()
        self.seen.zero_()
        self.tokens.zero_()
        torch.cuda.synchronize(self.device)

    def prepare_burst(self) -> None:
        """Multi-rank only: with the GPUs idle and a host barrier on both sides, re-arm the handoff
        flags (the token ring on rank 0 already holds the next input tokens)."""
        if self.world > 1:
            self.stream.synchronize()
            self.mesh.barrier()
            self.mesh.reset_flags(token_ready=True)
            torch.cuda.synchronize(self.device)
            self.mesh.barrier()
        self._ensure_graphs()

    def decode(self, n_steps: int, prepared: bool = False) -> None:
        """Enqueue ``n_steps`` decode steps for every group (no host synchronisation inside)."""
        if not prepared:
            self.prepare_burst()
        from ..utils.tracing import TRACER
        with TRACER.range(f"decode_burst[{n_steps}]", self.stream):
            self._enqueue_decode(n_steps)

    def _enqueue_decode(self, n_steps: int) -> None:
        with torch.cuda.stream(self.stream):
            for _ in range(n_steps):
                for g in range(self.groups):
                    if self.use_graphs:
                        self.graphs[g].replay()
                    else:
                        self._decode_group(g)
        self.kernel_launches += n_steps * self.groups * self.launches_per_decode_step()

    def sync(self) -> None:
        self.stream.synchronize()
        self.mesh.barrier()

    # ------------------------------------------------------------------- results
    def read_history(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(history [B, hist_len], hist_pos [B]) on the host.  hist_pos lives on the last rank; on a
        multi-rank mesh the caller tracks positions by step count instead."""
        return self.history.cpu(), self.hist_pos.cpu()

    def release(self, slots: Sequence[int]) -> None:
        with torch.cuda.stream(self.stream):
            for b in slots:
                self.q_len[b] = 0
                self.slots[b] = -1
                self.kv_len[b] = 0

    def close(self) -> None:
        self.graphs.clear()
        self.mesh.close()
