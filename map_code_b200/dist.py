"""Multi-GPU step: one process per GPU, batch split over R ranks, [V, .] tables row-sharded by
`owner(id) = id % R` (local row id // R), dense parameters replicated with one gradient all-reduce.

The reference initialises NCCL (code/arguments.py:74) but never wraps the model or issues a collective; this module is
the data-parallel path the north_star asks for.  Every exchange is a FIXED-SIZE collective (no host-side counts, no
synchronisation, nothing data-dependent in the launch parameters):

  embedding forward   all-gather(ids [B_l,F])            -> owner kernel emb_gather_owned (zeros for foreign rows)
                      reduce-scatter(rows [R*B_l*F, D])   -> X0 of the local batch
  embedding backward  all-gather(dE [B_l, F*D])           -> owned_keys + sort/dedup + segment sums + row-wise AdamW on the shard
  NCE ("move the query, not the rows": 128 B of query instead of 26 x 132 B of table rows per position)
                      all-gather(query [N_l,P], ids [N_l,K+1]) -> nce_scores_owned -> reduce-scatter(scores [R*N_l, K+1])
                      local loss / dz                      -> all-gather(dz) -> nce_dinput_owned -> reduce-scatter(d_query)
                      table gradients of owned rows through the same dedup pipeline
  dense parameters    all-reduce(flat gradient buffer, SUM); the loss is scaled by 1/N_global so the sum is the global mean

Philox counters are indexed by the GLOBAL row (row0 = rank * B_l), so masks, replacements and noise are identical for any R
and the R-rank run reproduces the single-GPU run on the concatenated batch.

`ShardExchange` holds the choreography and is backend-agnostic: on GPUs `kernels` is map_code_b200.ops (NCCL group); the CPU
tests drive the same code over gloo with a torch emulation of the five owner-side kernels (tests/test_dist_cpu.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import _lib
from .engine import FusedStep, _Table, is_no_decay


# ------------------------------------------------------------------------------------------------ shard layout
def shard_rows(V: int, R: int) -> int:
    """rows per shard including one trailing dummy row that absorbs the 'foreign id' sentinel segment"""
    return (V + R - 1) // R + 1


def shard_table(full: torch.Tensor, R: int, rank: int) -> torch.Tensor:
    """[V, D] -> this rank's [shard_rows, D]: local row i holds global row i*R + rank; the dummy row (and padding) is 0."""
    V, D = full.shape
    out = torch.zeros(shard_rows(V, R), D, dtype=full.dtype, device=full.device)
    mine = full[rank::R]
    out[:mine.shape[0]] = mine
    return out


def unshard_table(shards: List[torch.Tensor], V: int) -> torch.Tensor:
    """inverse of shard_table over the list of all R shards (checkpoint export in the reference's [V, D] layout)"""
    R = len(shards)
    D = shards[0].shape[1]
    full = torch.empty(V, D, dtype=shards[0].dtype, device=shards[0].device)
    for r, s in enumerate(shards):
        n = (V - r + R - 1) // R
        full[r::R] = s[:n]
    return full


class ShardExchange:
    """The collective choreography around the owner-side kernels.  `kernels` provides emb_gather_owned, owned_keys,
    nce_scores_owned, nce_dinput_owned (signatures of map_code_b200.ops)."""

    def __init__(self, kernels, world: int, rank: int, group=None):
        self.k, self.R, self.rank, self.group = kernels, world, rank, group

    # -- collectives (fixed sizes)
    def all_gather(self, local: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = torch.empty((self.R,) + tuple(local.shape), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=self.group)
        _lib.mark("nccl_all_gather", ("bytes_out", out.numel() * out.element_size()))
        return out

    def reduce_scatter(self, full: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        dist.reduce_scatter_tensor(out.view(-1), full.view(-1), op=dist.ReduceOp.SUM, group=self.group)
        _lib.mark("nccl_reduce_scatter", ("bytes_in", full.numel() * full.element_size()))
        return out

    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        _lib.mark("nccl_all_reduce", ("bytes", t.numel() * t.element_size()))
        return t

    # -- embedding
    def embed_forward(self, shard: torch.Tensor, ids_local: torch.Tensor, ids_g: torch.Tensor, rows_g: torch.Tensor,
                      out_local: torch.Tensor) -> torch.Tensor:
        """ids_local [B_l,F] -> out_local [B_l*F, D]; fills ids_g [R,B_l,F] (kept for the backward) and scratch rows_g."""
        self.all_gather(ids_local, ids_g)
        self.k.emb_gather_owned(shard, ids_g.view(-1), self.R, self.rank, rows_g)
        return self.reduce_scatter(rows_g, out_local)

    def local_keys(self, ids_g: torch.Tensor, sentinel: int, keys: torch.Tensor) -> torch.Tensor:
        return self.k.owned_keys(ids_g.view(-1), self.R, self.rank, sentinel, keys)

    # -- NCE
    def nce_scores(self, q_local, ids_local, emb_shard, bias_shard, q_g, ids_g, partial_g, scores_local):
        self.all_gather(q_local, q_g)
        self.all_gather(ids_local, ids_g)
        N_g = q_g.shape[0] * q_g.shape[1]
        self.k.nce_scores_owned(q_g.view(N_g, -1), ids_g.view(N_g, -1), emb_shard, bias_shard.view(-1), self.R, self.rank, partial_g)
        return self.reduce_scatter(partial_g, scores_local)

    def nce_dinput(self, dz_local, ids_g, emb_shard, dz_g, dq_partial_g, dq_local):
        self.all_gather(dz_local, dz_g)
        N_g = dz_g.shape[0] * dz_g.shape[1]
        self.k.nce_dinput_owned(dz_g.view(N_g, -1), ids_g.view(N_g, -1), emb_shard, self.R, self.rank, dq_partial_g)
        return self.reduce_scatter(dq_partial_g, dq_local)


# ------------------------------------------------------------------------------------------------ the sharded fused step
class CollectiveShardedFusedStep(FusedStep):
    """FusedStep with row-sharded tables, exchanges as fixed-size NCCL collectives (portable variant; moves R x the rows an
    exact exchange would).  `batch_size` is the per-rank batch; the global batch is batch_size * world."""

    def __init__(self, model, *, world: int, rank: int, group=None, **kw):
        from . import ops
        self.world, self.rank = world, rank
        self.xch = ShardExchange(ops, world, rank, group)
        B = kw["batch_size"]
        kw.setdefault("use_graph", False)      # NCCL collectives are issued eagerly between the kernels
        kw["multi_stream"] = False             # one stream: collectives order themselves with the kernels
        kw["row0"] = rank * B
        kw["global_batch"] = B * world
        self._full_shapes: Dict[str, int] = {}
        super().__init__(model, **kw)

    # -- tables become shards (the model's parameters are re-pointed at the shard; export with full_state_dict())
    def _shard_param(self, name: str, param: torch.nn.Parameter):
        V = param.shape[0]
        self._full_shapes[name] = V
        param.data = shard_table(param.data, self.world, self.rank)
        param._map_sharded = (self.world, self.rank, V)   # the module path refuses to read a shard (layers.TableEmbedding.forward)

    def _make_embed_table(self):
        name = "embed.embedding.weight"
        self._shard_param(name, self.embed_w)
        R, B, F, D, dev = self.world, self.B, self.F, self.D, self.dev
        n = R * B * F
        t = _Table(name, self.embed_w, n, 0.0 if is_no_decay(name) else self.wd, dev)
        from . import ops
        t.plan = ops.DedupPlan(n, self.embed_w.shape[0], dev)
        self.tables[name] = t
        self.ids_g = torch.empty(R, B, F, dtype=torch.int64, device=dev)
        self.keys_emb = torch.empty(n, dtype=torch.int64, device=dev)
        self.rows_g = torch.empty(n, D, dtype=torch.float32, device=dev)
        self.dE_g = torch.empty(R, B, F * D, dtype=torch.float32, device=dev)

    def _make_nce_tables(self):
        from . import ops
        crit = self.model.mfp_criterion
        R, N, K1, P, dev = self.world, self.N, self.K + 1, self.P, self.dev
        self._shard_param("mfp_criterion.emb.weight", crit.emb.weight)
        self._shard_param("mfp_criterion.bias.weight", crit.bias.weight)
        n_occ = R * N * K1
        plan = ops.DedupPlan(n_occ, crit.emb.weight.shape[0], dev)
        te = _Table("mfp_criterion.emb.weight", crit.emb.weight, n_occ, 0.0 if is_no_decay("mfp_criterion.emb.weight") else self.wd, dev)
        tb = _Table("mfp_criterion.bias.weight", crit.bias.weight, n_occ, 0.0 if is_no_decay("mfp_criterion.bias.weight") else self.wd, dev)
        te.plan = tb.plan = plan
        self.tables[te.name], self.tables[tb.name] = te, tb
        f32 = dict(dtype=torch.float32, device=dev)
        self.q_g = torch.empty(R, N, P, **f32)
        self.nce_ids_g = torch.empty(R, N, K1, dtype=torch.int64, device=dev)
        self.partial_g = torch.empty(R * N, K1, **f32)
        self.scores = torch.empty(N, K1, **f32)
        self.dz_g = torch.empty(R, N, K1, **f32)
        self.dq_partial_g = torch.empty(R * N, P, **f32)
        self.keys_nce = torch.empty(n_occ, dtype=torch.int64, device=dev)

    # -- hooks
    def _embed_lookup(self, ids):
        t = self.tables["embed.embedding.weight"]
        self._draw_noise()
        self.xch.embed_forward(self.embed_w.data, ids, self.ids_g, self.rows_g, self.X0.view(self.B * self.F, self.D))
        self._split(self.X0, self.X0p)
        self.xch.local_keys(self.ids_g, self.embed_w.shape[0] - 1, self.keys_emb)
        t.plan.run(self.keys_emb)

    def _embed_backward(self):
        t = self.tables["embed.embedding.weight"]
        self.xch.all_gather(self.dE, self.dE_g)
        t.plan.reduce_rows(self.dE_g.view(-1, self.D), self.D, out=t.grad)

    def _nce_core(self):
        from . import ops
        P, K, L, N = self.P, self.K, self.L, self.N
        crit = self.model.mfp_criterion
        torch.cat([self.labels.view(N, 1), self.noise], dim=1, out=self.ids_all)   # [target | noise], nce_loss.py:138
        self.xch.nce_scores(self.sel, self.ids_all, crit.emb.weight.data, crit.bias.weight.data, self.q_g, self.nce_ids_g, self.partial_g,
                            self.scores)
        self.acc_count.zero_()
        n_global = self.global_batch * L
        ops.nce_loss_from_scores(self.scores, self.ids_all, crit.logprob_noise, self.norm_term, self.loss_type, 1.0 / n_global, self.logits,
                                 self.loss_pos, self.dz, self.acc_count)
        ops.reduce_sum(self.loss_pos, 1.0 / n_global, out=self.loss, ws=self.red_ws)
        self.xch.nce_dinput(self.dz, self.nce_ids_g, crit.emb.weight.data, self.dz_g, self.dq_partial_g, self.d_sel)
        te, tb = self.tables["mfp_criterion.emb.weight"], self.tables["mfp_criterion.bias.weight"]
        self.xch.local_keys(self.nce_ids_g, crit.emb.weight.shape[0] - 1, self.keys_nce)
        te.plan.run(self.keys_nce)
        te.plan.reduce_rows(self.q_g.view(-1, P), P, scale=self.dz_g.view(-1), group=K + 1, out=te.grad, scalar_out=tb.grad)

    def reduce_gradients(self):
        self.xch.all_reduce(self.grad_flat)

    def outputs(self):
        """global metrics: sums of the per-rank partial means / counts (tiny all-reduces, issued only when metrics are read)"""
        if self.mode == "MFP":
            loss = self.xch.all_reduce(self.loss.clone())
            acc = self.xch.all_reduce(self.acc_count.clone())
            return (loss.view(()), self.global_batch * self.L, acc.view(()))
        st = self.xch.all_reduce(torch.stack([self.stats[0] * (self.B / self.global_batch), self.stats[1], self.stats[2]]))
        cnt = self.global_batch * self.F
        if self.mode == "RFD":
            return (st[0], cnt, st[1] / cnt, st[2] / cnt)
        return (st[0], self.ctr_logits)

    def full_state_dict(self) -> Dict[str, torch.Tensor]:
        """state_dict in the reference's layout: shards of every table are all-gathered and re-interleaved to [V, D]."""
        sd = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        for name, V in self._full_shapes.items():
            shard = sd[name]
            parts = self.xch.all_gather(shard)
            sd[name] = unshard_table([parts[r] for r in range(self.world)], V)
        return sd


# ------------------------------------------------------------------------------------------------ owner-side merge (layout)
def finished_suffix(grad_offsets: Dict[str, tuple], biases: set, done: set, hi: int) -> int:
    """Start of the longest suffix of the flat gradient buffer below `hi` whose weight gradients are all complete.
    grad_offsets: name -> (lo, hi) float range in grad_flat, laid out [biases | weights in reverse backward order]
    (engine.FusedStep._alloc); biases never join a bucket (their column sums finish on side streams: they go out last)."""
    lo_ok = hi
    for name, (lo, _) in sorted(grad_offsets.items(), key=lambda kv: -kv[1][0]):
        if lo >= hi:
            continue
        if name not in done or name in biases:
            break
        lo_ok = lo
    return lo_ok


def merge_key_layout(world: int, n_shard_rows: int):
    """(shift, key_bits) of the merge keys  ((id // R) << shift) | source_rank  built by map_owned_compact: the source rank
    sits in the low `shift` bits so that a stable sort orders the contributions of one row by source rank."""
    shift = max(0, (world - 1).bit_length())
    return shift, max(1, ((n_shard_rows << shift) - 1).bit_length())


def peer_merge_torch(uniq_lists: List[torch.Tensor], grad_lists: List[torch.Tensor], world: int, rank: int, n_shard_rows: int):
    """torch restatement of the owner-side pull (map_owned_compact -> map_dedup_ids_ex(seg_shift) -> map_segment_reduce_rows_ex
    over peer rows) for the CPU tests: uniq_lists[s] / grad_lists[s] are rank s's sorted unique ids and compact gradient rows.
    Returns (local_rows [U] ascending, merged_grad [U, D]); contributions of a row are summed in source-rank order."""
    shift, _ = merge_key_layout(world, n_shard_rows)
    keys, rows = [], []
    for s_, (u, g) in enumerate(zip(uniq_lists, grad_lists)):
        mine = (u % world) == rank
        keys.append(((u[mine] // world) << shift) | s_)
        rows.append(g[mine])
    keys, rows = torch.cat(keys), torch.cat(rows)
    order = torch.argsort(keys, stable=True)
    keys, rows = keys[order], rows[order]
    local = keys >> shift
    uniq, inv = torch.unique_consecutive(local, return_inverse=True)
    merged = torch.zeros(uniq.numel(), rows.shape[1], dtype=rows.dtype)
    for i in range(keys.numel()):  # explicit order: exactly the order the kernel adds in
        merged[inv[i]] += rows[i]
    return uniq, merged


# ------------------------------------------------------------------------------------------------ NVLink peer memory
class PeerBuffer:
    """One allocation per rank, mapped into every rank of the node: `local` is this rank's tensor, `ptrs` a ctypes array of the
    R device base pointers (valid in THIS process) that the *_sharded / *_ex kernels take."""

    def __init__(self, local: torch.Tensor, ptrs, raw_local: int, raw_peers: List[int]):
        self.local, self.ptrs, self._raw_local, self._raw_peers = local, ptrs, raw_local, raw_peers


class _CudaArrayView:
    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(ptr, False), version=2)


_TYPESTR = {torch.float32: "<f4", torch.int64: "<i8", torch.int32: "<i4"}


class PeerMemory:
    """CUDA-IPC peer memory of one NVSwitch node: map_p2p_alloc on every rank, handles exchanged through the process group,
    map_p2p_open for the other ranks' allocations (lazy peer access).  Collective: every rank calls alloc() in the same order."""

    def __init__(self, world: int, rank: int, device, group=None):
        self.R, self.rank, self.dev, self.group = world, rank, device, group
        self.buffers: List[PeerBuffer] = []

    def alloc(self, shape, dtype) -> PeerBuffer:
        import ctypes as C
        lib = _lib.load()
        n = 1
        for s_ in shape:
            n *= int(s_)
        nbytes = max(n, 1) * torch.empty(0, dtype=dtype).element_size()
        nbytes = (nbytes + 255) // 256 * 256
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        with torch.cuda.device(self.dev):
            _lib.call("map_p2p_alloc", nbytes, C.byref(ptr), handle)
        handles = [None] * self.R
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        raw = []
        for r in range(self.R):
            if r == self.rank:
                raw.append(int(ptr.value))
            else:
                p = C.c_void_p()
                h = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                with torch.cuda.device(self.dev):
                    _lib.call("map_p2p_open", h, C.byref(p))
                raw.append(int(p.value))
        local = torch.as_tensor(_CudaArrayView(raw[self.rank], (max(n, 1),), _TYPESTR[dtype]), device=self.dev)[:n].view(*shape)
        buf = PeerBuffer(local, (C.c_void_p * self.R)(*raw), raw[self.rank], raw)
        self.buffers.append(buf)
        return buf


class ShardedFusedStep(FusedStep):
    """FusedStep with row-sharded tables in NVLink peer memory (default multi-GPU path).

    forward   the embedding gather and the fused NCE kernel read rows of ANY rank's shard directly (peer loads over NVLink):
              no id exchange, no row exchange, no staging;
    backward  every rank reduces its own occurrences to a compact (unique id, gradient row) list exactly like the single-GPU
              step; after a barrier the OWNER of each row pulls the entries it owns from all R lists, sums them in source-rank
              order (deterministic) and applies row-wise AdamW to its shard;
    dense     parameters are replicated; the flat gradient buffer is all-reduced in buckets under the backward pass (finished
              weight gradients form a suffix of the buffer), the rest beside the embedding merge.
    `batch_size` is the per-rank batch; the global batch is batch_size * world.  Philox counters are indexed by the global
    row, so the R-rank run reproduces the single-GPU run on the concatenated batch."""

    # barrier sites of one step: ids sorted everywhere / NCE values complete / embedding values complete / end of step
    # (+ the two barriers of the peer-memory gradient all-reduce: copies complete / reduced slices complete)
    BAR_KEYS, BAR_NCE, BAR_EMBED, BAR_END, BAR_GRAD, BAR_GRAD2, N_BARRIER_SITES = 0, 1, 2, 3, 4, 5, 6

    def __init__(self, model, *, world: int, rank: int, group=None, **kw):
        self.world, self.rank, self.group = world, rank, group
        if world > 8:
            raise NotImplementedError("peer-memory sharding covers one NVSwitch node (<= 8 GPUs)")
        B = kw["batch_size"]
        kw.setdefault("use_graph", False)
        kw["row0"] = rank * B
        kw["global_batch"] = B * world
        self.pm = PeerMemory(world, rank, next(model.parameters()).device, group)
        self._full_shapes: Dict[str, int] = {}
        super().__init__(model, **kw)
        # Barriers: a one-warp kernel over peer-memory flags (map_p2p_barrier; one flag set per barrier site of the schedule, so
        # barriers on different streams may overlap).  MAP_B200_BARRIER=nccl keeps the earlier 4-byte all-reduce on its own
        # communicator (collectives of ONE communicator execute in issue order, and the barriers must not queue behind the
        # 23 MB dense all-reduce that overlaps the embedding merge).
        import os
        self.barrier_kind = os.environ.get("MAP_B200_BARRIER", "p2p")
        self.bar_flags = self.pm.alloc((self.N_BARRIER_SITES * 8,), torch.int32)
        self.bar_epochs = torch.zeros(self.N_BARRIER_SITES, dtype=torch.int32, device=self.dev)
        self.bar_error = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._xr_event = None
        self._bar = torch.zeros(1, dtype=torch.float32, device=self.dev)
        ranks = dist.get_process_group_ranks(group) if group is not None else list(range(world))
        self.bar_group = dist.new_group(ranks=ranks, backend="nccl")
        # the gradient buckets run under the backward GEMMs: on a normal-priority stream NCCL's CTAs queue behind every GEMM wave
        # (r01f 8-GPU timeline: ~100 us per 5.5 MB bucket), so the buckets get a communicator with a high-priority stream
        self.grad_group = group
        if os.environ.get("MAP_B200_NCCL_PRIO", "1") == "1":
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            self.grad_group = dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)
        # Dense gradients: NCCL all-reduce per bucket (default), or MAP_B200_GRAD_AR=p2p: our own all-reduce over peer memory
        # (map_p2p_reduce_f32 / map_p2p_gather_slices_f32: one-shot for small ranges, two-shot above; deterministic rank order, so
        # the replicas stay bit-identical).  Measured (profiles/r02f_*): equal at 2 GPUs (1.048 vs 1.050 ms / step), SLOWER at 8
        # (1.348 vs 1.218 ms): the persistent GEMM CTAs hold every register of every SM, so the copy / barrier / reduce launches of
        # a bucket only start in the gaps between two GEMM launches and the two-shot form needs two such gaps per bucket, while
        # NCCL reduces in the switch with one kernel.  Kept as an option and as the tested building block for a reduce fused into
        # the weight-gradient epilogues.
        self.grad_ar = os.environ.get("MAP_B200_GRAD_AR", "nccl")
        if self.grad_ar == "p2p":
            self.grad_send = self.pm.alloc((self.grad_flat.numel(),), torch.float32)
        if self.multi_stream:
            # all table-side streams of the sharded step are high-priority (the configuration the N > 1 numbers were measured in;
            # the 1-GPU step runs its 'tab' stream at normal priority, see engine.py)
            for k in ("tab", "keys", "nce", "comm"):
                self.streams[k] = torch.cuda.Stream(device=self.dev, priority=-1)

    # -- setup: shards, peer-visible compact gradients, merge plans
    def _shard_param(self, name: str, param: torch.nn.Parameter) -> PeerBuffer:
        V, D = param.shape
        self._full_shapes[name] = V
        rows = shard_rows(V, self.world)
        buf = self.pm.alloc((rows, D), torch.float32)
        buf.local.copy_(shard_table(param.data, self.world, self.rank))
        param.data = buf.local
        param._map_sharded = (self.world, self.rank, V)   # the module path refuses to read a shard (layers.TableEmbedding.forward)
        return buf

    def _peer_table(self, name, param, n_ids, plan):
        V_full = param.shape[0]
        shard = self._shard_param(name, param)
        t = _Table(name, param, n_ids, 0.0 if is_no_decay(name) else self.wd, self.dev)
        t.shard, t.V_full = shard, V_full
        t.gbuf = self.pm.alloc((n_ids, t.D), torch.float32)
        t.grad = t.gbuf.local
        t.plan = plan   # local dedup of this rank's occurrences; the merge reads its uniq / n_unique from every rank
        return t

    def _make_plan(self, n_ids, V_full):
        from . import ops
        bufs = []
        plan = ops.DedupPlan(n_ids, V_full, self.dev, alloc=lambda shape, dtype: self._alloc_peer(bufs, shape, dtype))
        plan.uniq_buf, plan.n_unique_buf = bufs
        return plan

    def _alloc_peer(self, sink, shape, dtype):
        b = self.pm.alloc(shape, dtype)
        sink.append(b)
        return b.local

    def _make_merge(self, plan, shard_rows_):
        """owner-side merge state of one id stream: compacted keys / sources (capacity R * n) and their dedup plan"""
        from . import ops
        R, cap = self.world, plan.n
        m = type("Merge", (), {})()
        m.cap = cap
        m.keys = torch.zeros(R * cap, dtype=torch.int64, device=self.dev)
        m.src = torch.zeros(R * cap, dtype=torch.int32, device=self.dev)
        m.n_owned = torch.zeros(1, dtype=torch.int32, device=self.dev)
        shift, bits = merge_key_layout(R, shard_rows_)
        m.plan = ops.DedupPlan(R * cap, 1 << bits, self.dev, seg_shift=shift, n_dev=m.n_owned)
        return m

    def _make_embed_table(self):
        name = "embed.embedding.weight"
        n = self.B * self.F
        plan = self._make_plan(n, self.embed_w.shape[0])
        t = self._peer_table(name, self.embed_w, n, plan)
        t.merge = self._make_merge(plan, t.p.shape[0])
        t.grad_owned = torch.zeros(self.world * n, t.D, dtype=torch.float32, device=self.dev)
        self.tables[name] = t

    def _make_lr_table(self):
        """DeepFM's first-order table lr_layer.embed_w [V,1]: same id stream as the embedding -> same plan and merge"""
        name = "lr_layer.embed_w.weight"
        te = self.tables["embed.embedding.weight"]
        t = self._peer_table(name, self.lr_w, self.B * self.F, te.plan)
        t.merge = te.merge
        t.grad_owned = torch.zeros(self.world * self.B * self.F, 1, dtype=torch.float32, device=self.dev)
        self.tables[name] = t
        self.w_occ = torch.zeros(self.B * self.F, dtype=torch.float32, device=self.dev)

    def _lr_fm_forward(self, ids, out, ld_out):
        from . import ops
        t = self.tables["lr_layer.embed_w.weight"]
        ops.emb_gather_sharded(t.shard.ptrs, self.world, t.V_full, 1, ids, out=self.w_occ)   # remote 4-byte rows over NVLink
        ops.fm_lr_fwd(self.X0.view(self.B, self.F, self.D), None, self.w_occ, self.lr_b.data, out=out, ld_out=ld_out)

    def _make_nce_tables(self):
        crit = self.model.mfp_criterion
        n_occ = max(self.N, 1) * (self.K + 1)
        plan = self._make_plan(n_occ, crit.emb.weight.shape[0])
        te = self._peer_table("mfp_criterion.emb.weight", crit.emb.weight, n_occ, plan)
        tb = self._peer_table("mfp_criterion.bias.weight", crit.bias.weight, n_occ, plan)
        te.merge = tb.merge = self._make_merge(plan, te.p.shape[0])
        te.grad_owned = torch.zeros(self.world * n_occ, te.D, dtype=torch.float32, device=self.dev)
        tb.grad_owned = torch.zeros(self.world * n_occ, tb.D, dtype=torch.float32, device=self.dev)
        self.tables[te.name], self.tables[tb.name] = te, tb

    # -- exchange primitives
    def _cross_rank(self):
        """Context for every operation that WAITS FOR OTHER RANKS inside a kernel (peer-flag barriers, NCCL collectives): each one
        starts after the previous one of the step has finished, on every rank in the same program order.  Without this chain two
        such kernels can be in flight at once on different streams; a captured graph may map their branches to one hardware
        queue in a different order on different ranks (A: barrier ahead of all-reduce, B: all-reduce ahead of barrier) and the
        ranks then wait for each other for ever (seen at N=2 with the gradient buckets next to the NCE barrier)."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            if self._xr_event is not None:
                torch.cuda.current_stream().wait_event(self._xr_event)
            yield
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._xr_event = ev
        return cm()

    def _barrier(self, site: int):
        """stream-ordered barrier over the ranks: everything every rank issued before it on the calling stream is complete and
        visible to peer loads after it"""
        from . import ops
        with self._cross_rank():
            if self.barrier_kind == "nccl":
                dist.all_reduce(self._bar, op=dist.ReduceOp.SUM, group=self.bar_group)
                _lib.mark("nccl_barrier")
            else:
                ops.p2p_barrier(self.bar_flags.ptrs, self.world, self.rank, site, self.N_BARRIER_SITES, self.bar_epochs, self.bar_error)

    def _merge_keys(self, table):
        """KEY side of the owner-side merge (needs only ids): pull the owned entries of all ranks' unique-id lists and sort them
        by (local row, source rank)"""
        from . import ops
        m, plan = table.merge, table.plan
        ops.owned_compact(plan.uniq_buf.ptrs, plan.n_unique_buf.ptrs, self.world, self.rank, m.cap, m.keys, m.src, m.n_owned)
        m.plan.run(m.keys)

    def _merge_values(self, tables):
        """VALUE side: tables share one id stream (plan + merge); reduce the ranks' compact gradient rows per local row"""
        m = tables[0].merge
        for t in tables:
            m.plan.reduce_peer_rows(t.gbuf.ptrs, self.world, m.cap, t.D, m.src, out=t.grad_owned)

    # -- hooks of FusedStep
    def _embed_lookup(self, ids):
        """Everything of the exchange that depends only on ids runs NOW on the 'keys' stream, under the forward GEMMs: the local
        sorts of both id streams, one barrier, the owners' pulls and their sorts.  What is left for the backward are the value
        reductions (local segment sums -> barrier -> owner-side sums over peer rows)."""
        from . import ops
        t = self.tables["embed.embedding.weight"]
        mfp = self.mode == "MFP"
        self._fork("tab")
        with self._on("tab"):
            self._draw_noise()
            if mfp:
                ops.nce_ids_concat(self.labels.view(-1), self.noise, out=self.ids_all)
            # every Philox consumer of the step has been issued (masks on the main stream before the fork, the noise draw above): the
            # step counter may advance and this step's AdamW coefficients exist from here on — the optimizer kernels of whatever
            # finishes early (NCE tables, reduced gradient buckets) wait for this event, not for the end of the backward
            self._hyper_step()
            self._ev_hyper = torch.cuda.Event()
            self._ev_hyper.record(torch.cuda.current_stream())
            self._fork("keys")
            with self._on("keys"):
                te = self.tables["mfp_criterion.emb.weight"] if mfp else None
                t.plan.run(ids.view(-1))
                if mfp:
                    te.plan.run(self.ids_all.view(-1))
                self._barrier(self.BAR_KEYS)      # every rank's unique-id lists are complete
                self._merge_keys(t)
                if mfp:
                    self._merge_keys(te)
        ops.emb_gather_sharded(t.shard.ptrs, self.world, t.V_full, self.D, ids, out=self.X0)
        self._split(self.X0, self.X0p)

    def _embed_backward(self):
        self._join("keys")
        t = self.tables["embed.embedding.weight"]
        t.plan.reduce_rows(self.dE, self.D, out=t.grad)
        if self.has_fm:
            tl = self.tables["lr_layer.embed_w.weight"]
            tl.plan.reduce_rows(self.d_w_occ, 1, out=tl.grad)

    def _nce_core(self):
        from . import ops
        P, K, L = self.P, self.K, self.L
        crit = self.model.mfp_criterion
        te, tb = self.tables["mfp_criterion.emb.weight"], self.tables["mfp_criterion.bias.weight"]
        self.acc_count.zero_()
        n_global = self.global_batch * L
        ops.nce_fwd(self.sel, self.labels.view(-1), self.noise, None, None, crit.logprob_noise, self.norm_term, self.loss_type,
                    grad_scale=1.0 / n_global, logits=self.logits, want_ids=False, loss_pos=self.loss_pos, dz=self.dz,
                    d_input=self.d_sel, acc_count=self.acc_count, shards=(te.shard.ptrs, tb.shard.ptrs, self.world))
        # the NCE table-gradient chain has its own stream: the embedding backward must not wait for it
        self._fork("nce")
        with self._on("nce"):
            ops.reduce_sum(self.loss_pos, 1.0 / n_global, out=self.loss, ws=self.red_ws)
            self._join("keys")
            te.plan.reduce_rows(self.sel, P, scale=self.dz.view(-1), group=K + 1, out=te.grad, scalar_out=tb.grad.view(-1))
            # every rank has finished reading the NCE tables (its nce_fwd precedes this point) and its compact gradients are
            # complete: owners pull now, off the critical path of the backward pass
            self._barrier(self.BAR_NCE)
            self._merge_values([te, tb])
            # no rank reads the NCE tables again in this step (the barrier above) and the owner-side sums are complete: update the
            # shard here, under the backward GEMMs, instead of at the end of the step
            torch.cuda.current_stream().wait_event(self._ev_hyper)
            for t in (te, tb):
                self._table_adamw(t)
                self._tables_done.add(t.name)

    def _join_streams(self):
        for name in self.streams:   # the NCE merge chain and the gradient buckets keep running; the optimizer joins them
            if name not in ("nce", "comm", "opt"):
                self._join(name)

    # -- dense gradients: bucketed all-reduce under the backward pass.  grad_flat is laid out so that the weight gradients the
    # backward has finished form a SUFFIX of the buffer (engine._alloc); after every grouped launch that completes weight
    # gradients, the newly finished part of the suffix goes out on the 'comm' stream once it is worth a collective.  What is left
    # for the end of the step are the biases and the first tower layers (~4 MB of the 23 MB at C2).
    BUCKET_MIN_FLOATS = 1 << 19   # 2 MB

    def forward_backward(self):
        self._ar_hi = self.grad_flat.numel()      # [_ar_hi, end) has been handed to NCCL
        self._ar_done = set()
        self._ar_wait = []                        # events of gradients finished on side streams (by-field encoder wgrad on 'dw')
        self._adam_done = set()                   # dense parameters already updated behind their gradient bucket
        self._ev_hyper = None
        self._xr_event = None                     # the chain of cross-rank operations restarts with every step
        super().forward_backward()

    def _all_reduce_range(self, lo: int, hi: int):
        self._fork("comm")
        with self._on("comm"), self._cross_rank():
            for ev in getattr(self, "_ar_wait", ()):
                torch.cuda.current_stream().wait_event(ev)
            self._ar_wait = []
            if self.grad_ar == "p2p":
                self._peer_all_reduce(lo, hi)
            else:
                dist.all_reduce(self.grad_flat[lo:hi], op=dist.ReduceOp.SUM, group=self.grad_group)
                _lib.mark("nccl_all_reduce", ("bytes", (hi - lo) * 4))
        if getattr(self, "_ev_hyper", None) is not None:
            # the parameters whose gradients this range holds are final and nothing of the step reads them again (their layers'
            # backward has run: that is why the range could go out): their AdamW follows the collective on the 'comm' stream, so
            # the end of the step only updates what the last bucket carries
            names = tuple(n for n, (plo, phi) in self.grad_offsets.items() if lo <= plo and phi <= hi and n not in self._adam_done)
            if names:
                from . import ops
                # (on the 'opt' stream: the 'comm' stream goes on with the next bucket's exchange at once)
                if self.multi_stream:
                    ev = torch.cuda.Event()
                    ev.record(self.streams["comm"])
                    self.streams["opt"].wait_event(ev)
                    self._forked.add("opt")
                with self._on("opt"):
                    torch.cuda.current_stream().wait_event(self._ev_hyper)
                    ops.adamw_multi_tensor(*self._adam_subset(names), self.hyper)
                self._adam_done.update(names)

    # one-shot (every rank sums all R copies) up to this many bytes of peer reads per rank, two-shot (reduce-scatter in place,
    # barrier, gather) above: R x bytes against 2 x bytes + one more barrier
    ONE_SHOT_MAX_PEER_BYTES = 6 << 20

    def _peer_all_reduce(self, lo: int, hi: int):
        """grad_flat[lo:hi] <- sum over the ranks, through the peer-visible send buffer (called on the 'comm' stream inside the
        cross-rank chain; the chain context is re-entered around each barrier by _barrier itself)."""
        from . import ops
        R, n = self.world, hi - lo
        send = self.grad_send
        send.local[lo:hi].copy_(self.grad_flat[lo:hi])
        self._barrier_nochain(self.BAR_GRAD)          # every rank's copy of the range is complete
        if n * 4 * (R - 1) <= self.ONE_SHOT_MAX_PEER_BYTES or n < 4 * R:
            ops.p2p_reduce(send.ptrs, R, lo, n, self.grad_flat[lo:hi])
        else:
            sl = ((n + R - 1) // R + 3) // 4 * 4      # slice of rank r: [lo + r * sl, lo + min(n, (r + 1) * sl))
            first = min(n, self.rank * sl)
            cnt = min(n, first + sl) - first
            if cnt > 0:
                ops.p2p_reduce(send.ptrs, R, lo + first, cnt, send.local[lo + first:lo + first + cnt])
            self._barrier_nochain(self.BAR_GRAD2)     # every rank's reduced slice is complete
            ops.p2p_gather_slices(send.ptrs, R, lo, n, sl, self.grad_flat[lo:hi])

    def _barrier_nochain(self, site: int):
        """barrier issued from INSIDE a _cross_rank() block (the caller already waited for the previous cross-rank operation and
        records the chain event when it leaves)"""
        from . import ops
        if self.barrier_kind == "nccl":
            dist.all_reduce(self._bar, op=dist.ReduceOp.SUM, group=self.bar_group)
            _lib.mark("nccl_barrier")
        else:
            ops.p2p_barrier(self.bar_flags.ptrs, self.world, self.rank, site, self.N_BARRIER_SITES, self.bar_epochs, self.bar_error)

    def _adam_subset(self, names):
        """device descriptor table of the dense AdamW restricted to `names` (built once per distinct set: the bucket boundaries
        are the same every step; the first step runs eagerly, so nothing is built under graph capture)"""
        from . import ops
        cache = self.__dict__.setdefault("_adam_subsets", {})
        if names not in cache:
            off = {n: self.grad_offsets[n][0] for n in names}
            entries = [(self.opt_param[n], self.grad_flat[off[n]:off[n] + self.opt_param[n].numel()].view_as(self.opt_param[n]),
                        self.exp_avg[n], self.exp_avg_sq[n], 0.0 if is_no_decay(n) else self.wd, None, self.wplanes.get(n)) for n in names]
            cache[names] = ops.make_adamw_tensor_list(entries, self.dev)
        return cache[names]

    def _table_adamw(self, t):
        from . import ops
        if self.optimizer_mode == "sparse":
            ops.adamw_sparse_rows(t.p.data, t.m, t.v, t.merge.plan, t.grad_owned, self.hyper, t.wd)
        else:
            ops.adamw_dense_rows_sparse_grad(t.p.data, t.m, t.v, t.merge.plan, t.grad_owned, self.hyper, t.wd)

    def _field_wgrad(self, W, Kd):
        """by-field encoder: its weight gradient is produced on the 'dw' stream, not by a grouped GEMM launch: mark it finished for
        the bucket logic below and make the next bucket wait for it"""
        super()._field_wgrad(W, Kd)
        if self.multi_stream and hasattr(self, "_ar_hi"):
            self._ar_done.add("feat_encoder.weight")
            ev = torch.cuda.Event()
            ev.record(self.streams["dw"])
            self._ar_wait.append(ev)

    def _gemm_group(self, problems):
        super()._gemm_group(problems)
        if not self.multi_stream or not hasattr(self, "_ar_hi"):
            return
        base, new = self.grad_flat.data_ptr(), False
        for pr in problems:
            if pr is None:
                continue
            off = (pr["C_out"].data_ptr() - base) // 4
            if 0 <= off < self.grad_flat.numel():
                for name, (lo, hi) in self.grad_offsets.items():
                    if lo <= off < hi:
                        self._ar_done.add(name)
                        new = True
        if not new:
            return
        lo_ok = finished_suffix(self.grad_offsets, {n for n, p in self.dense.items() if p.dim() == 1}, self._ar_done, self._ar_hi)
        if self._ar_hi - lo_ok >= self.BUCKET_MIN_FLOATS:
            self._all_reduce_range(lo_ok, self._ar_hi)
            self._ar_hi = lo_ok

    def reduce_gradients(self):
        # every side stream that produced gradients was joined into the current stream by forward_backward()
        self._barrier(self.BAR_EMBED)   # all ranks: embedding gathers done (tables may change), compact embedding gradients complete
        # the rest of the dense gradients (biases, first layers) on the 'comm' stream, beside the embedding merge below
        self._all_reduce_range(0, self._ar_hi)
        self._ar_hi = 0
        self._merge_values([self.tables[n] for n in ("embed.embedding.weight", "lr_layer.embed_w.weight") if n in self.tables])

    def optimizer_step(self):
        from . import ops
        rest = tuple(n for n in self.dense if n not in self._adam_done)
        if rest:                        # (everything went out with its bucket unless the step runs on a single stream)
            with self._on("comm"):
                ops.adamw_multi_tensor(*self._adam_subset(rest), self.hyper)
        self._join("nce")
        self._join("opt")
        for t in self.tables.values():
            if t.name not in self._tables_done:
                self._table_adamw(t)
        # end-of-step barrier: every rank's pulls are complete (compact gradients may be overwritten) and every shard is
        # up to date (the next step's gathers may read it)
        # (the 'comm' stream first: a rank that has passed the last barrier may overwrite its gradient send buffer in the next step,
        #  so every rank's peer reads of this step must be behind that barrier)
        self._join("comm")
        self._barrier(self.BAR_END)

    def dense_table_grad(self, name: str) -> torch.Tensor:
        """[shard_rows, D] dense view of the merged gradient of this rank's shard"""
        t = self.tables[name]
        dense = torch.zeros(t.p.shape[0], t.D, dtype=torch.float32, device=self.dev)
        t.merge.plan.scatter_dense(t.grad_owned, t.D, dense)
        return dense

    def _all_reduce_small(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def outputs(self):
        """global metrics: sums of the per-rank partial means / counts (tiny all-reduces, issued only when metrics are read)"""
        if self.mode == "MFP":
            loss = self._all_reduce_small(self.loss.clone())
            acc = self._all_reduce_small(self.acc_count.clone())
            return (loss.view(()), self.global_batch * self.L, acc.view(()))
        st = self._all_reduce_small(torch.stack([self.stats[0] * (self.B / self.global_batch), self.stats[1], self.stats[2]]))
        cnt = self.global_batch * self.F
        if self.mode == "RFD":
            return (st[0], cnt, st[1] / cnt, st[2] / cnt)
        return (st[0], self.ctr_logits)

    def full_state_dict(self) -> Dict[str, torch.Tensor]:
        """state_dict in the reference's layout: shards of every table are all-gathered and re-interleaved to [V, D]."""
        sd = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        for name, V in self._full_shapes.items():
            shard = sd[name].contiguous()
            parts = torch.empty((self.world,) + tuple(shard.shape), dtype=shard.dtype, device=shard.device)
            dist.all_gather_into_tensor(parts.view(-1), shard.view(-1), group=self.group)
            sd[name] = unshard_table([parts[r] for r in range(self.world)], V)
        return sd


def make_sharded_step(trainer, total_steps: int, warmup_steps: int, world: int, rank: int, group=None) -> ShardedFusedStep:
    """Builds the sharded step from a Trainer's arguments and installs it as the trainer's fused step."""
    cfg, a = trainer.model_config, trainer.args
    beta1, beta2 = (float(b) for b in a.adam_betas.split(","))
    need_x = cfg.pretrain and cfg.pt_type == "RFD" and a.RFD_replace in ("Unigram", "Whole-Unigram")
    lo, hi = getattr(cfg, "idx_low", None), getattr(cfg, "idx_high", None)
    eng = ShardedFusedStep(
        trainer.model, world=world, rank=rank, group=group, batch_size=a.per_gpu_train_batch_size, mask_ratio=a.mask_ratio,
        sampling_method=a.sampling_method, lr=a.learning_rate, weight_decay=a.weight_decay, betas=(beta1, beta2), eps=a.adam_epsilon,
        sched=a.lr_sched, warmup_steps=warmup_steps, total_steps=total_steps, seed=a.seed,
        optimizer_mode=getattr(a, "optimizer_mode", "sparse"), x_train=trainer._train_matrix() if need_x else None,
        idx_low=None if lo is None else lo.to(trainer.device), idx_high=None if hi is None else hi.to(trainer.device))
    trainer._fused = eng
    return eng
