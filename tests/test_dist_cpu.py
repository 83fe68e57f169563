"""world_size-2 gloo tests (CPU) of the multi-GPU choreography in map_code_b200/dist.py: ShardExchange drives fixed-size
collectives around five owner-side kernels.  Here the kernels are emulated with torch ops (this file, test infrastructure)
and the result of the 2-rank sharded exchange is compared with the single-process oracle on the concatenated batch —
the parity oracle of SURVEY.md §8(e)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import map_oracle as O


class CpuKernels:
    """torch emulation of the C-ABI owner-side kernels (same signatures as map_code_b200.ops)."""

    @staticmethod
    def emb_gather_owned(shard, ids, R, rank, out):
        own = (ids % R) == rank
        rows = shard[torch.where(own, ids // R, torch.zeros_like(ids))]
        out.copy_(torch.where(own[:, None], rows, torch.zeros_like(rows)))
        return out

    @staticmethod
    def owned_keys(ids, R, rank, sentinel, out):
        out.copy_(torch.where((ids % R) == rank, ids // R, torch.full_like(ids, sentinel)))
        return out

    @staticmethod
    def nce_scores_owned(q, idx, emb_shard, bias_shard, R, rank, out):
        own = (idx % R) == rank
        loc = torch.where(own, idx // R, torch.zeros_like(idx))
        s = (q[:, None, :] * emb_shard[loc]).sum(-1) + bias_shard[loc]
        out.copy_(torch.where(own, s, torch.zeros_like(s)))
        return out

    @staticmethod
    def nce_dinput_owned(dz, idx, emb_shard, R, rank, out):
        own = (idx % R) == rank
        loc = torch.where(own, idx // R, torch.zeros_like(idx))
        out.copy_((torch.where(own, dz, torch.zeros_like(dz))[:, :, None] * emb_shard[loc]).sum(1))
        return out


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world)
    finally:
        dist.destroy_process_group()


def run2(fn, world=2):
    mp.spawn(_worker, args=(world, _free_port(), fn), nprocs=world, join=True)


def _problem():
    g = torch.Generator().manual_seed(0)
    V, D, P, K, F, Bl, L = 103, 8, 4, 5, 6, 10, 2
    table = torch.randn(V, D, generator=g)
    emb = torch.randn(V, P, generator=g)
    bias = torch.randn(V, 1, generator=g)
    fc = torch.randint(1, 50, (V,), generator=g).float()
    return g, V, D, P, K, F, Bl, L, table, emb, bias, fc


def _embedding_case(rank, world):
    from map_code_b200.dist import ShardExchange, shard_rows, shard_table, unshard_table
    g, V, D, P, K, F, Bl, L, table, emb, bias, fc = _problem()
    ids_all = torch.randint(0, V, (world * Bl, F), generator=g)
    ids_all[:, 0] = 3  # the hot <mask> row lives on one rank only
    dE_all = torch.randn(world * Bl, F * D, generator=g)
    ids = ids_all[rank * Bl:(rank + 1) * Bl].contiguous()
    xch = ShardExchange(CpuKernels, world, rank)
    shard = shard_table(table, world, rank)
    assert shard.shape[0] == shard_rows(V, world)
    ids_g = torch.empty(world, Bl, F, dtype=torch.int64)
    rows_g = torch.empty(world * Bl * F, D)
    out = torch.empty(Bl * F, D)
    xch.embed_forward(shard, ids, ids_g, rows_g, out)
    assert torch.equal(out.view(Bl, F, D), O.embeddings_forward(table, ids))           # lookup: bit-exact
    assert torch.equal(ids_g.view(-1, F), ids_all)
    # backward: every rank ends up with the gradient rows it owns, equal to the dense reference gradient of the GLOBAL batch
    dE_g = xch.all_gather(dE_all[rank * Bl:(rank + 1) * Bl].contiguous())
    keys = torch.empty(world * Bl * F, dtype=torch.int64)
    sentinel = shard.shape[0] - 1
    xch.local_keys(ids_g, sentinel, keys)
    grad_shard = torch.zeros(shard.shape[0], D).index_add_(0, keys, dE_g.view(-1, D))
    ref = torch.zeros(V, D).index_add_(0, ids_all.view(-1), dE_all.view(-1, D))
    n_mine = (V - rank + world - 1) // world
    torch.testing.assert_close(grad_shard[:n_mine], ref[rank::world])
    # the sentinel row collected exactly the foreign occurrences
    foreign = (ids_all.view(-1) % world) != rank
    torch.testing.assert_close(grad_shard[sentinel], dE_all.view(-1, D)[foreign].sum(0), rtol=1e-5, atol=1e-5)
    # checkpoint export: all shards -> the reference's [V, D] layout
    parts = xch.all_gather(shard)
    assert torch.equal(unshard_table([parts[r] for r in range(world)], V), table)


def _peer_merge_case(rank, world):
    """The default multi-GPU path (tables in NVLink peer memory): every rank dedups its OWN occurrences, the owner of a row
    pulls the matching entries of all ranks' compact lists and sums them in source-rank order.  Peer memory is emulated by an
    all_gather_object of the lists; the result must equal the dense reference gradient of the global batch."""
    from map_code_b200.dist import merge_key_layout, peer_merge_torch, shard_rows
    g, V, D, P, K, F, Bl, L, table, emb, bias, fc = _problem()
    ids_all = torch.randint(0, V, (world * Bl, F), generator=g)
    ids_all[:, 0] = 3
    dE_all = torch.randn(world * Bl, F * D, generator=g)
    ids = ids_all[rank * Bl:(rank + 1) * Bl].reshape(-1)
    dE = dE_all[rank * Bl:(rank + 1) * Bl].reshape(-1, D)
    uniq, inv = torch.unique(ids, return_inverse=True)                       # local dedup (sorted ascending, like the radix sort)
    compact = torch.zeros(uniq.numel(), D).index_add_(0, inv, dE)
    lists = [None] * world
    dist.all_gather_object(lists, (uniq, compact))                           # "peer loads"
    n_rows = shard_rows(V, world)
    shift, bits = merge_key_layout(world, n_rows)
    assert (1 << shift) >= world and ((n_rows - 1) << shift | (world - 1)) < (1 << bits)
    rows, merged = peer_merge_torch([l[0] for l in lists], [l[1] for l in lists], world, rank, n_rows)
    ref = torch.zeros(V, D).index_add_(0, ids_all.view(-1), dE_all.view(-1, D))[rank::world]
    touched = torch.zeros(V, dtype=torch.bool)
    touched[ids_all.view(-1)] = True
    want_rows = torch.nonzero(touched[rank::world]).view(-1)
    assert torch.equal(rows, want_rows)                                       # exactly the owned rows the global batch touched
    torch.testing.assert_close(merged, ref[want_rows], rtol=1e-5, atol=1e-5)


def _nce_case(rank, world):
    from map_code_b200.dist import ShardExchange, shard_table
    g, V, D, P, K, F, Bl, L, table, emb, bias, fc = _problem()
    N = Bl * L
    _, logq, norm = O.nce_noise_distribution(fc)
    q_all = torch.randn(world * N, P, generator=g)
    target_all = torch.randint(0, V, (world * N, 1), generator=g)
    noise_all = torch.randint(0, V, (world * N, K), generator=g)
    idx_all = torch.cat([target_all, noise_all], 1)
    sl = slice(rank * N, (rank + 1) * N)
    xch = ShardExchange(CpuKernels, world, rank)
    emb_s, bias_s = shard_table(emb, world, rank), shard_table(bias, world, rank)
    q_g, ids_g = torch.empty(world, N, P), torch.empty(world, N, K + 1, dtype=torch.int64)
    partial, scores = torch.empty(world * N, K + 1), torch.empty(N, K + 1)
    xch.nce_scores(q_all[sl].contiguous(), idx_all[sl].contiguous(), emb_s, bias_s, q_g, ids_g, partial, scores)
    # oracle on the concatenated batch (global mean over world*N positions)
    embr, biasr, qr = emb.clone().requires_grad_(True), bias.clone().requires_grad_(True), q_all.clone().requires_grad_(True)
    loss, logits, ids = O.nce_forward(embr, biasr, logq, norm, target_all, noise_all.view(world * N, 1, K), qr.view(world * N, 1, P))
    loss.backward()
    torch.testing.assert_close(scores - norm, logits.view(world * N, K + 1)[sl].detach(), rtol=1e-5, atol=1e-5)
    # local loss / dz from complete scores (what map_nce_loss_from_scores computes), scaled by 1/N_global
    z = (scores - norm) - logq[idx_all[sl]] - torch.log(torch.tensor(float(K)))
    y = torch.zeros_like(z)
    y[:, 0] = 1
    dz = (torch.sigmoid(z) - y) / (world * N)
    part = torch.nn.functional.binary_cross_entropy_with_logits(z, y, reduction="none").sum() / (world * N)
    tot = part.clone()
    dist.all_reduce(tot)
    torch.testing.assert_close(tot, loss.detach(), rtol=1e-5, atol=1e-6)               # sum of partial means == global mean
    dz_g, dq_partial, dq = torch.empty(world, N, K + 1), torch.empty(world * N, P), torch.empty(N, P)
    xch.nce_dinput(dz.contiguous(), ids_g, emb_s, dz_g, dq_partial, dq)
    torch.testing.assert_close(dq, qr.grad[sl], rtol=1e-4, atol=1e-7)
    # table gradients of the owned rows
    keys = torch.empty(world * N * (K + 1), dtype=torch.int64)
    sentinel = emb_s.shape[0] - 1
    xch.local_keys(ids_g, sentinel, keys)
    contrib = dz_g.view(-1, 1) * q_g.view(world * N, P).repeat_interleave(K + 1, 0)
    g_emb = torch.zeros(emb_s.shape[0], P).index_add_(0, keys, contrib)
    g_bias = torch.zeros(emb_s.shape[0]).index_add_(0, keys, dz_g.view(-1))
    n_mine = (V - rank + world - 1) // world
    torch.testing.assert_close(g_emb[:n_mine], embr.grad[rank::world], rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(g_bias[:n_mine], biasr.grad[rank::world, 0], rtol=1e-4, atol=1e-7)


def _rng_sharding_case(rank, world):
    """Philox counters are indexed by the global row: rank shards of the draws concatenate to the single-process draws."""
    Bl, L, F, K = 16, 3, 39, 5
    mi = O.draw_masked_index(42, 8 * 7, Bl, L, F, "randint", row0=rank * Bl)
    parts = [torch.empty_like(mi) for _ in range(world)]
    dist.all_gather(parts, mi)
    assert torch.equal(torch.cat(parts), O.draw_masked_index(42, 8 * 7, world * Bl, L, F, "randint"))
    prob, alias = O.alias_build(torch.full((50,), 0.02))
    nz = O.alias_draw(prob, alias, 42, 8 * 7 + 2, Bl * L * K, elem0=rank * Bl * L * K)
    parts = [torch.empty_like(nz) for _ in range(world)]
    dist.all_gather(parts, nz)
    assert torch.equal(torch.cat(parts), O.alias_draw(prob, alias, 42, 8 * 7 + 2, world * Bl * L * K))


def test_sharded_embedding_exchange_world2():
    run2(_embedding_case)


def test_sharded_nce_exchange_world2():
    run2(_nce_case)


def test_sharded_embedding_exchange_world3_uneven():
    run2(_embedding_case, world=3)  # V = 103 is not a multiple of 3: shards have different numbers of live rows


def test_philox_streams_shard_by_global_row():
    run2(_rng_sharding_case)


def test_shard_table_roundtrip():
    from map_code_b200.dist import shard_table, unshard_table
    t = torch.arange(35.0).view(7, 5)
    for R in (1, 2, 3, 4, 8):
        assert torch.equal(unshard_table([shard_table(t, R, r) for r in range(R)], 7), t)


def test_peer_memory_merge_world2():
    run2(_peer_merge_case)


def test_peer_memory_merge_world3_uneven():
    run2(_peer_merge_case, world=3)


def test_gradient_bucket_suffix_logic():
    """dist.finished_suffix: which part of the flat gradient buffer may go to NCCL after each level of the backward pass
    (layout of engine.FusedStep._alloc for DCNv2-MFP: biases first, then weights in reverse backward order)."""
    from map_code_b200.dist import finished_suffix
    from map_code_b200.engine import FusedStep
    names = ["cross_net.cross_layers.0.bias", "parallel_dnn.dnn.0.bias", "feat_encoder.bias",
             "cross_net.cross_layers.0.weight", "parallel_dnn.dnn.0.weight", "cross_net.cross_layers.1.weight", "parallel_dnn.dnn.3.weight",
             "cross_net.cross_layers.2.weight", "parallel_dnn.dnn.6.weight", "feat_encoder.weight"]
    # the sort key of the engine puts them in exactly this order
    depth = [FusedStep._layer_depth(n) for n in names[3:]]
    assert depth == sorted(depth) and depth[-1] == 99
    offs, o = {}, 0
    for i, n in enumerate(names):
        offs[n] = (o, o + 100 * (i + 1))
        o += 100 * (i + 1)
    biases = set(names[:3])
    hi = o
    done = set()
    assert finished_suffix(offs, biases, done, hi) == hi                          # nothing finished yet
    done.add("feat_encoder.weight")
    hi1 = finished_suffix(offs, biases, done, hi)
    assert hi1 == offs["feat_encoder.weight"][0]
    done.add("parallel_dnn.dnn.6.weight")                                          # cross layer 2 still missing: no progress past dnn.6
    assert finished_suffix(offs, biases, done, hi1) == offs["parallel_dnn.dnn.6.weight"][0]
    done.add("cross_net.cross_layers.1.weight")                                    # out of order: a hole below dnn.6 -> no further progress
    assert finished_suffix(offs, biases, done, hi1) == offs["parallel_dnn.dnn.6.weight"][0]
    done.update(names[3:])
    assert finished_suffix(offs, biases, done, hi1) == offs["cross_net.cross_layers.0.weight"][0]   # stops at the biases
    done.update(names)
    assert finished_suffix(offs, biases, done, hi1) == offs["cross_net.cross_layers.0.weight"][0]
