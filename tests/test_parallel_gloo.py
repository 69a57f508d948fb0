"""CPU, world_size = 2, gloo: the data-parallel gradient exchange (clg_vqa_b200.parallel.GradExchange).
N-rank result must equal the single-process result on the concatenated batch: dense ranges are
sum-all-reduced in place, the word-embedding gradient is exchanged as (token ids, rows) and reduced from one list,
no_sync() suppresses communication during gradient accumulation."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clg_vqa_b200.parallel import GradExchange
        torch.manual_seed(100 + rank)
        V, D, NTOK = 50, 8, 12
        dense = torch.randn(4096)                       # stand-in for the dense part of the gradient arena
        ids = torch.randint(0, V, (NTOK,))
        rows = torch.randn(NTOK, D)
        table = torch.zeros(V, D).index_add_(0, ids, rows)      # this rank's own word-table gradient
        ex = GradExchange(world)
        mine_dense = dense.clone()
        # per-"layer" buckets launched early, remainder at the end
        ex.reduce_range(mine_dense, 1024, 2048)
        ex.reduce_range(mine_dense, 2048, 3072)
        ex.reduce_range(mine_dense, 0, 1024)
        ex.reduce_range(mine_dense, 3072, 4096)
        table.zero_()                                           # data-parallel mode: the rank does not scatter its own rows
        ex.add_rows(ids, rows)
        ex.exchange_rows_deterministic(lambda i, r: table.index_add_(0, i, r))
        ex.finish()
        # reference: gather everything and sum
        all_dense = [torch.zeros_like(dense) for _ in range(world)]
        dist.all_gather(all_dense, dense)
        all_ids = [torch.zeros_like(ids) for _ in range(world)]
        all_rows = [torch.zeros_like(rows) for _ in range(world)]
        dist.all_gather(all_ids, ids)
        dist.all_gather(all_rows, rows)
        ref_table = torch.zeros(V, D)
        for i, r in zip(all_ids, all_rows):
            ref_table.index_add_(0, i, r)
        ok = torch.allclose(mine_dense, sum(all_dense), atol=1e-6) and torch.allclose(table, ref_table, atol=1e-6)
        # gradient accumulation: no communication inside no_sync()
        with ex.no_sync():
            keep = dense.clone()
            ex.reduce_range(keep, 0, 4096)
            ex.finish()
        ok = ok and torch.equal(keep, dense)
        # deterministic row exchange: rows of two micro-steps (the first inside no_sync) are stashed, all-gathered once and
        # reduced from the same rank-ordered list on every rank -> bit-identical tables
        ids2 = torch.randint(0, V, (NTOK,)); rows2 = torch.randn(NTOK, D)
        with ex.no_sync():
            ex.add_rows(ids, rows)
        ex.add_rows(ids2, rows2)
        t2 = torch.zeros(V, D)
        seen = {}

        def reduce_rows(i, r):
            seen["n"] = i.numel()
            order = torch.argsort(i, stable=True)
            t2.index_add_(0, i[order], r[order])
        ex.exchange_rows_deterministic(reduce_rows)
        ok = ok and seen.get("n") == 2 * NTOK * world and ex.row_ids == []
        all_i2 = [torch.zeros_like(ids2) for _ in range(world)]; all_r2 = [torch.zeros_like(rows2) for _ in range(world)]
        dist.all_gather(all_i2, ids2); dist.all_gather(all_r2, rows2)
        ref2 = ref_table.clone()
        for i, r in zip(all_i2, all_r2):
            ref2.index_add_(0, i, r)
        ok = ok and torch.allclose(t2, ref2, atol=1e-5)
        all_t2 = [torch.zeros_like(t2) for _ in range(world)]
        dist.all_gather(all_t2, t2)
        ok = ok and all(torch.equal(all_t2[0], t) for t in all_t2)       # replicas bit-identical
        # forward-time plan of the row exchange (plan_rows): duplicated ids keep one carrier row per rank, the padding id
        # is never sent, and the gathered + sorted list reduces to the same table
        PAD = 1
        ids3 = torch.randint(0, 6, (NTOK,)); rows3 = torch.randn(NTOK, D)        # many duplicates, some padding ids
        ex.compact_granule = 1                                   # (256 in production: few distinct all-gather sizes)
        ex.plan_rows(ids3, PAD)
        ex.finalize_plan()
        plan = ex.plan
        K = plan["K"]
        local = torch.zeros(V, D).index_add_(0, ids3, rows3)
        send = plan["all_ids"].view(world, NTOK)[rank]
        keep = send != PAD
        ok = ok and bool((send[keep] == ids3[keep]).all()) and send[keep].unique().numel() == int(keep.sum())
        ok = ok and set(send[keep].tolist()) == set(ids3[ids3 != PAD].tolist())
        ok = ok and K < NTOK and K >= int(keep.sum())            # duplicates and padding do not travel
        carried = torch.where(keep[:, None], local[ids3], torch.full_like(rows3, float("nan")))     # what the local pre-reduction leaves
        carried_c = carried[plan["mine"].long()]                 # compaction: the carriers first, in position order
        ok = ok and bool(torch.isfinite(carried_c[:int(keep.sum())]).all())
        gathered = [torch.zeros_like(carried_c) for _ in range(world)]
        dist.all_gather(gathered, carried_c)
        all_rows = torch.cat(gathered)
        t3 = torch.zeros(V, D)
        so, od = plan["all_sorted"], plan["all_order"]
        live = so != PAD
        t3.index_add_(0, so[live], all_rows[od[live]])
        all_i3 = [torch.zeros_like(ids3) for _ in range(world)]; all_r3 = [torch.zeros_like(rows3) for _ in range(world)]
        dist.all_gather(all_i3, ids3); dist.all_gather(all_r3, rows3)
        ref3 = torch.zeros(V, D)
        for i, r in zip(all_i3, all_r3):
            ref3.index_add_(0, i[i != PAD], r[i != PAD])
        ok = ok and torch.allclose(t3, ref3, atol=1e-5) and bool(torch.equal(so, torch.sort(plan["ids_c"], stable=True).values))
        ok = ok and plan["ids_c"].numel() == world * K
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_grad_exchange_world2_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
