"""Worker for the multi-process tests (launched by torch.distributed.run / mp.spawn).

mode 'plan' (CPU, gloo): every rank fabricates per-partition counts, all_gathers them, runs the planning step of the
    distributed count and checks the tiling invariants of the exchange layout.
mode 'gpu' (NCCL, one GPU per rank): distributed count of a sharded read set; rank 0 reassembles the buckets and
    compares them with the oracle's count of the whole set.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_plan(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spades_b200.distributed import owner_bounds, plan_host
    B, rA, W = 37, 2, 16
    npart = B << rA
    rng = np.random.default_rng(100 + rank)
    local = rng.integers(0, 5000, size=npart).astype(np.uint64)
    local[rng.integers(0, npart, 10)] = 0
    t = torch.from_numpy(local.view(np.int64))
    g = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    allc = torch.stack(g).numpy().view(np.uint64)
    tot = allc.sum(axis=0)
    budget = int(tot.sum() * W * 2 / 3)          # forces several passes
    npass, bounds, mx = plan_host(world, B, rA, allc, budget, W)
    ok = npass >= 1 and bounds[0] == 0 and bounds[-1] == B and all(bounds[i] < bounds[i + 1] for i in range(npass))
    # exchange layout: for every pass and owner the (source, partition) pieces tile [0, recv) without gaps or overlaps
    worst = 0
    for p in range(npass):
        ob = owner_bounds(int(bounds[p]), int(bounds[p + 1]), world)
        for gidx in range(world):
            qlo, qhi = ob[gidx] << rA, ob[gidx + 1] << rA
            pieces = []
            run = 0
            for q in range(qlo, qhi):
                off = run
                for s in range(world):
                    pieces.append((off, int(allc[s, q])))
                    off += int(allc[s, q])
                run += int(tot[q])
            pos = 0
            for off, n in pieces:
                ok &= off == pos
                pos += n
            ok &= pos == int(tot[qlo:qhi].sum())
            worst = max(worst, pos)
    ok &= worst == mx
    # every rank must have computed the same plan
    sig = torch.tensor([npass, mx] + [int(x) for x in bounds], dtype=torch.int64)
    sigs = [torch.empty_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    ok &= all(bool((s == sig).all()) for s in sigs)
    out[rank] = bool(ok) and npass > 1
    dist.destroy_process_group()


def run_gpu():
    import torch
    import torch.distributed as dist
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    import oracle as O
    from spades_b200.distributed import DistributedKMerCounter
    from spades_b200.kmer_index import Context, KMerIndexBuilder
    from spades_b200.packing import pack_reads, synthetic_reads
    ctx = Context(lrank)
    fails = []
    counters = {}
    for (K, B, n, budget) in ((56, 40, 6000, None), (22, 7, 4000, None), (78, 64, 3000, (192 << 20) + 1_500_000), (56, 16, 6000, (192 << 20) + 2_000_000)):
        reads = synthetic_reads(n, 150, 4000, 0.01, seed=K + B)
        mine = reads[rank::world]
        ctx.set_reads(*pack_reads(mine))
        cnt = counters.setdefault(K, DistributedKMerCounter(ctx, K))
        st = cnt.Count(B, budget_bytes=budget)
        keys, counts, bsz = st.kmers(), st.counts(), st.bucket_sizes()
        idx = KMerIndexBuilder(ctx).BuildIndex(st)            # per-rank index over its own buckets
        ids = idx.seq_idx(keys) if len(keys) else np.zeros(0, np.uint64)
        perfect = len(np.unique(ids)) == len(keys)
        gathered = [None] * world
        dist.all_gather_object(gathered, (keys, counts, bsz, perfect, cnt.npass, st.checksum()))
        if rank == 0:
            words, offs, lens = pack_reads(reads)
            ks = O.count(words, offs, lens, K, B, 0)
            tot_bsz = sum(g[2] for g in gathered)
            ok = np.array_equal(tot_bsz, ks.bsz) and all(g[3] for g in gathered)
            # every bucket lives on exactly one rank; reassemble in bucket order
            parts_k, parts_c = [], []
            offs_r = [0] * world
            for b in range(B):
                owners = [r for r in range(world) if gathered[r][2][b] > 0]
                ok &= len(owners) <= 1
                for r in owners:
                    nb = int(gathered[r][2][b])
                    parts_k.append(gathered[r][0][offs_r[r]:offs_r[r] + nb]); parts_c.append(gathered[r][1][offs_r[r]:offs_r[r] + nb])
                    offs_r[r] += nb
            allk = np.concatenate(parts_k) if parts_k else np.zeros((0, ks.nw), np.uint64)
            allc = np.concatenate(parts_c) if parts_c else np.zeros(0, np.uint32)
            ok &= np.array_equal(allk.ravel(), ks.keys.ravel()) and np.array_equal(allc, ks.counts)
            if budget is not None:
                ok &= gathered[0][4] > 1          # the per-pass budget must have forced several passes (192 MB are the planner's fixed reserve)
            # the order-independent device checksums (what bench.py's multi-GPU self check uses) must add / xor up to the union's
            cs = [g[5] for g in gathered]
            tot = [sum(c[0] for c in cs), sum(c[1] for c in cs) & ((1 << 64) - 1), 0, sum(c[3] for c in cs) & ((1 << 64) - 1)]
            for c in cs:
                tot[2] ^= c[2]
            words_sum = int((ks.keys.astype(np.uint64) * (2 * np.arange(ks.nw, dtype=np.uint64) + 1)[None, :]).sum(dtype=np.uint64)) if ks.n else 0
            ok &= tot[0] == ks.n and tot[1] == words_sum and tot[3] == int(ks.counts.astype(np.uint64).sum())
            if not ok:
                fails.append((K, B, n))
            print("dist case K=%d B=%d reads=%d passes=%d distinct=%d %s" % (K, B, n, gathered[0][4], ks.n, "OK" if ok else "MISMATCH"), flush=True)
        idx.free(); st.free()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and fails:
        sys.exit(1)


if __name__ == "__main__":
    try:
        run_gpu()
    except Exception:
        import traceback
        traceback.print_exc()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(1)
