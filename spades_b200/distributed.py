"""Multi-GPU k-mer counting: one process per GPU, torch.distributed for the small tables and the barriers,
the record exchange itself is one kernel over NVLink peer memory (sgpu_dist_exchange: pull + merge).

Mirrors what hpcspades distributes with MPI tasks + a shared filesystem (projects/hpcspades/mpi/stages/construction_mpi.cpp:222-300):
every rank reads its own slice of the reads; afterwards every bucket lives on exactly one rank.
"""
import ctypes as C

import numpy as np

from .kmer_index import KMerDiskStorage, SGPU_CANONICAL

SGPU_IPC_BYTES = 96


class DistributedKMerCounter:
    """KMerDiskCounter over a read set sharded across the ranks of a torch.distributed process group."""

    def __init__(self, ctx, K, mode=SGPU_CANONICAL, group=None):
        self.ctx, self.K, self.mode, self.group = ctx, K, mode, group
        self.npass = 0

    def close(self):
        pass

    def Count(self, num_buckets, budget_bytes=None):
        import torch
        import torch.distributed as dist
        ctx, L = self.ctx, self.ctx.L
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        backend_dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        h = C.c_void_p()
        ctx.check(L.sgpu_dist_begin(ctx.h, self.K, num_buckets, self.mode, world, rank, C.byref(h)))
        try:
            npart = L.sgpu_dist_num_partitions(h)
            local = np.zeros(npart, np.uint64)
            ctx.check(L.sgpu_dist_local_counts(h, local.ctypes.data_as(C.c_void_p)))
            # all_gather of the per-partition record counts (world x npart x 8 bytes: small)
            t_local = torch.from_numpy(local.view(np.int64)).to(backend_dev)
            gathered = [torch.empty_like(t_local) for _ in range(world)]
            dist.all_gather(gathered, t_local, group=self.group)
            all_counts = np.ascontiguousarray(torch.stack(gathered).cpu().numpy().view(np.uint64))
            total = C.c_uint64()
            ctx.check(L.sgpu_dist_plan(h, all_counts.ctypes.data_as(C.c_void_p), C.byref(total)))
            npass = 0
            while True:
                # every pass is planned against what ALL ranks can allocate right now (earlier passes' outputs are resident)
                if budget_bytes is None:
                    fb = C.c_uint64()
                    ctx.check(L.sgpu_dist_free_bytes(h, C.byref(fb)))
                    tb = torch.tensor([fb.value], dtype=torch.int64, device=backend_dev)
                    dist.all_reduce(tb, op=dist.ReduceOp.MIN, group=self.group)
                    budget = int(int(tb.item()) * 0.90)
                else:
                    budget = int(budget_bytes)                 # tests: a fixed per-pass budget
                p = C.c_int()
                ctx.check(L.sgpu_dist_next_pass(h, budget, C.byref(p)))
                if p.value < 0:
                    break
                desc = np.zeros(SGPU_IPC_BYTES, np.uint8)
                ctx.check(L.sgpu_dist_ipc_handle(h, desc.ctypes.data_as(C.c_void_p)))
                t_h = torch.from_numpy(desc).to(backend_dev)
                hs = [torch.empty_like(t_h) for _ in range(world)]
                dist.all_gather(hs, t_h, group=self.group)
                descs = np.ascontiguousarray(torch.stack(hs).cpu().numpy())
                ctx.check(L.sgpu_dist_open_peers(h, descs.ctypes.data_as(C.c_void_p)))     # peers' arenas are mapped once per process
                ctx.check(L.sgpu_dist_scatter(h, p.value))    # local partition into the staging buffer
                dist.barrier(group=self.group)                # every rank's staging buffer is complete
                ctx.check(L.sgpu_dist_exchange(h, p.value))   # one kernel: pull my pieces from all peers over NVLink + merge
                dist.barrier(group=self.group)                # nobody reads my staging buffer any more (it becomes the sort's partner)
                ctx.check(L.sgpu_dist_sort(h, p.value))
                npass += 1
            dist.barrier(group=self.group)
            ks = C.c_void_p()
            ctx.check(L.sgpu_dist_end(h, C.byref(ks)))
            self.npass = npass
            return KMerDiskStorage(ctx, ks)
        finally:
            L.sgpu_dist_free(h)


def plan_host(world, num_buckets, key_bits, all_counts, budget_bytes, record_bytes):
    """The pass / ownership planning alone (pure host arithmetic; used by the CPU gloo tests)."""
    from . import _lib
    L = _lib.load()
    all_counts = np.ascontiguousarray(all_counts, np.uint64)
    bounds = np.zeros(num_buckets + 2, np.int32)
    mx = C.c_uint64()
    n = L.sgpu_dist_plan_host(world, num_buckets, key_bits, all_counts.ctypes.data_as(C.c_void_p), budget_bytes, record_bytes,
                              bounds.ctypes.data_as(C.c_void_p), C.byref(mx))
    if n < 0:
        raise RuntimeError("sgpu_dist_plan_host failed")
    return n, bounds[: n + 1].copy(), int(mx.value)


def owner_bounds(pass_lo, pass_hi, world):
    """contiguous ownership split of a pass's bucket range (same formula as DistPlan::own_lo)."""
    nb = pass_hi - pass_lo
    return [pass_lo + (nb * g) // world for g in range(world + 1)]
