// pair_mailbox.cuh -- sector pairing for scattered 16-byte record stores (NW == 2).
//
// ncu (profiles/r01c_ncu_full_staged_*): the partition kernel moves 2.4x its algorithmic DRAM bytes because a (CTA, partition)
// stream receives a record only every few microseconds; its half-written 32-byte sector leaves the L2 before the other half
// arrives and then costs a fill read plus a second write. A CTA therefore pairs the two records of a sector in shared memory
// and emits ONE 32-byte store (st.global.v4.u64 = STG.E.256 on sm_100a) per pair.
//
// Per stream (partition) a small ring of mailboxes, indexed by the low bits of the sector number. Positions are known up front
// (pos = base[stream] + slot, slot from the stream's running counter), so a record knows its sector and which half it is:
//   EMPTY              -> deposit: claim (WRITING), copy the record in, publish (FULL)
//   FULL, my partner   -> take: claim (TAKING), copy it out, release (EMPTY), store both halves as one 32-byte sector
//   FULL, other sector -> evict: claim (WRITING), copy the old record out, deposit mine, store the old one alone (16 bytes)
//   anything else      -> (a claim in progress, or a lost race) store my record alone
// Nobody ever waits, every record is stored exactly once whatever the interleaving, and what is still FULL at the end is flushed
// as single records. Every state change is a compare-and-swap from the exact value observed, so two claimants cannot both win.
//
// The protocol is exercised on the host by sgpu_selftest op 11 (real threads hammering the same mailboxes; every position must be
// written exactly once with the right record) -- the atomics are macros so that the same source compiles for both.
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define PM_CAS(p, cmp, val) atomicCAS((p), (cmp), (val))
#define PM_XCHG(p, val) atomicExch((p), (val))
#define PM_LOAD(p) (*(volatile uint32_t *)(p))
#define PM_FENCE() __threadfence_block()
#define PM_HD __device__ __forceinline__
#else
static inline uint32_t pm_cas_host(uint32_t *p, uint32_t cmp, uint32_t val) {
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE);
    return cmp;
}
#define PM_CAS(p, cmp, val) pm_cas_host((p), (cmp), (val))
#define PM_XCHG(p, val) __atomic_exchange_n((p), (val), __ATOMIC_ACQ_REL)
#define PM_LOAD(p) __atomic_load_n((p), __ATOMIC_ACQUIRE)
#define PM_FENCE() __atomic_thread_fence(__ATOMIC_SEQ_CST)
#define PM_HD static inline
#endif

namespace sg {

static const int kPmDepth = 2;                   // mailboxes per stream in the level-A kernel (the refinement kernel uses 1)
struct PmBox { uint64_t w0, w1; uint32_t slot; uint32_t state; };      // 24 bytes
// state: 0 = EMPTY, else ((slot + 1) << 2) | phase of the record the box holds / is receiving
enum { kPmWriting = 1, kPmFull = 2, kPmTaking = 3 };

// Sink: void pair(uint64_t first_pos, a0, a1, b0, b1)  -- records at first_pos (even) and first_pos + 1
//       void single(uint64_t pos, w0, w1)
template <int DEPTH, class Sink>
PM_HD void pm_put(PmBox *boxes /* this stream's DEPTH boxes (DEPTH a power of two) */, uint64_t base, uint32_t slot, uint64_t w0, uint64_t w1, Sink &sink) {
    const uint64_t pos = base + slot;
    PmBox *bx = boxes + ((pos >> 1) & (uint64_t)(DEPTH - 1));
    const uint32_t mine_w = ((slot + 1u) << 2) | kPmWriting, mine_f = ((slot + 1u) << 2) | kPmFull;
    // partner = the other half of my sector; it exists in this stream only if its slot is >= 0
    const bool odd = (pos & 1) != 0;
    const bool has_partner = odd ? (slot >= 1u) : true;          // an even record's partner is slot + 1 (may never come: flushed at the end)
    const uint32_t pslot = odd ? slot - 1u : slot + 1u;
    const uint32_t partner_f = ((pslot + 1u) << 2) | kPmFull;
    const uint32_t s = PM_LOAD(&bx->state);
    if (s == 0u) {
        if (PM_CAS(&bx->state, 0u, mine_w) == 0u) {
            bx->w0 = w0; bx->w1 = w1; bx->slot = slot;
            PM_FENCE();
            PM_XCHG(&bx->state, mine_f);
            return;
        }
    } else if (has_partner && s == partner_f) {
        if (PM_CAS(&bx->state, s, (s & ~3u) | kPmTaking) == s) {
            PM_FENCE();
            const uint64_t p0 = bx->w0, p1 = bx->w1;
            PM_FENCE();
            PM_XCHG(&bx->state, 0u);
            if (odd) sink.pair(pos - 1, p0, p1, w0, w1); else sink.pair(pos, w0, w1, p0, p1);
            return;
        }
    } else if ((s & 3u) == (uint32_t)kPmFull) {
        if (PM_CAS(&bx->state, s, mine_w) == s) {
            PM_FENCE();
            const uint64_t o0 = bx->w0, o1 = bx->w1;
            const uint32_t oslot = bx->slot;
            bx->w0 = w0; bx->w1 = w1; bx->slot = slot;
            PM_FENCE();
            PM_XCHG(&bx->state, mine_f);
            sink.single(base + oslot, o0, o1);
            return;
        }
    }
    sink.single(pos, w0, w1);
}

// after every producer is done (barrier): what is still deposited goes out as single records
template <class Sink>
PM_HD void pm_flush_box(PmBox *bx, uint64_t base, Sink &sink) {
    if ((bx->state & 3u) == (uint32_t)kPmFull) { sink.single(base + bx->slot, bx->w0, bx->w1); bx->state = 0u; }
}

}  // namespace sg
