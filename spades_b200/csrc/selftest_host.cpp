// selftest_host.cpp -- plain host C++ (compiled without the CUDA front end): self tests that need real host threads.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/spades_b200.h"
#include "pair_mailbox.cuh"

using namespace sg;

// op 11 (host only): the sector-pairing mailbox protocol of pair_mailbox.cuh under real concurrency. arg = (threads << 32) | streams,
// n = records per thread. Every thread appends records to random streams exactly like the partition kernel does (slot from the
// stream's counter, position = base + slot) and pushes them through pm_put; afterwards the leftovers are flushed. Every position
// must have been written exactly once with its own record. out[0] = errors, out[1] = pair stores, out[2] = single stores.
namespace {
struct PmTestSink {
    uint64_t *X; uint32_t *written; uint64_t *pairs, *singles;
    static uint64_t f0(uint64_t pos) { return pos * 0x9E3779B185EBCA87ull + 1; }
    static uint64_t f1(uint64_t pos) { return ~pos ^ 0x5bd1e995ull; }
    void put(uint64_t pos, uint64_t a, uint64_t b) { X[2 * pos] = a; X[2 * pos + 1] = b; __atomic_fetch_add(&written[pos], 1u, __ATOMIC_RELAXED); }
    void pair(uint64_t pos, uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1) { put(pos, a0, a1); put(pos + 1, b0, b1); __atomic_fetch_add(pairs, 1ull, __ATOMIC_RELAXED); }
    void single(uint64_t pos, uint64_t a0, uint64_t a1) { put(pos, a0, a1); __atomic_fetch_add(singles, 1ull, __ATOMIC_RELAXED); }
};
}
extern "C" int sg_selftest_pair_mailbox(uint64_t arg, int64_t per_thread, uint64_t *out) {
    const int T = (int)(arg >> 32), S = (int)(arg & 0xffffffffu);
    if (!(T >= 1 && T <= 256 && S >= 1 && S <= (1 << 20) && per_thread >= 1)) return SGPU_EINVAL;
    std::vector<std::vector<uint32_t>> plan(T);
    std::vector<uint64_t> count(S, 0), base(S + 1, 0);
    for (int t = 0; t < T; ++t) {
        uint64_t x = 0x243F6A8885A308D3ull ^ (uint64_t)(t + 1) * 0x9E3779B97F4A7C15ull;
        plan[t].resize((size_t)per_thread);
        for (int64_t i = 0; i < per_thread; ++i) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            // skewed: a few hot streams (long runs -> many pairs) and many cold ones (leftovers, evictions)
            const uint32_t p = (x & 3) ? (uint32_t)((x >> 8) % (uint64_t)std::max(1, S / 16)) : (uint32_t)((x >> 8) % (uint64_t)S);
            plan[t][(size_t)i] = p; count[p]++;
        }
    }
    uint64_t run = 0;
    for (int p = 0; p < S; ++p) { run += (uint64_t)(p % 3 == 1);  base[p] = run; run += count[p]; }     // bases of either parity, small gaps
    const uint64_t total = run + 2;
    std::vector<uint64_t> X(2 * total, 0);
    std::vector<uint32_t> written(total, 0), cnt((size_t)S, 0);
    std::vector<PmBox> boxes((size_t)S * kPmDepth);
    memset(boxes.data(), 0, boxes.size() * sizeof(PmBox));
    uint64_t pairs = 0, singles = 0;
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t]() {
            PmTestSink sink{X.data(), written.data(), &pairs, &singles};
            for (uint32_t p : plan[t]) {
                const uint32_t slot = __atomic_fetch_add(&cnt[p], 1u, __ATOMIC_RELAXED);
                const uint64_t pos = base[p] + slot;
                pm_put<kPmDepth>(boxes.data() + (size_t)p * kPmDepth, base[p], slot, PmTestSink::f0(pos), PmTestSink::f1(pos), sink);
            }
        });
    for (auto &t : th) t.join();
    PmTestSink sink{X.data(), written.data(), &pairs, &singles};
    for (int p = 0; p < S; ++p)
        for (int d = 0; d < kPmDepth; ++d) pm_flush_box(&boxes[(size_t)p * kPmDepth + d], base[p], sink);
    uint64_t errors = 0;
    for (int p = 0; p < S; ++p) {
        if (cnt[p] != count[p]) ++errors;
        for (uint64_t i = 0; i < count[p]; ++i) {
            const uint64_t pos = base[p] + i;
            if (written[pos] != 1 || X[2 * pos] != PmTestSink::f0(pos) || X[2 * pos + 1] != PmTestSink::f1(pos)) ++errors;
        }
    }
    uint64_t stray = 0;
    for (uint64_t pos = 0; pos < total; ++pos) stray += written[pos];
    if (stray != (uint64_t)T * (uint64_t)per_thread) ++errors;          // nothing written outside the streams
    out[0] = errors; out[1] = pairs; out[2] = singles;
    return SGPU_OK;
}

