// scatter_bench.cu -- microbenchmark behind DESIGN.md 6.1: what bounds scattered record stores on a B200?
//
// Every CTA appends records to S private streams, picking the stream of each record pseudo-randomly (like the level-A partition
// kernel: a shared-memory cursor per stream, slot = atomicAdd). Knobs:
//   -s S        streams per CTA                       (open 128-byte lines per CTA; 2 CTAs of 512 threads per SM)
//   -w 16|32    bytes per store                       (32 = two records of a stream written with one st.global.v4.u64)
//   -l cta|part layout: CTA-major (a CTA's streams are adjacent: its stores stay inside total/G bytes) or partition-major
//               (stream p of every CTA adjacent: a CTA's stores spread over the whole buffer -> TLB reach)
//   -g GB       total bytes written
// Prints GB/s for each configuration so that "streams per CTA", "store width" and "layout" can be separated. Build + run:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scatter_bench scatter_bench.cu && ./scatter_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int WIDTH>
__global__ void __launch_bounds__(512, 2) scatter_k(uint64_t *out, const uint64_t *stream_base /*[G][S] in 16-byte records*/, int S, uint32_t per_stream,
                                                   uint32_t records_per_cta) {
    extern __shared__ uint32_t cur[];          // S cursors
    for (int i = threadIdx.x; i < S; i += blockDim.x) cur[i] = 0;
    __syncthreads();
    const uint64_t *mybase = stream_base + (size_t)blockIdx.x * S;
    constexpr uint32_t RPS = WIDTH / 16;        // records per store
    for (uint32_t i = threadIdx.x; i < records_per_cta / RPS; i += blockDim.x) {
        const uint32_t h = mix(i * 2654435761u + blockIdx.x * 40503u);
        const uint32_t s = h % (uint32_t)S;
        const uint32_t slot = atomicAdd(&cur[s], RPS);
        if (slot + RPS > per_stream) continue;                   // stream full (the random pick is only balanced on average)
        uint64_t *dst = out + (mybase[s] + slot) * 2;
        const uint64_t a = h, b = ~(uint64_t)h;
        if (WIDTH == 16) asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(a), "l"(b) : "memory");
        else asm volatile("st.global.L1::no_allocate.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "l"(a), "l"(b), "l"(a + 1), "l"(b + 1) : "memory");
    }
}

static float run(int width, int S, bool cta_major, double gb, int G, uint64_t *d_out, size_t out_bytes) {
    const uint64_t total_rec = (uint64_t)(gb * 1e9 / 16);
    const uint32_t per_cta = (uint32_t)(total_rec / G);
    uint32_t per_stream = (uint32_t)((uint64_t)per_cta * 5 / 4 / S + 8) & ~1u;      // 25 % slack, even (32-byte alignment of every stream)
    if ((uint64_t)per_stream * S * G * 16 > out_bytes) { fprintf(stderr, "buffer too small\n"); exit(1); }
    std::vector<uint64_t> base((size_t)G * S);
    for (int g = 0; g < G; ++g)
        for (int s = 0; s < S; ++s) base[(size_t)g * S + s] = cta_major ? ((uint64_t)g * S + s) * per_stream : ((uint64_t)s * G + g) * per_stream;
    uint64_t *d_base;
    CK(cudaMalloc(&d_base, base.size() * 8));
    CK(cudaMemcpy(d_base, base.data(), base.size() * 8, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        if (width == 16) scatter_k<16><<<G, 512, S * 4>>>(d_out, d_base, S, per_stream, per_cta);
        else scatter_k<32><<<G, 512, S * 4>>>(d_out, d_base, S, per_stream, per_cta);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    CK(cudaFree(d_base));
    return (float)((double)per_cta * G * 16 / 1e9 / (best / 1e3));
}

int main(int argc, char **argv) {
    double gb = 8.0;
    for (int i = 1; i + 1 < argc; i += 2) if (!strcmp(argv[i], "-g")) gb = atof(argv[i + 1]);
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int G = p.multiProcessorCount * 2;
    const size_t out_bytes = (size_t)(gb * 1.4e9) + (64u << 20);
    uint64_t *d_out;
    CK(cudaMalloc(&d_out, out_bytes));
    CK(cudaMemset(d_out, 0, out_bytes));
    CK(cudaFuncSetAttribute(scatter_k<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 << 10));
    CK(cudaFuncSetAttribute(scatter_k<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 << 10));
    printf("%s, %d CTAs x 512 threads, %.1f GB per run; GB/s of record bytes\n", p.name, G, gb);
    printf("%8s | %12s %12s | %12s %12s\n", "streams", "16B cta-maj", "32B cta-maj", "16B part-maj", "32B part-maj");
    const int Ss[] = {32, 64, 128, 256, 512, 1024, 2048, 4096, 8192};
    for (int S : Ss) {
        printf("%8d | %12.0f %12.0f | %12.0f %12.0f\n", S, run(16, S, true, gb, G, d_out, out_bytes), run(32, S, true, gb, G, d_out, out_bytes),
               run(16, S, false, gb, G, d_out, out_bytes), run(32, S, false, gb, G, d_out, out_bytes));
        fflush(stdout);
    }
    CK(cudaFree(d_out));
    return 0;
}
