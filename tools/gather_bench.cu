// tools/gather_bench.cu -- measures the B200's random-access roofline for the occurrence-lookup unit of the search
// kernel: uniform random gathers of one aligned 32-byte (or 64-byte) element from an array much larger than L2.
//   mode "indep": every load address comes from a counter hash (unbounded memory-level parallelism)
//   mode "chain": every load address depends on the previous load's data (one dependent chain per lane, like one
//                 search-tree path per lane), with 1..4 chains per lane
// Prints one JSON line per configuration: sectors/s, GB/s.  Timed with CUDA events, 3 warm-ups.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct __align__(32) E32 { uint32_t w[8]; };

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__device__ __forceinline__ void ld32(const E32* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}

template <int BYTES, int UNROLL>
__global__ void indep_kernel(const E32* __restrict__ a, uint64_t n_elems, uint32_t iters, uint32_t* out) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
    uint32_t ctr = tid * 2654435761u;
    for (uint32_t it = 0; it < iters; it++) {
        uint32_t r[UNROLL][8];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            ctr += 0x9e3779b9u;
            uint64_t idx = ((uint64_t)mix(ctr) * n_elems) >> 32;
            if (BYTES == 64) idx &= ~1ull;
            ld32(a + idx, r[u]);
            if (BYTES == 64) { uint32_t t[8]; ld32(a + idx + 1, t); r[u][7] ^= t[0]; }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) acc ^= r[u][0] ^ r[u][7];
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int CHAINS>
__global__ void chain_kernel(const E32* __restrict__ a, uint64_t n_elems, uint32_t iters, uint32_t* out) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t cur[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) cur[c] = mix(tid * CHAINS + c + 1);
    for (uint32_t it = 0; it < iters; it++) {
        uint32_t r[CHAINS][8];
#pragma unroll
        for (int c = 0; c < CHAINS; c++) { uint64_t idx = ((uint64_t)cur[c] * n_elems) >> 32; ld32(a + idx, r[c]); }
#pragma unroll
        for (int c = 0; c < CHAINS; c++) cur[c] = mix(cur[c] + r[c][3] + it);     // data dependent next address
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc ^= cur[c];
    if (acc == 0x12345678u) out[0] = acc;
}

// one 64-byte (PAIR = 2) or 128-byte (PAIR = 4) element per lane group, fetched by ONE warp-level load instruction:
// lane j of a group loads sector j of the group's element, so the request carries 2 (4) sectors of one 128 B line
template <int PAIR>
__global__ void grouped_kernel(const E32* __restrict__ a, uint64_t n_elems, uint32_t iters, uint32_t* out) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
    uint32_t grp = tid / PAIR, sub = tid % PAIR;
    uint32_t ctr = grp * 2654435761u;
    for (uint32_t it = 0; it < iters; it++) {
        uint32_t r[4][8];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            ctr += 0x9e3779b9u;
            uint64_t idx = (((uint64_t)mix(ctr) * (n_elems / PAIR)) >> 32) * PAIR + sub;
            ld32(a + idx, r[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc ^= r[u][0] ^ r[u][7];
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__device__ __forceinline__ void ld32_hint256(const E32* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.nc.L2::256B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__global__ void hint_kernel(const E32* __restrict__ a, uint64_t n_elems, uint32_t iters, uint32_t* out) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
    uint32_t ctr = tid * 2654435761u;
    for (uint32_t it = 0; it < iters; it++) {
        uint32_t r[4][8];
#pragma unroll
        for (int u = 0; u < 4; u++) { ctr += 0x9e3779b9u; uint64_t idx = ((uint64_t)mix(ctr) * n_elems) >> 32; ld32_hint256(a + idx, r[u]); }
#pragma unroll
        for (int u = 0; u < 4; u++) acc ^= r[u][0] ^ r[u][7];
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__global__ void fill_kernel(uint32_t* p, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = mix((uint32_t)i);
}

template <class F> static float time_ms(F f) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; i++) f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < 5; i++) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms; }
    return best;
}

int main(int argc, char** argv) {
    int sms = 148; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); sms = prop.multiProcessorCount;
    uint32_t* out; CK(cudaMalloc(&out, 4));
    const double sizes_gb[] = {0.05, 1.55, 3.1, 12.0};
    for (double gb : sizes_gb) {
        uint64_t n_elems = (uint64_t)(gb * 1e9 / 32);
        E32* a; CK(cudaMalloc(&a, n_elems * 32));
        fill_kernel<<<sms * 8, 256>>>((uint32_t*)a, n_elems * 8); CK(cudaDeviceSynchronize());
        const uint32_t iters = 256;
        for (int tps : {512, 1024, 1536, 2048}) {
            int blocks = sms * tps / 256;
            {
                float ms = time_ms([&] { indep_kernel<32, 4><<<blocks, 256>>>(a, n_elems, iters, out); });
                double loads = (double)blocks * 256 * iters * 4;
                printf("{\"mode\":\"indep\",\"bytes\":32,\"array_gb\":%.2f,\"threads_per_sm\":%d,\"unroll\":4,\"ms\":%.3f,\"gsectors_per_s\":%.2f,\"gb_per_s\":%.1f}\n", gb, tps, ms, loads / ms / 1e6, loads * 32 / ms / 1e6);
            }
            {
                float ms = time_ms([&] { indep_kernel<64, 4><<<blocks, 256>>>(a, n_elems, iters, out); });
                double loads = (double)blocks * 256 * iters * 4;
                printf("{\"mode\":\"indep\",\"bytes\":64,\"array_gb\":%.2f,\"threads_per_sm\":%d,\"unroll\":4,\"ms\":%.3f,\"gelems_per_s\":%.2f,\"gb_per_s\":%.1f}\n", gb, tps, ms, loads / ms / 1e6, loads * 64 / ms / 1e6);
            }
            if (tps == 1024) {
                float ms = time_ms([&] { grouped_kernel<2><<<blocks, 256>>>(a, n_elems, iters, out); });
                double elems = (double)blocks * 256 * iters * 4 / 2;
                printf("{\"mode\":\"grouped\",\"bytes\":64,\"array_gb\":%.2f,\"threads_per_sm\":%d,\"ms\":%.3f,\"gelems_per_s\":%.2f,\"gb_per_s\":%.1f}\n", gb, tps, ms, elems / ms / 1e6, elems * 64 / ms / 1e6);
                ms = time_ms([&] { grouped_kernel<4><<<blocks, 256>>>(a, n_elems, iters, out); });
                elems = (double)blocks * 256 * iters * 4 / 4;
                printf("{\"mode\":\"grouped\",\"bytes\":128,\"array_gb\":%.2f,\"threads_per_sm\":%d,\"ms\":%.3f,\"gelems_per_s\":%.2f,\"gb_per_s\":%.1f}\n", gb, tps, ms, elems / ms / 1e6, elems * 128 / ms / 1e6);
                ms = time_ms([&] { hint_kernel<<<blocks, 256>>>(a, n_elems, iters, out); });
                double loads = (double)blocks * 256 * iters * 4;
                printf("{\"mode\":\"indep_L2_256B_hint\",\"bytes\":32,\"array_gb\":%.2f,\"threads_per_sm\":%d,\"ms\":%.3f,\"gsectors_per_s\":%.2f,\"gb_per_s\":%.1f}\n", gb, tps, ms, loads / ms / 1e6, loads * 32 / ms / 1e6);
            }
            auto chain = [&](int chains, auto kern) {
                float ms = time_ms([&] { kern<<<blocks, 256>>>(a, n_elems, iters, out); });
                double loads = (double)blocks * 256 * iters * chains;
                printf("{\"mode\":\"chain\",\"bytes\":32,\"array_gb\":%.2f,\"threads_per_sm\":%d,\"chains_per_lane\":%d,\"ms\":%.3f,\"gsectors_per_s\":%.2f,\"gb_per_s\":%.1f,\"ns_per_step\":%.1f}\n",
                       gb, tps, chains, ms, loads / ms / 1e6, loads * 32 / ms / 1e6, ms * 1e6 / iters);
            };
            chain(1, chain_kernel<1>); chain(2, chain_kernel<2>); chain(4, chain_kernel<4>);
        }
        CK(cudaFree(a));
        fflush(stdout);
    }
    return 0;
}
