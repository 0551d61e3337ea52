// tools/gather_policy.cu -- which load instruction / L2 fetch granularity keeps DRAM traffic of random 32 B gathers at
// 32 B per lookup?  Uniform random aligned 32-byte gathers over a 3.1 GB array with different PTX load forms; run plain
// for throughput and under `ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum` for traffic.
//   usage: gather_policy [l2_fetch_granularity_bytes]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
struct __align__(32) E32 { uint32_t w[8]; };
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int P> __device__ __forceinline__ void ld32(const E32* p, uint32_t (&r)[8]) {
#define OUTS "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
    if (P == 0) asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : OUTS : "l"(p));
    if (P == 1) asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : OUTS : "l"(p));
    if (P == 2) asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : OUTS : "l"(p));
    if (P == 3) asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : OUTS : "l"(p));
    if (P == 4) asm volatile("ld.global.cv.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : OUTS : "l"(p));
    if (P == 5) {
        asm volatile("ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(p));
        asm volatile("ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(reinterpret_cast<const char*>(p) + 16));
    }
    if (P == 6) asm volatile("ld.global.nc.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : OUTS : "l"(p));
    if (P == 7) asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : OUTS : "l"(p));
    if (P == 8) asm volatile("ld.global.L1::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : OUTS : "l"(p));
#undef OUTS
}

template <int P>
__global__ void k(const E32* __restrict__ a, uint64_t n_elems, uint32_t iters, uint32_t* out) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
    uint32_t ctr = tid * 2654435761u;
    for (uint32_t it = 0; it < iters; it++) {
        uint32_t r[4][8];
#pragma unroll
        for (int u = 0; u < 4; u++) { ctr += 0x9e3779b9u; uint64_t idx = ((uint64_t)mix(ctr) * n_elems) >> 32; ld32<P>(a + idx, r[u]); }
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
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < 3; i++) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms; }
    return best;
}
int main(int argc, char** argv) {
    if (argc > 1) { CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]))); }
    size_t gran = 0; CK(cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); int sms = prop.multiProcessorCount;
    uint32_t* out; CK(cudaMalloc(&out, 4));
    uint64_t n_elems = (uint64_t)(3.1e9 / 32);
    E32* a; CK(cudaMalloc(&a, n_elems * 32));
    fill_kernel<<<sms * 8, 256>>>((uint32_t*)a, n_elems * 8); CK(cudaDeviceSynchronize());
    const uint32_t iters = 128; int blocks = sms * 4;
    const char* names[] = {"ld.global.nc", "ld.global (ca)", "ld.global.cg", "ld.global.nc.L1::no_allocate", "ld.global.cv", "2 x ld.global.nc.v4", "ld.global.nc.L2::64B", "ld.global.L1::no_allocate", "ld.global.L1::evict_first"};
#define RUN(P) { float ms = time_ms([&] { k<P><<<blocks, 256>>>(a, n_elems, iters, out); }); double loads = (double)blocks * 256 * iters * 4; \
    printf("{\"l2_fetch_granularity\":%zu,\"policy\":\"%s\",\"ms\":%.3f,\"glookups_per_s\":%.2f,\"useful_gb_per_s\":%.1f}\n", gran, names[P], ms, loads / ms / 1e6, loads * 32 / ms / 1e6); fflush(stdout); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
    return 0;
}
