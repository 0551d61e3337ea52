// tools/gather_width.cu -- DRAM / L2 sectors fetched per random lookup as a function of the load width.
// One random 32-byte-aligned element per lookup over a 3.1 GB array; W selects how it is read:
//   0: one 32-bit load        1: one 64-bit load      2: one 128-bit load       3: two 128-bit loads (whole 32 B)
//   4: one 256-bit load       5: four 64-bit loads    6: eight 32-bit loads
// Run under: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
struct __align__(32) E32 { uint32_t w[8]; };
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int W> __device__ __forceinline__ uint32_t ld(const E32* p) {
    uint32_t r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const char* c = reinterpret_cast<const char*>(p);
    if (W == 0) asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(r[0]) : "l"(c));
    if (W == 1) asm volatile("ld.global.nc.v2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "l"(c));
    if (W == 2) asm volatile("ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(c));
    if (W == 3) {
        asm volatile("ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(c));
        asm volatile("ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(c + 16));
    }
    if (W == 4) asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(c));
    if (W == 5) for (int i = 0; i < 4; i++) asm volatile("ld.global.nc.v2.b32 {%0,%1}, [%2];" : "=r"(r[2 * i]), "=r"(r[2 * i + 1]) : "l"(c + 8 * i));
    if (W == 6) for (int i = 0; i < 8; i++) asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(r[i]) : "l"(c + 4 * i));
    return r[0] ^ r[1] ^ r[2] ^ r[3] ^ r[4] ^ r[5] ^ r[6] ^ r[7];
}
template <int W>
__global__ void kw(const E32* __restrict__ a, uint64_t n_elems, uint32_t iters, uint32_t* out) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
    uint32_t ctr = tid * 2654435761u;
    for (uint32_t it = 0; it < iters; it++) {
        uint32_t v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { ctr += 0x9e3779b9u; uint64_t idx = ((uint64_t)mix(ctr) * n_elems) >> 32; v[u] = ld<W>(a + idx); }
        acc ^= v[0] ^ v[1] ^ v[2] ^ v[3];
    }
    if (acc == 0x12345678u) out[0] = acc;
}
__global__ void fill_kernel(uint32_t* p, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = mix((uint32_t)i);
}
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); int sms = prop.multiProcessorCount;
    uint32_t* out; CK(cudaMalloc(&out, 4));
    uint64_t n_elems = (uint64_t)(3.1e9 / 32);
    E32* a; CK(cudaMalloc(&a, n_elems * 32));
    fill_kernel<<<sms * 8, 256>>>((uint32_t*)a, n_elems * 8); CK(cudaDeviceSynchronize());
    const uint32_t iters = 128; int blocks = sms * 4;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
#define RUN(W) { kw<W><<<blocks, 256>>>(a, n_elems, iters, out); CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0)); kw<W><<<blocks, 256>>>(a, n_elems, iters, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); \
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); double loads = (double)blocks * 256 * iters * 4; printf("{\"width_mode\":%d,\"ms\":%.3f,\"glookups_per_s\":%.2f}\n", W, ms, loads / ms / 1e6); fflush(stdout); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
    return 0;
}
