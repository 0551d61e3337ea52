// tools/slice_gather.cu -- random gathers confined to an L2-sized slice that moves across a large array.
// Question: if the search visits the index slice by slice (all guides' lookups that fall into one slice, then the next
// slice), how many random lookups per second does the B200 sustain, as a function of slice size, touches per line and
// sectors read per lookup?  Compare with the uniform-random figure over the whole array (tools/gather_bench.cu).
//   array: 6.2 GB of 128-byte lines; slice = S MB; per slice every thread issues random line reads inside it;
//   reads per slice = touches * lines_in_slice;  SECT = 32-byte sectors read from each line (1, 2 or 4)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__device__ __forceinline__ uint32_t ld32(const char* c) {
    uint32_t r[8];
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(c));
    return r[0] ^ r[1] ^ r[2] ^ r[3] ^ r[4] ^ r[5] ^ r[6] ^ r[7];
}
template <int SECT>
__global__ void slice_kernel(const char* __restrict__ a, uint64_t n_lines, uint64_t lines_per_slice, uint32_t reads_per_thread_per_slice, uint32_t* out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0, ctr = tid * 2654435761u;
    const uint64_t n_slices = (n_lines + lines_per_slice - 1) / lines_per_slice;
    for (uint64_t s = 0; s < n_slices; s++) {
        const uint64_t base = s * lines_per_slice;
        const uint64_t cnt = (base + lines_per_slice <= n_lines) ? lines_per_slice : (n_lines - base);
        for (uint32_t it = 0; it < reads_per_thread_per_slice; it += 4) {
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                ctr += 0x9e3779b9u;
                const uint64_t line = base + (((uint64_t)mix(ctr) * cnt) >> 32);
                const char* p = a + line * 128;
                uint32_t x = ld32(p);
                if (SECT >= 2) x ^= ld32(p + 32);
                if (SECT >= 4) { x ^= ld32(p + 64); x ^= ld32(p + 96); }
                v[u] = x;
            }
            acc ^= v[0] ^ v[1] ^ v[2] ^ v[3];
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
__global__ void fill_kernel(uint32_t* p, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = mix((uint32_t)i);
}
template <int SECT>
static void run(const char* a, uint64_t n_lines, double slice_mb, double touches, int sms, int tpsm, uint32_t* out) {
    const int threads = 256, blocks = sms * (tpsm / threads);
    const uint64_t lps = slice_mb <= 0 ? n_lines : (uint64_t)(slice_mb * 1e6 / 128);
    const uint64_t n_slices = (n_lines + lps - 1) / lps;
    uint32_t rpt = (uint32_t)(touches * (double)lps / ((double)blocks * threads));
    rpt = (rpt + 3) & ~3u; if (rpt < 4) rpt = 4;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    slice_kernel<SECT><<<blocks, threads>>>(a, n_lines, lps, rpt, out); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    slice_kernel<SECT><<<blocks, threads>>>(a, n_lines, lps, rpt, out);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double reads = (double)blocks * threads * rpt * (double)n_slices;
    printf("{\"slice_mb\":%.0f,\"touches_per_line\":%.1f,\"sectors_per_lookup\":%d,\"threads_per_sm\":%d,\"ms\":%.2f,\"glookups_per_s\":%.2f,\"useful_gb_per_s\":%.0f}\n",
           slice_mb, (double)rpt * blocks * threads / (double)lps, SECT, tpsm, ms, reads / ms / 1e6, reads * 32.0 * SECT / ms / 1e6);
    fflush(stdout);
}
int main(int argc, char** argv) {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); const int sms = prop.multiProcessorCount;
    uint32_t* out; CK(cudaMalloc(&out, 4));
    const uint64_t n_lines = (uint64_t)(6.2e9 / 128);
    char* a; CK(cudaMalloc(&a, n_lines * 128));
    fill_kernel<<<sms * 8, 256>>>((uint32_t*)a, n_lines * 32); CK(cudaDeviceSynchronize());
    const double slices[] = {0, 8, 16, 32, 48, 64, 96};
    const double touches[] = {4, 11, 22};
    for (double s : slices)
        for (double t : touches) {
            if (s == 0 && t != 4) continue;
            const double tt = s == 0 ? 1.0 : t;
            run<1>(a, n_lines, s, tt, sms, 1024, out);
            run<2>(a, n_lines, s, tt, sms, 1024, out);
            run<4>(a, n_lines, s, tt, sms, 1024, out);
        }
    run<1>(a, n_lines, 32, 11, sms, 2048, out);
    run<4>(a, n_lines, 32, 11, sms, 2048, out);
    run<1>(a, n_lines, 32, 11, sms, 512, out);
    return 0;
}
