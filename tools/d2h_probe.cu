// tools/d2h_probe.cu -- measurement helper: device-to-host copy rate into pinned memory for the copy sizes of a result readback
// (64 MB .. 1 GB), alone and while host threads stream through memory the way the formatter does.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_build/d2h_probe tools/d2h_probe.cu -lpthread
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
    const size_t G = 1ull << 30;
    void *d, *h; CK(cudaMalloc(&d, G)); CK(cudaMemset(d, 1, G));
    auto t0 = std::chrono::steady_clock::now();
    CK(cudaHostAlloc(&h, G, cudaHostAllocDefault));
    printf("{\"cudaHostAlloc_1GiB_seconds\": %.3f", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    cudaStream_t s; CK(cudaStreamCreate(&s)); cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    std::atomic<bool> stop{false}; std::vector<std::thread> load;
    for (int phase = 0; phase < 2; phase++) {
        if (phase == 1) for (int t = 0; t < 12; t++) load.emplace_back([&] { std::vector<char> x(256u << 20, 1), y(256u << 20); while (!stop) memcpy(y.data(), x.data(), x.size()); });
        if (phase == 1) std::this_thread::sleep_for(std::chrono::milliseconds(300));
        for (size_t mb : {64, 256, 1024}) {
            float best = 1e9f;
            for (int rep = 0; rep < 4; rep++) {
                CK(cudaEventRecord(a, s)); CK(cudaMemcpyAsync(h, d, mb << 20, cudaMemcpyDeviceToHost, s)); CK(cudaEventRecord(b, s)); CK(cudaStreamSynchronize(s));
                float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
            }
            printf(", \"d2h_%zuMB_%s_GBps\": %.1f", mb, phase ? "under_host_memory_load" : "alone", (double)(mb << 20) / best / 1e6);
        }
        // a result readback: twelve arrays back to back (1-8 bytes per hit)
        CK(cudaEventRecord(a, s));
        for (int k = 0; k < 12; k++) CK(cudaMemcpyAsync((char*)h + (size_t)k * (80u << 20), (char*)d + (size_t)k * (80u << 20), 80u << 20, cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(b, s)); CK(cudaStreamSynchronize(s));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        printf(", \"d2h_12x80MB_%s_GBps\": %.1f", phase ? "under_host_memory_load" : "alone", 12.0 * (80u << 20) / ms / 1e6);
    }
    stop = true; for (auto& t : load) t.join();
    printf("}\n");
    return 0;
}
