// gsx_arrange.cu -- ordering of the matches of a batch whose guides have thousands of matches each (bulges).
//
// The reference collects the matches of one guide in std::set<match> per strand index and mismatch count, ordered by the
// match string (process.hpp:21-23,78-79, structures.hpp:41-43), and walks them bucket by bucket, forward index first
// (process.hpp:100-115).  order_matches_kernel (gsx_kernels.cu) restates that as a per-guide rank sort, quadratic in the
// matches of a guide: right for the ~10 matches of a mismatch-only search, hopeless for the ~18 000 of a search with
// bulges.  Here the whole arena is ordered at once by (guide, mismatches, strand index, string key) with three stable
// least-significant-first radix sorts (CUB) over a permutation, duplicates of a string are flagged by comparing neighbours,
// and the per-guide hit offsets come from one prefix sum over the interval widths.  Outputs are those of launch_order.
#include "gsx_kernels.h"
#include "gsx_core.h"
#include <cub/cub.cuh>

namespace gsx {

namespace {

constexpr size_t kAlign = 256;
inline size_t up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

size_t cub_temp_bytes(uint32_t n) {
    size_t a = 0, b = 0;
    cub::DoubleBuffer<uint64_t> k(nullptr, nullptr); cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, a, k, v, (int)n, 0, 64, (cudaStream_t)0);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)n + 1, (cudaStream_t)0);
    return a > b ? a : b;
}

// pass: 0 = low key word, 1 = high key word, 2 = guide | mismatches | strand index
__global__ void arrange_keys_kernel(const MatchRec* __restrict__ m, const uint32_t* perm, uint32_t n, int pass,
                                    uint64_t* __restrict__ keys, uint32_t* vals) {      // (perm and vals are the same array)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t e = pass == 0 ? i : perm[i];
        const MatchRec& r = m[e];
        keys[i] = pass == 0 ? r.key_lo : pass == 1 ? r.key_hi : (((uint64_t)(r.task >> 1) << 4) | ((uint64_t)(r.info & 7u) << 1) | (r.task & 1u));
        if (pass == 0) vals[i] = i;
    }
}

// neighbours with the same guide, bucket and string are one match (std::set): all but the first are flagged
__global__ void arrange_flag_kernel(const MatchRec* __restrict__ m, const uint32_t* __restrict__ perm, uint32_t n,
                                    uint32_t* __restrict__ sorted, uint64_t* __restrict__ width) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
        if (i == n) { width[n] = 0; break; }
        const uint32_t e = perm[i];
        bool dup = false;
        if (i) dup = match_cmp(m[perm[i - 1]], m[e]) == 0 && (m[perm[i - 1]].task >> 1) == (m[e].task >> 1);
        sorted[i] = dup ? (e | 0x80000000u) : e;
        width[i] = dup ? 0ull : (uint64_t)m[e].width;
    }
}

__global__ void arrange_offsets_kernel(const MatchRec* __restrict__ m, const uint32_t* __restrict__ sorted, const uint64_t* __restrict__ scan,
                                       const uint32_t* __restrict__ moff, uint32_t n, uint32_t* __restrict__ sorted_off) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t g = m[sorted[i] & 0x7FFFFFFFu].task >> 1;
        sorted_off[i] = (uint32_t)(scan[i] - scan[moff[g]]);
    }
}

// per (guide, mismatch count): hits = widths between the bucket's boundaries in the sorted key array
__global__ void arrange_counts_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ scan, const uint32_t* __restrict__ moff,
                                      uint32_t n, uint32_t n_guides, uint32_t n_dist, uint32_t* __restrict__ nhits, uint32_t* __restrict__ cbd) {
    const uint32_t total = n_guides * (n_dist + 1u);
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const uint32_t g = t / (n_dist + 1u), d = t - g * (n_dist + 1u);
        const uint32_t b = moff[g], e = moff[g + 1];
        if (d == n_dist) { nhits[g] = (uint32_t)(scan[e] - scan[b]); continue; }
        const uint64_t k0 = ((uint64_t)g << 4) | ((uint64_t)d << 1), k1 = k0 + 2u;
        uint32_t lo = b, hi = e;                                      // first index in [b, e) with key >= k0
        while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (keys[mid] < k0) lo = mid + 1; else hi = mid; }
        const uint32_t s0 = lo;
        hi = e;
        while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (keys[mid] < k1) lo = mid + 1; else hi = mid; }
        cbd[(size_t)g * n_dist + d] = (uint32_t)(scan[lo] - scan[s0]);
    }
}

inline int grid_of(uint64_t n) { uint64_t b = (n + 255) / 256; if (b < 1) b = 1; if (b > 148 * 16) b = 148 * 16; return (int)b; }

}  // namespace

size_t order_sorted_scratch_bytes(uint32_t n) {
    return 2 * up((size_t)n * 8) + 2 * up((size_t)n * 4) + up(((size_t)n + 1) * 8) + up(((size_t)n + 1) * 8) + up(cub_temp_bytes(n)) + kAlign;
}

cudaError_t launch_order_sorted(const MatchRec* m, uint32_t n, const uint32_t* moff, uint32_t n_guides, uint32_t n_dist,
                                uint32_t* sorted, uint32_t* sorted_off, uint32_t* nhits, uint32_t* cbd, void* scratch, size_t scratch_bytes, cudaStream_t s) {
    if (scratch_bytes < order_sorted_scratch_bytes(n)) return cudaErrorInvalidValue;
    unsigned char* p = (unsigned char*)scratch;
    p = (unsigned char*)(((uintptr_t)p + kAlign - 1) / kAlign * kAlign);
    uint64_t* k0 = (uint64_t*)p; p += up((size_t)n * 8);
    uint64_t* k1 = (uint64_t*)p; p += up((size_t)n * 8);
    uint32_t* v0 = (uint32_t*)p; p += up((size_t)n * 4);
    uint32_t* v1 = (uint32_t*)p; p += up((size_t)n * 4);
    uint64_t* width = (uint64_t*)p; p += up(((size_t)n + 1) * 8);
    uint64_t* scan = (uint64_t*)p; p += up(((size_t)n + 1) * 8);
    void* temp = p; size_t temp_bytes = cub_temp_bytes(n);
    cub::DoubleBuffer<uint64_t> keys(k0, k1); cub::DoubleBuffer<uint32_t> vals(v0, v1);
    int guide_bits = 1; while (guide_bits < 32 && (1ull << guide_bits) < (uint64_t)n_guides) guide_bits++;
    cudaError_t e;
    if (n) {
        for (int pass = 0; pass < 3; pass++) {
            // keys of this pass in the current order (values: the permutation so far), written over the spent key buffer
            arrange_keys_kernel<<<grid_of(n), 256, 0, s>>>(m, vals.Current(), n, pass, keys.Current(), vals.Current());
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, vals, (int)n, 0, pass == 2 ? 4 + guide_bits : 64, s);
            if (e != cudaSuccess) return e;
        }
    }
    arrange_flag_kernel<<<grid_of((uint64_t)n + 1), 256, 0, s>>>(m, vals.Current(), n, sorted, width);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    e = cub::DeviceScan::ExclusiveSum(temp, temp_bytes, width, scan, (int)n + 1, s);
    if (e != cudaSuccess) return e;
    if (n) {
        arrange_offsets_kernel<<<grid_of(n), 256, 0, s>>>(m, sorted, scan, moff, n, sorted_off);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    arrange_counts_kernel<<<grid_of((uint64_t)n_guides * (n_dist + 1)), 256, 0, s>>>(keys.Current(), scan, moff, n, n_guides, n_dist, nhits, cbd);
    return cudaGetLastError();
}

}  // namespace gsx
