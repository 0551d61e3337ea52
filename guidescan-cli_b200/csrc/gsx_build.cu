// gsx_build.cu -- GPU index construction: suffix array by prefix doubling on the device, then BWT -> OccBlocks.
//
// Replaces the reference's single-threaded CPU build (sdsl::construct: libdivsufsort suffix sorting +
// BWT streaming + wavelet tree, reference sdsl/include/sdsl/construct.hpp:121-165, construct_sa.hpp:78-117,
// construct_bwt.hpp:49-78).  The suffix array of text+'\0' is unique, so the rows, intervals and SA samples are
// identical to the reference's; only the encoding differs (gsx_types.h).
// Index construction is not the timed hot path; the key/value sorts use CUB's device radix sort.
#include "gsx_host.h"
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstring>
#include <numeric>
#include <stdexcept>

namespace gsx {
namespace {

struct BuildError : std::runtime_error { using std::runtime_error::runtime_error; };
#define BK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) throw BuildError(std::string(#call) + ": " + cudaGetErrorString(e__)); } while (0)

struct CodeTable { uint8_t code[256]; };

// key[i] = the first `chars` symbol codes of suffix i, `bits` bits each, most significant first (0 beyond the end)
__global__ void make_keys_kernel(const uint8_t* __restrict__ text, uint64_t len, uint64_t n, CodeTable tab, int bits, int chars,
                                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t k = 0;
        for (int j = 0; j < chars; j++) {
            uint64_t p = i + j;
            k = (k << bits) | (p < len ? (uint64_t)tab.code[text[p]] : 0ull);
        }
        keys[i] = k; vals[i] = (uint32_t)i;
    }
}

// head[i] = i if sorted element i starts a new key group, else 0
__global__ void mark_heads_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t* __restrict__ head, unsigned long long* n_groups) {
    unsigned long long local = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        bool h = i == 0 || keys[i] != keys[i - 1];
        head[i] = h ? (uint32_t)i : 0u;
        local += h;
    }
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(n_groups, local);
}

__global__ void scatter_rank_kernel(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ group_start, uint64_t n, uint32_t* __restrict__ rank) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) rank[sa[i]] = group_start[i];
}

__global__ void doubling_keys_kernel(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ rank, uint64_t n, uint64_t h, uint64_t* __restrict__ keys) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t p = sa[i];
        uint64_t r1 = rank[p];
        uint64_t r2 = p + h < n ? (uint64_t)rank[p + h] + 1ull : 0ull;
        keys[i] = (r1 << 32) | r2;
    }
}

// one warp per 32 rows, two iterations per 64-row block: planes + per-block symbol counts; exceptions appended
__global__ void bwt_blocks_kernel(const uint8_t* __restrict__ text, const uint32_t* __restrict__ sa, uint64_t n, uint64_t n_blocks,
                                  OccBlock* __restrict__ blocks, uint32_t* __restrict__ exc_row, uint8_t* __restrict__ exc_sym,
                                  unsigned long long* exc_count, unsigned long long exc_cap) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t b = warp; b < n_blocks; b += n_warps) {
        uint64_t hi = 0, lo = 0; uint32_t cnt[4] = {0, 0, 0, 0};
        for (int half = 0; half < 2; half++) {
            uint64_t row = b * 64 + half * 32 + lane;
            int code = -1; uint8_t sym = 0; bool valid = row < n;
            if (valid) {
                uint32_t p = sa[row];
                sym = p ? text[p - 1] : 0;
                code = sym == 'A' ? 0 : sym == 'C' ? 1 : sym == 'G' ? 2 : sym == 'T' ? 3 : -1;
                if (code < 0) {
                    unsigned long long slot = atomicAdd(exc_count, 1ull);
                    if (slot < exc_cap) { exc_row[slot] = (uint32_t)row; exc_sym[slot] = sym; }
                }
            }
            uint32_t mh = __ballot_sync(0xffffffffu, valid && code >= 2);
            uint32_t ml = __ballot_sync(0xffffffffu, valid && (code == 1 || code == 3));
            uint32_t mv = __ballot_sync(0xffffffffu, valid && code >= 0);
            hi |= (uint64_t)mh << (32 * half); lo |= (uint64_t)ml << (32 * half);
            cnt[3] += __popc(mh & ml); cnt[2] += __popc(mh & ~ml); cnt[1] += __popc(~mh & ml & mv); cnt[0] += __popc(~mh & ~ml & mv);
        }
        if (lane == 0) { OccBlock o; o.cnt[0] = cnt[0]; o.cnt[1] = cnt[1]; o.cnt[2] = cnt[2]; o.cnt[3] = cnt[3]; o.hi = hi; o.lo = lo; blocks[b] = o; }
    }
}

struct MaxU32 { __host__ __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } };
struct Cnt4 { uint32_t c[4]; };
struct Cnt4Add { __host__ __device__ Cnt4 operator()(const Cnt4& a, const Cnt4& b) const { Cnt4 r; for (int i = 0; i < 4; i++) r.c[i] = a.c[i] + b.c[i]; return r; } };

__global__ void extract_counts_kernel(const OccBlock* __restrict__ blocks, uint64_t n_blocks, Cnt4* __restrict__ out) {
    for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b < n_blocks; b += (uint64_t)gridDim.x * blockDim.x) { Cnt4 c; for (int i = 0; i < 4; i++) c.c[i] = blocks[b].cnt[i]; out[b] = c; }
}
__global__ void apply_counts_kernel(OccBlock* __restrict__ blocks, uint64_t n_blocks, const Cnt4* __restrict__ pre) {
    for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b < n_blocks; b += (uint64_t)gridDim.x * blockDim.x) for (int i = 0; i < 4; i++) blocks[b].cnt[i] = pre[b].c[i];
}
__global__ void sample_sa_kernel(const uint32_t* __restrict__ sa, uint64_t n_samples, uint32_t shift, uint32_t* __restrict__ out) {
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n_samples; k += (uint64_t)gridDim.x * blockDim.x) out[k] = sa[k << shift];
}

struct Bufs {
    std::vector<void*> p;
    template <class T> T* get(uint64_t n) { void* d = nullptr; BK(cudaMalloc(&d, std::max<uint64_t>(n, 1) * sizeof(T))); p.push_back(d); return (T*)d; }
    void drop(void* d) { auto it = std::find(p.begin(), p.end(), d); if (it != p.end()) { cudaFree(d); p.erase(it); } }
    ~Bufs() { for (void* d : p) cudaFree(d); }
};

}  // namespace

bool build_strand_gpu(int device, const uint8_t* text, uint64_t len, uint32_t sa_shift, HostStrand& out, std::string& err) {
    try {
        BK(cudaSetDevice(device));
        const uint64_t n = len + 1;
        if (n > 0xFFFFFFFFull) throw BuildError("text longer than 2^32 - 2");
        const int grid = 148 * 16, threads = 256;
        // alphabet
        uint64_t hist[256]; memset(hist, 0, sizeof hist);
        for (uint64_t i = 0; i < len; i++) hist[text[i]]++;
        if (hist[0]) throw BuildError("text contains a NUL byte");
        CodeTable tab; memset(&tab, 0, sizeof tab);
        int sigma = 0;
        for (int c = 1; c < 256; c++) if (hist[c]) tab.code[c] = (uint8_t)(++sigma);
        int bits = 1; while ((1 << bits) < sigma + 1) bits++;
        const int chars = 64 / bits;

        Bufs B;
        uint8_t* d_text = B.get<uint8_t>(len);
        BK(cudaMemcpy(d_text, text, len, cudaMemcpyHostToDevice));
        uint64_t* keys[2] = {B.get<uint64_t>(n), B.get<uint64_t>(n)};
        uint32_t* vals[2] = {B.get<uint32_t>(n), B.get<uint32_t>(n)};
        uint32_t* d_rank = B.get<uint32_t>(n);
        uint32_t* d_tmp = B.get<uint32_t>(n);
        unsigned long long* d_cnt = B.get<unsigned long long>(2);

        size_t sort_bytes = 0, scan_bytes = 0;
        {
            cub::DoubleBuffer<uint64_t> dk(keys[0], keys[1]); cub::DoubleBuffer<uint32_t> dv(vals[0], vals[1]);
            BK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, dk, dv, (int64_t)n, 0, 64));
            BK(cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, d_tmp, d_tmp, MaxU32(), (int64_t)n));
        }
        void* d_temp = B.get<uint8_t>(std::max(sort_bytes, scan_bytes));
        size_t temp_bytes = std::max(sort_bytes, scan_bytes);

        make_keys_kernel<<<grid, threads>>>(d_text, len, n, tab, bits, chars, keys[0], vals[0]);
        BK(cudaGetLastError());
        int cur = 0;
        uint64_t h = (uint64_t)chars;
        int end_bit = bits * chars;
        for (int round = 0;; round++) {
            if (round > 40) throw BuildError("prefix doubling did not converge");
            cub::DoubleBuffer<uint64_t> dk(keys[cur], keys[cur ^ 1]); cub::DoubleBuffer<uint32_t> dv(vals[cur], vals[cur ^ 1]);
            size_t tb = temp_bytes;
            BK(cub::DeviceRadixSort::SortPairs(d_temp, tb, dk, dv, (int64_t)n, 0, end_bit));
            cur = dk.Current() == keys[0] ? 0 : 1;
            if (dv.Current() != vals[cur]) throw BuildError("CUB double buffers out of step");
            BK(cudaMemset(d_cnt, 0, 16));
            mark_heads_kernel<<<grid, threads>>>(keys[cur], n, d_tmp, d_cnt);
            BK(cudaGetLastError());
            unsigned long long groups = 0;
            BK(cudaMemcpy(&groups, d_cnt, 8, cudaMemcpyDeviceToHost));
            if (groups == n) break;                                        // every suffix has a distinct key: vals[cur] is the SA
            tb = temp_bytes;
            BK(cub::DeviceScan::InclusiveScan(d_temp, tb, d_tmp, d_tmp, MaxU32(), (int64_t)n));
            scatter_rank_kernel<<<grid, threads>>>(vals[cur], d_tmp, n, d_rank);
            doubling_keys_kernel<<<grid, threads>>>(vals[cur], d_rank, n, h, keys[cur]);
            BK(cudaGetLastError());
            h *= 2; end_bit = 64;
        }
        uint32_t* d_sa = vals[cur];
        // free what the second phase does not need
        B.drop(keys[0]); B.drop(keys[1]); B.drop(vals[cur ^ 1]); B.drop(d_rank); B.drop(d_tmp);

        const uint64_t n_blocks = n / 64 + 1;
        OccBlock* d_blocks = B.get<OccBlock>(n_blocks);
        BK(cudaMemset(d_blocks, 0, n_blocks * sizeof(OccBlock)));
        uint64_t exc_cap = 1;                                               // the sentinel row
        for (int c = 1; c < 256; c++) if (c != 'A' && c != 'C' && c != 'G' && c != 'T') exc_cap += hist[c];
        uint32_t* d_exc_row = B.get<uint32_t>(exc_cap); uint8_t* d_exc_sym = B.get<uint8_t>(exc_cap);
        BK(cudaMemset(d_cnt, 0, 16));
        bwt_blocks_kernel<<<grid, threads>>>(d_text, d_sa, n, n_blocks, d_blocks, d_exc_row, d_exc_sym, d_cnt, exc_cap);
        BK(cudaGetLastError());
        Cnt4* d_c = B.get<Cnt4>(n_blocks); Cnt4* d_pre = B.get<Cnt4>(n_blocks);
        extract_counts_kernel<<<grid, threads>>>(d_blocks, n_blocks, d_c);
        size_t sb = 0; Cnt4 zero{};
        BK(cub::DeviceScan::ExclusiveScan(nullptr, sb, d_c, d_pre, Cnt4Add(), zero, (int64_t)n_blocks));
        void* d_t2 = B.get<uint8_t>(sb);
        BK(cub::DeviceScan::ExclusiveScan(d_t2, sb, d_c, d_pre, Cnt4Add(), zero, (int64_t)n_blocks));
        apply_counts_kernel<<<grid, threads>>>(d_blocks, n_blocks, d_pre);
        const uint64_t n_samples = ((n - 1) >> sa_shift) + 1;
        uint32_t* d_samples = B.get<uint32_t>(n_samples);
        sample_sa_kernel<<<grid, threads>>>(d_sa, n_samples, sa_shift, d_samples);
        BK(cudaGetLastError());
        BK(cudaDeviceSynchronize());

        // to the host
        out.n = n; out.sa_shift = sa_shift;
        out.blocks.resize(n_blocks); out.sa_samples.resize(n_samples);
        BK(cudaMemcpy(out.blocks.data(), d_blocks, n_blocks * sizeof(OccBlock), cudaMemcpyDeviceToHost));
        BK(cudaMemcpy(out.sa_samples.data(), d_samples, n_samples * 4, cudaMemcpyDeviceToHost));
        unsigned long long n_exc = 0; BK(cudaMemcpy(&n_exc, d_cnt, 8, cudaMemcpyDeviceToHost));
        if (n_exc != exc_cap) throw BuildError("exception count mismatch");
        std::vector<uint32_t> er(n_exc); std::vector<uint8_t> es(n_exc);
        BK(cudaMemcpy(er.data(), d_exc_row, n_exc * 4, cudaMemcpyDeviceToHost));
        BK(cudaMemcpy(es.data(), d_exc_sym, n_exc, cudaMemcpyDeviceToHost));
        std::vector<uint32_t> order(n_exc); std::iota(order.begin(), order.end(), 0u);
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return er[a] < er[b]; });
        uint64_t Cb[257]; uint64_t acc = 0; hist[0] = 1;
        for (int c = 0; c < 256; c++) { Cb[c] = acc; acc += hist[c]; }
        out.C[0] = (uint32_t)Cb['A']; out.C[1] = (uint32_t)Cb['C']; out.C[2] = (uint32_t)Cb['G']; out.C[3] = (uint32_t)Cb['T']; out.C[4] = (uint32_t)Cb['N'];
        out.exc_rows.resize(n_exc); out.exc_lf.resize(n_exc); out.exc_sym.resize(n_exc); out.n_rows.clear();
        uint64_t seen[256]; memset(seen, 0, sizeof seen);
        for (uint64_t i = 0; i < n_exc; i++) {
            uint32_t row = er[order[i]]; uint8_t sym = es[order[i]];
            out.exc_rows[i] = row; out.exc_sym[i] = sym; out.exc_lf[i] = (uint32_t)(Cb[sym] + seen[sym]++);       // LF = C[c] + rank_c(row)
            if (sym == 'N') out.n_rows.push_back(row);
        }
        return true;
    } catch (const std::exception& e) { err = std::string("gsx_index_build: ") + e.what(); return false; }
}

}  // namespace gsx
