// gsx_sdsl_write.cpp -- the index written back in the REFERENCE's own on-disk format, so that an index built on the GPU
// (gsx_index_build: seconds of suffix sorting instead of the reference's hour of divsufsort at 3.1 Gb) can be opened by the
// unmodified reference binary, and by every other tool that reads GuideScan2 indices.
//
// Output = what `guidescan index` stores with sdsl::store_to_file (reference src/guidescan.cxx:167-175), byte for byte:
//   csa_wt<wt_huff<>,64,8192>::serialize (sdsl csa_wt.hpp:372-382)
//     wt_pc::serialize (wt_pc.hpp:656-671): size, sigma, bit vector, rank_support_v, select_support_mcl<1>, <0>, tree
//     SA samples every 64 rows (csa_sampling_strategy.hpp:85-111), ISA samples every 8192 text positions (:626-648)
//     byte_alphabet (lib/csa_alphabet_strategy.cpp:25-55,103-121)
// Field layout: SURVEY.md App. B.  What each structure must CONTAIN is restated here from the reference's constructors (cited
// at each function); how it is computed is ours: the bit vector is filled by row ranges in parallel from the 2-bit planes with
// the node cursors of a range start taken from the block counters, the select directories are produced by one streaming pass
// over the words, and the ISA samples come from the nearest SA sample and a few LF steps instead of a full suffix array.
//
// tests/test_sdsl_writer.py: files of the reference rewritten through this writer are identical to the originals; the select and
// rank directories are compared with sdsl's own on crafted bit vectors (oracle/_ref/sdsl_probe); the unmodified reference
// enumerates over a rewritten index.
#include "gsx_host.h"
#include "gsx_core.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace gsx {
namespace {

inline uint32_t top_bit(uint64_t x) { return x ? 63u - (uint32_t)__builtin_clzll(x) : 0u; }      // sdsl bits::hi: hi(0) = 0

struct Sink {
    FILE* f = nullptr; bool ok = true;
    void bytes(const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) ok = false; }
    void u64(uint64_t v) { bytes(&v, 8); }
    void u16(uint16_t v) { bytes(&v, 2); }
    void u8(uint8_t v) { bytes(&v, 1); }
};

// sdsl int_vector<0>: u64 size in bits, u8 width, ceil(bits / 64) words, value i in bits [i * width, (i + 1) * width)
// (int_vector.hpp:593-609,1563-1595).  A default-constructed one has no values and width 64.
struct Packed {
    uint64_t n = 0; uint32_t width = 64; std::vector<uint64_t> w;
    Packed() {}
    Packed(uint64_t n_, uint32_t width_) : n(n_), width(width_), w((n_ * width_ + 63) / 64, 0ull) {}
    void set(uint64_t i, uint64_t v) {
        if (width < 64) v &= (1ull << width) - 1ull;
        const uint64_t bit = i * width, at = bit >> 6; const uint32_t sh = (uint32_t)(bit & 63);
        w[at] |= v << sh;
        if (sh + width > 64) w[at + 1] |= v >> (64 - sh);
    }
    void put(Sink& s) const { s.u64(n * width); s.u8((uint8_t)width); s.bytes(w.data(), w.size() * 8); }
    void append_to(std::vector<uint8_t>& out) const {
        const uint64_t bits = n * width; const uint8_t wd = (uint8_t)width;
        const uint8_t* p = (const uint8_t*)&bits; out.insert(out.end(), p, p + 8); out.push_back(wd);
        const uint8_t* q = (const uint8_t*)w.data(); out.insert(out.end(), q, q + w.size() * 8);
    }
};

// ---- tree shape ---------------------------------------------------------------------------------------------------------
// Huffman shape as sdsl builds it (wt_huff.hpp:92-117): one leaf per byte that occurs, in byte order; the two lightest roots
// (ties: the older node first) become child 0 and child 1 of a new node until one root is left.  Stored breadth first with the
// root at index 0 (wt_helper.hpp:166-233): an inner node's bits start at bv_pos, a leaf carries its byte in bv_pos_rank and the
// running offset in bv_pos; path[c] = the root-to-leaf turns of c, first turn in bit 0, length in bits 56..63; a byte that does
// not occur gets length 0 and the value of the last byte below it that does.
struct WtShape {
    struct Node { uint64_t bv_pos = 0, bv_pos_rank = 0, weight = 0; uint16_t parent = 0xFFFF, child[2] = {0xFFFF, 0xFFFF}; };
    std::vector<Node> nodes;
    uint16_t leaf_of[256]; uint64_t path[256];
    uint64_t bv_bits = 0; uint32_t sigma = 0;
};

WtShape huffman_shape(const uint64_t count[256]) {
    struct Tmp { uint64_t weight; uint32_t sym; int kid[2]; };
    std::vector<Tmp> tmp; std::vector<int> roots;
    for (uint32_t c = 0; c < 256; c++) if (count[c]) { roots.push_back((int)tmp.size()); tmp.push_back({count[c], c, {-1, -1}}); }
    WtShape t; t.sigma = (uint32_t)tmp.size();
    auto lighter = [&](int a, int b) { return tmp[a].weight != tmp[b].weight ? tmp[a].weight < tmp[b].weight : a < b; };
    while (roots.size() > 1) {
        size_t i0 = 0; for (size_t i = 1; i < roots.size(); i++) if (lighter(roots[i], roots[i0])) i0 = i;
        const int a = roots[i0]; roots.erase(roots.begin() + (long)i0);
        size_t i1 = 0; for (size_t i = 1; i < roots.size(); i++) if (lighter(roots[i], roots[i1])) i1 = i;
        const int b = roots[i1]; roots.erase(roots.begin() + (long)i1);
        roots.push_back((int)tmp.size()); tmp.push_back({tmp[a].weight + tmp[b].weight, 0, {a, b}});
    }
    for (int c = 0; c < 256; c++) { t.leaf_of[c] = 0xFFFF; t.path[c] = 0; }
    if (tmp.empty()) return t;
    // breadth-first numbering; `from[i]` = the construction-order node stored at i
    std::vector<int> from{roots[0]}; std::vector<uint64_t> code{0}; std::vector<uint32_t> depth{0};
    t.nodes.resize(tmp.size());
    for (size_t i = 0; i < from.size(); i++) {
        const Tmp& s = tmp[from[i]]; WtShape::Node& nd = t.nodes[i];
        nd.weight = s.weight; nd.bv_pos = t.bv_bits;
        if (s.kid[0] < 0) { nd.bv_pos_rank = s.sym; t.leaf_of[s.sym] = (uint16_t)i; t.path[s.sym] = code[i] | ((uint64_t)depth[i] << 56); continue; }
        t.bv_bits += s.weight;
        for (int k = 0; k < 2; k++) {
            nd.child[k] = (uint16_t)from.size(); t.nodes[from.size()].parent = (uint16_t)i;
            from.push_back(s.kid[k]); code.push_back(code[i] | ((uint64_t)k << depth[i])); depth.push_back(depth[i] + 1);
        }
    }
    for (uint32_t c = 0, below = 0; c < 256; c++) { if (t.leaf_of[c] != 0xFFFF) below = c; else t.path[c] = below; }
    return t;
}

// ---- bit vector ---------------------------------------------------------------------------------------------------------
// Node v's bits, in row order, for the rows whose byte lies below v: 0 = the byte is below child 0 (wt_pc.hpp:110-123,225-247).
// Rows [r0, r1) of the BWT are independent of the others once each node's cursor at r0 is known, and that is bv_pos plus the
// occurrences before r0 of the bytes below the node: the block counters for A/C/G/T, a running tally for the exception rows.
const uint8_t kAcgt[4] = {'A', 'C', 'G', 'T'};

void byte_counts(const HostStrand& h, uint64_t count[256]) {
    memset(count, 0, 256 * sizeof(uint64_t));
    for (uint64_t b = 0; b < h.blocks.size(); b++) {                 // planes of the rows of block b, exception rows hold code 0
        const uint64_t rows = std::min<uint64_t>(64, h.n - std::min<uint64_t>(h.n, b * 64));
        if (!rows) break;
        const uint64_t m = rows == 64 ? ~0ull : (1ull << rows) - 1ull, hi = h.blocks[b].hi & m, lo = h.blocks[b].lo & m;
        const uint64_t t = (uint64_t)__builtin_popcountll(hi & lo), g = (uint64_t)__builtin_popcountll(hi) - t, c = (uint64_t)__builtin_popcountll(lo) - t;
        count['T'] += t; count['G'] += g; count['C'] += c; count['A'] += rows - t - g - c;
    }
    count['A'] -= h.exc_rows.size();
    for (uint8_t s : h.exc_sym) count[s]++;
}

std::vector<uint64_t> fill_bit_vector(const HostStrand& h, const WtShape& t, unsigned threads) {
    const uint64_t words = (t.bv_bits + 63) / 64;
    std::vector<uint64_t> bv(words, 0ull);
    const size_t n_nodes = t.nodes.size();
    // per byte: the inner nodes on its path and the turn taken at each
    struct Path { uint32_t len = 0; uint16_t node[56]; uint8_t turn[56]; };
    std::vector<Path> paths(256);
    for (uint32_t c = 0; c < 256; c++) {
        if (t.leaf_of[c] == 0xFFFF) continue;
        Path& p = paths[c]; p.len = (uint32_t)(t.path[c] >> 56);
        uint16_t v = 0;
        for (uint32_t l = 0; l < p.len; l++) { p.node[l] = v; p.turn[l] = (uint8_t)((t.path[c] >> l) & 1u); v = t.nodes[v].child[p.turn[l]]; }
    }
    const uint64_t n_blocks = (h.n + 63) / 64;
    threads = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(threads, n_blocks / 1024 + 1));
    // range starts (in blocks) and the exception tallies there
    std::vector<uint64_t> start(threads + 1);
    for (unsigned i = 0; i <= threads; i++) start[i] = n_blocks * i / threads;
    std::vector<std::vector<uint64_t>> exc_before(threads, std::vector<uint64_t>(256, 0));
    std::vector<size_t> exc_at(threads + 1, 0);
    {
        std::vector<uint64_t> tally(256, 0); size_t e = 0;
        for (unsigned i = 0; i < threads; i++) {
            while (e < h.exc_rows.size() && h.exc_rows[e] < start[i] * 64) tally[h.exc_sym[e++]]++;
            exc_before[i] = tally; exc_at[i] = e;
        }
        exc_at[threads] = h.exc_rows.size();
    }
    auto work = [&](unsigned i) {
        struct Cur { uint64_t pos = 0, acc = 0; };
        std::vector<Cur> cur(n_nodes);
        for (size_t v = 0; v < n_nodes; v++) cur[v].pos = t.nodes[v].bv_pos;
        const uint64_t b0 = start[i], b1 = start[i + 1];
        if (b0 == b1) return;
        uint64_t before[256]; memcpy(before, exc_before[i].data(), sizeof before);
        for (int s = 0; s < 4; s++) before[kAcgt[s]] = h.blocks[b0].cnt[s];
        for (uint32_t c = 0; c < 256; c++) if (before[c]) for (uint32_t l = 0; l < paths[c].len; l++) cur[paths[c].node[l]].pos += before[c];
        // a range may begin in the middle of a word of a node's bits: those words are shared with the neighbour range
        auto flush = [&](uint64_t word, uint64_t bits) { if (bits) __atomic_fetch_or(&bv[word], bits, __ATOMIC_RELAXED); };
        size_t e = exc_at[i];
        for (uint64_t b = b0; b < b1; b++) {
            const uint64_t hi = h.blocks[b].hi, lo = h.blocks[b].lo;
            const uint64_t row0 = b * 64, rows = std::min<uint64_t>(64, h.n - row0);
            for (uint64_t r = 0; r < rows; r++) {
                uint8_t c = kAcgt[(((hi >> r) & 1ull) << 1) | ((lo >> r) & 1ull)];
                if (e < h.exc_rows.size() && h.exc_rows[e] == row0 + r) c = h.exc_sym[e++];
                const Path& p = paths[c];
                for (uint32_t l = 0; l < p.len; l++) {
                    Cur& k = cur[p.node[l]];
                    k.acc |= (uint64_t)p.turn[l] << (k.pos & 63);
                    if ((++k.pos & 63) == 0) { flush((k.pos >> 6) - 1, k.acc); k.acc = 0; }
                }
            }
        }
        for (size_t v = 0; v < n_nodes; v++) if (cur[v].pos & 63) flush(cur[v].pos >> 6, cur[v].acc);
    };
    std::vector<std::thread> pool;
    for (unsigned i = 1; i < threads; i++) pool.emplace_back(work, i);
    work(0);
    for (auto& th : pool) th.join();
    return bv;
}

// ---- rank directory -----------------------------------------------------------------------------------------------------
// rank_support_v (rank_support_v.hpp:65-108): per 512 bits two words, the ones before the block and, 9 bits each from the top,
// the ones of the block's first 1..7 words -- as far as those words exist, plus one field (or, when the vector fills its last
// block, one whole entry) for the end of the vector.
void put_rank_directory(Sink& s, const std::vector<uint64_t>& bv) {
    const uint64_t W = bv.size(), n_blk = (W >> 3) + 1;
    std::vector<uint64_t> dir(2 * n_blk, 0ull);
    uint64_t total = 0;
    for (uint64_t b = 0; b < n_blk; b++) {
        dir[2 * b] = total;
        uint64_t in_block = 0, rel = 0;
        for (uint64_t k = 1; k <= 8 && 8 * b + k <= W; k++) {
            in_block += (uint64_t)__builtin_popcountll(bv[8 * b + k - 1]);
            if (k < 8) rel |= in_block << (63 - 9 * k);
        }
        dir[2 * b + 1] = rel; total += in_block;
    }
    s.u64(dir.size() * 64); s.bytes(dir.data(), dir.size() * 8);
}

// ---- select directories -------------------------------------------------------------------------------------------------
// select_support_mcl (select_support_mcl.hpp:108-116,203-343,425-462).  The marked bits ("args": ones, or zeros) are cut into
// superblocks of 4096.  superblock[k] = position of the block's first arg; a block whose args span more than log^4 bits stores
// all 4096 positions ("long"), any other the offsets of every 64th arg ("mini").  Two constructions exist and their results
// differ in documented-nowhere details that a byte-identical file has to follow:
//   vectors below 100 000 bits (init_slow): exactly the above; the span of a block is last arg - first arg, a partly filled
//     last block is judged like the others and its mini directory is filled as far as it has args;
//   larger vectors (init_fast) work on whole words, so for zeros the padding bits of the last word count as args ("phantoms")
//     wherever the code looks at words and not where it looks at bits:
//       * a block is closed when its 4033rd arg (word view) is met; its span then runs to the last real arg among the NEXT 64
//         args, which reaches one arg into the following block; width of a long block = bits of that position;
//       * whatever is left at the end (1..4032 args, word view) becomes a long block of width bits(size - 1) holding the real
//         args only, with superblock[k] left 0;
//       * a phantom-only block past the last real one is built but not stored; it still makes the long/mini flags appear.
struct SelectDir {
    uint64_t n_args = 0, n_sb = 0; bool any_long = false;
    Packed superblock; std::vector<uint8_t> is_mini; std::vector<uint8_t> blocks;      // serialized blocks, in order
};

// The blocks are independent of each other but for one look at the first arg of the next block, so ranges of blocks are built by
// one thread each: a range starts at the word that holds its first arg, found through the arg counts of the 512-bit chunks.
SelectDir build_select(const std::vector<uint64_t>& bv, uint64_t n_bits, bool ones, unsigned threads) {
    SelectDir d;
    const uint64_t capacity = bv.size() * 64;
    const bool fast = n_bits >= 100000;
    const uint64_t view_bits = fast ? capacity : n_bits;            // where the args are looked for
    auto args_of = [&](uint64_t w) {                                // the args of word w as set bits
        uint64_t x = ones ? bv[w] : ~bv[w];
        if ((w + 1) * 64 > view_bits) { const uint64_t keep = view_bits - w * 64; x &= keep ? (~0ull >> (64 - keep)) : 0ull; }
        return x;
    };
    std::vector<uint64_t> before(bv.size() / 8 + 2, 0);             // args (word view) before each chunk of 8 words
    uint64_t n_ones = 0;
    for (uint64_t w = 0; w < bv.size(); w++) {
        if ((w & 7) == 0) before[w / 8 + 1] = before[w / 8];
        before[w / 8 + 1] += (uint64_t)__builtin_popcountll(args_of(w)); n_ones += (uint64_t)__builtin_popcountll(bv[w]);
    }
    const uint64_t n_view = before[(bv.size() + 7) / 8];            // args in the word view (phantoms included)
    d.n_args = ones ? n_ones : n_bits - n_ones;
    if (!d.n_args) return d;
    d.n_sb = (d.n_args + 4095) >> 12;
    const uint32_t logn = top_bit(capacity) + 1; const uint64_t logn4 = (uint64_t)logn * logn * logn * logn;
    d.superblock = Packed(d.n_sb, logn);
    d.is_mini.assign(d.n_sb, 0);
    const uint64_t n_view_blocks = (n_view + 4095) >> 12;
    const unsigned T = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(threads ? threads : 1, n_view_blocks / 4 + 1));
    struct Part { std::vector<uint64_t> first; std::vector<uint8_t> is_mini, blocks; bool any_long = false; };
    std::vector<Part> parts(T);
    auto build = [&](unsigned t) {
        Part& part = parts[t];
        const uint64_t k0 = n_view_blocks * t / T, k1 = n_view_blocks * (t + 1) / T;
        if (k0 == k1) return;
        std::vector<uint64_t> P(4097); uint64_t have = 0, k = k0;
        auto store = [&](const Packed& v, bool mini, uint64_t first) {
            part.first.push_back(first); part.is_mini.push_back(mini ? 1 : 0); if (!mini) part.any_long = true; v.append_to(part.blocks);
        };
        // `have` args buffered, at most 4096 of them belong to block k, P[4096] (if present) is the first of the next block
        auto close_block = [&]() {
            const uint64_t c = std::min<uint64_t>(have, 4096);
            if (k >= d.n_sb) { part.any_long = true; return; }
            if (!fast) {
                const uint64_t first = P[0], last = P[c - 1], span = last - first;
                if (span > logn4) { Packed v(4096, top_bit(last) + 1); for (uint64_t j = 0; j < c; j++) v.set(j, P[j]); store(v, false, first); }
                else { Packed v(64, top_bit(span) + 1); for (uint64_t j = 0; j < c; j += 64) v.set(j / 64, P[j] - first); store(v, true, first); }
                return;
            }
            if (c >= 4033) {
                const uint64_t first = P[0]; uint64_t last = P[4032];
                for (uint64_t j = have - 1; j > 4032; j--) if (P[j] < n_bits) { last = P[j]; break; }
                const uint64_t span = last - first;
                if (span > logn4) { Packed v(4096, top_bit(last) + 1); for (uint64_t j = 0; j < c && P[j] <= last; j++) v.set(j, P[j]); store(v, false, first); }
                else { Packed v(64, top_bit(span) + 1); for (uint64_t j = 0; j < 4096; j += 64) v.set(j / 64, P[j] - first); store(v, true, first); }
            } else {
                Packed v(4096, top_bit(n_bits - 1) + 1);
                for (uint64_t j = 0; j < c && P[j] < n_bits; j++) v.set(j, P[j]);
                store(v, false, 0);                                 // (superblock[k] stays 0)
            }
        };
        // the word of arg number 4096 k0 (counted from 0), and how many args of that word belong to the block before
        const uint64_t a0 = k0 << 12;
        uint64_t chunk = (uint64_t)(std::upper_bound(before.begin(), before.begin() + (long)((bv.size() + 7) / 8 + 1), a0) - before.begin()) - 1;
        uint64_t w = chunk * 8, seen = before[chunk];
        for (;; w++) { const uint64_t c = (uint64_t)__builtin_popcountll(args_of(w)); if (seen + c > a0) break; seen += c; }
        uint64_t skip = a0 - seen;
        for (; w < bv.size() && k < k1; w++) {
            uint64_t x = args_of(w);
            for (; skip && x; skip--) x &= x - 1;
            while (x && k < k1) {
                P[have++] = w * 64 + (uint64_t)__builtin_ctzll(x); x &= x - 1;
                if (have == 4097) { close_block(); k++; P[0] = P[4096]; have = 1; }
            }
        }
        if (k < k1 && have) { close_block(); k++; }                  // the end of the vector (last range only)
    };
    {
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < T; t++) pool.emplace_back(build, t);
        build(0);
        for (auto& th : pool) th.join();
    }
    uint64_t k = 0;
    for (const Part& part : parts) {
        d.any_long = d.any_long || part.any_long;
        for (size_t i = 0; i < part.first.size(); i++, k++) { d.superblock.set(k, part.first[i]); d.is_mini[k] = part.is_mini[i]; }
        d.blocks.insert(d.blocks.end(), part.blocks.begin(), part.blocks.end());
    }
    return d;
}

void put_select(Sink& s, const SelectDir& d) {
    s.u64(d.n_args);
    if (!d.n_args) return;
    d.superblock.put(s);
    if (d.any_long) {                                               // bit i set: block i is a mini block
        std::vector<uint64_t> w((d.n_sb + 63) / 64, 0ull);
        for (uint64_t i = 0; i < d.n_sb; i++) if (d.is_mini[i]) w[i >> 6] |= 1ull << (i & 63);
        s.u64(d.n_sb); s.bytes(w.data(), w.size() * 8);
    } else s.u64(0);
    s.bytes(d.blocks.data(), d.blocks.size());
}

// ---- ISA samples --------------------------------------------------------------------------------------------------------
// entry k = the row of the suffix that starts at text position 8192 k (csa_sampling_strategy.hpp:626-648).  Every SA sample
// (row r, position p) knows ISA[p] = r, and one LF step goes from the row of p to the row of p - 1: for each k the sample with
// the smallest p >= 8192 k inside the window is walked down p - 8192 k steps.  Windows without a sample (tiny or very repetitive
// texts) continue the walk of the window above; the topmost one starts from row 0, the empty suffix at position n - 1.
// GSX_ISA_FROM_SAMPLES=0 (tests) ignores the samples, so that every window takes that second way.
uint32_t lf_step(const DevStrand& st, uint32_t r) {
    if (st.n_exc && r >= st.exc_lo && r <= st.exc_hi) {
        const uint32_t k = lower_bound_u32(st.exc_rows, st.n_exc, r);
        if (k < st.n_exc && st.exc_rows[k] == r) return st.exc_lf[k];
    }
    const OccBlock& b = st.blocks[r >> 6]; uint32_t o[4];
    block_occ(st, b.cnt, b.hi, b.lo, r, o);
    const uint32_t sy = block_sym(b.hi, b.lo, r);
    return st.C[sy] + o[sy];
}

std::vector<uint64_t> isa_samples(const HostStrand& h, unsigned threads) {
    const DevStrand st = host_view(h);
    const uint64_t n = h.n, n_k = (n - 1) / 8192 + 1, none = ~0ull;
    std::vector<uint64_t> best(n_k, none);                          // (p mod 8192) << 32 | row
    const uint64_t n_samp = h.sa_samples.size();
    threads = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(threads, n_samp / 65536 + 1));
    const char* knob = getenv("GSX_ISA_FROM_SAMPLES");
    if (!knob || atoi(knob) != 0) {
        std::vector<std::vector<uint64_t>> part(threads);
        auto scan = [&](unsigned i) {
            std::vector<uint64_t>& mine = part[i]; mine.assign(n_k, none);
            for (uint64_t s = n_samp * i / threads, e = n_samp * (i + 1) / threads; s < e; s++) {
                const uint64_t p = h.sa_samples[s], key = ((p & 8191ull) << 32) | (s << h.sa_shift);
                if (key < mine[p >> 13]) mine[p >> 13] = key;
            }
        };
        std::vector<std::thread> pool;
        for (unsigned i = 1; i < threads; i++) pool.emplace_back(scan, i);
        scan(0);
        for (auto& th : pool) th.join();
        for (auto& mine : part) for (uint64_t k = 0; k < n_k; k++) best[k] = std::min(best[k], mine[k]);
    }
    std::vector<uint64_t> isa(n_k, 0);
    std::atomic<uint64_t> next{0};
    auto walk = [&]() {
        for (;;) {
            const uint64_t k0 = next.fetch_add(256);
            if (k0 >= n_k) return;
            for (uint64_t k = k0; k < std::min(n_k, k0 + 256); k++) {
                if (best[k] == none) continue;
                uint32_t r = (uint32_t)best[k];
                for (uint64_t steps = best[k] >> 32; steps; steps--) r = lf_step(st, r);
                isa[k] = r;
            }
        }
    };
    {
        std::vector<std::thread> pool;
        for (unsigned i = 1; i < threads; i++) pool.emplace_back(walk);
        walk();
        for (auto& th : pool) th.join();
    }
    for (uint64_t k = n_k; k-- > 0;) {
        if (best[k] != none) continue;
        uint32_t r = k + 1 == n_k ? 0u : (uint32_t)isa[k + 1];
        for (uint64_t steps = (k + 1 == n_k ? n - 1 : 8192 * (k + 1)) - 8192 * k; steps; steps--) r = lf_step(st, r);
        isa[k] = r;
    }
    return isa;
}

}  // namespace

// ---- test access (tests/sdsl_write_check.cpp; not part of the C ABI) ------------------------------------------------------
bool sdsl_write_bit_vector_supports(const std::string& path, const std::vector<uint64_t>& words, uint64_t n_bits, std::string& err) {
    Sink s; s.f = fopen(path.c_str(), "wb");
    if (!s.f) { err = "cannot write " + path; return false; }
    put_rank_directory(s, words);
    put_select(s, build_select(words, n_bits, true, 3));
    put_select(s, build_select(words, n_bits, false, 3));
    const bool ok = s.ok; fclose(s.f);
    if (!ok) err = "short write to " + path;
    return ok;
}

// size, sigma, bits, rank and select directories, tree (wt_pc.hpp:656-671)
static bool put_wavelet_tree(Sink& s, const HostStrand& h, const uint64_t count[256], unsigned threads) {
    const WtShape t = huffman_shape(count);
    const bool timing = getenv("GSX_SDSL_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (timing) { auto t1 = std::chrono::steady_clock::now(); fprintf(stderr, "  %-28s %.2f s\n", what, std::chrono::duration<double>(t1 - t0).count()); t0 = t1; } };
    const std::vector<uint64_t> bv = fill_bit_vector(h, t, threads);
    lap("bit vector");
    s.u64(h.n); s.u64(t.sigma);
    s.u64(t.bv_bits); s.bytes(bv.data(), bv.size() * 8);
    put_rank_directory(s, bv);
    lap("bits + rank directory out");
    SelectDir sel1, sel0;
    {
        std::thread other([&] { sel0 = build_select(bv, t.bv_bits, false, (threads + 1) / 2); });
        sel1 = build_select(bv, t.bv_bits, true, (threads + 1) / 2);
        other.join();
    }
    lap("select directories");
    put_select(s, sel1); sel1 = SelectDir();
    put_select(s, sel0); sel0 = SelectDir();
    s.u64(t.nodes.size());
    // an inner node's bv_pos_rank = the ones before its first bit (wt_helper.hpp:235-241); nodes come in bit order
    uint64_t total = 0, w = 0;
    for (const WtShape::Node& nd : t.nodes) {
        uint64_t second = nd.bv_pos_rank;
        if (nd.child[0] != 0xFFFF) {
            while (w < (nd.bv_pos >> 6)) total += (uint64_t)__builtin_popcountll(bv[w++]);
            const uint32_t off = (uint32_t)(nd.bv_pos & 63);
            second = total + (off ? (uint64_t)__builtin_popcountll(bv[w] & (~0ull >> (64 - off))) : 0ull);
        }
        s.u64(nd.bv_pos); s.u64(second); s.u16(nd.parent); s.u16(nd.child[0]); s.u16(nd.child[1]);
    }
    s.bytes(t.leaf_of, sizeof t.leaf_of); s.bytes(t.path, sizeof t.path);
    return s.ok;
}

bool sdsl_write_wavelet_tree(const std::string& path, const HostStrand& h, unsigned threads, std::string& err) {
    Sink s; s.f = fopen(path.c_str(), "wb");
    if (!s.f) { err = "cannot write " + path; return false; }
    uint64_t count[256]; byte_counts(h, count);
    const bool ok = put_wavelet_tree(s, h, count, threads ? threads : 1); fclose(s.f);
    if (!ok) err = "short write to " + path;
    return ok;
}

bool save_sdsl_strand(const std::string& path, const HostStrand& h, unsigned threads, std::string& err) {
    if (h.sa_shift > 6) { err = "the reference format needs an SA sample every 64 rows; this index samples every " + std::to_string(1u << h.sa_shift); return false; }
    if (h.n < 2 || h.blocks.size() != h.n / 64 + 1) { err = "no host copy of the index to write"; return false; }
    if (!threads) threads = std::max(1u, std::thread::hardware_concurrency());
    uint64_t count[256]; byte_counts(h, count);
    Sink s; s.f = fopen(path.c_str(), "wb");
    if (!s.f) { err = "cannot write " + path; return false; }
    std::vector<char> iobuf(8u << 20); setvbuf(s.f, iobuf.data(), _IOFBF, iobuf.size());
    put_wavelet_tree(s, h, count, threads);
    // SA samples: entry k = SA[64 k], width = bits of n (csa_sampling_strategy.hpp:85-101)
    const uint32_t width = top_bit(h.n) + 1;
    {
        const uint64_t n_samp = (h.n + 63) / 64, stride = 1ull << (6 - h.sa_shift);
        Packed sa(n_samp, width);
        for (uint64_t k = 0; k < n_samp; k++) sa.set(k, h.sa_samples[k * stride]);
        sa.put(s);
    }
    {
        const std::vector<uint64_t> isa = isa_samples(h, threads);
        Packed v(isa.size(), width);
        for (uint64_t k = 0; k < isa.size(); k++) v.set(k, isa[k]);
        v.put(s);
    }
    // byte_alphabet (lib/csa_alphabet_strategy.cpp:25-55): char2comp over all 256 bytes, comp2char and the cumulative counts over
    // the bytes that occur
    {
        uint8_t char2comp[256] = {0}; std::vector<uint8_t> comp2char; std::vector<uint64_t> C{0};
        for (uint32_t c = 0; c < 256; c++) if (count[c]) { char2comp[c] = (uint8_t)comp2char.size(); comp2char.push_back((uint8_t)c); C.push_back(C.back() + count[c]); }
        s.u64(256 * 8); s.bytes(char2comp, 256);
        std::vector<uint8_t> padded((comp2char.size() + 7) / 8 * 8, 0); memcpy(padded.data(), comp2char.data(), comp2char.size());
        s.u64(comp2char.size() * 8); s.bytes(padded.data(), padded.size());
        s.u64(C.size() * 64); s.bytes(C.data(), C.size() * 8);
        s.u16((uint16_t)comp2char.size());
    }
    const bool ok = s.ok && fflush(s.f) == 0; fclose(s.f);
    if (!ok) err = "short write to " + path;
    return ok;
}

bool save_sdsl_index(const std::string& prefix, const HostIndex& ix, std::string& err) {
    const unsigned cores = std::max(2u, std::thread::hardware_concurrency());
    bool ok[2] = {false, false}; std::string e[2];
    std::thread rev([&] { ok[1] = save_sdsl_strand(prefix + ".reverse", ix.st[1], cores / 2, e[1]); });
    ok[0] = save_sdsl_strand(prefix + ".forward", ix.st[0], cores / 2, e[0]);
    rev.join();
    if (!ok[0] || !ok[1]) { err = ok[0] ? e[1] : e[0]; return false; }
    FILE* gs = fopen((prefix + ".gs").c_str(), "w");                 // seq_io.cxx:112-122: name and length on alternating lines
    if (!gs) { err = "cannot write " + prefix + ".gs"; return false; }
    for (size_t i = 0; i < ix.chr_names.size(); i++) fprintf(gs, "%s\n%llu\n", ix.chr_names[i].c_str(), (unsigned long long)ix.chr_lens[i]);
    const bool gok = fflush(gs) == 0; fclose(gs);
    if (!gok) err = "short write to " + prefix + ".gs";
    return gok;
}

}  // namespace gsx
