// tests/host_core_check.cpp -- UNIT-TEST HARNESS (not part of the product, never shipped in libgsx.so).
//
// Runs the per-node arithmetic that the CUDA kernels are built from (guidescan-cli_b200/csrc/gsx_core.h: occurrence
// lookup in the 32-byte block layout, child generation, string keys, LF-walk locate, coordinates, CFD, specificity)
// as a plain sequential depth-first search on the host, over an index laid out by the product's own loader, and
// writes CSV / SAM through the product's own formatter.  tests/test_host_core.py diffs that text against the
// reference's golden output.  This pins every piece of device arithmetic on a machine without a GPU; what remains
// GPU-only (warp stacks, ballots, atomics, arenas) is covered by the -m gpu parity tests.
//
// usage: host_core_check <index prefix> <guides.csv> <out> [-m N] [--rna N] [--dna N] [-t N] [--start] [--max N]
//                        [--sam] [--succinct] [-a PAM]...
#include "../include/gsx.h"
#include "../guidescan-cli_b200/csrc/gsx_host.h"
#include "../guidescan-cli_b200/csrc/gsx_core.h"
#include "../guidescan-cli_b200/csrc/cfd_tables.h"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

using namespace gsx;

struct HostBlockLoader {
    void operator()(const OccBlock* p, uint32_t c[4], uint64_t& hi, uint64_t& lo) const { for (int i = 0; i < 4; i++) c[i] = p->cnt[i]; hi = p->hi; lo = p->lo; }
};

static DevStrand view_of(const HostStrand& h) {
    DevStrand d{};
    d.blocks = h.blocks.data(); d.sa_samples = h.sa_samples.data(); d.exc_rows = h.exc_rows.data(); d.exc_lf = h.exc_lf.data();
    d.n_rows = h.n_rows.data(); d.n = (uint32_t)h.n; d.n_exc = (uint32_t)h.exc_rows.size(); d.n_nrows = (uint32_t)h.n_rows.size();
    d.sa_shift = h.sa_shift; for (int c = 0; c < 5; c++) d.C[c] = h.C[c];
    d.exc_lo = h.exc_rows.empty() ? 0xFFFFFFFFu : h.exc_rows.front(); d.exc_hi = h.exc_rows.empty() ? 0 : h.exc_rows.back();
    d.blk_shift = 5; d.lines = nullptr; d.ftab = nullptr; d.ftab_L = 0;
    return d;
}

// look-ahead planes t1..t6 per 64-row block (what build_lookahead_kernel writes behind each OccBlock), built on the host
static std::vector<uint64_t> g_look[2];     // 12 words per block: hi1, lo1, ..., hi6, lo6
static bool g_prune = false;
static bool g_fused = false;     // alternative PAMs in one pass (gsx_core.h fused_pam_ok), as gsx_enumerate does under GSX_FUSED_PAMS=1
static bool g_forced = false;    // the sweep skips patterns that substitute an inserted position of an edited guide (sweep_kernel<..., FORCED>)
static std::vector<uint32_t> g_vfmask;
static int g_parts = 1;          // split the result into this many parts, as gsx_enumerate does over several devices (formatter / merged view)
static bool g_variants = false;  // bulges through edited guides (gsx_core.h variant_rewrite), as gsx_enumerate does when Prepared::variant_ok

static std::vector<uint64_t> g_tail[2];     // 4 words per block: hi7, lo7, hi8, lo8 (the scratch planes the summaries' sum2 is built from)
static void build_look(const DevStrand& st, std::vector<uint64_t>& look, std::vector<uint64_t>& tail) {
    const uint32_t nb = st.n / 64 + 1;
    look.assign((size_t)nb * 12, 0); tail.assign((size_t)nb * 4, 0);
    for (uint32_t row = 0; row < st.n; row++) {
        uint32_t cur = row;
        for (int j = 1; j <= 8; j++) {
            // cur = LF(cur)
            bool exc = false;
            if (st.n_exc && cur >= st.exc_lo && cur <= st.exc_hi) {
                uint32_t k = lower_bound_u32(st.exc_rows, st.n_exc, cur);
                if (k < st.n_exc && st.exc_rows[k] == cur) { cur = st.exc_lf[k]; exc = true; }
            }
            if (!exc) {
                const OccBlock& b = st.blocks[cur >> 6]; uint32_t o[4];
                block_occ(st, b.cnt, b.hi, b.lo, cur, o);
                uint32_t sy = block_sym(b.hi, b.lo, cur);
                cur = st.C[sy] + o[sy];
            }
            const OccBlock& c = st.blocks[cur >> 6];
            uint32_t sy = block_sym(c.hi, c.lo, cur);       // exception rows read as code 0, as on the device
            uint64_t* dst = j <= 6 ? &look[(size_t)(row >> 6) * 12 + 2 * (j - 1)] : &tail[(size_t)(row >> 6) * 4 + 2 * (j - 7)];
            dst[0] |= (uint64_t)(sy >> 1) << (row & 63);
            dst[1] |= (uint64_t)(sy & 1) << (row & 63);
        }
    }
}

// k-mer jump table (what build_ftab does on the device), built on the host level by level
static uint32_t g_ftab_L = 0;
static std::vector<FtabEntry> g_ftab[2];
static uint64_t g_pow5[32];

static void build_ftab_host(const DevStrand& st, uint32_t L, std::vector<FtabEntry>& tab) {
    std::vector<FtabEntry> cur(1); cur[0] = {0, st.n};
    for (uint32_t d = 0; d < L; d++) {
        std::vector<FtabEntry> nxt(cur.size() * 4);
        for (size_t e = 0; e < cur.size(); e++) {
            uint32_t os[4] = {0, 0, 0, 0}, oe[4] = {0, 0, 0, 0};
            if (cur[e].width) {
                uint32_t sp = cur[e].sp, e1 = sp + cur[e].width;
                const OccBlock& b0 = st.blocks[sp >> 6]; const OccBlock& b1 = st.blocks[e1 >> 6];
                block_occ(st, b0.cnt, b0.hi, b0.lo, sp, os); block_occ(st, b1.cnt, b1.hi, b1.lo, e1, oe);
            }
            for (uint32_t s = 0; s < 4; s++) nxt[e + s * cur.size()] = {st.C[s] + os[s], oe[s] - os[s]};
        }
        cur.swap(nxt);
    }
    tab.swap(cur);
}

// ---- independent statement of the sweep kernel's row filter, row block by row block over the 128-byte look-ahead lines
// (the product evaluates the same predicate from its pattern summaries: gsx_core.h summary_eval*; the harness checks that
// both agree on every pattern it visits) --------------------------------------------------------------------------------
namespace gsx {
// can ANY row of [sp, ep] (inside one block or two adjacent ones) still reach the
// final level?  ld(block, k, w) fetches 32-byte sector k of the block's 128-byte line as four 64-bit words: k = 0 the
// OccBlock {cnt01, cnt23, hi, lo}, k = 1..3 the planes {hi_(2k-1), lo_(2k-1), hi_2k, lo_2k}.  Sectors are fetched in
// pairs (0,1) then (2,3) -- two independent loads in flight -- and the second pair only while some row is alive.
// State: u[r] = rows with at most budget - r mismatches so far (r = 0 .. NB-1, NB > budget), so u[0] = rows still alive;
// one plane costs one and-or per mask.  Intervals spanning more than two blocks are not examined (true).
template <int NB, class LoadSector>
GSX_HD bool node_viable(LoadSector ld, uint32_t sp, uint32_t ep, uint32_t lvl, uint32_t qlen, uint32_t total, uint64_t q,
                        uint32_t pampack, uint32_t budget, uint32_t& sectors) {
    const uint32_t e1 = ep + 1u, bs = sp >> 6, be = e1 >> 6;
    if (be - bs > 1u) return true;
    const uint32_t left = total - lvl;
    for (uint32_t part = 0; part < 2u; part++) {
        if (part == 1u && (be == bs || (e1 & 63u) == 0u)) break;
        const uint32_t r0 = part ? 0u : (sp & 63u), r1 = part ? (ep & 63u) : (be != bs ? 63u : (ep & 63u));
        const uint64_t rows = (r1 == 63u ? ~0ull : ((1ull << (r1 + 1u)) - 1ull)) & ~((1ull << r0) - 1ull);
        uint64_t u[NB];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int r = 0; r < NB; r++) u[r] = budget >= (uint32_t)r ? rows : 0ull;
        for (uint32_t pair = 0; pair < 2u; pair++) {
            const uint32_t j0 = pair ? 3u : 0u;
            if (j0 >= left || !u[0]) break;
            uint64_t wa[4], wb[4] = {0, 0, 0, 0};
            ld(bs + part, 2u * pair, wa); sectors++;
            if ((pair ? 5u : 1u) < left) { ld(bs + part, 2u * pair + 1u, wb); sectors++; }
            // planes of this pair: pair 0 -> t0 (OccBlock), t1, t2 ; pair 1 -> t3, t4, t5, t6
            const uint64_t ph[4] = {pair ? wa[0] : wa[2], pair ? wa[2] : wb[0], pair ? wb[0] : wb[2], wb[2]};
            const uint64_t pl[4] = {pair ? wa[1] : wa[3], pair ? wa[3] : wb[1], pair ? wb[1] : wb[3], wb[3]};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (uint32_t t = 0; t < 4u; t++) {
                const uint32_t j = j0 + t;
                if (j >= left || (pair == 0u && t == 3u)) break;
                const uint32_t Lv = lvl + j;
                uint32_t sym; const bool proto = Lv < qlen; bool wild = false, kill = false;
                if (proto) sym = (uint32_t)(q >> (2u * Lv)) & 3u;
                else { const uint32_t pc = (pampack >> (3u * (Lv - qlen))) & 7u; sym = pc & 3u; wild = pc == 4u; kill = pc > 4u; }
                const uint64_t eq = ~(ph[t] ^ ((sym & 2u) ? ~0ull : 0ull)) & ~(pl[t] ^ ((sym & 1u) ? ~0ull : 0ull));
                if (proto) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                    for (int r = 0; r + 1 < NB; r++) u[r] = (u[r] & eq) | u[r + 1];
                    u[NB - 1] &= eq;
                } else if (!wild) {
                    const uint64_t keep = kill ? 0ull : eq;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                    for (int r = 0; r < NB; r++) u[r] &= keep;
                }
            }
        }
        if (u[0]) return true;
    }
    return false;
}

// The same test for a node with no budget left (9 of 10 level-L nodes at m = 3): a row survives a protospacer plane only
// if its symbol equals the query character, so one mask is enough.
template <class LoadSector>
GSX_HD bool node_viable_exact(LoadSector ld, uint32_t sp, uint32_t ep, uint32_t lvl, uint32_t qlen, uint32_t total, uint64_t q,
                              uint32_t pampack, uint32_t& sectors) {
    const uint32_t e1 = ep + 1u, bs = sp >> 6, be = e1 >> 6;
    if (be - bs > 1u) return true;
    const uint32_t left = total - lvl;
    for (uint32_t part = 0; part < 2u; part++) {
        if (part == 1u && (be == bs || (e1 & 63u) == 0u)) break;
        const uint32_t r0 = part ? 0u : (sp & 63u), r1 = part ? (ep & 63u) : (be != bs ? 63u : (ep & 63u));
        uint64_t alive = (r1 == 63u ? ~0ull : ((1ull << (r1 + 1u)) - 1ull)) & ~((1ull << r0) - 1ull);
        uint64_t wa[4], wb[4] = {0, 0, 0, 0};
#define GSX_EXACT_PLANE(J, HI, LO)                                                                                     \
        if ((J) < left) {                                                                                             \
            const uint32_t Lv = lvl + (J);                                                                            \
            uint32_t pc = Lv < qlen ? ((uint32_t)(q >> (2u * Lv)) & 3u) : ((pampack >> (3u * (Lv - qlen))) & 7u);       \
            if (pc < 4u) alive &= ~((HI) ^ ((pc & 2u) ? ~0ull : 0ull)) & ~((LO) ^ ((pc & 1u) ? ~0ull : 0ull));         \
            else if (pc > 4u) alive = 0;                                                                              \
        }
        ld(bs + part, 0u, wa); sectors++;
        if (1u < left) { ld(bs + part, 1u, wb); sectors++; }
        GSX_EXACT_PLANE(0u, wa[2], wa[3]) GSX_EXACT_PLANE(1u, wb[0], wb[1]) GSX_EXACT_PLANE(2u, wb[2], wb[3])
        if (alive && 3u < left) {
            ld(bs + part, 2u, wa); sectors++;
            if (5u < left) { ld(bs + part, 3u, wb); sectors++; }
            GSX_EXACT_PLANE(3u, wa[0], wa[1]) GSX_EXACT_PLANE(4u, wa[2], wa[3]) GSX_EXACT_PLANE(5u, wb[0], wb[1]) GSX_EXACT_PLANE(6u, wb[2], wb[3])
        }
#undef GSX_EXACT_PLANE
        if (alive) return true;
    }
    return false;
}

}  // namespace gsx

// slice-major front end (what sweep_kernel does): per task, the level-L nodes that pass node_viable
static uint32_t g_sweep_sb = 0;
static unsigned long long g_shape_checks[4] = {0, 0, 0, 0};      // patterns checked per plane layout of sweep_lean_kernel ([3]: other layouts)
static std::map<uint32_t, std::vector<std::vector<Node>>> g_seeds_by_M;
struct HostSectorLoader {            // whole 64-row lines (node_viable / node_viable_exact)
    const DevStrand* st; const std::vector<uint64_t>* look;
    void operator()(uint32_t b, uint32_t k, uint64_t w[4]) const {
        if (k == 0) { const OccBlock& o = st->blocks[b]; w[0] = ((uint64_t)o.cnt[1] << 32) | o.cnt[0]; w[1] = ((uint64_t)o.cnt[3] << 32) | o.cnt[2]; w[2] = o.hi; w[3] = o.lo; }
        else for (int u = 0; u < 4; u++) w[u] = (*look)[(size_t)b * 12 + 4 * (k - 1) + u];      // hi_(2k-1), lo_(2k-1), hi_2k, lo_2k
    }
};
// pattern summaries (what build_summary_kernel writes), built on the host through the same summary_build
static std::vector<uint32_t> g_sum0[2], g_sum1[2], g_sum2[2];
struct HostPlane {
    const DevStrand* st; const std::vector<uint64_t>* look; const std::vector<uint64_t>* tail;
    uint64_t operator()(uint32_t b, uint32_t j, bool hi) const {
        if (j == 0) return hi ? st->blocks[b].hi : st->blocks[b].lo;
        if (j >= 7) return (*tail)[(size_t)b * 4 + 2 * (j - 7) + (hi ? 0 : 1)];
        return (*look)[(size_t)b * 12 + 2 * (j - 1) + (hi ? 0 : 1)];
    }
};
static void build_summaries_host(const DevStrand& st, const std::vector<uint64_t>& look, const std::vector<uint64_t>& tail, const std::vector<FtabEntry>& tab,
                                 std::vector<uint32_t>& s0, std::vector<uint32_t>& s1, std::vector<uint32_t>& s2) {
    s0.assign(tab.size() * 8, 0); s1.assign(tab.size() * 8, 0); s2.assign(tab.size() * 4, 0);
    HostPlane plane{&st, &look, &tail};
    for (size_t e = 0; e < tab.size(); e++) summary_build(plane, true, tab[e].sp, tab[e].width, &s0[e * 8], &s1[e * 8], &s2[e * 4]);
}
struct HostSummaryLoader {
    const std::vector<uint32_t>* s0; const std::vector<uint32_t>* s1; const std::vector<uint32_t>* s2;
    void operator()(uint32_t stage, uint32_t idx, uint32_t w[8]) const {
        if (stage == 2) { for (int i = 0; i < 4; i++) w[i] = (*s2)[(size_t)idx * 4 + i]; return; }
        for (int i = 0; i < 8; i++) w[i] = (stage ? *s1 : *s0)[(size_t)idx * 8 + i];
    }
};
// independent statement of what the full filter decides: follow ONE row for up to n_levels characters (LF walk over the
// packed blocks; a non-ACGT row reads as code 0 and steps through its precomputed LF target, as on the device)
static bool row_reaches(const DevStrand& st, uint32_t row, uint32_t lvl, uint32_t n_levels, uint64_t q, uint32_t plen, uint32_t pampack, uint32_t budget) {
    const uint32_t qlen = (uint32_t)(q >> 58);
    uint32_t cur = row, mm = 0;
    for (uint32_t j = 0; j < n_levels && lvl + j < qlen + plen; j++) {
        const uint32_t Lv = lvl + j;
        const OccBlock& b = st.blocks[cur >> 6];
        const uint32_t sy = block_sym(b.hi, b.lo, cur);
        if (Lv < qlen) { if (sy != ((uint32_t)(q >> (2 * Lv)) & 3u) && ++mm > budget) return false; }
        else { const uint32_t pc = (pampack >> (3 * (Lv - qlen))) & 7u; if (pc > 4u || (pc < 4u && pc != sy)) return false; }
        bool exc = false;
        if (st.n_exc && cur >= st.exc_lo && cur <= st.exc_hi) {
            uint32_t k = lower_bound_u32(st.exc_rows, st.n_exc, cur);
            if (k < st.n_exc && st.exc_rows[k] == cur) { cur = st.exc_lf[k]; exc = true; }
        }
        if (!exc) { uint32_t o[4]; block_occ(st, b.cnt, b.hi, b.lo, cur, o); cur = st.C[sy] + o[sy]; }
    }
    return true;
}
static void sweep_host(const DevStrand st[2], const Prepared& prep, uint32_t M) {
    const uint32_t L = g_ftab_L, sb = g_sweep_sb; const size_t n = prep.recs.size();
    SweepPlan plan; std::vector<uint32_t> masks; sweep_make_plan(L, sb, M, plan, masks);
    std::vector<std::vector<Node>>& g_seeds = g_seeds_by_M[M]; g_seeds.assign(2 * n, {});
    static std::vector<uint64_t> combos; combos = ftab_combos(L - 2, M);
    for (uint32_t strand = 0; strand < 2; strand++) {
        HostSectorLoader ld{&st[strand], &g_look[strand]}; HostSummaryLoader lds{&g_sum0[strand], &g_sum1[strand], &g_sum2[strand]};
        std::vector<std::vector<std::pair<uint32_t, uint32_t>>> seen(n);
        for (uint32_t beta = 0; beta < (1u << (2 * sb)); beta++)
            for (size_t g = 0; g < n; g++) {
                const uint64_t q = prep.gq[g]; const uint32_t qlen = (uint32_t)(q >> 58);
                if (qlen < L) { fprintf(stderr, "--sweep needs guides of at least L characters\n"); exit(2); }
                const uint32_t h = sweep_slice_distance(q, L, sb, beta);
                if (h > M) continue;
                const uint32_t B = M - h;
                const uint32_t fm = (g_forced && g < g_vfmask.size() ? g_vfmask[g] : 0u) & (L >= 16 ? ~0u : ((1u << (2 * L)) - 1u));
                for (int zero = 1; zero >= 0; zero--) {
                    const uint32_t n_pat = plan.xcnt[zero][B];
                    for (uint32_t t = 0; t < n_pat; t++) {
                        uint32_t used = B;
                        const uint32_t idx = sweep_pattern(plan, masks.data(), (uint32_t)zero, q, beta, B, t, used);
                        if (zero && used != B) { fprintf(stderr, "pass 1 pattern keeps budget\n"); exit(3); }
                        const uint32_t mm = h + used;
                        if (!forced_kept(idx, q, fm)) continue;              // (the kernel: slice check + xor word check)
                        seen[g].push_back({idx, mm});
                        const FtabEntry& e = g_ftab[strand][idx];
                        if (!e.width) continue;
                        uint32_t sectors = 0;
                        const bool ok = zero ? node_viable_exact(ld, e.sp, e.sp + e.width - 1, L, qlen, qlen + prep.plen, q, prep.pampack, sectors)
                                             : node_viable<kMaxDist>(ld, e.sp, e.sp + e.width - 1, L, qlen, qlen + prep.plen, q, prep.pampack, M - mm, sectors);
                        // what the kernel runs (summary_step0/1 over the pattern summaries) must agree with the row-by-row forms
                        const uint32_t codes = sweep_codes(q, L, prep.plen, prep.pampack);
                        const uint32_t codes2 = sweep_codes2(q, L, prep.plen, prep.pampack);
                        const bool ok7 = summary_viable<kMaxDist>(lds, idx, codes, 0x77u, M - mm);          // seven levels only
                        const bool ok2 = summary_viable<kMaxDist>(lds, idx, codes, codes2, M - mm);         // + levels L+7, L+8 (sum2)
                        if (e.width <= 32) {                                                                // against single-row LF walks
                            bool any9 = false, any7 = false;
                            for (uint32_t r = 0; r < e.width; r++) { any9 = any9 || row_reaches(st[strand], e.sp + r, L, 9, q, prep.plen, prep.pampack, M - mm);
                                                                     any7 = any7 || row_reaches(st[strand], e.sp + r, L, 7, q, prep.plen, prep.pampack, M - mm); }
                            if (any9 != ok2 || any7 != ok7) { fprintf(stderr, "summary filter disagrees with the row walks (idx %u: %d/%d vs %d/%d)\n", idx, (int)ok7, (int)ok2, (int)any7, (int)any9); exit(3); }
                        } else if (!ok2) { fprintf(stderr, "summary filter dropped a node of more than 32 rows (idx %u)\n", idx); exit(3); }
                        if (e.width <= 32 ? ok7 != ok : !ok7) { fprintf(stderr, "summary filter disagrees with node_viable (idx %u zero %d ok %d ok7 %d sp %u w %u budget %u)\n", idx, zero, (int)ok, (int)ok7, e.sp, e.width, M - mm); exit(3); }
                        {   // the hoisted-mask forms the kernel's main loop uses
                            uint32_t gm[15], w[8], u1[kMaxDist], u2[kMaxDist];
                            summary_masks(codes, gm);
                            for (uint32_t stage = 0; stage < 2; stage++) {
                                lds(stage, idx, w);
                                summary_eval<kMaxDist>(lds, stage, idx, codes, M - mm, u1);
                                summary_eval_masks<kMaxDist>(w, gm, M - mm, u2);
                                for (int r = 0; r < kMaxDist; r++) if ((u1[r] & 0xFFFFu) != (u2[r] & 0xFFFFu)) { fprintf(stderr, "summary_eval_masks disagrees (idx %u stage %u r %d)\n", idx, stage, r); exit(3); }
                                if (zero && summary_eval_exact(w, gm) != (u1[0] & 0xFFFFu)) { fprintf(stderr, "summary_eval_exact disagrees (idx %u stage %u)\n", idx, stage); exit(3); }
                                // the per-layout forms of sweep_lean_kernel
                                const int sh = sweep_shape_of(codes);
                                if (sh >= 0) {
                                    uint32_t u3[kMaxDist]; uint32_t ex;
                                    if (sh == 0) { summary_masks_shape<0x3Fu, 0u, kMaxDist>(w, gm + 7, M - mm, u3); ex = summary_exact_shape<0x3Fu>(w, gm + 7); }
                                    else if (sh == 1) { summary_masks_shape<0x7Fu, 0u, kMaxDist>(w, gm + 7, M - mm, u3); ex = summary_exact_shape<0x7Fu>(w, gm + 7); }
                                    else { summary_masks_shape<0x1Fu, 0x40u, kMaxDist>(w, gm + 7, M - mm, u3); ex = summary_exact_shape<0x5Fu>(w, gm + 7); }
                                    if (sweep_shape_proto(sh) != gm[14]) { fprintf(stderr, "sweep_shape_of disagrees with summary_masks (codes %x)\n", codes); exit(3); }
                                    for (int r = 0; r < kMaxDist; r++) if ((u1[r] & 0xFFFFu) != (u3[r] & 0xFFFFu)) { fprintf(stderr, "summary_masks_shape disagrees (idx %u stage %u r %d shape %d)\n", idx, stage, r, sh); exit(3); }
                                    if (zero && ex != (u1[0] & 0xFFFFu)) { fprintf(stderr, "summary_exact_shape disagrees (idx %u stage %u shape %d)\n", idx, stage, sh); exit(3); }
                                    g_shape_checks[sh]++;
                                } else g_shape_checks[3]++;
                            }
                        }
                        if (zero && summary_viable<1>(lds, idx, codes, codes2, 0) != ok2) { fprintf(stderr, "summary filter <1> disagrees (idx %u)\n", idx); exit(3); }
                        if (!ok2) continue;
                        Node nd{}; nd.sp = e.sp; nd.ep = e.sp + e.width - 1; nd.key_lo = ftab_key(idx, q, L); nd.task = (uint32_t)(2 * g + strand);
                        nd.meta = meta_make(L, mm, 0, 0, 0, 0, 0);
                        g_seeds[2 * g + strand].push_back(nd);
                    }
                }
            }
        // the slice-major enumeration must visit exactly the patterns (and mismatch counts) of the per-guide enumeration
        for (size_t g = 0; g < n; g++) {
            std::vector<std::pair<uint32_t, uint32_t>> want; const uint64_t q = prep.gq[g];
            for (uint64_t combo : combos) {
                uint32_t idx, j; uint64_t key; ftab_apply(combo, q, L, ftab_exact_index(q, L), g_pow5, idx, key, j);
                for (uint32_t e = 0; e < 16; e++) { uint64_t k2 = key; uint32_t extra = ftab_beginning(e, q, L, g_pow5, k2);
                    const uint32_t fmg = (g_forced && g < g_vfmask.size() ? g_vfmask[g] : 0u) & (L >= 16 ? ~0u : ((1u << (2 * L)) - 1u));
                    if (j + extra <= M && forced_kept((idx & ~15u) | e, q, fmg)) want.push_back({(idx & ~15u) | e, j + extra}); }
            }
            std::sort(want.begin(), want.end()); std::sort(seen[g].begin(), seen[g].end());
            if (want != seen[g]) { fprintf(stderr, "sweep enumeration differs from the per-guide enumeration for guide %zu (%zu vs %zu patterns)\n", g, seen[g].size(), want.size()); exit(3); }
        }
    }
}

template <bool WIDE>
static void dfs(const DevStrand st[2], const Prepared& prep, uint32_t task, uint32_t M, uint32_t R, uint32_t D, bool counting,
                uint64_t* count, std::vector<MatchRec>& out, uint64_t* nodes) {
    const DevStrand& s = st[task & 1];
    const GuideRec& g = prep.recs[task >> 1];
    ExpandCtx cx{&s, &g, &prep.pamsets[g.pamset], M, R, D};
    std::vector<Node> stack;
    if (!WIDE && g_sweep_sb && prep.fast_ok) {
        if (!g_seeds_by_M.count(M)) sweep_host(st, prep, M);
        stack = g_seeds_by_M[M][task];
    } else if (!WIDE && g_ftab_L && g.qlen >= g_ftab_L && prep.fast_ok) {
        // the table phase of search_fast_kernel: every pattern within the budget of the first L characters
        const uint32_t L = g_ftab_L; const uint64_t q = prep.gq[task >> 1];
        const uint32_t gidx = ftab_exact_index(q, L);
        static std::vector<uint64_t> combos; static uint32_t combos_M = ~0u;
        if (combos_M != M) { combos = ftab_combos(L - 2, M); combos_M = M; }
        for (uint64_t combo : combos) {
            uint32_t idx, j; uint64_t key;
            ftab_apply(combo, q, L, gidx, g_pow5, idx, key, j);
            for (uint32_t e = 0; e < 16; e++) {
                uint64_t k2 = key; uint32_t extra = ftab_beginning(e, q, L, g_pow5, k2);
                if (j + extra > M) continue;
                const FtabEntry& t = g_ftab[task & 1][(idx & ~15u) | e];
                if (!t.width) continue;
                Node nd{}; nd.sp = t.sp; nd.ep = t.sp + t.width - 1; nd.key_lo = k2; nd.task = task; nd.meta = meta_make(L, j + extra, 0, 0, 0, 0, 0);
                stack.push_back(nd);
            }
        }
    } else {
        Node root{}; root.sp = 0; root.ep = s.n - 1; root.task = task;
        stack.push_back(root);
    }
    while (!stack.empty()) {
        Node nd = stack.back(); stack.pop_back();
        (*nodes)++;
        uint32_t os[4], oe[4];
        const OccBlock& b0 = s.blocks[nd.sp >> 6]; const OccBlock& b1 = s.blocks[(nd.ep + 1) >> 6];
        block_occ(s, b0.cnt, b0.hi, b0.lo, nd.sp, os);
        block_occ(s, b1.cnt, b1.hi, b1.lo, nd.ep + 1, oe);
        uint32_t vmask = 15;
        if (g_prune && !WIDE && ((nd.ep + 1) >> 6) - (nd.sp >> 6) <= 1u) {       // the pruning step of search_fast_kernel
            const std::vector<uint64_t>& lk = g_look[task & 1];
            const uint32_t bs = nd.sp >> 6, e1 = nd.ep + 1, be = e1 >> 6;
            vmask = 0;
            for (uint32_t part = 0; part < (be != bs ? 2u : 1u); part++) {
                if (part == 1 && (e1 & 63u) == 0) break;
                const uint32_t b = bs + part;
                const uint32_t r_lo = part ? 0u : nd.sp, r_hi = part ? nd.ep : (be != bs ? 63u : nd.ep);
                uint64_t hi[7], lo[7]; hi[0] = s.blocks[b].hi; lo[0] = s.blocks[b].lo;
                for (int j = 1; j <= 6; j++) { hi[j] = lk[(size_t)b * 12 + 2 * (j - 1)]; lo[j] = lk[(size_t)b * 12 + 2 * (j - 1) + 1]; }
                vmask |= viable_children(hi, lo, r_lo, r_hi, meta_lvl(nd.meta), g.qlen, g.qlen + prep.plen, prep.gq[task >> 1], prep.pampack, M - meta_mm(nd.meta));
            }
        }
        for (int cand = 0; cand < CAND_END; cand++) {
            if (!WIDE && cand > CAND_FORK) break;
            if (cand < 4 && !((vmask >> cand) & 1)) continue;
            Node ch; bool emit;
            if (!make_child<WIDE>(cand, nd, cx, os, oe, ch, emit)) continue;
            if (emit) {
                if (counting) *count += ch.ep - ch.sp + 1;
                else { MatchRec m; fill_match(m, ch, WIDE); out.push_back(m); }
            } else stack.push_back(ch);
        }
    }
}

// the per-layout row filters of sweep_lean_kernel against the general forms, on random sectors and random guides of every layout the
// kernel is compiled for (the golden genomes are too small for a jump table deep enough to produce layouts 0 and 2 naturally)
static int shape_selftest() {
    uint64_t rs = 0x243F6A8885A308D3ull;
    auto rnd = [&]() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 16); };
    unsigned long long checked = 0;
    for (int sh = 0; sh < kSweepShapes; sh++) {
        const uint32_t proto = sweep_shape_proto(sh), pam = sweep_shape_pam(sh);
        for (int it = 0; it < 200000; it++) {
            uint32_t codes = 0;
            for (uint32_t j = 0; j < 7; j++) {
                uint32_t c;
                if (proto & (1u << j)) c = rnd() & 3u; else if (pam & (1u << j)) c = 8u + (rnd() & 3u); else c = (rnd() & 1u) ? 12u : 7u;
                codes |= c << (4u * j);
            }
            if (sweep_shape_of(codes) != sh) { fprintf(stderr, "sweep_shape_of(%x) != %d\n", codes, sh); return 3; }
            uint32_t gm[15], w[8];
            summary_masks(codes, gm);
            // sectors as summary_build writes them: a row either follows the guide closely or is random
            w[0] = (rnd() & 0xFFFFu) | ((rnd() & 7u) << 16);
            for (int j = 1; j < 8; j++) {
                const uint32_t noise = (it & 3) == 0 ? rnd() : (rnd() & rnd() & rnd());
                w[j] = gm[7 + j - 1] ^ noise;
            }
            const uint32_t budget = rnd() % 5u;
            uint32_t u1[kMaxDist], u3[kMaxDist], ex;
            summary_eval_masks<kMaxDist>(w, gm, budget, u1);
            if (sh == 0) { summary_masks_shape<0x3Fu, 0u, kMaxDist>(w, gm + 7, budget, u3); ex = summary_exact_shape<0x3Fu>(w, gm + 7); }
            else if (sh == 1) { summary_masks_shape<0x7Fu, 0u, kMaxDist>(w, gm + 7, budget, u3); ex = summary_exact_shape<0x7Fu>(w, gm + 7); }
            else { summary_masks_shape<0x1Fu, 0x40u, kMaxDist>(w, gm + 7, budget, u3); ex = summary_exact_shape<0x5Fu>(w, gm + 7); }
            for (int r = 0; r < kMaxDist; r++) if ((u1[r] & 0xFFFFu) != (u3[r] & 0xFFFFu)) { fprintf(stderr, "summary_masks_shape disagrees (shape %d codes %x r %d)\n", sh, codes, r); return 3; }
            if (ex != summary_eval_exact(w, gm)) { fprintf(stderr, "summary_exact_shape disagrees (shape %d codes %x)\n", sh, codes); return 3; }
            // two-mask form of the unsubstituted-pattern step: budgets 0 and 1
            if (budget <= 1u) { uint32_t u2[2]; summary_eval_masks<2>(w, gm, budget, u2); if ((u2[0] & 0xFFFFu) != (u1[0] & 0xFFFFu)) { fprintf(stderr, "summary_eval_masks<2> disagrees\n"); return 3; } }
            checked++;
        }
    }
    // layouts the lean loops are not compiled for are recognised as such
    if (sweep_shape_of(0x7777777u) >= 0 || sweep_shape_of(0x0D00000u | 0x0012301u) >= 0 || sweep_shape_of(0xC333333u) != 0) { fprintf(stderr, "sweep_shape_of classifies wrongly\n"); return 3; }
    printf("shape selftest ok: %llu sectors\n", checked);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && std::string(argv[1]) == "--shape-selftest") return shape_selftest();
    if (argc < 4) { fprintf(stderr, "usage: host_core_check <prefix> <guides.csv> <out> [options]\n"); return 2; }
    std::string prefix = argv[1], guides_csv = argv[2], out_path = argv[3];
    gsx_params p; gsx_params_default(&p);
    bool sam = false, complete = true; std::vector<const char*> alts;
    for (int i = 4; i < argc; i++) {
        std::string a = argv[i];
        if (a == "-m") p.mismatches = atoi(argv[++i]); else if (a == "--rna") p.rna_bulges = atoi(argv[++i]);
        else if (a == "--dna") p.dna_bulges = atoi(argv[++i]); else if (a == "-t") p.threshold = atoi(argv[++i]);
        else if (a == "--start") p.start = 1; else if (a == "--max") p.max_off_targets = atoll(argv[++i]);
        else if (a == "--lookahead") g_prune = true;
        else if (a == "--ftab") g_ftab_L = atoi(argv[++i]);
        else if (a == "--sweep") g_sweep_sb = atoi(argv[++i]);
        else if (a == "--variants") g_variants = true;
        else if (a == "--parts") g_parts = atoi(argv[++i]);
        else if (a == "--fused") g_fused = true;
        else if (a == "--forced") g_forced = true;
        else if (a == "--sam") sam = true; else if (a == "--succinct") complete = false; else if (a == "-a") alts.push_back(argv[++i]);
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    p.alt_pams = alts.data(); p.n_alt_pams = (uint32_t)alts.size(); p.sam_scoring = sam;

    gsx_index ix; std::string err;
    if (!load_genome_structure(prefix + ".gs", ix.host, err) || !load_sdsl_strand(prefix + ".forward", ix.host.st[0], err) ||
        !load_sdsl_strand(prefix + ".reverse", ix.host.st[1], err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    uint64_t start = 0;
    for (auto l : ix.host.chr_lens) { ix.chroms.push_back({start, l}); start += l; }
    DevStrand st[2] = {view_of(ix.host.st[0]), view_of(ix.host.st[1])};

    // the mapped exception count of search_fast_kernel<..., EXC> against the table search of block_occ, for every row
    for (int sidx = 0; sidx < 2; sidx++) {
        const std::vector<uint32_t> map = build_exc_map(ix.host.st[sidx]);
        const DevStrand& d = st[sidx];
        for (uint64_t i = 0; i <= d.n; i++) {
            const uint32_t got = exc_before(map.data(), d.exc_rows, d.n_exc, (uint32_t)i), want = exc_in(d, (uint32_t)i - ((uint32_t)i & 63u), (uint32_t)i);
            if (got != want) { fprintf(stderr, "exc_before disagrees with exc_in at row %llu of strand %d (%u vs %u)\n", (unsigned long long)i, sidx, got, want); return 3; }
        }
    }

    // guides file: id,sequence,pam,chromosome,position,sense (fixed column order is enough for the harness)
    std::vector<std::string> ids, seqs, pams; std::vector<int> pos;
    { std::ifstream f(guides_csv); std::string line; std::getline(f, line);
      while (std::getline(f, line)) { if (line.empty()) continue; std::vector<std::string> fl; size_t b = 0;
        for (;;) { size_t e = line.find(',', b); fl.push_back(line.substr(b, e == std::string::npos ? e : e - b)); if (e == std::string::npos) break; b = e + 1; }
        ids.push_back(fl[0]); seqs.push_back(fl[1]); pams.push_back(fl[2]); pos.push_back(fl[5] == "+"); } }
    const size_t n = ids.size();
    std::vector<gsx_guide> gg(n); std::vector<gsx_guide_row> rows(n);
    for (size_t i = 0; i < n; i++) { gg[i] = {seqs[i].c_str(), pams[i].c_str()}; rows[i] = {ids[i].c_str(), seqs[i].c_str(), pams[i].c_str(), pos[i]}; }
    Prepared prep;
    if (gsx_prepare_guides(gg.data(), n, &p, prep)) { fprintf(stderr, "%s\n", gsx_last_error()); return 1; }
    const uint32_t n_dist = p.mismatches + 1;
    // bulges as edited guides: one mismatch-only batch holding every variant of every guide
    Prepared vprep; std::vector<uint32_t> vdesc, voff(n + 1, 0);
    if (g_variants) {
        if (!prep.variant_ok) { fprintf(stderr, "--variants: not a variant batch (wide %d, max_qlen %u, min_qlen %u, PAM passes %u)\n", (int)prep.wide, prep.max_qlen, prep.min_qlen, prep.n_fast_pams); return 4; }
        memcpy(vprep.pamsets, prep.pamsets, sizeof prep.pamsets); vprep.max_pams = prep.max_pams; vprep.wide = false; vprep.fast_ok = true;
        vprep.pampack = prep.pampack; vprep.plen = prep.plen; vprep.n_fast_pams = prep.n_fast_pams;
        for (int k = 0; k < kMaxPams; k++) { vprep.pampacks[k] = prep.pampacks[k]; vprep.plens[k] = prep.plens[k]; }
        for (size_t g = 0; g < n; g++) {
            const GuideRec& r = prep.recs[g];
            if (bulge_variant_count(r.qlen, p.rna_bulges, p.dna_bulges) != bulge_variants(r.qlen, p.rna_bulges, p.dna_bulges).size()) { fprintf(stderr, "variant count formula is off\n"); return 3; }
            for (uint32_t desc : bulge_variants(r.qlen, p.rna_bulges, p.dna_bulges)) {
                const uint64_t v = variant_pack(r.q, r.qlen, desc);
                GuideRec e = r; e.qlen = (uint8_t)(v >> 58);
                for (uint32_t l = 0; l < e.qlen; l++) e.q[l] = (uint8_t)((v >> (2 * l)) & 3u);
                vprep.recs.push_back(e); vprep.gq.push_back(v); vdesc.push_back(desc);
                g_vfmask.push_back(variant_forced_mask(r.qlen, desc));
            }
            voff[g + 1] = (uint32_t)vdesc.size();
        }
        vprep.min_qlen = 255; for (const GuideRec& e : vprep.recs) vprep.min_qlen = std::min<uint32_t>(vprep.min_qlen, e.qlen);
    }
    // alternative PAMs in one pass: the batch is searched with the single filter PAM, finished alignments are kept if their PAM
    // characters spell one of the real PAMs
    Prepared mprep; uint32_t fused_n = 0, fused_packs[kMaxPams] = {0}, fused_plen = 0;
    {
        const Prepared& src = g_variants ? vprep : prep;
        if (g_fused && src.n_fast_pams > 1 && (src.fast_ok || g_variants)) {
            mprep = src;
            fused_n = src.n_fast_pams; fused_plen = src.plens[0];
            for (uint32_t k = 0; k < fused_n; k++) fused_packs[k] = src.pampacks[k];
            const uint32_t filt = fused_filter_pampack(fused_packs, fused_n, fused_plen);
            PamSet& ps = mprep.pamsets[0];
            ps.n_pams = 1; ps.plen[0] = (uint8_t)fused_plen;
            for (uint32_t j = 0; j < fused_plen; j++) ps.sym[0][j] = (uint8_t)((filt >> (3 * j)) & 7u);
            mprep.max_pams = 1; mprep.n_fast_pams = 1; mprep.pampack = filt; mprep.pampacks[0] = filt; mprep.plen = fused_plen; mprep.fast_ok = true;
        } else if (g_fused) { fprintf(stderr, "--fused: needs a fast-path batch with alternative PAMs\n"); return 4; }
    }
    const Prepared& fprep = fused_n ? mprep : (g_variants ? vprep : prep);          // the batch the specialised-kernel mirrors (--lookahead / --ftab / --sweep) work on
    if (g_ftab_L) {
        g_pow5[0] = 1; for (int i = 1; i < 32; i++) g_pow5[i] = g_pow5[i - 1] * 5ull;
        build_ftab_host(st[0], g_ftab_L, g_ftab[0]); build_ftab_host(st[1], g_ftab_L, g_ftab[1]);
    }
    if ((g_prune || g_ftab_L || g_sweep_sb) && fprep.n_fast_pams > 1) { fprintf(stderr, "--lookahead / --ftab / --sweep mirror the single-PAM passes: no -a here\n"); return 2; }
    if (g_sweep_sb && (!g_prune || !g_ftab_L || !fprep.fast_ok || g_sweep_sb + 3 > g_ftab_L)) { fprintf(stderr, "--sweep SB needs --lookahead, --ftab L >= SB + 3 and a fast-path batch\n"); return 2; }
    if (g_prune) {
        if (!fprep.fast_ok) { fprintf(stderr, "--lookahead needs a fast-path batch (one PAM, ACGT guides, no bulges)\n"); return 2; }
        build_look(st[0], g_look[0], g_tail[0]); build_look(st[1], g_look[1], g_tail[1]);
    }
    if (g_sweep_sb) for (int s = 0; s < 2; s++) build_summaries_host(st[s], g_look[s], g_tail[s], g_ftab[s], g_sum0[s], g_sum1[s], g_sum2[s]);

    // ---- search + order + expand (what search_kernel / order_matches_kernel / expand_hits_kernel do) ----------------
    std::vector<uint8_t> dropped(n, 0);
    std::vector<MatchRec> sorted; std::vector<uint32_t> hit_match, hit_row, hoff(n + 1, 0), nhits(n, 0), cbd(n * n_dist, 0);
    uint64_t nodes = 0;
    for (size_t g = 0; g < n; g++) {
        hoff[g] = (uint32_t)hit_row.size();
        if (p.threshold > 0) {
            uint64_t count = 0; std::vector<MatchRec> dummy;
            dfs<false>(st, prep, (uint32_t)(2 * g), p.threshold, 0, 0, true, &count, dummy, &nodes);
            dfs<false>(st, prep, (uint32_t)(2 * g + 1), p.threshold, 0, 0, true, &count, dummy, &nodes);
            if (count > 1) { dropped[g] = 1; continue; }
        }
        std::vector<MatchRec> ms;
        if (g_variants) {
            std::vector<MatchRec> tmp;
            for (uint32_t v = voff[g]; v < voff[g + 1]; v++)
                for (uint32_t s = 0; s < 2; s++) {
                    tmp.clear();
                    dfs<false>(st, fprep, 2 * v + s, p.mismatches, 0, 0, false, nullptr, tmp, &nodes);
                    for (const MatchRec& m : tmp) {
                        if (fused_n && !fused_pam_ok(m.key_lo, fused_plen, fused_packs, fused_n)) continue;
                        MatchRec o; if (variant_rewrite(m, prep.recs[g], vdesc[v], (uint32_t)(2 * g + s), o)) ms.push_back(o);
                    }
                }
        } else
        for (uint32_t s = 0; s < 2; s++) {
            if (prep.wide) dfs<true>(st, prep, (uint32_t)(2 * g + s), p.mismatches, p.rna_bulges, p.dna_bulges, false, nullptr, ms, &nodes);
            else if (fused_n) {
                std::vector<MatchRec> tmp;
                dfs<false>(st, fprep, (uint32_t)(2 * g + s), p.mismatches, 0, 0, false, nullptr, tmp, &nodes);
                for (const MatchRec& m : tmp) if (fused_pam_ok(m.key_lo, fused_plen, fused_packs, fused_n)) ms.push_back(m);
            }
            else dfs<false>(st, prep, (uint32_t)(2 * g + s), p.mismatches, p.rna_bulges, p.dna_bulges, false, nullptr, ms, &nodes);
        }
        std::stable_sort(ms.begin(), ms.end(), [](const MatchRec& a, const MatchRec& b) { return match_cmp(a, b) < 0; });
        for (size_t i = 0; i < ms.size(); i++) {
            if (i && match_cmp(ms[i - 1], ms[i]) == 0) continue;
            uint32_t mi = (uint32_t)sorted.size(); sorted.push_back(ms[i]);
            for (uint32_t r = 0; r < ms[i].width; r++) { hit_match.push_back(mi); hit_row.push_back(ms[i].sp + r); }
            cbd[g * n_dist + (ms[i].info & 0xff)] += ms[i].width; nhits[g] += ms[i].width;
        }
    }
    hoff[n] = (uint32_t)hit_row.size();
    const uint32_t nh = hoff[n];

    // ---- locate + score + specificity through the same HD functions the kernels call -------------------------------
    gsx_result res; HostArrays H; H.n_guides = n; H.n_hits = nh;
    std::vector<int64_t> abs_pos(nh + 1); std::vector<int32_t> chr(nh + 1); std::vector<uint32_t> pos1(nh + 1);
    std::vector<uint8_t> strand(nh + 1), distance(nh + 1), dna(nh + 1), rna(nh + 1), index_id(nh + 1), flags(nh + 1), counted(nh + 1, 0), perfect(n + 1);
    std::vector<float> cfd(nh + 1), spec(n + 1);
    LocateArgs L{}; L.st[0] = st[0]; L.st[1] = st[1]; L.matches = sorted.data(); L.guides = prep.recs.data(); L.pamsets = prep.pamsets;
    L.chroms = ix.chroms.data(); L.hit_match = hit_match.data(); L.hit_row = hit_row.data(); L.n_hits = nh; L.n_chr = (uint32_t)ix.chroms.size();
    L.wide = prep.wide; L.genome_length = ix.host.genome_length; L.abs_pos = abs_pos.data(); L.chr = chr.data(); L.pos1 = pos1.data();
    L.strand = strand.data(); L.distance = distance.data(); L.dna = dna.data(); L.rna = rna.data(); L.index_id = index_id.data();
    L.cfd = cfd.data(); L.flags = flags.data();
    std::vector<uint64_t> key_lo(nh + 1), key_hi(nh + 1); std::vector<uint8_t> mlen(nh + 1);
    L.key_lo = key_lo.data(); L.key_hi = prep.wide ? key_hi.data() : nullptr; L.mlen = mlen.data();
    uint32_t steps = 0;
    for (uint32_t h = 0; h < nh; h++) {
        const MatchRec& m = sorted[hit_match[h]];
        uint32_t sa = locate_row_t(st[m.task & 1], hit_row[h], &steps, HostBlockLoader());
        score_hit(L, h, m, sa, &GSX_CFD_MM[0][0][0], &GSX_CFD_PAM[0][0]);
    }
    SpecArgs S{}; S.guide_hoff = hoff.data(); S.count_by_distance = cbd.data(); S.chr = chr.data(); S.cfd = cfd.data(); S.flags = flags.data();
    S.counted = counted.data(); S.specificity = spec.data(); S.perfect = perfect.data(); S.n_guides = (uint32_t)n; S.n_dist = n_dist;
    S.sam_rule = sam; S.max_off_targets = p.max_off_targets;
    for (uint32_t g = 0; g < n; g++) guide_specificity(S, g);

    H.dropped = dropped.data(); H.n_hits_of = nhits.data(); H.hoff = hoff.data(); H.specificity = spec.data(); H.perfect = perfect.data(); H.cbd = cbd.data();
    H.abs_pos = abs_pos.data(); H.sa_row = hit_row.data(); H.chr = chr.data(); H.pos1 = pos1.data(); H.strand = strand.data(); H.distance = distance.data();
    H.rna = rna.data(); H.dna = dna.data(); H.index_id = index_id.data(); H.cfd = cfd.data(); H.counted = counted.data();
    H.key_lo = key_lo.data(); H.key_hi = prep.wide ? key_hi.data() : nullptr; H.mlen = mlen.data();
    res.n_dist = n_dist; res.wide = prep.wide; res.guides = prep.recs;
    std::vector<std::vector<uint32_t>> part_hoff;
    if (g_parts <= 1) { res.parts.push_back(H); res.part_g0.push_back(0); res.part_h0.push_back(0); }
    else {
        // the same arrays cut into contiguous guide shards, one HostArrays each with its own guide and hit numbering: what
        // gsx_enumerate assembles from several devices (some shards empty when there are fewer guides than parts)
        part_hoff.resize(g_parts);
        for (int k = 0; k < g_parts; k++) {
            const size_t g0 = n * k / g_parts, g1 = n * (k + 1) / g_parts; const uint32_t h0 = hoff[g0], h1 = hoff[g1];
            HostArrays P = H;
            P.n_guides = g1 - g0; P.n_hits = h1 - h0;
            part_hoff[k].resize(g1 - g0 + 1);
            for (size_t i = 0; i <= g1 - g0; i++) part_hoff[k][i] = hoff[g0 + i] - h0;
            P.hoff = part_hoff[k].data(); P.dropped += g0; P.n_hits_of += g0; P.specificity += g0; P.perfect += g0; P.cbd += g0 * n_dist;
            P.abs_pos += h0; P.sa_row += h0; P.chr += h0; P.pos1 += h0; P.strand += h0; P.distance += h0; P.rna += h0; P.dna += h0; P.index_id += h0;
            P.cfd += h0; P.counted += h0; P.key_lo += h0; if (P.key_hi) P.key_hi += h0; P.mlen += h0;
            res.parts.push_back(P); res.part_g0.push_back(g0); res.part_h0.push_back(h0);
        }
    }
    gsx_build_view(&res);
    if (g_parts > 1) {
        // the merged view (built on request) must be the unsplit arrays again
        gsx_result_view v;
        if (gsx_result_view_get(&res, &v)) { fprintf(stderr, "view failed\n"); return 1; }
        bool same = v.n_guides == n && v.n_hits == nh && !memcmp(v.specificity, spec.data(), n * 4) && !memcmp(v.n_hits_of, nhits.data(), n * 4) &&
                    !memcmp(v.count_by_distance, cbd.data(), n * n_dist * 4) && !memcmp(v.dropped, dropped.data(), n);
        same = same && (nh == 0 || (!memcmp(v.abs_pos, abs_pos.data(), nh * 8) && !memcmp(v.chr, chr.data(), nh * 4) && !memcmp(v.pos1, pos1.data(), nh * 4) &&
                                    !memcmp(v.cfd, cfd.data(), nh * 4) && !memcmp(v.counted, counted.data(), nh) && !memcmp(v.distance, distance.data(), nh) &&
                                    !memcmp(v.sa_row, hit_row.data(), nh * 4) && !memcmp(v.strand, strand.data(), nh)));
        for (size_t g = 0; same && g < n; g++) same = v.first_hit[g] == hoff[g];
        char a1[64], a2[64];
        for (uint32_t h = 0; same && h < nh; h += 7) {                      // match strings through the part lookup of the ABI call
            gsx_result_match_sequence(&res, h, a1, sizeof a1);
            MatchRec m{}; m.key_lo = key_lo[h]; m.key_hi = prep.wide ? key_hi[h] : 0; m.info = (uint32_t)mlen[h] << 24;
            size_t gi = (size_t)(std::upper_bound(hoff.begin(), hoff.end() - 1, h) - hoff.begin()) - 1;
            uint32_t len = decode_match(m, prep.recs[gi], prep.wide, a2); for (uint32_t i = 0; i < len; i++) a2[i] = complement_char(a2[i]); a2[len] = 0;
            same = !strcmp(a1, a2);
        }
        if (!same) { fprintf(stderr, "merged view of %d parts differs from the unsplit arrays\n", g_parts); return 3; }
    }

    FILE* out = fopen(out_path.c_str(), "wb");
    char* buf; size_t len;
    gsx_format_header(&ix, sam, complete, &buf, &len); fwrite(buf, 1, len, out); gsx_free(buf);
    if (gsx_format_rows(&ix, &res, rows.data(), 0, n, &p, sam, complete, &buf, &len)) { fprintf(stderr, "format failed\n"); return 1; }
    fwrite(buf, 1, len, out); gsx_free(buf); fclose(out);
    if (const char* reps = getenv("GSX_TIME_FORMAT")) {          // formatter throughput on one host thread (rows are those of this run, repeated)
        const int R = atoi(reps); size_t bytes = 0, lines = 0;
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < R; i++) { gsx_format_rows(&ix, &res, rows.data(), 0, n, &p, sam, complete, &buf, &len); bytes += len; if (i == 0) for (size_t k = 0; k < len; k++) lines += buf[k] == '\n'; gsx_free(buf); }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "format: %.1f MB/s, %.2f M lines/s, %.2f M guides/s on one thread (%zu lines per pass)\n", bytes / sec / 1e6, lines * (double)R / sec / 1e6, n * (double)R / sec / 1e6, lines);
    }
    fprintf(stderr, "host_core_check: %zu guides, %llu nodes, %u hits, %u LF steps\n", n, (unsigned long long)nodes, nh, steps);
    return 0;
}
