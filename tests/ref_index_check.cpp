// tests/ref_index_check.cpp -- TEST TOOL (not part of the product).
//
// Pushes an index written by the UNMODIFIED reference (`guidescan index`: <prefix>.forward / .reverse, sdsl
// csa_wt<wt_huff<>,64,8192>, reference src/guidescan.cxx:167-175) through the host half of gsx_index_open -- the product's own
// parser of those files (guidescan-cli_b200/csrc/gsx_index.cpp load_sdsl_strand) -- and checks what comes out against the
// genome text itself, with no suffix array of our own in between:
//   * inverts the BWT held in the 32-byte blocks by walking LF from the row of the empty suffix: every character must be the
//     text's, back to front, and the walk must end on the sentinel row after exactly n - 1 steps (one cycle = a valid BWT of
//     exactly this text);
//   * on the way, every row that is a multiple of 64 carries an SA sample: it must equal the text position the walk is at.
// That pins the 2-bit planes, the four checkpoint counters of every block (through LF), C[], the exception tables (sentinel,
// genome N) and every SA sample.  Used on the golden genomes by tests/test_ref_index.py and on the 3.1 Gb reference index of
// tools/ref_3100mb.py (about ten minutes per strand, one host thread each).
//
// usage: ref_index_check <index prefix> <genome text file: upper-case concatenated chromosomes, no separators> [--dump DIR]
//   --dump DIR: also write what the parser extracted -- DIR/bwt.<strand> (one byte per row, 0 in the sentinel row) and
//   DIR/sa64.<strand> (u32 SA samples every 64 rows) -- the form the CPU oracle imports (oracle.py Index.from_bwt), so that the
//   port can be timed on the reference's own index
#include "../guidescan-cli_b200/csrc/gsx_host.h"
#include "../guidescan-cli_b200/csrc/gsx_core.h"
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace gsx;

static DevStrand view_of(const HostStrand& h) {
    DevStrand d{};
    d.blocks = h.blocks.data(); d.sa_samples = h.sa_samples.data(); d.exc_rows = h.exc_rows.data(); d.exc_lf = h.exc_lf.data();
    d.n_rows = h.n_rows.data(); d.n = (uint32_t)h.n; d.n_exc = (uint32_t)h.exc_rows.size(); d.n_nrows = (uint32_t)h.n_rows.size();
    d.sa_shift = h.sa_shift; for (int c = 0; c < 5; c++) d.C[c] = h.C[c];
    d.exc_lo = h.exc_rows.empty() ? 0xFFFFFFFFu : h.exc_rows.front(); d.exc_hi = h.exc_rows.empty() ? 0 : h.exc_rows.back();
    d.blk_shift = 5;
    return d;
}

struct Report { bool ok = false; std::string msg; uint64_t steps = 0, samples = 0, exc = 0; double load_s = 0, walk_s = 0; };

// text_at(p): character p of the text this strand indexes
static std::string g_dump_dir;
template <class TextAt>
static void check_strand(const std::string& path, uint64_t G, TextAt text_at, Report& rep, const char* strand_name) {
    HostStrand h; std::string err;
    auto t0 = std::chrono::steady_clock::now();
    if (!load_sdsl_strand(path, h, err)) { rep.msg = err; return; }
    rep.load_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (h.n != G + 1) { rep.msg = "row count " + std::to_string(h.n) + " != genome length + 1"; return; }
    const DevStrand st = view_of(h);
    static const char SYM[4] = {'A', 'C', 'G', 'T'};
    if (!g_dump_dir.empty()) {
        std::vector<uint8_t> bwt(h.n);
        for (uint64_t row = 0; row < h.n; row++) { const OccBlock& b = h.blocks[row >> 6]; bwt[row] = (uint8_t)SYM[block_sym(b.hi, b.lo, (uint32_t)row)]; }
        for (size_t i = 0; i < h.exc_rows.size(); i++) bwt[h.exc_rows[i]] = h.exc_sym[i];
        FILE* f = fopen((g_dump_dir + "/bwt." + strand_name).c_str(), "wb");
        if (!f || fwrite(bwt.data(), 1, bwt.size(), f) != bwt.size()) { rep.msg = "cannot write the BWT dump"; if (f) fclose(f); return; }
        fclose(f);
        f = fopen((g_dump_dir + "/sa64." + strand_name).c_str(), "wb");
        if (!f || fwrite(h.sa_samples.data(), 4, h.sa_samples.size(), f) != h.sa_samples.size()) { rep.msg = "cannot write the SA sample dump"; if (f) fclose(f); return; }
        fclose(f);
    }
    t0 = std::chrono::steady_clock::now();
    uint64_t p = G; uint32_t r = 0;                              // row 0 = the empty suffix, SA[0] = G
    for (;;) {
        if ((r & ((1u << h.sa_shift) - 1u)) == 0) {
            if (h.sa_samples[r >> h.sa_shift] != (uint32_t)p) { rep.msg = "SA sample of row " + std::to_string(r) + " is " + std::to_string(h.sa_samples[r >> h.sa_shift]) + ", text position " + std::to_string(p); return; }
            rep.samples++;
        }
        // BWT[r] and LF(r)
        bool exc = false; uint8_t c = 0; uint32_t nxt = 0;
        if (st.n_exc && r >= st.exc_lo && r <= st.exc_hi) {
            const uint32_t k = lower_bound_u32(st.exc_rows, st.n_exc, r);
            if (k < st.n_exc && st.exc_rows[k] == r) { exc = true; c = h.exc_sym[k]; nxt = st.exc_lf[k]; rep.exc++; }
        }
        if (!exc) {
            const OccBlock& b = st.blocks[r >> 6]; uint32_t o[4];
            block_occ(st, b.cnt, b.hi, b.lo, r, o);
            const uint32_t sy = block_sym(b.hi, b.lo, r);
            c = (uint8_t)SYM[sy]; nxt = st.C[sy] + o[sy];
        }
        if (p == 0) {
            if (!(exc && c == 0)) { rep.msg = "the walk reached text position 0 on a row whose BWT symbol is not the sentinel"; return; }
            break;
        }
        if (exc && c == 0) { rep.msg = "sentinel met at text position " + std::to_string(p); return; }
        if (c != text_at(p - 1)) { rep.msg = "BWT character differs from the text at position " + std::to_string(p - 1); return; }
        r = nxt; p--; rep.steps++;
    }
    if (rep.steps != G) { rep.msg = "walk length differs from the genome length"; return; }
    rep.walk_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    rep.ok = true;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: ref_index_check <index prefix> <genome text file> [--dump DIR]\n"); return 2; }
    const std::string prefix = argv[1];
    if (argc >= 5 && std::string(argv[3]) == "--dump") g_dump_dir = argv[4];
    std::vector<uint8_t> text;
    {
        FILE* f = fopen(argv[2], "rb");
        if (!f) { fprintf(stderr, "cannot open %s\n", argv[2]); return 2; }
        fseeko(f, 0, SEEK_END); const off_t sz = ftello(f); fseeko(f, 0, SEEK_SET);
        text.resize((size_t)sz);
        if (sz && fread(text.data(), 1, (size_t)sz, f) != (size_t)sz) { fprintf(stderr, "short read of %s\n", argv[2]); return 2; }
        fclose(f);
        while (!text.empty() && (text.back() == '\n' || text.back() == '\r')) text.pop_back();
    }
    const uint64_t G = text.size();
    Report rep[2];
    std::thread t0([&] { check_strand(prefix + ".forward", G, [&](uint64_t p) { return text[p]; }, rep[0], "forward"); });
    // the reverse index is over the reverse complement of the whole concatenated genome (reference src/genomics/seq_io.cxx:65-72)
    std::thread t1([&] { check_strand(prefix + ".reverse", G, [&](uint64_t p) { return (uint8_t)complement_char((char)text[G - 1 - p]); }, rep[1], "reverse"); });
    t0.join(); t1.join();
    int rc = 0;
    for (int s = 0; s < 2; s++) {
        printf("{\"strand\": \"%s\", \"ok\": %s, \"genome_length\": %llu, \"lf_steps_checked\": %llu, \"sa_samples_checked\": %llu, \"exception_rows_met\": %llu, "
               "\"load_seconds\": %.2f, \"walk_seconds\": %.2f, \"message\": \"%s\"}\n", s ? "reverse" : "forward", rep[s].ok ? "true" : "false",
               (unsigned long long)G, (unsigned long long)rep[s].steps, (unsigned long long)rep[s].samples, (unsigned long long)rep[s].exc,
               rep[s].load_s, rep[s].walk_s, rep[s].msg.c_str());
        if (!rep[s].ok) rc = 1;
    }
    return rc;
}
