// tests/format_bench.cpp -- TEST TOOL: the host formatter alone on a fabricated result (G guides x H hits each), one pass timed through the
// library's test hook gsx_internal_format_rate.  usage: format_bench <guides> <hits per guide> <1 = SAM, 0 = CSV>   (GSX_FORMAT_THREADS=n; SUCC=1: succinct)
// build: g++ -O2 -std=c++17 -I/usr/local/cuda/include -o tests/_build/format_bench tests/format_bench.cpp -Lguidescan-cli_b200 -lgsx -Wl,-rpath,$PWD/guidescan-cli_b200 -lpthread
#include "../guidescan-cli_b200/csrc/gsx_host.h"
#include <chrono>
#include <cstdio>
#include <cstring>
#include <random>
#include <thread>
using namespace gsx;
void gsx_build_view(gsx_result* r);
extern "C" int gsx_internal_format_rate(const gsx_index*, const gsx_result*, const gsx_guide_row*, size_t, const gsx_params*, int, int, size_t, double*, size_t*);
struct gsx_result_fwd;
int main(int argc, char** argv) {
    const size_t G = argc > 1 ? atol(argv[1]) : 20000, HPG = argc > 2 ? atol(argv[2]) : 300; const int sam = argc > 3 ? atoi(argv[3]) : 1;
    const uint32_t n_dist = 5;
    gsx_index ix; ix.host.genome_length = 3100000000ull; for (int c = 0; c < 24; c++) { ix.host.chr_names.push_back("chr" + std::to_string(c + 1)); ix.host.chr_lens.push_back(129166666); }
    const size_t nh = G * HPG;
    std::vector<uint8_t> dropped(G, 0), perfect(G, 1); std::vector<uint32_t> nhits(G, HPG), hoff(G + 1), cbd(G * n_dist); std::vector<float> spec(G, 0.123456f);
    std::vector<int64_t> abs_pos(nh); std::vector<uint32_t> sa_row(nh), pos1(nh); std::vector<int32_t> chr(nh); std::vector<uint8_t> strand(nh), distance(nh), rna(nh, 0), dna(nh, 0), idx(nh, 0), counted(nh, 1), mlen(nh, 23);
    std::vector<float> cfd(nh, 0.5f); std::vector<uint64_t> key_lo(nh, 12345);
    std::mt19937_64 rng(1);
    for (size_t g = 0; g <= G; g++) hoff[g] = g * HPG;
    for (size_t g = 0; g < G; g++) {
        uint32_t c[5] = {1, 1, (uint32_t)HPG / 30, (uint32_t)HPG / 6, 0}; c[4] = HPG - c[0] - c[1] - c[2] - c[3];
        size_t h = g * HPG;
        for (uint32_t d = 0; d < 5; d++) { cbd[g * 5 + d] = c[d]; for (uint32_t j = 0; j < c[d]; j++, h++) { distance[h] = d; abs_pos[h] = (int64_t)(rng() % 3100000000ull); chr[h] = (int32_t)(rng() % 24); pos1[h] = rng() % 100000000; strand[h] = rng() & 1; } }
    }
    gsx_result res; HostArrays H; H.n_guides = G; H.n_hits = nh;
    H.dropped = dropped.data(); H.n_hits_of = nhits.data(); H.hoff = hoff.data(); H.specificity = spec.data(); H.perfect = perfect.data(); H.cbd = cbd.data();
    H.abs_pos = abs_pos.data(); H.sa_row = sa_row.data(); H.chr = chr.data(); H.pos1 = pos1.data(); H.strand = strand.data(); H.distance = distance.data();
    H.rna = rna.data(); H.dna = dna.data(); H.index_id = idx.data(); H.cfd = cfd.data(); H.counted = counted.data(); H.key_lo = key_lo.data(); H.key_hi = nullptr; H.mlen = mlen.data();
    res.n_dist = n_dist; res.wide = false; res.guides.resize(G);
    for (auto& r : res.guides) { memset(&r, 0, sizeof r); r.qlen = r.seqlen = 20; memcpy(r.seq, "ACGTACGTACGTACGTACGT", 20); }
    res.parts.push_back(H); res.part_g0.push_back(0); res.part_h0.push_back(0);
    gsx_build_view(&res);
    std::vector<gsx_guide_row> rows(G); for (auto& r : rows) { r.id = "chr1:12345678:+"; r.seq = "ACGTACGTACGTACGTACGT"; r.pam = "NGG"; r.sense_positive = 1; }
    gsx_params p; gsx_params_default(&p); p.mismatches = 4;
    {
        double sec; size_t len;
        if (gsx_internal_format_rate(&ix, &res, rows.data(), G, &p, sam, getenv("SUCC") ? 0 : 1, getenv("ROUNDS") ? atoi(getenv("ROUNDS")) : 4, &sec, &len)) { fprintf(stderr, "format failed\n"); return 1; }
        printf("%s: %zu guides x %zu hits: %.3f s, %.1f MB, %.0f MB/s, %.2f M guides/s, %.1f ns/hit\n", sam ? "SAM" : "CSV", G, HPG, sec, len / 1e6, len / sec / 1e6, G / sec / 1e6, sec / nh * 1e9);
    }
    res.parts.clear();
    return 0;
}
