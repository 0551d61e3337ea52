// tests/sdsl_write_check.cpp -- TEST TOOL (not part of the product): the pieces of guidescan-cli_b200/csrc/gsx_sdsl_write.cpp
// on their own, without a GPU.
//   sdsl_write_check rewrite <strand file of the reference> <out>   parse it (load_sdsl_strand) and write it back (save_sdsl_strand)
//   sdsl_write_check wt <file of BWT bytes, 0 = sentinel> <out>     wavelet tree section only
//   sdsl_write_check bv <file of u64 words> <n_bits> <out>          rank directory + select directories (ones, zeros) of a bit vector
#include "../guidescan-cli_b200/csrc/gsx_host.h"
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace gsx;

static bool slurp(const char* path, std::vector<uint8_t>& out) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseeko(f, 0, SEEK_END); const off_t sz = ftello(f); fseeko(f, 0, SEEK_SET);
    out.resize((size_t)sz);
    const bool ok = sz == 0 || fread(out.data(), 1, (size_t)sz, f) == (size_t)sz;
    fclose(f); return ok;
}

int main(int argc, char** argv) {
    std::string err;
    if (argc >= 4 && !strcmp(argv[1], "rewrite")) {
        HostStrand h;
        const auto t0 = std::chrono::steady_clock::now();
        if (!load_sdsl_strand(argv[2], h, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        const auto t1 = std::chrono::steady_clock::now();
        const unsigned threads = argc >= 5 ? (unsigned)atoi(argv[4]) : 0;
        if (!save_sdsl_strand(argv[3], h, threads, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        printf("{\"rows\": %llu, \"load_seconds\": %.2f, \"write_seconds\": %.2f}\n", (unsigned long long)h.n,
               std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
        return 0;
    }
    if (argc >= 4 && !strcmp(argv[1], "wt")) {
        std::vector<uint8_t> bwt;
        if (!slurp(argv[2], bwt) || bwt.empty()) { fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
        HostStrand h; StrandBuilder b(&h, bwt.size());
        for (uint8_t c : bwt) b.push(c);
        b.finish();
        const unsigned threads = argc >= 5 ? (unsigned)atoi(argv[4]) : 1;
        if (!sdsl_write_wavelet_tree(argv[3], h, threads, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        return 0;
    }
    if (argc >= 5 && !strcmp(argv[1], "bv")) {
        std::vector<uint8_t> raw;
        if (!slurp(argv[2], raw)) { fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
        const uint64_t n_bits = std::stoull(argv[3]);
        std::vector<uint64_t> words((n_bits + 63) / 64, 0ull);
        if (raw.size() < words.size() * 8) { fprintf(stderr, "too few words in %s\n", argv[2]); return 1; }
        memcpy(words.data(), raw.data(), words.size() * 8);
        if (!sdsl_write_bit_vector_supports(argv[4], words, n_bits, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        return 0;
    }
    fprintf(stderr, "usage: sdsl_write_check rewrite <strand file> <out> [threads] | wt <bwt bytes> <out> [threads] | bv <words> <n_bits> <out>\n");
    return 2;
}
