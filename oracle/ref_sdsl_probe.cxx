// oracle/ref_sdsl_probe.cxx -- TEST INFRASTRUCTURE ONLY (our file; compiled against the reference's vendored sdsl headers and
// the sdsl objects of oracle/_ref by Makefile.ref, output oracle/_ref/sdsl_probe).
//
// Lets tests/test_sdsl_writer.py put sdsl's own structures next to the ones guidescan-cli_b200/csrc/gsx_sdsl_write.cpp writes,
// on inputs a whole `guidescan index` run cannot be steered to (bit vectors whose zero count, length and padding hit the corner
// cases of select_support_mcl; BWTs with arbitrary byte alphabets):
//   sdsl_probe bv <file of u64 words> <n_bits> <out>   rank_support_v<1>, select_support_mcl<1>, select_support_mcl<0> of the
//                                                      bit vector, serialized one after the other (the three directories that
//                                                      follow the bits inside wt_pc::serialize, sdsl wt_pc.hpp:656-671)
//   sdsl_probe wt <file of bytes> <out>                wt_huff<> over the byte sequence, serialized
#include <sdsl/wavelet_trees.hpp>
#include <sdsl/rank_support_v.hpp>
#include <sdsl/select_support_mcl.hpp>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

int main(int argc, char** argv) {
    if (argc >= 5 && !strcmp(argv[1], "bv")) {
        const uint64_t n_bits = std::stoull(argv[3]);
        sdsl::bit_vector bv(n_bits, 0);
        std::ifstream in(argv[2], std::ios::binary);
        in.read((char*)bv.data(), (std::streamsize)(((n_bits + 63) / 64) * 8));
        if (!in) { fprintf(stderr, "short read of %s\n", argv[2]); return 2; }
        sdsl::rank_support_v<1> rank(&bv);
        sdsl::select_support_mcl<1, 1> sel1(&bv);
        sdsl::select_support_mcl<0, 1> sel0(&bv);
        std::ofstream out(argv[4], std::ios::binary);
        rank.serialize(out); sel1.serialize(out); sel0.serialize(out);
        return out ? 0 : 2;
    }
    if (argc >= 4 && !strcmp(argv[1], "wt")) {
        sdsl::wt_huff<> wt;
        sdsl::construct(wt, argv[2], 1);
        std::ofstream out(argv[3], std::ios::binary);
        wt.serialize(out);
        return out ? 0 : 2;
    }
    fprintf(stderr, "usage: sdsl_probe bv <words> <n_bits> <out> | wt <bytes> <out>\n");
    return 2;
}
