// gsx_kmers.cu -- genome-wide guide generation on the GPU: what the reference's
// scripts/generate_kmers.py does in Python (reference scripts/generate_kmers.py:55-136), as a PAM scan kernel + ordered
// compaction per chromosome, with the CSV rows formatted on the host.  Output is byte-identical to the script's stdout.
//
// Semantics restated (not copied) from the script:
//   * chromosome = first word of the FASTA header; sequence upper-cased; records shorter than min_chr_length are skipped
//   * the PAM pattern's N's are expanded one at a time, first N first, in the order A,C,T,G, breadth first -- e.g. NGG gives
//     AGG, CGG, TGG, GGG (generate_kmers.py:55-68); the "-" strand uses the reverse complements of the same list
//   * for each concrete PAM, every (overlapping) occurrence i: PAM at the end (default): "+" k-mer = chr[i-k, i), position
//     i-k; "-" k-mer = revcomp(chr[i+|pam|, i+|pam|+k)), position i.  PAM at the start (--start): "+" k-mer =
//     chr[i+|pam|, i+|pam|+k), position i; "-" k-mer = revcomp(chr[i-k, i)), position i-k (generate_kmers.py:70-100)
//   * dropped: negative position, k-mer running off the chromosome, any base outside ACGT (generate_kmers.py:96-117)
//   * row: id = prefix + chr:pos1:sense, sequence, pam PATTERN, chr, pos1 (1-based), sense; per chromosome all "+" rows by
//     concrete PAM then position, then all "-" rows; duplicates are kept (generate_kmers.py:119-136)
#include "../../include/gsx.h"
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <thread>
#include <vector>

int gsx_set_error(int code, const std::string& msg);      // gsx_api.cpp

namespace {

// flags[i] = 1 iff a k-mer is generated for PAM occurrence i
//   kmer_off: offset of the k-mer relative to i (may be negative); the reported position is i + pos_off
__global__ void kmer_scan_kernel(const unsigned char* __restrict__ chr, uint32_t len, uint64_t pam_packed, uint32_t plen,
                                 int32_t kmer_off, uint32_t k, unsigned char* __restrict__ flags) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
        bool ok = i + plen <= len;
        for (uint32_t j = 0; ok && j < plen; j++) ok = chr[i + j] == (unsigned char)(pam_packed >> (8 * j));
        const int64_t a = (int64_t)i + kmer_off;
        ok = ok && a >= 0 && a + k <= len;
        for (uint32_t j = 0; ok && j < k; j++) { const unsigned char c = chr[a + j]; ok = c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
        flags[i] = ok ? 1 : 0;
    }
}
__global__ void upper_kernel(unsigned char* s, uint32_t len) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) { unsigned char c = s[i]; if (c >= 'a' && c <= 'z') s[i] = c - 32; }
}

struct Fail { int code; std::string msg; };
#define KCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw Fail{GSX_ERR_CUDA, std::string(#x ": ") + cudaGetErrorString(e_)}; } while (0)

char comp(char c) { return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : 0; }

std::vector<std::string> expand_pam(const std::string& pam) {
    std::deque<std::string> q{pam};
    auto any_n = [&]() { for (auto& s : q) if (s.find('N') != std::string::npos) return true; return false; };
    while (any_n()) {
        std::string s = q.front(); q.pop_front();
        const size_t at = s.find('N');
        if (at == std::string::npos) { q.push_back(s); continue; }
        for (char c : {'A', 'C', 'T', 'G'}) { std::string t = s; t[at] = c; q.push_back(t); }
    }
    return std::vector<std::string>(q.begin(), q.end());
}

struct DeviceScratch {
    unsigned char* chr = nullptr; unsigned char* flags = nullptr; uint32_t* pos = nullptr; uint32_t* count = nullptr; void* tmp = nullptr;
    size_t cap = 0, tmp_bytes = 0;
    ~DeviceScratch() { cudaFree(chr); cudaFree(flags); cudaFree(pos); cudaFree(count); cudaFree(tmp); }
    void reserve(size_t len) {
        if (len <= cap) return;
        cudaFree(chr); cudaFree(flags); cudaFree(pos); cudaFree(tmp); chr = flags = nullptr; pos = nullptr; tmp = nullptr;
        cap = len + len / 4 + 1024;
        KCK(cudaMalloc(&chr, cap)); KCK(cudaMalloc(&flags, cap)); KCK(cudaMalloc(&pos, cap * 4));
        if (!count) KCK(cudaMalloc(&count, 4));
        tmp_bytes = 0;
        thrust::counting_iterator<uint32_t> it(0);
        KCK(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, it, flags, pos, count, (int)cap));
        KCK(cudaMalloc(&tmp, tmp_bytes));
    }
};

}  // namespace

extern "C" int gsx_generate_kmers(const char* fasta_path, const char* out_csv_path, const char* pam_c, uint32_t k, uint64_t min_chr_length,
                                  const char* prefix_c, int start, int device, uint64_t* n_kmers) {
    if (!fasta_path || !out_csv_path || !pam_c) return gsx_set_error(GSX_ERR_ARG, "null argument");
    const std::string pam(pam_c), prefix(prefix_c ? prefix_c : "");
    if (pam.empty() || pam.size() > 8 || k == 0 || k > 64) return gsx_set_error(GSX_ERR_ARG, "PAM must have 1..8 characters and the k-mer 1..64");
    for (char c : pam) if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N')) return gsx_set_error(GSX_ERR_ARG, "PAM characters must be A, C, G, T or N");
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have <= 0) return gsx_set_error(GSX_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= have) return gsx_set_error(GSX_ERR_ARG, "device ordinal out of range");
    FILE* in = fopen(fasta_path, "rb");
    if (!in) return gsx_set_error(GSX_ERR_IO, std::string("cannot open ") + fasta_path);
    FILE* out = fopen(out_csv_path, "wb");
    if (!out) { fclose(in); return gsx_set_error(GSX_ERR_IO, std::string("cannot write ") + out_csv_path); }
    uint64_t total = 0;
    int rc = GSX_OK;
    try {
        KCK(cudaSetDevice(device));
        const std::vector<std::string> pams = expand_pam(pam);
        std::vector<std::string> rpams;
        for (auto& p : pams) { std::string r(p.rbegin(), p.rend()); for (auto& c : r) c = comp(c); rpams.push_back(r); }
        const uint32_t plen = (uint32_t)pam.size();
        DeviceScratch D;
        std::vector<uint32_t> hpos;
        std::string buf; buf.reserve(1 << 22);
        fputs("id,sequence,pam,chromosome,position,sense\n", out);
        size_t rows_per_thread = 16384;                                                      // (tests lower it to reach the threaded writer on small inputs)
        if (const char* e = getenv("GSX_KMERS_ROWS_PER_THREAD")) if (*e && atoll(e) > 0) rows_per_thread = (size_t)atoll(e);

        auto process = [&](const std::string& name, std::string& seq) {
            if (seq.size() < min_chr_length || seq.empty()) return;
            if (seq.size() >= (1ull << 32) - 64) throw Fail{GSX_ERR_ARG, "chromosome longer than 2^32 bases"};
            const uint32_t len = (uint32_t)seq.size();
            D.reserve(len);
            KCK(cudaMemcpy(D.chr, seq.data(), len, cudaMemcpyHostToDevice));
            const int blocks = (int)std::min<uint64_t>((len + 255) / 256, 148 * 16);
            upper_kernel<<<blocks, 256>>>(D.chr, len);
            KCK(cudaMemcpy(&seq[0], D.chr, len, cudaMemcpyDeviceToHost));                 // the rows print the upper-cased bases
            for (int sense = 0; sense < 2; sense++) {
                const bool fwd = sense == 0;
                // where the k-mer lies relative to the PAM occurrence, and which end is reported (find_kmers in the script)
                const bool before = (fwd == (start == 0));                                    // k-mer in front of the PAM occurrence
                const int32_t kmer_off = before ? -(int32_t)k : (int32_t)plen;
                for (const std::string& p : fwd ? pams : rpams) {
                    uint64_t packed = 0; for (uint32_t j = 0; j < plen; j++) packed |= (uint64_t)(unsigned char)p[j] << (8 * j);
                    kmer_scan_kernel<<<blocks, 256>>>(D.chr, len, packed, plen, kmer_off, k, D.flags);
                    thrust::counting_iterator<uint32_t> it(0);
                    size_t tb = D.tmp_bytes;
                    KCK(cub::DeviceSelect::Flagged(D.tmp, tb, it, D.flags, D.pos, D.count, (int)len));
                    uint32_t n = 0; KCK(cudaMemcpy(&n, D.count, 4, cudaMemcpyDeviceToHost));
                    hpos.resize(n);
                    if (n) KCK(cudaMemcpy(hpos.data(), D.pos, (size_t)n * 4, cudaMemcpyDeviceToHost));
                    // rows of positions [lo, hi) appended to o (find_kmers + the print loop of the script)
                    auto format_rows = [&](size_t lo, size_t hi, std::string& o, bool flush) {
                        for (size_t q = lo; q < hi; q++) {
                            const uint32_t i = hpos[q];
                            const uint32_t a = (uint32_t)((int64_t)i + kmer_off);              // first base of the k-mer on the + strand
                            const uint32_t pos1 = a + 1u - (before ? 0u : plen);                 // = (before ? i - k : i) + 1
                            char num[16]; const int nl = snprintf(num, sizeof num, "%u", pos1);
                            o += prefix; o += name; o += ':'; o.append(num, nl); o += ':'; o += fwd ? '+' : '-'; o += ',';
                            if (fwd) o.append(seq, a, k);
                            else for (uint32_t j = 0; j < k; j++) o += comp(seq[a + k - 1 - j]);
                            o += ','; o += pam; o += ','; o += name; o += ','; o.append(num, nl); o += ','; o += fwd ? '+' : '-'; o += '\n';
                            if (flush && o.size() > (1u << 22) - 256) { fwrite(o.data(), 1, o.size(), out); o.clear(); }
                        }
                    };
                    const size_t nt = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), n / rows_per_thread);
                    if (nt < 2) format_rows(0, n, buf, true);
                    else {
                        // many rows (a whole chromosome of PAM sites): slices on host threads, written in order behind what is buffered
                        fwrite(buf.data(), 1, buf.size(), out); buf.clear();
                        std::vector<std::string> parts(nt); std::vector<std::thread> th;
                        const size_t row_bytes = 2 * (prefix.size() + name.size()) + 2 * 11 + k + pam.size() + 16;
                        for (size_t t = 0; t < nt; t++)
                            th.emplace_back([&, t] { const size_t lo = (size_t)n * t / nt, hi = (size_t)n * (t + 1) / nt; parts[t].reserve((hi - lo) * row_bytes); format_rows(lo, hi, parts[t], false); });
                        for (auto& x : th) x.join();
                        for (const std::string& o : parts) fwrite(o.data(), 1, o.size(), out);
                    }
                    total += n;
                }
            }
        };

        // FASTA records: name = first word of the header, sequence = the lines joined without white space
        std::string name, seq, line; bool have_rec = false;
        char chunk[1 << 16];
        auto handle_line = [&](const std::string& l) {
            if (!l.empty() && l[0] == '>') {
                if (have_rec) process(name, seq);
                size_t b = 1; while (b < l.size() && isspace((unsigned char)l[b])) b++;
                size_t e = b; while (e < l.size() && !isspace((unsigned char)l[e])) e++;
                name = l.substr(b, e - b); seq.clear(); have_rec = true;
            } else if (have_rec) {
                size_t e = l.size(); while (e && isspace((unsigned char)l[e - 1])) e--;
                bool inner = false; for (size_t q = 0; q < e; q++) if (isspace((unsigned char)l[q])) { inner = true; break; }
                if (!inner) seq.append(l, 0, e);                                            // the usual line: one block
                else for (size_t q = 0; q < e; q++) if (!isspace((unsigned char)l[q])) seq += l[q];
            }
        };
        while (fgets(chunk, sizeof chunk, in)) {
            line += chunk;
            if (!line.empty() && line.back() == '\n') { handle_line(line); line.clear(); }
        }
        if (!line.empty()) handle_line(line);
        if (have_rec) process(name, seq);
        fwrite(buf.data(), 1, buf.size(), out);
    } catch (const Fail& f) { rc = gsx_set_error(f.code, f.msg); }
    catch (const std::bad_alloc&) { rc = gsx_set_error(GSX_ERR_NOMEM, "out of host memory"); }
    fclose(in); fclose(out);
    if (n_kmers) *n_kmers = total;
    return rc;
}
