// gsx_format.cpp -- host text output: reproduces the reference's CSV / SAM writers byte for byte
// (include/genomics/printer.hpp:18-360) from the result arrays of gsx_enumerate, plus the whole-file driver that
// mirrors do_enumerate_cmd (src/guidescan.cxx:181-258).  Text formatting stays on the host.
#include "../../include/gsx.h"
#include "gsx_host.h"
#include "gsx_core.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace gsx;

namespace {
thread_local std::string t_err;

inline void put_u64(std::string& s, uint64_t v) { char t[24]; int n = snprintf(t, sizeof t, "%llu", (unsigned long long)v); s.append(t, n); }
inline void put_float(std::string& s, float f) { char t[64]; int n = snprintf(t, sizeof t, "%f", (double)f); s.append(t, n); }   // std::to_string(float)
inline void put_hex_le64(std::string& s, uint64_t v) {                                      // printer.hpp:18-88
    static const char H[] = "0123456789abcdef";
    for (int i = 0; i < 8; i++) { unsigned b = (unsigned)(v & 255); v >>= 8; s.push_back(H[b >> 4]); s.push_back(H[b & 15]); }
}
std::string revcomp(const std::string& x) { std::string r(x.size(), 'N'); for (size_t i = 0; i < x.size(); i++) r[i] = complement_char(x[x.size() - 1 - i]); return r; }

void format_csv_guide(const gsx_index* ix, const gsx_result* r, const gsx_guide_row& row, size_t g, const gsx_params* p, bool complete, std::string& out) {
    const gsx_result_view& v = r->view;
    if (v.dropped[g]) return;                                                                // process.hpp:68-70: nothing is printed
    std::string sequence = p->start ? std::string(row.pam) + row.seq : std::string(row.seq) + row.pam;
    const uint64_t b = v.first_hit[g]; const uint32_t n = v.n_hits_of[g];
    if (n == 0) {                                                                            // printer.hpp:190-199
        out += row.id; out += ","; out += sequence; out += ",NA,NA,NA,0";
        if (complete) out += ",NA,NA,NA";
        out += ",1.0\n";
        return;
    }
    char sp[64]; int spn = snprintf(sp, sizeof sp, "%f", (double)v.specificity[g]);
    char ms[64];
    for (uint64_t h = b; h < b + n; h++) {
        if (!v.counted[h]) continue;
        out += row.id; out += ","; out += sequence; out += ",";
        out += ix->host.chr_names[v.chr[h]]; out += ","; put_u64(out, v.pos1[h]); out += ",";
        out.push_back((char)v.strand[h]); out += ","; put_u64(out, v.distance[h]);
        if (complete) {
            gsx_result_match_sequence(r, h, ms, sizeof ms);
            out += ","; out += ms; out += ","; put_u64(out, v.rna_bulges[h]); out += ","; put_u64(out, v.dna_bulges[h]);
        }
        out += ","; out.append(sp, spn); out += "\n";
    }
}

void format_sam_guide(const gsx_index* ix, const gsx_result* r, const gsx_guide_row& row, size_t g, const gsx_params* p, bool complete, std::string& out) {
    const gsx_result_view& v = r->view;
    if (v.dropped[g]) return;
    const uint64_t b = v.first_hit[g]; const uint32_t n = v.n_hits_of[g];
    if (n == 0) return;
    bool any0 = false;
    for (uint64_t h = b; h < b + n && v.distance[h] == 0; h++) any0 = true;
    if (!any0) return;                                                                       // rows exist only for 0-mismatch alignments
    std::string sequence = p->start ? std::string(row.pam) + row.seq : std::string(row.seq) + row.pam;
    std::string seq_out = row.sense_positive ? sequence : revcomp(sequence);
    std::string hex;
    if (complete) {                                                                          // off_target_fields, printer.hpp:115-170
        int64_t delim = -((int64_t)ix->host.genome_length + 1);
        uint64_t h = b;
        for (uint32_t d = 0; d < v.n_dist; d++) {
            uint32_t cnt = v.count_by_distance[g * v.n_dist + d];
            for (uint32_t j = 0; j < cnt; j++, h++) if (v.counted[h]) put_hex_le64(hex, (uint64_t)v.abs_pos[h]);
            put_hex_le64(hex, d); put_hex_le64(hex, (uint64_t)delim);
        }
    }
    char sp[64]; int spn = snprintf(sp, sizeof sp, "%f", (double)v.specificity[g]);
    for (uint64_t h = b; h < b + n; h++) {
        if (v.distance[h] != 0) continue;
        out += row.id; out += "\t"; out += row.sense_positive ? "0" : "16"; out += "\t";
        if (v.chr[h] >= 0) { out += ix->host.chr_names[v.chr[h]]; out += "\t"; put_u64(out, v.pos1[h]); }
        else out += "\t0";                                                                   // sentinel coordinates: chr "" offset 0
        out += "\t100\t"; put_u64(out, sequence.size()); out += "M\t*\t0\t0\t"; out += seq_out; out += "\t*";
        for (uint32_t d = 0; d < v.n_dist; d++) { out += "\tk"; put_u64(out, d); out += ":i:"; put_u64(out, v.count_by_distance[g * v.n_dist + d]); }
        if (complete) { out += "\tof:H:"; out += hex; }
        out += "\tsp:f:"; out.append(sp, spn); out += "\n";
    }
}

int ret_buf(const std::string& s, char** buf, size_t* len) {
    char* b = (char*)malloc(s.size() + 1);
    if (!b) { t_err = "out of memory"; return GSX_ERR_NOMEM; }
    memcpy(b, s.data(), s.size()); b[s.size()] = 0;
    *buf = b; *len = s.size();
    return GSX_OK;
}

// guides CSV (reference src/genomics/kmer.cxx:9-25 on fast-cpp-csv-parser with trim_chars<' ','\t'>, no quoting):
// the header must name all six columns (any order); position is parsed but unused downstream.
struct GuideTable { std::vector<std::string> id, seq, pam; std::vector<uint8_t> positive; };

std::string trim(const std::string& s) {
    size_t b = 0, e = s.size();
    while (b < e && (s[b] == ' ' || s[b] == '\t')) b++;
    while (e > b && (s[e - 1] == ' ' || s[e - 1] == '\t' || s[e - 1] == '\r')) e--;
    return s.substr(b, e - b);
}

bool read_guides_csv(const char* path, GuideTable& t, std::string& err) {
    FILE* f = fopen(path, "r");
    if (!f) { err = std::string("cannot open kmers file ") + path; return false; }
    char* line = nullptr; size_t cap = 0; ssize_t l;
    static const char* want[6] = {"id", "sequence", "pam", "chromosome", "position", "sense"};
    int col_of[6] = {-1, -1, -1, -1, -1, -1}; int ncol = 0;
    if ((l = getline(&line, &cap, f)) < 0) { fclose(f); free(line); err = "empty kmers file"; return false; }
    {
        std::string h(line, l); while (!h.empty() && (h.back() == '\n' || h.back() == '\r')) h.pop_back();
        size_t b = 0;
        for (;;) {
            size_t e = h.find(',', b); std::string name = trim(h.substr(b, e == std::string::npos ? std::string::npos : e - b));
            bool known = false;
            for (int k = 0; k < 6; k++) if (name == want[k]) { col_of[k] = ncol; known = true; }
            if (!known) { fclose(f); free(line); err = "Extra column \"" + name + "\" in header of kmers file"; return false; }
            ncol++;
            if (e == std::string::npos) break;
            b = e + 1;
        }
        for (int k = 0; k < 6; k++) if (col_of[k] < 0) { fclose(f); free(line); err = std::string("Missing column \"") + want[k] + "\" in header of kmers file"; return false; }
    }
    std::vector<std::string> fields;
    while ((l = getline(&line, &cap, f)) >= 0) {
        std::string s(line, l); while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back();
        if (s.empty()) continue;
        fields.clear(); size_t b = 0;
        for (;;) { size_t e = s.find(',', b); fields.push_back(trim(s.substr(b, e == std::string::npos ? std::string::npos : e - b))); if (e == std::string::npos) break; b = e + 1; }
        if ((int)fields.size() != ncol) { fclose(f); free(line); err = "wrong number of columns in kmers file line: " + s; return false; }
        t.id.push_back(fields[col_of[0]]); t.seq.push_back(fields[col_of[1]]); t.pam.push_back(fields[col_of[2]]);
        t.positive.push_back(fields[col_of[5]] == "+");
    }
    fclose(f); free(line);
    return true;
}
}  // namespace

extern "C" int gsx_format_header(const gsx_index* ix, int format_sam, int complete, char** buf, size_t* len) {
    if (!ix || !buf || !len) return GSX_ERR_ARG;
    std::string s;
    if (format_sam) {                                                                        // printer.hpp:173-179
        s += "@HD\tVN:1.0\tSO:unknown\n@PG\tID:Guidescan\tVN:2.0.0\n";
        for (size_t i = 0; i < ix->host.chr_names.size(); i++) { s += "@SQ\tSN:" + ix->host.chr_names[i] + "\tLN:"; put_u64(s, ix->host.chr_lens[i]); s += "\n"; }
    } else {                                                                                 // printer.hpp:181-187
        s += "id,sequence,match_chrm,match_position,match_strand,match_distance";
        if (complete) s += ",match_sequence,rna_bulges,dna_bulges";
        s += ",specificity\n";
    }
    return ret_buf(s, buf, len);
}

extern "C" int gsx_format_rows(const gsx_index* ix, const gsx_result* r, const gsx_guide_row* rows, size_t g0, size_t g1,
                               const gsx_params* p, int format_sam, int complete, char** buf, size_t* len) {
    if (!ix || !r || !rows || !p || !buf || !len || g1 > r->view.n_guides || g0 > g1) return GSX_ERR_ARG;
    // guides are independent: format slices on host threads, concatenate in order
    size_t n = g1 - g0;
    unsigned nt = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), std::max<size_t>(1, n / 2048));
    std::vector<std::string> parts(nt);
    auto work = [&](unsigned t) {
        size_t a = g0 + n * t / nt, b = g0 + n * (t + 1) / nt;
        std::string& o = parts[t]; o.reserve((b - a) * 256);
        for (size_t g = a; g < b; g++) {
            if (format_sam) format_sam_guide(ix, r, rows[g - g0], g, p, complete != 0, o);
            else format_csv_guide(ix, r, rows[g - g0], g, p, complete != 0, o);
        }
    };
    if (nt == 1) work(0);
    else { std::vector<std::thread> th; for (unsigned t = 0; t < nt; t++) th.emplace_back(work, t); for (auto& x : th) x.join(); }
    size_t tot = 0; for (auto& s : parts) tot += s.size();
    char* b = (char*)malloc(tot + 1);
    if (!b) return GSX_ERR_NOMEM;
    size_t o = 0; for (auto& s : parts) { memcpy(b + o, s.data(), s.size()); o += s.size(); }
    b[tot] = 0; *buf = b; *len = tot;
    return GSX_OK;
}

extern "C" int gsx_enumerate_file(const gsx_index* ix, const char* kmers_csv, const char* out_path, const gsx_params* p,
                                  int format_sam, int complete, size_t batch_guides, size_t* n_guides, gsx_counters* counters) {
    if (!ix || !kmers_csv || !out_path || !p) return GSX_ERR_ARG;
    GuideTable t; std::string err;
    if (!read_guides_csv(kmers_csv, t, err)) { fprintf(stderr, "gsx: %s\n", err.c_str()); return GSX_ERR_IO; }
    FILE* out = fopen(out_path, "wb");
    if (!out) { fprintf(stderr, "gsx: cannot write %s\n", out_path); return GSX_ERR_IO; }
    char* buf = nullptr; size_t len = 0;
    gsx_format_header(ix, format_sam, complete, &buf, &len); fwrite(buf, 1, len, out); free(buf);
    gsx_params pp = *p; pp.sam_scoring = format_sam ? 1 : 0;
    if (batch_guides == 0) batch_guides = 1u << 20;
    gsx_counters total{}; const size_t n = t.id.size();
    for (size_t b0 = 0; b0 < n; b0 += batch_guides) {
        size_t b1 = std::min(n, b0 + batch_guides);
        std::vector<gsx_guide> g(b1 - b0); std::vector<gsx_guide_row> rows(b1 - b0);
        for (size_t i = b0; i < b1; i++) {
            g[i - b0] = {t.seq[i].c_str(), t.pam[i].c_str()};
            rows[i - b0] = {t.id[i].c_str(), t.seq[i].c_str(), t.pam[i].c_str(), (int)t.positive[i]};
        }
        gsx_result* r = nullptr;
        int rc = gsx_enumerate(ix, g.data(), g.size(), &pp, &r);
        if (rc) { fclose(out); return rc; }
        rc = gsx_format_rows(ix, r, rows.data(), 0, g.size(), &pp, format_sam, complete, &buf, &len);
        if (rc) { gsx_result_free(r); fclose(out); return rc; }
        fwrite(buf, 1, len, out); free(buf);
        gsx_counters c; gsx_result_counters(r, &c);
        total.nodes += c.nodes; total.lookups += c.lookups; total.matches += c.matches; total.hits += c.hits; total.lf_steps += c.lf_steps; total.spills += c.spills; total.launches += c.launches;
        total.ms_search += c.ms_search; total.ms_arrange += c.ms_arrange; total.ms_locate += c.ms_locate; total.ms_score += c.ms_score;
        total.ms_total_device += c.ms_total_device; total.ms_h2d += c.ms_h2d; total.ms_d2h += c.ms_d2h; total.ms_sweep += c.ms_sweep; total.seeds += c.seeds; total.ms_prepare += c.ms_prepare; total.ms_wall += c.ms_wall; total.sectors += c.sectors; total.edited_guides += c.edited_guides;
        gsx_result_free(r);
    }
    fclose(out);
    if (n_guides) *n_guides = n;
    if (counters) *counters = total;
    return GSX_OK;
}
