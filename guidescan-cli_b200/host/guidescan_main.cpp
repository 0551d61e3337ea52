// guidescan -- host CLI with the reference's `index` / `enumerate` command line (reference src/guidescan.cxx:28-95,
// 316-358; option spellings and defaults: SURVEY.md section 8(b)), driving the GPU path through the C ABI.
// Text formatting, file I/O and option parsing are host C++; all arithmetic of the hot path runs on the device.
#include "../../include/gsx.h"
#include <algorithm>
#include <chrono>
#include <fcntl.h>
#include <sys/stat.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static void usage() {
    fprintf(stderr,
            "Guidescan all-in-one interface (B200 build).\n"
            "Usage: guidescan [--version] SUBCOMMAND ...\n\n"
            "  index [--index PREFIX] [--reference-format] GENOME.fa\n"
            "      Builds a genomic index over a FASTA file (GPU suffix sorting; writes PREFIX.gsx + PREFIX.gs).\n"
            "      --reference-format          Also write PREFIX.forward / PREFIX.reverse, the files of the original guidescan (extension).\n"
            "  enumerate [OPTIONS] INDEX_PREFIX\n"
            "      --start                     Match PAM at start of kmer instead at end (default).\n"
            "      --max-off-targets INT=-1    Maximum number of off-targets to store for each number of mismatches.\n"
            "      -n,--threads UINT           Host worker threads that format the output (default: all cores); the search runs on the GPU(s).\n"
            "      -a,--alt-pam TEXT ...       Alternative PAMs used to find off-targets\n"
            "      -m,--mismatches UINT=3      Number of mismatches to allow when finding off-targets\n"
            "      --rna-bulges UINT=0         Max number of RNA bulges to allow when finding off-targets\n"
            "      --dna-bulges UINT=0         Number of DNA bulges to allow when finding off-targets\n"
            "      -t,--threshold INT=-1       Filters gRNAs with off-targets at a distance at or below this threshold\n"
            "      --format TEXT:{csv,sam}     File format for output.\n"
            "      --mode TEXT:{succinct,complete}  Information to output.\n"
            "      -f,--kmers-file FILE        REQUIRED  File containing kmers to build gRNA database over\n"
            "      -o,--output TEXT            REQUIRED  Output file.\n"
            "      --gpus INT=1                Number of GPUs to shard the guides over (extension): devices 0 .. INT-1.\n"
            "      --devices LIST              The same with explicit device ordinals, e.g. 0,2,3 (extension).\n");
}

static std::string lower(std::string s) { std::transform(s.begin(), s.end(), s.begin(), ::tolower); return s; }

static int do_index(int argc, char** argv) {
    std::string fasta, prefix; bool reference_format = false;
    for (int i = 0; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--index" && i + 1 < argc) prefix = argv[++i];
        else if (a == "--reference-format") reference_format = true;
        else if (a.rfind("--index=", 0) == 0) prefix = a.substr(8);
        else fasta = a;
    }
    if (fasta.empty()) { fprintf(stderr, "genome is required\n"); return 106; }
    if (prefix.empty()) { prefix = fasta + ".index"; printf("Index file prefix not specified. Using default: %s\n", prefix.c_str()); }
    gsx_index* ix = nullptr;
    int rc = gsx_index_build(fasta.c_str(), prefix.c_str(), nullptr, 0, &ix);
    if (rc) { fprintf(stderr, "ERROR: %s\n", gsx_last_error()); return 1; }
    if (reference_format) {
        const auto t0 = std::chrono::steady_clock::now();
        if (gsx_index_save_reference_format(ix, prefix.c_str())) { fprintf(stderr, "ERROR: %s\n", gsx_last_error()); gsx_index_close(ix); return 1; }
        utimensat(AT_FDCWD, (prefix + ".gsx").c_str(), nullptr, 0);     // gsx_index_open takes the newer format: keep that the one without conversion
        printf("Wrote %s.forward and %s.reverse in %.2f s.\n", prefix.c_str(), prefix.c_str(), std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    printf("Index construction complete.\n");
    gsx_index_close(ix);
    return 0;
}

static int do_enumerate(int argc, char** argv) {
    gsx_params p; gsx_params_default(&p);
    std::string index, kmers, output, format = "csv", mode = "complete";
    std::vector<std::string> alts; int gpus = 1; std::vector<int> devs;
    for (int i = 0; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&](const char* name) -> const char* { if (i + 1 >= argc) { fprintf(stderr, "%s: 1 required\n", name); exit(106); } return argv[++i]; };
        if (a == "--start") p.start = 1;
        else if (a == "--max-off-targets") p.max_off_targets = atoll(need("--max-off-targets"));
        else if (a == "-n" || a == "--threads") { const int t = atoi(need("--threads")); if (t > 0) setenv("GSX_FORMAT_THREADS", std::to_string(t).c_str(), 1); }
        else if (a == "-a" || a == "--alt-pam") {          // greedy multi-value, as CLI11 parses it
            while (i + 1 < argc && argv[i + 1][0] != '-') alts.push_back(argv[++i]);
        }
        else if (a == "-m" || a == "--mismatches") p.mismatches = (uint32_t)atoi(need("--mismatches"));
        else if (a == "--rna-bulges") p.rna_bulges = (uint32_t)atoi(need("--rna-bulges"));
        else if (a == "--dna-bulges") p.dna_bulges = (uint32_t)atoi(need("--dna-bulges"));
        else if (a == "-t" || a == "--threshold") p.threshold = atoi(need("--threshold"));
        else if (a == "--format") format = lower(need("--format"));
        else if (a == "--mode") mode = lower(need("--mode"));
        else if (a == "-f" || a == "--kmers-file") kmers = need("--kmers-file");
        else if (a == "-o" || a == "--output") output = need("--output");
        else if (a == "--gpus") gpus = atoi(need("--gpus"));
        else if (a == "--devices") {
            std::string l = need("--devices");
            for (size_t b = 0; b < l.size();) { size_t e = l.find(',', b); if (e == std::string::npos) e = l.size(); if (e > b) devs.push_back(atoi(l.substr(b, e - b).c_str())); b = e + 1; }
        }
        else if (a == "-h" || a == "--help") { usage(); return 0; }
        else if (!a.empty() && a[0] == '-') { fprintf(stderr, "The following argument was not expected: %s\n", a.c_str()); return 109; }
        else index = a;
    }
    if (index.empty()) { fprintf(stderr, "index is required\n"); return 106; }
    if (kmers.empty()) { fprintf(stderr, "--kmers-file is required\n"); return 106; }
    if (output.empty()) { fprintf(stderr, "--output is required\n"); return 106; }
    if (format != "csv" && format != "sam") { fprintf(stderr, "--format: %s not in {csv,sam}\n", format.c_str()); return 105; }
    if (mode != "succinct" && mode != "complete") { fprintf(stderr, "--mode: %s not in {succinct,complete}\n", mode.c_str()); return 105; }
    std::vector<const char*> ap; for (auto& s : alts) ap.push_back(s.c_str());
    p.alt_pams = ap.data(); p.n_alt_pams = (uint32_t)ap.size();
    if (devs.empty()) for (int d = 0; d < std::max(1, gpus); d++) devs.push_back(d);
    printf("Loading genome index at \"%s\".\n", index.c_str());
    gsx_index* ix = nullptr;
    int rc = gsx_index_open(index.c_str(), devs.data(), (int)devs.size(), &ix);
    if (rc) { fprintf(stderr, "%s\n", gsx_last_error()); return 1; }
    {
        double t[3] = {0, 0, 0}; gsx_index_open_seconds(ix, t);
        printf("Successfully loaded genome index. (%zu device(s): files %.2f s, device layout %.2f s, replication %.2f s)\n", devs.size(), t[0], t[1], t[2]);
    }
    auto t0 = std::chrono::steady_clock::now();
    size_t n = 0; gsx_counters c;
    rc = gsx_enumerate_file(ix, kmers.c_str(), output.c_str(), &p, format == "sam", mode == "complete", 0, &n, &c);
    if (rc) { fprintf(stderr, "%s\n", gsx_last_error()); gsx_index_close(ix); return 1; }
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("Read in %zu kmer(s).\n", n);
    printf("Processed %zu kmers in %lld seconds. (%.3f s, %.1f kmers/s; device %.1f ms; %llu nodes, %llu hits)\n", n, (long long)secs, secs,
           secs > 0 ? n / secs : 0.0, c.ms_total_device, (unsigned long long)c.nodes, (unsigned long long)c.hits);
    gsx_index_close(ix);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { usage(); return 106; }
    std::string cmd = argv[1];
    if (cmd == "--version") { printf("%s\n", gsx_version()); return 0; }
    if (cmd == "-h" || cmd == "--help") { usage(); return 0; }
    if (cmd == "index") return do_index(argc - 2, argv + 2);
    if (cmd == "enumerate") return do_enumerate(argc - 2, argv + 2);
    if (cmd == "download") { fprintf(stderr, "download: not part of the GPU hot path (no network); use the reference binary\n"); return 1; }
    usage();
    return 106;
}
