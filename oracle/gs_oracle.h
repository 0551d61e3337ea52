/*
 * gs_oracle.h -- CPU restatement of GuideScan2's off-target enumeration hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under guidescan-cli_b200/ (the product) may include, link or call
 * this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it, as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement byte-for-byte against outputs of the
 * unmodified reference binary (oracle/_ref/guidescan, built by oracle/Makefile.ref) committed under
 * tests/golden/ by tests/golden/make_golden.py, and -- when oracle/_ref/guidescan is present -- against
 * fresh runs of that binary on seeded inputs.  The reference itself ships no tests of this path
 * (SURVEY.md section 4).
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 */
#ifndef GS_ORACLE_H
#define GS_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct gso_index gso_index;

typedef struct {
    int32_t mismatches;        /* -m, default 3            src/guidescan.cxx:39-76 */
    int32_t rna_bulges;        /* --rna-bulges, default 0 */
    int32_t dna_bulges;        /* --dna-bulges, default 0 */
    int32_t threshold;         /* -t, default -1 (active only if > 0; process.hpp:66) */
    int32_t start;             /* --start flag */
    int32_t format_sam;        /* 0 csv, 1 sam */
    int32_t complete;          /* --mode complete (1) / succinct (0) */
    int32_t n_alt_pams;
    int64_t max_off_targets;   /* --max-off-targets, default -1 */
    const char* alt_pams[16];
} gso_opts;

typedef struct {
    int64_t  abs_pos;          /* process.hpp:104,111 sign/strand encoded absolute coordinate */
    uint64_t sa_row;
    uint32_t distance, rna, dna, index_id;   /* index_id 0 = forward index ("-"), 1 = reverse ("+") */
    char     seq[48];          /* match.sequence as the reference holds it (pre-complement) */
} gso_hit;

typedef struct {
    uint64_t nodes;            /* recursive calls of the protospacer-stage search (visited nodes) */
    uint64_t pam_nodes;        /* recursive calls of the PAM-stage search */
    uint64_t rank_calls;       /* csa.rank_bwt invocations */
    uint64_t lf_steps;         /* LF steps spent in locate */
    uint64_t hits;
} gso_counters;

/* index over text (forward strand, upper-case) + chromosome table; builds both strand FM-indexes */
gso_index* gso_index_from_text(const uint8_t* fwd, uint64_t G, int n_chr,
                               const char* const* names, const uint64_t* lens);
gso_index* gso_index_from_fasta(const char* fasta_path);
gso_index* gso_index_from_bwt(const uint8_t* bwt_fwd, const uint32_t* sa64_fwd, const uint8_t* bwt_rev, const uint32_t* sa64_rev,
                              uint64_t n, int n_chr, const char* const* names, const uint64_t* lens);
void       gso_index_free(gso_index*);
uint64_t   gso_index_n(const gso_index*);                       /* csa.size() = G + 1 */
int        gso_index_n_chr(const gso_index*);
const char* gso_index_chr_name(const gso_index*, int i);
uint64_t   gso_index_chr_len(const gso_index*, int i);
uint64_t   gso_rank_bwt(const gso_index*, int strand, uint64_t i, int c);
uint64_t   gso_sa(const gso_index*, int strand, uint64_t row);  /* csa[row] via LF walk to a sample */
uint64_t   gso_sa_direct(const gso_index*, int strand, uint64_t row); /* from the full SA (self-check) */
uint8_t    gso_bwt(const gso_index*, int strand, uint64_t row);
uint64_t   gso_C(const gso_index*, int strand, int c);

/* one guide: returns malloc'ed CSV/SAM text (possibly empty) -- process.hpp:35-128 */
char* gso_process_kmer(const gso_index*, const gso_opts*, const char* id, const char* seq, const char* pam,
                       int sense_positive, gso_counters* ctr);
/* one guide: the ordered hit list of process.hpp:100-115, plus specificity (CSV rule) and dropped flag */
int   gso_enumerate_hits(const gso_index*, const gso_opts*, const char* seq, const char* pam,
                         gso_hit** hits, uint64_t* n_hits, float* specificity, int* dropped,
                         gso_counters* ctr);
/* whole file, header included, guides in input order; returns number of guides or <0 on error */
int64_t gso_enumerate_file(const gso_index*, const gso_opts*, const char* kmers_csv, const char* out_path,
                           int nthreads, gso_counters* ctr);
float gso_calculate_cfd(const char* sgrna, const char* sequence, const char* pam);
void  gso_free(void*);

#ifdef __cplusplus
}
#endif
#endif
