/*
 * gs_oracle.c -- CPU restatement of GuideScan2's off-target enumeration hot path (see gs_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into, called from or shipped with the product.
 * Parity status: PINNED against the unmodified reference binary (tests/test_oracle.py, tests/golden/).
 *
 * The FM-index here is deliberately the plainest possible one (byte BWT + occurrence checkpoints every 64
 * rows + the full suffix array): the suffix array of text+'\0' is unique, so SA intervals, SA rows and
 * located positions are identical to those of the reference's sdsl::csa_wt<wt_huff<>,64,8192>
 * (reference src/guidescan.cxx:24-27) whatever the encoding.
 */
#define _GNU_SOURCE
#include "gs_oracle.h"
#include "cfd_tables.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <pthread.h>
#include <sys/types.h>

/* ------------------------------------------------------------------------------------------------ */
/* FM-index                                                                                          */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    uint64_t  n;            /* csa.size() = text length + 1 */
    uint8_t*  bwt;          /* n bytes, 0 in the row whose suffix starts at text position 0 */
    uint32_t* sa;           /* full suffix array (self-check and sampling source); NULL for an imported index */
    uint32_t* sa_samples;   /* SA[64 k] */
    uint64_t  C[257];       /* C[c] = number of symbols smaller than c in text+'\0'  (csa.C[char2comp[c]]) */
    int       present[256];
    int       slot[256];
    int       nslot;
    uint64_t* ckpt;         /* (n/64 + 1) * nslot occurrence counts before row 64*b */
} fm_t;

struct gso_index {
    fm_t      fm[2];        /* 0 = forward text, 1 = reverse-complement text  (src/guidescan.cxx:119-124,167-175) */
    int       n_chr;
    char**    names;
    uint64_t* lens;
    uint64_t  G;
};

typedef struct { uint64_t key; uint32_t pos; } kp_t;
static const uint8_t* g_sort_text; static uint64_t g_sort_n; /* text incl. sentinel, its length */

static int cmp_suffix_tail(const void* a, const void* b) {
    /* both suffixes share their first 21 symbols; the unique sentinel guarantees a difference */
    uint32_t pa = ((const kp_t*)a)->pos + 21, pb = ((const kp_t*)b)->pos + 21;
    uint64_t la = g_sort_n - pa, lb = g_sort_n - pb, l = la < lb ? la : lb;
    int r = memcmp(g_sort_text + pa, g_sort_text + pb, l);
    if (r) return r;
    return la < lb ? -1 : (la > lb ? 1 : 0);
}

/* Suffix array of t[0..n) where t[n-1] == 0 is the unique smallest symbol.  (The reference gets the same
 * array from libdivsufsort: sdsl/include/sdsl/construct_sa.hpp:78-117.) */
static uint32_t* build_sa(const uint8_t* t, uint64_t n) {
    int code[256]; memset(code, 0, sizeof code);
    int seen[256]; memset(seen, 0, sizeof seen);
    for (uint64_t i = 0; i < n; i++) seen[t[i]] = 1;
    int nc = 0;
    for (int c = 0; c < 256; c++) if (seen[c]) code[c] = nc++;
    if (nc > 8) { fprintf(stderr, "gs_oracle: more than 8 distinct symbols\n"); return NULL; }
    kp_t* a = (kp_t*)malloc(sizeof(kp_t) * n), *b = (kp_t*)malloc(sizeof(kp_t) * n);
    /* key = first 21 symbols, 3 bits each, rolling */
    uint64_t key = 0;
    for (int j = 0; j < 21; j++) key = (key << 3) | (uint64_t)((uint64_t)j < n ? code[t[j]] : 0);
    for (uint64_t i = 0; i < n; i++) {
        a[i].key = key; a[i].pos = (uint32_t)i;
        uint64_t nx = i + 21 < n ? (uint64_t)code[t[i + 21]] : 0;
        key = ((key << 3) | nx) & ((1ULL << 63) - 1);
    }
    for (int pass = 0; pass < 8; pass++) {
        size_t cnt[257]; memset(cnt, 0, sizeof cnt);
        int sh = pass * 8;
        for (uint64_t i = 0; i < n; i++) cnt[((a[i].key >> sh) & 255) + 1]++;
        for (int c = 0; c < 256; c++) cnt[c + 1] += cnt[c];
        for (uint64_t i = 0; i < n; i++) b[cnt[(a[i].key >> sh) & 255]++] = a[i];
        kp_t* tmp = a; a = b; b = tmp;
    }
    g_sort_text = t; g_sort_n = n;
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i + 1;
        while (j < n && a[j].key == a[i].key) j++;
        if (j - i > 1) qsort(a + i, j - i, sizeof(kp_t), cmp_suffix_tail);
        i = j;
    }
    uint32_t* sa = (uint32_t*)malloc(sizeof(uint32_t) * n);
    for (uint64_t i = 0; i < n; i++) sa[i] = a[i].pos;
    free(a); free(b);
    return sa;
}

/* occurrence checkpoints + C from a finished BWT (f->bwt, f->n set) */
static void fm_finish(fm_t* f) {
    uint64_t cnt[256]; memset(cnt, 0, sizeof cnt);
    for (uint64_t i = 0; i < f->n; i++) cnt[f->bwt[i]]++;      /* the BWT is a permutation of text+sentinel */
    uint64_t acc = 0;
    for (int c = 0; c < 256; c++) { f->C[c] = acc; acc += cnt[c]; f->present[c] = cnt[c] > 0; }
    f->C[256] = acc;
    for (int c = 0; c < 256; c++) if (f->present[c]) f->slot[c] = f->nslot++;
    uint64_t nb = f->n / 64 + 1;
    f->ckpt = (uint64_t*)calloc(nb * f->nslot, sizeof(uint64_t));
    uint64_t run[8]; memset(run, 0, sizeof run);
    for (uint64_t i = 0; i < f->n; i++) {
        if ((i & 63) == 0) memcpy(f->ckpt + (i / 64) * f->nslot, run, sizeof(uint64_t) * f->nslot);
        run[f->slot[f->bwt[i]]]++;
    }
    if ((f->n & 63) == 0) memcpy(f->ckpt + (f->n / 64) * f->nslot, run, sizeof(uint64_t) * f->nslot);
}

static int fm_build(fm_t* f, const uint8_t* text, uint64_t len) {
    memset(f, 0, sizeof *f);
    f->n = len + 1;
    uint8_t* t = (uint8_t*)malloc(f->n);
    memcpy(t, text, len); t[len] = 0;                 /* sdsl appends the 0 sentinel: construct.hpp:121-165 */
    f->sa = build_sa(t, f->n);
    if (!f->sa) { free(t); return -1; }
    f->bwt = (uint8_t*)malloc(f->n);
    for (uint64_t i = 0; i < f->n; i++) f->bwt[i] = f->sa[i] ? t[f->sa[i] - 1] : 0;
    uint64_t ns = (f->n - 1) / 64 + 1;
    f->sa_samples = (uint32_t*)malloc(ns * sizeof(uint32_t));
    for (uint64_t k = 0; k < ns; k++) f->sa_samples[k] = f->sa[k * 64];
    fm_finish(f);
    free(t);
    return 0;
}

static void fm_free(fm_t* f) { free(f->bwt); free(f->sa); free(f->sa_samples); free(f->ckpt); }

/* csa.rank_bwt(i, c): occurrences of c in BWT[0, i); 0 for a symbol the text does not contain.
 * reference sdsl/include/sdsl/csa_wt.hpp:270-273 -> wt_pc.hpp:360-384 */
static inline uint64_t rank_bwt(const fm_t* f, uint64_t i, int c, gso_counters* ctr) {
    if (ctr) ctr->rank_calls++;
    c &= 255;
    if (!f->present[c]) return 0;
    uint64_t b = i >> 6, r = f->ckpt[b * f->nslot + f->slot[c]];
    for (uint64_t j = b << 6; j < i; j++) r += f->bwt[j] == c;
    return r;
}

/* csa[i]: LF-walk until the row is a multiple of 64, then sample + steps (mod n).
 * reference csa_wt.hpp:333-346, suffix_array_helper.hpp:337-349, csa_sampling_strategy.hpp:64-112 */
static uint64_t sa_walk(const fm_t* f, uint64_t i, gso_counters* ctr) {
    uint64_t off = 0;
    while (i & 63) {
        int c = f->bwt[i];
        i = f->C[c] + rank_bwt(f, i, c, NULL);
        off++;
        if (ctr) ctr->lf_steps++;
    }
    uint64_t r = (uint64_t)f->sa_samples[i >> 6] + off;
    return r < f->n ? r : r - f->n;
}

/* ------------------------------------------------------------------------------------------------ */
/* sequences (reference src/genomics/sequences.cxx:14-46)                                            */
/* ------------------------------------------------------------------------------------------------ */
static char complement_c(char c) {
    switch (c) {
    case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
    case 'a': return 't'; case 't': return 'a'; case 'c': return 'g'; case 'g': return 'c';
    default: return c;
    }
}
static void complement_s(const char* in, char* out) { size_t n = strlen(in); for (size_t i = 0; i < n; i++) out[i] = complement_c(in[i]); out[n] = 0; }
static void revcomp_s(const char* in, char* out) { size_t n = strlen(in); for (size_t i = 0; i < n; i++) out[i] = complement_c(in[n - 1 - i]); out[n] = 0; }

/* ------------------------------------------------------------------------------------------------ */
/* search (reference include/genomics/index.hpp)                                                     */
/* ------------------------------------------------------------------------------------------------ */
#define MAXSEQ 47
typedef struct { char seq[MAXSEQ + 1]; uint64_t sp, ep; uint32_t mm, dna, rna; } match_t;
typedef struct { match_t* v; size_t n, cap; } mvec_t;
static void mvec_push(mvec_t* m, const match_t* x) {
    if (m->n == m->cap) { m->cap = m->cap ? m->cap * 2 : 64; m->v = (match_t*)realloc(m->v, m->cap * sizeof(match_t)); }
    m->v[m->n++] = *x;
}

typedef struct {
    const fm_t* f;
    const char* query; size_t qlen;
    char (*pams)[8]; int n_pams;
    size_t M, R, D, B;
    int counting; uint64_t* count;      /* off_target_counter (process.hpp:27-29) */
    mvec_t* out;                        /* off_target_enumerator (process.hpp:21-23), insertion order */
    gso_counters* ctr;
} sctx_t;

static const char SEARCH_ALPHABET[] = "ATCG";      /* index.hpp:31 */

static void emit(sctx_t* s, const char* str, int len, uint64_t sp, uint64_t ep, uint32_t mm, uint32_t dna, uint32_t rna) {
    if (s->counting) { *s->count += ep - sp + 1; return; }
    match_t m; memset(&m, 0, sizeof m);
    if (len > MAXSEQ) len = MAXSEQ;
    memcpy(m.seq, str, len); m.sp = sp; m.ep = ep; m.mm = mm; m.dna = dna; m.rna = rna;
    mvec_push(s->out, &m);
}

/* PAM / wildcard stage: index.hpp:125-170.  Called with mismatches = 0, k = 0. */
static void pam_search(sctx_t* s, const char* pam, int end, uint64_t sp, uint64_t ep, char* str, int len,
                       size_t mismatches, size_t k, uint32_t mm, uint32_t dna, uint32_t rna) {
    if (s->ctr) s->ctr->pam_nodes++;
    if (end == 0) { emit(s, str, len, sp, ep, mm, dna, rna); return; }
    char c = pam[end - 1];
    uint64_t occ_before = rank_bwt(s->f, sp, c, s->ctr);
    uint64_t occ_within = rank_bwt(s->f, ep + 1, c, s->ctr) - occ_before;
    if (occ_within > 0) {
        uint64_t sp2 = s->f->C[(uint8_t)c] + occ_before, ep2 = sp2 + occ_within - 1;
        str[len] = c;
        pam_search(s, pam, end - 1, sp2, ep2, str, len + 1, mismatches, k, mm, dna, rna);
    }
    size_t cost = 1;
    if (k >= mismatches && c != 'N') return;
    if (c == 'N') cost = 0;
    for (int i = 0; i < 4; i++) {
        char a = SEARCH_ALPHABET[i];
        if (a == c) continue;
        occ_before = rank_bwt(s->f, sp, a, s->ctr);
        occ_within = rank_bwt(s->f, ep + 1, a, s->ctr) - occ_before;
        if (occ_within > 0) {
            uint64_t sp2 = s->f->C[(uint8_t)a] + occ_before, ep2 = sp2 + occ_within - 1;
            str[len] = a;
            pam_search(s, pam, end - 1, sp2, ep2, str, len + 1, mismatches, k + cost, mm, dna, rna);
        }
    }
}

/* mismatch-only variant: index.hpp:182-248 */
static void search_mm(sctx_t* s, ssize_t position, uint64_t sp, uint64_t ep, char* str, int len, size_t k) {
    if (s->ctr) s->ctr->nodes++;
    if (position < 0) {
        for (int p = 0; p < s->n_pams; p++)
            pam_search(s, s->pams[p], (int)strlen(s->pams[p]), sp, ep, str, len, 0, 0, (uint32_t)k, 0, 0);
        return;
    }
    char c = s->query[position];
    uint64_t occ_before = rank_bwt(s->f, sp, c, s->ctr);
    uint64_t occ_within = rank_bwt(s->f, ep + 1, c, s->ctr) - occ_before;
    if (occ_within > 0) {
        uint64_t sp2 = s->f->C[(uint8_t)c] + occ_before, ep2 = sp2 + occ_within - 1;
        str[len] = c;
        search_mm(s, position - 1, sp2, ep2, str, len + 1, k);
    }
    if (k >= s->M) return;
    for (int i = 0; i < 4; i++) {
        char a = SEARCH_ALPHABET[i];
        if (a == c) continue;
        occ_before = rank_bwt(s->f, sp, a, s->ctr);
        occ_within = rank_bwt(s->f, ep + 1, a, s->ctr) - occ_before;
        if (occ_within > 0) {
            uint64_t sp2 = s->f->C[(uint8_t)a] + occ_before, ep2 = sp2 + occ_within - 1;
            str[len] = (char)tolower(a);
            search_mm(s, position - 1, sp2, ep2, str, len + 1, k + 1);
        }
    }
}

typedef struct { uint64_t mm, dna, rna; int state; uint64_t curr; } aff_t;  /* index.hpp:12-20; state 0 none 1 dna 2 rna */

/* bulge-aware variant: index.hpp:250-375 */
static void search_bulge(sctx_t* s, ssize_t position, uint64_t sp, uint64_t ep, char* str, int len, aff_t aff) {
    if (s->ctr) s->ctr->nodes++;
    if (len >= MAXSEQ - 4) return;    /* unreachable with max_bulge_size = 1 and the option ranges tested */
    aff_t d = aff;
    if (s->D > aff.dna) {
        if (aff.state != 1 || d.curr == s->B) { d.state = 1; d.curr = 0; d.dna += 1; }
    }
    if (d.state == 1 && d.curr < s->B && (size_t)position != s->qlen - 1) {
        d.curr += 1;
        for (int i = 0; i < 4; i++) {
            char a = SEARCH_ALPHABET[i];
            uint64_t occ_before = rank_bwt(s->f, sp, a, s->ctr);
            uint64_t occ_within = rank_bwt(s->f, ep + 1, a, s->ctr) - occ_before;
            if (occ_within > 0) {
                uint64_t sp2 = s->f->C[(uint8_t)a] + occ_before, ep2 = sp2 + occ_within - 1;
                str[len] = (char)tolower(a);
                search_bulge(s, position, sp2, ep2, str, len + 1, d);
            }
        }
    }
    if (position < 0) {
        for (int p = 0; p < s->n_pams; p++)
            pam_search(s, s->pams[p], (int)strlen(s->pams[p]), sp, ep, str, len, 0, 0,
                       (uint32_t)aff.mm, (uint32_t)aff.dna, (uint32_t)aff.rna);
        return;
    }
    char c = s->query[position];
    uint64_t occ_before = rank_bwt(s->f, sp, c, s->ctr);
    uint64_t occ_within = rank_bwt(s->f, ep + 1, c, s->ctr) - occ_before;
    if (occ_within > 0) {
        uint64_t sp2 = s->f->C[(uint8_t)c] + occ_before, ep2 = sp2 + occ_within - 1;
        aff_t o = aff; o.state = 0;
        str[len] = c;
        search_bulge(s, position - 1, sp2, ep2, str, len + 1, o);
    }
    if (s->M > aff.mm) {
        for (int i = 0; i < 4; i++) {
            char a = SEARCH_ALPHABET[i];
            if (a == c) continue;
            occ_before = rank_bwt(s->f, sp, a, s->ctr);
            occ_within = rank_bwt(s->f, ep + 1, a, s->ctr) - occ_before;
            if (occ_within > 0) {
                uint64_t sp2 = s->f->C[(uint8_t)a] + occ_before, ep2 = sp2 + occ_within - 1;
                aff_t m = aff; m.state = 0; m.mm += 1;
                str[len] = (char)tolower(a);
                search_bulge(s, position - 1, sp2, ep2, str, len + 1, m);
            }
        }
    }
    aff_t r = aff;
    if (s->R > aff.rna) {
        if (aff.state != 2 || r.curr == s->B) { r.state = 2; r.curr = 0; r.rna += 1; }
    }
    if (r.state == 2 && r.curr < s->B && (size_t)position != s->qlen - 1) {
        r.curr += 1;
        str[len] = '.';
        search_bulge(s, position - 1, sp, ep, str, len + 1, r);
    }
}

/* public entry: index.hpp:377-398 */
static void inexact_search(const fm_t* f, const char* query, char (*pams)[8], int n_pams, size_t M, size_t R, size_t D,
                           size_t B, int counting, uint64_t* count, mvec_t* out, gso_counters* ctr) {
    sctx_t s; memset(&s, 0, sizeof s);
    s.f = f; s.query = query; s.qlen = strlen(query); s.pams = pams; s.n_pams = n_pams;
    s.M = M; s.R = R; s.D = D; s.B = B; s.counting = counting; s.count = count; s.out = out; s.ctr = ctr;
    char str[MAXSEQ + 8]; memset(str, 0, sizeof str);
    if (R == 0 && D == 0) { search_mm(&s, (ssize_t)s.qlen - 1, 0, f->n - 1, str, 0, 0); return; }
    aff_t a = {0, 0, 0, 0, 0};
    search_bulge(&s, (ssize_t)s.qlen - 1, 0, f->n - 1, str, 0, a);
}

/* std::set<match> semantics (structures.hpp:33-43): ordered and de-duplicated on `sequence` alone, the
 * first inserted of equal strings is kept.  Returns, per mismatch bucket, the ordered survivors. */
typedef struct { const match_t* m; size_t ord; } mref_t;
static int cmp_mref(const void* a, const void* b) {
    const mref_t* x = (const mref_t*)a, *y = (const mref_t*)b;
    int r = strcmp(x->m->seq, y->m->seq);        /* all characters are ASCII < 128 */
    if (r) return r;
    return x->ord < y->ord ? -1 : (x->ord > y->ord ? 1 : 0);
}
static size_t bucket_set(const mvec_t* all, uint32_t mm, mref_t* out) {
    size_t n = 0;
    for (size_t i = 0; i < all->n; i++) if (all->v[i].mm == mm) { out[n].m = &all->v[i]; out[n].ord = i; n++; }
    qsort(out, n, sizeof(mref_t), cmp_mref);
    size_t w = 0;
    for (size_t i = 0; i < n; i++) if (w == 0 || strcmp(out[w - 1].m->seq, out[i].m->seq) != 0) out[w++] = out[i];
    return w;
}

/* ------------------------------------------------------------------------------------------------ */
/* coordinates (reference src/genomics/structures.cxx:7-52)                                          */
/* ------------------------------------------------------------------------------------------------ */
/* returns chromosome index or -1 for the sentinel; *offset = 1-based start; *strand = '+'/'-' */
static int resolve_absolute(const gso_index* ix, int64_t abs, size_t seq_len, size_t pam_len, uint64_t* offset, char* strand) {
    *strand = '+';
    if (abs < 0) { abs = -abs; *strand = '-'; }
    int ci = -1; uint64_t clen = 0;
    for (int i = 0; i < ix->n_chr; i++) {
        if (abs <= (int64_t)(ix->lens[i] - 1)) { ci = i; clen = ix->lens[i]; break; }
        abs -= (int64_t)ix->lens[i];
    }
    int64_t start, end;
    if (*strand == '+') { end = abs + 1; start = end - (int64_t)seq_len - (int64_t)pam_len + 1; }
    else { start = abs + 1; end = start + (int64_t)seq_len + (int64_t)pam_len - 1; }
    /* no chromosome found: the reference keeps c = {"",0}; end > 0 = c.length then yields the sentinel for
     * every realistic case (an assert guards it in debug builds) */
    if (start < 0 || end > (int64_t)clen) return -1;
    if (ci < 0) return -1;
    *offset = (uint64_t)start;
    return ci;
}

/* ------------------------------------------------------------------------------------------------ */
/* scoring (reference include/genomics/printer.hpp:98-113)                                           */
/* ------------------------------------------------------------------------------------------------ */
static int base_idx(char c) { switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; } }

float gso_calculate_cfd(const char* sgrna, const char* sequence, const char* pam) {
    if (strlen(sgrna) != 20 || strlen(pam) != 3) return 1.0f;
    float cfd = 1.0f;
    for (int i = 0; i < 20; i++) {
        char g = sgrna[i], q = sequence[i];
        if (g != q) {
            /* key "r<g with T->U>:d<upper(complement(q))>,<i+1>"; a key the table lacks evaluates to 0.0 */
            int r = g == 'U' ? 3 : base_idx(g);      /* 'T' is rewritten to 'U'; both address the rU rows */
            int d = base_idx((char)toupper(complement_c(q)));
            double v = (r >= 0 && d >= 0) ? GSX_CFD_MM[r][d][i] : 0.0;
            cfd = (float)(cfd * v);             /* float *= double: product in double, rounded to float */
        }
    }
    int a = base_idx(pam[1]), b = base_idx(pam[2]);
    /* pam_scores keys are upper-case ACGT dinucleotides */
    double pv = (a >= 0 && b >= 0) ? GSX_CFD_PAM[a][b] : 0.0;
    cfd = (float)(cfd * pv);
    return cfd;
}

/* ------------------------------------------------------------------------------------------------ */
/* per-guide pipeline (reference include/genomics/process.hpp:35-128)                                */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { int64_t abs; const match_t* m; uint64_t row; int index_id; } ot_t;
typedef struct { ot_t* v; size_t n, cap; } otvec_t;
static void ot_push(otvec_t* o, ot_t x) {
    if (o->n == o->cap) { o->cap = o->cap ? o->cap * 2 : 64; o->v = (ot_t*)realloc(o->v, o->cap * sizeof(ot_t)); }
    o->v[o->n++] = x;
}

typedef struct { char* s; size_t n, cap; } sbuf_t;
static void sb_add(sbuf_t* b, const char* s) {
    size_t l = strlen(s);
    if (b->n + l + 1 > b->cap) { b->cap = (b->cap ? b->cap * 2 : 256) + l; b->s = (char*)realloc(b->s, b->cap); }
    memcpy(b->s + b->n, s, l + 1); b->n += l;
}
static void sb_addu(sbuf_t* b, uint64_t v) { char t[32]; snprintf(t, sizeof t, "%llu", (unsigned long long)v); sb_add(b, t); }

typedef struct {
    int dropped;                 /* threshold prefilter returned early: nothing is printed */
    size_t n_dist;               /* opts.mismatches + 1 */
    otvec_t* off;                /* off_targets[d] */
    mvec_t fwd, rev;             /* backing storage of the matches */
} guide_res_t;

static void guide_res_free(guide_res_t* r) {
    for (size_t i = 0; i < r->n_dist; i++) free(r->off[i].v);
    free(r->off); free(r->fwd.v); free(r->rev.v);
}

static void run_guide(const gso_index* ix, const gso_opts* o, const char* seq, const char* kpam, guide_res_t* res, gso_counters* ctr) {
    memset(res, 0, sizeof *res);
    /* pams = alt_pams + k.pam, or {""} when k.pam is empty (alt PAMs dropped): process.hpp:51-56 */
    char pams[17][8], pams_c[17][8]; int n_pams = 0;
    if (kpam[0] == 0) { pams[0][0] = 0; n_pams = 1; }
    else {
        for (int i = 0; i < o->n_alt_pams && i < 16; i++) { strncpy(pams[n_pams], o->alt_pams[i], 7); pams[n_pams][7] = 0; n_pams++; }
        strncpy(pams[n_pams], kpam, 7); pams[n_pams][7] = 0; n_pams++;
    }
    for (int i = 0; i < n_pams; i++) revcomp_s(pams[i], pams_c[i]);
    char kmer[64];
    if (!o->start) revcomp_s(seq, kmer); else { strncpy(kmer, seq, 63); kmer[63] = 0; }
    char (*use)[8] = o->start ? pams : pams_c;

    if (o->threshold > 0) {                                 /* process.hpp:66-76 */
        uint64_t count = 0;
        inexact_search(&ix->fm[0], kmer, use, n_pams, (size_t)o->threshold, 0, 0, 0, 1, &count, NULL, ctr);
        if (count > 1) { res->dropped = 1; return; }
        inexact_search(&ix->fm[1], kmer, use, n_pams, (size_t)o->threshold, 0, 0, 0, 1, &count, NULL, ctr);
        if (count > 1) { res->dropped = 1; return; }
    }
    inexact_search(&ix->fm[0], kmer, use, n_pams, (size_t)o->mismatches, (size_t)o->rna_bulges, (size_t)o->dna_bulges, 1, 0, NULL, &res->fwd, ctr);
    inexact_search(&ix->fm[1], kmer, use, n_pams, (size_t)o->mismatches, (size_t)o->rna_bulges, (size_t)o->dna_bulges, 1, 0, NULL, &res->rev, ctr);

    uint64_t genome_length = 0;
    for (int i = 0; i < ix->n_chr; i++) genome_length += ix->lens[i];

    res->n_dist = (size_t)o->mismatches + 1;
    res->off = (otvec_t*)calloc(res->n_dist, sizeof(otvec_t));
    size_t mx = res->fwd.n > res->rev.n ? res->fwd.n : res->rev.n;
    mref_t* tmp = (mref_t*)malloc(sizeof(mref_t) * (mx + 1));
    for (size_t d = 0; d < res->n_dist; d++) {              /* process.hpp:100-115 */
        size_t k = bucket_set(&res->fwd, (uint32_t)d, tmp);
        for (size_t i = 0; i < k; i++)
            for (uint64_t j = tmp[i].m->sp; j <= tmp[i].m->ep; j++) {
                ot_t t = { -(int64_t)sa_walk(&ix->fm[0], j, ctr), tmp[i].m, j, 0 };
                ot_push(&res->off[d], t);
            }
        k = bucket_set(&res->rev, (uint32_t)d, tmp);
        for (size_t i = 0; i < k; i++)
            for (uint64_t j = tmp[i].m->sp; j <= tmp[i].m->ep; j++) {
                ot_t t = { (int64_t)(genome_length - (sa_walk(&ix->fm[1], j, ctr) + 1)), tmp[i].m, j, 1 };
                ot_push(&res->off[d], t);
            }
        if (ctr) ctr->hits += res->off[d].n;
    }
    free(tmp);
}

static void match_pam(const char* match_sequence, char* pam) {          /* printer.hpp:134-139,265-270 */
    size_t l = strlen(match_sequence);
    if (l < 20) { pam[0] = 0; return; }
    size_t k = l - 20 < 3 ? l - 20 : 3;
    memcpy(pam, match_sequence + 20, k); pam[k] = 0;
}

static void fmt_float(float f, char* out, size_t n) { snprintf(out, n, "%f", (double)f); }   /* std::to_string(float) */

/* printer.hpp:244-300 (+ 189-242) */
static float csv_lines(const gso_index* ix, const gso_opts* o, const char* id, const char* seq, const char* kpam,
                       const guide_res_t* r, sbuf_t* out, int* any) {
    float cfd_sum = 0.0f; int perfect = 0, none = 1;
    size_t sl = strlen(seq), pl = strlen(kpam);
    sbuf_t lines = {0, 0, 0}; size_t n_lines = 0;
    size_t* line_end = NULL; size_t le_cap = 0;
    char sequence[80];
    if (o->start) snprintf(sequence, sizeof sequence, "%s%s", kpam, seq); else snprintf(sequence, sizeof sequence, "%s%s", seq, kpam);
    for (size_t d = 0; d < r->n_dist; d++) {
        for (int64_t i = 0; i < (int64_t)r->off[d].n; i++) {
            none = 0;
            if (o->max_off_targets != -1 && i >= o->max_off_targets) break;
            const ot_t* t = &r->off[d].v[i];
            char ms[MAXSEQ + 1], pam[8];
            complement_s(t->m->seq, ms);
            match_pam(ms, pam);
            if (t->m->mm == 0 && strlen(pam) == 3 && pam[1] == 'G' && pam[2] == 'G') perfect = 1;
            uint64_t off; char strand;
            int ci = resolve_absolute(ix, t->abs, sl, pl, &off, &strand);
            if (ci < 0) continue;
            sb_add(&lines, id); sb_add(&lines, ","); sb_add(&lines, sequence); sb_add(&lines, ",");
            sb_add(&lines, ix->names[ci]); sb_add(&lines, ","); sb_addu(&lines, off); sb_add(&lines, ",");
            char st[2] = { strand, 0 }; sb_add(&lines, st); sb_add(&lines, ","); sb_addu(&lines, t->m->mm);
            if (o->complete) {
                sb_add(&lines, ","); sb_add(&lines, ms); sb_add(&lines, ","); sb_addu(&lines, t->m->rna);
                sb_add(&lines, ","); sb_addu(&lines, t->m->dna);
            }
            if (n_lines == le_cap) { le_cap = le_cap ? le_cap * 2 : 64; line_end = (size_t*)realloc(line_end, le_cap * sizeof(size_t)); }
            line_end[n_lines++] = lines.n;
            cfd_sum += gso_calculate_cfd(seq, ms, pam);
        }
    }
    float specificity = 0.0f;
    if (none) {
        sb_add(out, id); sb_add(out, ","); sb_add(out, sequence); sb_add(out, ",NA,NA,NA,0");
        if (o->complete) sb_add(out, ",NA,NA,NA");
        sb_add(out, ",1.0\n");
        specificity = 1.0f;
    } else {
        if (!perfect) cfd_sum += 1;
        if (cfd_sum > 0) specificity = 1 / cfd_sum;
        char sp[64]; fmt_float(specificity, sp, sizeof sp);
        size_t b = 0;
        for (size_t i = 0; i < n_lines; i++) {
            char save = lines.s[line_end[i]]; lines.s[line_end[i]] = 0;
            sb_add(out, lines.s + b); sb_add(out, ","); sb_add(out, sp); sb_add(out, "\n");
            lines.s[line_end[i]] = save; b = line_end[i];
        }
    }
    if (any) *any = !none;
    free(lines.s); free(line_end);
    return specificity;
}

static void hex_le64(sbuf_t* b, uint64_t v) {                            /* printer.hpp:18-88 */
    static const char H[] = "0123456789abcdef";
    char t[17];
    for (int i = 0; i < 8; i++) { unsigned char lo = (unsigned char)(v & 255); v >>= 8; t[2 * i] = H[lo >> 4]; t[2 * i + 1] = H[lo & 15]; }
    t[16] = 0; sb_add(b, t);
}

/* printer.hpp:115-170 and 302-360 */
static void sam_lines(const gso_index* ix, const gso_opts* o, const char* id, const char* seq, const char* kpam,
                      int sense_positive, const guide_res_t* r, sbuf_t* out) {
    size_t sl = strlen(seq), pl = strlen(kpam);
    int64_t delim = 0; for (int i = 0; i < ix->n_chr; i++) delim += (int64_t)ix->lens[i];
    delim = -(delim + 1);
    sbuf_t hex = {0, 0, 0}; sb_add(&hex, "");
    float cfd_sum = 0.0f; int perfect = 0;
    for (size_t d = 0; d < r->n_dist; d++) {
        int64_t n_off = 0;
        for (size_t i = 0; i < r->off[d].n; i++) {
            if (o->max_off_targets != -1 && n_off >= o->max_off_targets) break;
            const ot_t* t = &r->off[d].v[i];
            char ms[MAXSEQ + 1], pam[8];
            complement_s(t->m->seq, ms); match_pam(ms, pam);
            if (t->m->mm == 0 && strlen(pam) == 3 && pam[1] == 'G' && pam[2] == 'G') perfect = 1;
            uint64_t off; char strand;
            if (resolve_absolute(ix, t->abs, sl, pl, &off, &strand) < 0) continue;
            hex_le64(&hex, (uint64_t)t->abs);
            cfd_sum += gso_calculate_cfd(seq, ms, pam);
            n_off++;
        }
        hex_le64(&hex, (uint64_t)d);
        hex_le64(&hex, (uint64_t)delim);
    }
    float specificity = 0.0f;
    if (!perfect) cfd_sum += 1;
    if (cfd_sum > 0) specificity = 1 / cfd_sum;
    char sp[64]; fmt_float(specificity, sp, sizeof sp);
    char sequence[80], rc[80];
    if (o->start) snprintf(sequence, sizeof sequence, "%s%s", kpam, seq); else snprintf(sequence, sizeof sequence, "%s%s", seq, kpam);
    revcomp_s(sequence, rc);
    for (size_t d = 0; d < r->n_dist; d++)
        for (size_t i = 0; i < r->off[d].n; i++) {
            const ot_t* t = &r->off[d].v[i];
            if (t->m->mm != 0) continue;
            uint64_t off = 0; char strand;
            int ci = resolve_absolute(ix, t->abs, sl, pl, &off, &strand);
            if (ci < 0) off = 0;
            sb_add(out, id); sb_add(out, "\t"); sb_add(out, sense_positive ? "0" : "16"); sb_add(out, "\t");
            sb_add(out, ci < 0 ? "" : ix->names[ci]); sb_add(out, "\t"); sb_addu(out, off); sb_add(out, "\t100\t");
            sb_addu(out, strlen(sequence)); sb_add(out, "M\t*\t0\t0\t"); sb_add(out, sense_positive ? sequence : rc);
            sb_add(out, "\t*");
            for (size_t k = 0; k < r->n_dist; k++) { sb_add(out, "\tk"); sb_addu(out, k); sb_add(out, ":i:"); sb_addu(out, r->off[k].n); }
            if (o->complete) { sb_add(out, "\tof:H:"); sb_add(out, hex.s); }
            sb_add(out, "\tsp:f:"); sb_add(out, sp); sb_add(out, "\n");
        }
    free(hex.s);
}

char* gso_process_kmer(const gso_index* ix, const gso_opts* o, const char* id, const char* seq, const char* pam,
                       int sense_positive, gso_counters* ctr) {
    guide_res_t r; run_guide(ix, o, seq, pam, &r, ctr);
    sbuf_t out = {0, 0, 0}; sb_add(&out, "");
    if (!r.dropped) {
        if (o->format_sam) sam_lines(ix, o, id, seq, pam, sense_positive, &r, &out);
        else csv_lines(ix, o, id, seq, pam, &r, &out, NULL);
    }
    guide_res_free(&r);
    return out.s;
}

int gso_enumerate_hits(const gso_index* ix, const gso_opts* o, const char* seq, const char* pam, gso_hit** hits,
                       uint64_t* n_hits, float* specificity, int* dropped, gso_counters* ctr) {
    guide_res_t r; run_guide(ix, o, seq, pam, &r, ctr);
    *dropped = r.dropped; *n_hits = 0; *hits = NULL; *specificity = 0.0f;
    if (r.dropped) { guide_res_free(&r); return 0; }
    size_t tot = 0; for (size_t d = 0; d < r.n_dist; d++) tot += r.off[d].n;
    gso_hit* h = (gso_hit*)calloc(tot + 1, sizeof(gso_hit)); size_t w = 0;
    for (size_t d = 0; d < r.n_dist; d++)
        for (size_t i = 0; i < r.off[d].n; i++) {
            const ot_t* t = &r.off[d].v[i];
            h[w].abs_pos = t->abs; h[w].sa_row = t->row; h[w].distance = t->m->mm; h[w].rna = t->m->rna; h[w].dna = t->m->dna;
            h[w].index_id = (uint32_t)t->index_id; memcpy(h[w].seq, t->m->seq, sizeof h[w].seq - 1); w++;
        }
    sbuf_t tmp = {0, 0, 0}; sb_add(&tmp, "");
    gso_opts oc = *o; oc.format_sam = 0;
    *specificity = csv_lines(ix, &oc, "x", seq, pam, &r, &tmp, NULL);
    free(tmp.s);
    *hits = h; *n_hits = tot;
    guide_res_free(&r);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* file level: guides CSV (reference src/genomics/kmer.cxx:9-25), headers (printer.hpp:173-187)      */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { char* id; char* seq; char* pam; int positive; } guide_t;

static char* trim(char* s) { while (*s == ' ' || *s == '\t') s++; size_t l = strlen(s); while (l && (s[l - 1] == ' ' || s[l - 1] == '\t' || s[l - 1] == '\r' || s[l - 1] == '\n')) s[--l] = 0; return s; }

static int64_t read_guides(const char* path, guide_t** out) {
    FILE* f = fopen(path, "r"); if (!f) return -1;
    char* line = NULL; size_t cap = 0; ssize_t l;
    int col_of[6] = {-1, -1, -1, -1, -1, -1}; const char* want[6] = {"id", "sequence", "pam", "chromosome", "position", "sense"};
    if ((l = getline(&line, &cap, f)) < 0) { fclose(f); free(line); return -2; }
    { int c = 0; char* save; for (char* tok = strtok_r(line, ",", &save); tok; tok = strtok_r(NULL, ",", &save), c++) { char* t = trim(tok); for (int k = 0; k < 6; k++) if (!strcmp(t, want[k])) col_of[k] = c; } }
    for (int k = 0; k < 6; k++) if (col_of[k] < 0) { fclose(f); free(line); return -3; }
    guide_t* g = NULL; int64_t n = 0, gcap = 0;
    while ((l = getline(&line, &cap, f)) >= 0) {
        char* fields[16]; int nf = 0; char* p = line;
        while (l && (line[l - 1] == '\n' || line[l - 1] == '\r')) line[--l] = 0;
        if (l == 0) continue;
        fields[nf++] = p;
        for (; *p && nf < 16; p++) if (*p == ',') { *p = 0; fields[nf++] = p + 1; }
        if (nf < 6) { fclose(f); free(line); return -4; }
        if (n == gcap) { gcap = gcap ? gcap * 2 : 1024; g = (guide_t*)realloc(g, gcap * sizeof(guide_t)); }
        g[n].id = strdup(trim(fields[col_of[0]])); g[n].seq = strdup(trim(fields[col_of[1]])); g[n].pam = strdup(trim(fields[col_of[2]]));
        g[n].positive = strcmp(trim(fields[col_of[5]]), "+") == 0;
        n++;
    }
    fclose(f); free(line); *out = g; return n;
}

typedef struct { const gso_index* ix; const gso_opts* o; guide_t* g; int64_t n; char** res; int64_t next; pthread_mutex_t mu; gso_counters ctr; } job_t;
static void* worker(void* p) {
    job_t* j = (job_t*)p; gso_counters local; memset(&local, 0, sizeof local);
    for (;;) {
        pthread_mutex_lock(&j->mu); int64_t i = j->next++; pthread_mutex_unlock(&j->mu);
        if (i >= j->n) break;
        j->res[i] = gso_process_kmer(j->ix, j->o, j->g[i].id, j->g[i].seq, j->g[i].pam, j->g[i].positive, &local);
    }
    pthread_mutex_lock(&j->mu);
    j->ctr.nodes += local.nodes; j->ctr.pam_nodes += local.pam_nodes; j->ctr.rank_calls += local.rank_calls; j->ctr.lf_steps += local.lf_steps; j->ctr.hits += local.hits;
    pthread_mutex_unlock(&j->mu);
    return NULL;
}

int64_t gso_enumerate_file(const gso_index* ix, const gso_opts* o, const char* kmers_csv, const char* out_path, int nthreads, gso_counters* ctr) {
    guide_t* g = NULL; int64_t n = read_guides(kmers_csv, &g);
    if (n < 0) return n;
    FILE* out = fopen(out_path, "w"); if (!out) return -10;
    if (o->format_sam) {
        fprintf(out, "@HD\tVN:1.0\tSO:unknown\n@PG\tID:Guidescan\tVN:2.0.0\n");
        for (int i = 0; i < ix->n_chr; i++) fprintf(out, "@SQ\tSN:%s\tLN:%llu\n", ix->names[i], (unsigned long long)ix->lens[i]);
    } else {
        fprintf(out, "id,sequence,match_chrm,match_position,match_strand,match_distance");
        if (o->complete) fprintf(out, ",match_sequence,rna_bulges,dna_bulges");
        fprintf(out, ",specificity\n");
    }
    job_t j; memset(&j, 0, sizeof j); j.ix = ix; j.o = o; j.g = g; j.n = n; j.res = (char**)calloc(n + 1, sizeof(char*)); pthread_mutex_init(&j.mu, NULL);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, worker, &j);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    for (int64_t i = 0; i < n; i++) { fputs(j.res[i], out); free(j.res[i]); free(g[i].id); free(g[i].seq); free(g[i].pam); }
    if (ctr) *ctr = j.ctr;
    free(j.res); free(g); fclose(out);
    return n;
}

/* ------------------------------------------------------------------------------------------------ */
/* index construction front-ends                                                                     */
/* ------------------------------------------------------------------------------------------------ */
gso_index* gso_index_from_text(const uint8_t* fwd, uint64_t G, int n_chr, const char* const* names, const uint64_t* lens) {
    gso_index* ix = (gso_index*)calloc(1, sizeof *ix);
    ix->G = G; ix->n_chr = n_chr; ix->names = (char**)calloc(n_chr + 1, sizeof(char*)); ix->lens = (uint64_t*)calloc(n_chr + 1, sizeof(uint64_t));
    for (int i = 0; i < n_chr; i++) { ix->names[i] = strdup(names[i]); ix->lens[i] = lens[i]; }
    uint8_t* rc = (uint8_t*)malloc(G + 1);
    for (uint64_t i = 0; i < G; i++) rc[i] = (uint8_t)complement_c((char)fwd[G - 1 - i]);     /* seq_io.cxx:65-72 */
    int e0 = fm_build(&ix->fm[0], fwd, G), e1 = fm_build(&ix->fm[1], rc, G);
    free(rc);
    if (e0 || e1) { gso_index_free(ix); return NULL; }
    return ix;
}

/* FASTA -> raw upper-case sequence + chromosome table: seq_io.cxx:57-63 and 74-110 */
gso_index* gso_index_from_fasta(const char* path) {
    FILE* f = fopen(path, "r"); if (!f) return NULL;
    char* line = NULL; size_t cap = 0; ssize_t l;
    uint8_t* seq = NULL; uint64_t n = 0, scap = 0;
    char** names = NULL; uint64_t* lens = NULL; int n_chr = 0, ccap = 0;
    while ((l = getline(&line, &cap, f)) >= 0) {
        while (l && (line[l - 1] == '\n')) line[--l] = 0;
        if (l > 0 && line[0] == '>') {
            char* s = line + 1; while (*s && isspace((unsigned char)*s)) s++;
            char* e = s; while (*e && *e != ' ') e++;     /* first space-delimited word (after trimming) */
            size_t tl = strlen(s); while (tl && isspace((unsigned char)s[tl - 1])) s[--tl] = 0;
            if (e > s + tl) e = s + tl;
            *e = 0;
            if (n_chr == ccap) { ccap = ccap ? ccap * 2 : 16; names = (char**)realloc(names, ccap * sizeof(char*)); lens = (uint64_t*)realloc(lens, ccap * sizeof(uint64_t)); }
            names[n_chr] = strdup(s); lens[n_chr] = 0; n_chr++;
            continue;
        }
        /* chromosome length counts the raw line length (seq_io.cxx:105), the sequence is trimmed + upper-cased */
        if (n_chr) lens[n_chr - 1] += (uint64_t)l;
        char* s = line; while (*s && isspace((unsigned char)*s)) s++;
        size_t tl = strlen(s); while (tl && isspace((unsigned char)s[tl - 1])) tl--;
        if (n + tl + 1 > scap) { scap = (scap ? scap * 2 : (1 << 20)) + tl; seq = (uint8_t*)realloc(seq, scap); }
        for (size_t i = 0; i < tl; i++) seq[n + i] = (uint8_t)toupper((unsigned char)s[i]);
        n += tl;
    }
    fclose(f); free(line);
    gso_index* ix = gso_index_from_text(seq, n, n_chr, (const char* const*)names, lens);
    for (int i = 0; i < n_chr; i++) free(names[i]);
    free(names); free(lens); free(seq);
    return ix;
}

/* An index over an existing BWT (bytes, 0 = sentinel row) and SA samples every 64 rows, e.g. exported by the GPU index
 * builder -- lets the CPU port run on genomes whose suffix array the simple sorter above could not build in time. */
gso_index* gso_index_from_bwt(const uint8_t* bwt_fwd, const uint32_t* sa64_fwd, const uint8_t* bwt_rev, const uint32_t* sa64_rev,
                              uint64_t n, int n_chr, const char* const* names, const uint64_t* lens) {
    gso_index* ix = (gso_index*)calloc(1, sizeof *ix);
    ix->G = n - 1; ix->n_chr = n_chr; ix->names = (char**)calloc(n_chr + 1, sizeof(char*)); ix->lens = (uint64_t*)calloc(n_chr + 1, sizeof(uint64_t));
    for (int i = 0; i < n_chr; i++) { ix->names[i] = strdup(names[i]); ix->lens[i] = lens[i]; }
    const uint8_t* b[2] = {bwt_fwd, bwt_rev}; const uint32_t* sm[2] = {sa64_fwd, sa64_rev};
    uint64_t ns = (n - 1) / 64 + 1;
    for (int s = 0; s < 2; s++) {
        fm_t* f = &ix->fm[s]; memset(f, 0, sizeof *f);
        f->n = n; f->bwt = (uint8_t*)malloc(n); memcpy(f->bwt, b[s], n);
        f->sa_samples = (uint32_t*)malloc(ns * sizeof(uint32_t)); memcpy(f->sa_samples, sm[s], ns * sizeof(uint32_t));
        fm_finish(f);
    }
    return ix;
}

void gso_index_free(gso_index* ix) {
    if (!ix) return;
    fm_free(&ix->fm[0]); fm_free(&ix->fm[1]);
    for (int i = 0; i < ix->n_chr; i++) free(ix->names[i]);
    free(ix->names); free(ix->lens); free(ix);
}
uint64_t gso_index_n(const gso_index* ix) { return ix->fm[0].n; }
int gso_index_n_chr(const gso_index* ix) { return ix->n_chr; }
const char* gso_index_chr_name(const gso_index* ix, int i) { return ix->names[i]; }
uint64_t gso_index_chr_len(const gso_index* ix, int i) { return ix->lens[i]; }
uint64_t gso_rank_bwt(const gso_index* ix, int s, uint64_t i, int c) { return rank_bwt(&ix->fm[s], i, c, NULL); }
uint64_t gso_sa(const gso_index* ix, int s, uint64_t row) { return sa_walk(&ix->fm[s], row, NULL); }
uint64_t gso_sa_direct(const gso_index* ix, int s, uint64_t row) { return ix->fm[s].sa ? ix->fm[s].sa[row] : sa_walk(&ix->fm[s], row, NULL); }
uint8_t gso_bwt(const gso_index* ix, int s, uint64_t row) { return ix->fm[s].bwt[row]; }
uint64_t gso_C(const gso_index* ix, int s, int c) { return ix->fm[s].C[c & 255]; }
void gso_free(void* p) { free(p); }
