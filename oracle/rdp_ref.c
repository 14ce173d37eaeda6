/*
 * rdp_ref.c -- CPU ORACLE for Stage A (RDP Classifier 2.5 semantics).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (pangea-plus_b200/)
 * may include, link or execute this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker
 * or as the timed CPU baseline.
 *
 * PARITY UNPINNED: the algorithm lives in an un-vendored third-party
 * dependency of the reference -- RDP Classifier 2.5 (rdp_classifier_2.5.zip,
 * SourceForge project rdp-classifier), downloaded by
 * Classify/RunRDP/install_RDPClassifier.sh:61 and invoked only at
 * README.md:119 (`java -Xmx1g -jar rdp_classifier-2.5.jar -q in -o out`).
 * The reference stores no RDP output, bootstrap value, word count or score,
 * and no JVM exists in the build container, so this file restates the
 * PUBLISHED algorithm (Wang et al. 2007; upstream edu.msu.cme.rdp.classifier)
 * row by row as listed in SURVEY.md section 8(a) A1-A9.  What it is pinned
 * against: the java.util.Random known-answer streams of row A8, hand-checked
 * 8-mer ids, and a hand-computed toy model (tests/test_oracle_rdp.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define RDP_WORDSIZE      8
#define RDP_NWORDS        65536          /* 4^8 */
#define RDP_MASK          0xFFFF
#define RDP_NUM_OF_RUNS   100            /* A8: bootstrap replicates */
#define RDP_MIN_SEQ_LEN   50             /* A2 */
#define RDP_SEED          1ULL           /* A8: setSeed(1) per read */

/* ------------------------------------------------------------------ A1 */
/* GoodWordIterator: A/a=0, T/t/U/u=1, G/g=2, C/c=3; anything else restarts
 * the run so every 8-mer touching it is skipped.  Duplicates kept, sequence
 * order.  Returns n (<= len-7). */
static int base_code(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'T': case 't': case 'U': case 'u': return 1;
    case 'G': case 'g': return 2;
    case 'C': case 'c': return 3;
    default: return -1;
    }
}

int rdp_words(const char *seq, int len, int32_t *words)
{
    int n = 0, run = 0;
    uint32_t w = 0;
    for (int i = 0; i < len; i++) {
        int c = base_code((unsigned char)seq[i]);
        if (c < 0) { run = 0; w = 0; continue; }
        w = ((w << 2) & RDP_MASK) | (uint32_t)c;
        if (++run >= RDP_WORDSIZE) words[n++] = (int32_t)w;
    }
    return n;
}

/* A3: reverse complement of one word; complement lookup {1,0,3,2} = XOR 1. */
int32_t rdp_revcomp_word(int32_t w)
{
    uint32_t r = 0, x = (uint32_t)w;
    for (int i = 0; i < RDP_WORDSIZE; i++) {
        r = (r << 2) | ((x & 3u) ^ 1u);
        x >>= 2;
    }
    return (int32_t)r;
}

/* ------------------------------------------------------------------ A8 RNG */
/* java.util.Random, bit for bit. */
typedef struct { uint64_t s; } jrandom;
#define JR_MULT 0x5DEECE66DULL
#define JR_MASK ((1ULL << 48) - 1)

void jr_set_seed(jrandom *r, uint64_t seed) { r->s = (seed ^ JR_MULT) & JR_MASK; }

int32_t jr_next(jrandom *r, int bits)
{
    r->s = (r->s * JR_MULT + 0xBULL) & JR_MASK;
    return (int32_t)(r->s >> (48 - bits));   /* (int)(seed >>> (48-bits)) */
}

int32_t jr_next_int(jrandom *r, int32_t n)
{
    if ((n & -n) == n)                       /* power of two */
        return (int32_t)(((int64_t)n * (int64_t)jr_next(r, 31)) >> 31);
    int32_t bits, val;
    do {
        bits = jr_next(r, 31);
        val = bits % n;
        /* Java: while (bits - val + (n-1) < 0)  i.e. 32-bit overflow */
    } while ((int64_t)bits - val + (n - 1) > 0x7FFFFFFFLL);
    return val;
}

/* exported for the known-answer tests */
void rdp_jrandom_stream(uint64_t seed, int32_t n, int count, int32_t *out)
{
    jrandom r; jr_set_seed(&r, seed);
    for (int i = 0; i < count; i++) out[i] = jr_next_int(&r, n);
}
void rdp_jrandom_ints(uint64_t seed, int count, int32_t *out)
{
    jrandom r; jr_set_seed(&r, seed);
    for (int i = 0; i < count; i++) out[i] = jr_next(&r, 32);
}

/* ------------------------------------------------------------------ model */
typedef struct {
    int      G;
    int64_t  N;            /* training sequences */
    int32_t *m;            /* [65536][G] sequences of genus g containing w  (A5) */
    int32_t *nw;           /* [65536]    sequences containing w             (A5) */
    int32_t *M;            /* [G]        sequences per genus (leaveCount)   (A5) */
    float   *logPrior;     /* [65536]                                        (A6) */
    float   *logLeave;     /* [G]                                            (A6) */
    float   *logP;         /* [65536][G] dense; absent := prior - leave      (A4) */
    uint8_t *seen;         /* scratch bitmap for per-sequence dedupe */
} rdp_model;

rdp_model *rdp_model_new(int G)
{
    rdp_model *md = (rdp_model *)calloc(1, sizeof *md);
    md->G = G;
    md->m  = (int32_t *)calloc((size_t)RDP_NWORDS * G, sizeof(int32_t));
    md->nw = (int32_t *)calloc(RDP_NWORDS, sizeof(int32_t));
    md->M  = (int32_t *)calloc(G, sizeof(int32_t));
    md->seen = (uint8_t *)calloc(RDP_NWORDS, 1);
    return md;
}

void rdp_model_free(rdp_model *md)
{
    if (!md) return;
    free(md->m); free(md->nw); free(md->M); free(md->logPrior);
    free(md->logLeave); free(md->logP); free(md->seen); free(md);
}

/* A5: one training sequence, forward strand only, DISTINCT words. */
void rdp_train_add(rdp_model *md, const char *seq, int len, int genus)
{
    int32_t *words = (int32_t *)malloc(sizeof(int32_t) * (size_t)(len > 0 ? len : 1));
    int n = rdp_words(seq, len, words);
    for (int i = 0; i < n; i++) {
        int32_t w = words[i];
        if (md->seen[w]) continue;
        md->seen[w] = 1;
        md->m[(size_t)w * md->G + genus]++;
        md->nw[w]++;
    }
    for (int i = 0; i < n; i++) md->seen[words[i]] = 0;
    md->M[genus]++;
    md->N++;
    free(words);
}

/* A6: derive the log tables.  Upstream keeps the word prior and the
 * conditional quotient in Java `float`, widens to double only for Math.log and
 * narrows the result; this restatement does the same (named choice, see
 * DESIGN.md "A6 arithmetic").  All fp32 ops below are single IEEE operations
 * (compile with -ffp-contract=off; there is no a*b+c here anyway). */
void rdp_train_finish(rdp_model *md)
{
    int G = md->G;
    free(md->logPrior); free(md->logLeave); free(md->logP);
    md->logPrior = (float *)malloc(sizeof(float) * RDP_NWORDS);
    md->logLeave = (float *)malloc(sizeof(float) * G);
    md->logP     = (float *)malloc(sizeof(float) * (size_t)RDP_NWORDS * G);
    float Nf1 = (float)md->N + 1.0f;
    for (int g = 0; g < G; g++)
        md->logLeave[g] = (float)log((double)((float)md->M[g] + 1.0f));
    for (int w = 0; w < RDP_NWORDS; w++) {
        float Pw = ((float)md->nw[w] + 0.5f) / Nf1;
        float lp = (float)log((double)Pw);
        md->logPrior[w] = lp;
        const int32_t *mw = md->m + (size_t)w * G;
        float *row = md->logP + (size_t)w * G;
        for (int g = 0; g < G; g++) {
            if (mw[g] > 0) {
                float q = ((float)mw[g] + Pw) / ((float)md->M[g] + 1.0f);
                row[g] = (float)log((double)q);
            } else {
                row[g] = lp - md->logLeave[g];      /* A4 default fill */
            }
        }
    }
}

/* parity hooks */
int          rdp_model_G(const rdp_model *md)        { return md->G; }
int64_t      rdp_model_N(const rdp_model *md)        { return md->N; }
const int32_t *rdp_model_m(const rdp_model *md)      { return md->m; }
const int32_t *rdp_model_nw(const rdp_model *md)     { return md->nw; }
const int32_t *rdp_model_M(const rdp_model *md)      { return md->M; }
const float *rdp_model_logPrior(const rdp_model *md) { return md->logPrior; }
const float *rdp_model_logLeave(const rdp_model *md) { return md->logLeave; }
const float *rdp_model_logP(const rdp_model *md)     { return md->logP; }

/* ------------------------------------------------------------------ A3 */
/* upstream TrainingInfo.isSeqReversed: one float accumulator over wordPairPriorDiffArr[w] = logWordPrior[w] -
 * logWordPrior[reverse complement of w], in word order; the query is reversed iff the sum is negative. */
int rdp_is_reversed(const rdp_model *md, const int32_t *words, int n)
{
    float prior = 0.0f;
    for (int i = 0; i < n; i++) {
        const float diff = md->logPrior[words[i]] - md->logPrior[rdp_revcomp_word(words[i])];
        prior += diff;
    }
    return prior < 0.0f;
}

static char comp_base(char c)
{
    switch (c) {
    case 'A': return 'T'; case 'a': return 't';
    case 'T': case 'U': return 'A'; case 't': case 'u': return 'a';
    case 'G': return 'C'; case 'g': return 'c';
    case 'C': return 'G'; case 'c': return 'g';
    default: return c;                       /* N / IUPAC: still breaks words */
    }
}

/* ------------------------------------------------------------------ A4,A7,A8 */
typedef struct {
    int32_t genus;                 /* A7 winner, -1 if short */
    int32_t n_words;
    float   score;                 /* A7 summed log posterior of the winner */
    int32_t reversed;
    int32_t status;                /* 0 ok, 1 ShortSequenceException (A2) */
    int32_t boot[RDP_NUM_OF_RUNS]; /* A8 winner of each replicate */
} rdp_result;

/* min_boot_words: 0 for 2.5 (k = n/8); later releases use max(n/8, 5). */
void rdp_classify(const rdp_model *md, const char *seq, int len, int min_boot_words,
                  rdp_result *out)
{
    int G = md->G;
    memset(out, 0, sizeof *out);
    out->genus = -1;
    if (len < RDP_MIN_SEQ_LEN) { out->status = 1; return; }

    int32_t *words = (int32_t *)malloc(sizeof(int32_t) * (size_t)len);
    char *rc = NULL;
    int n = rdp_words(seq, len, words);
    if (rdp_is_reversed(md, words, n)) {
        rc = (char *)malloc((size_t)len);
        for (int i = 0; i < len; i++) rc[i] = comp_base(seq[len - 1 - i]);
        n = rdp_words(rc, len, words);       /* getReversedSeq + recompute */
        out->reversed = 1;
    }
    out->n_words = n;

    /* A4: rows of the per-read matrix are rows of the dense table. */
    float *acc = (float *)malloc(sizeof(float) * (size_t)G);

    /* A7: sequential fp32 adds in word order, first strict max wins. */
    for (int g = 0; g < G; g++) acc[g] = 0.0f;
    for (int j = 0; j < n; j++) {
        const float *row = md->logP + (size_t)words[j] * G;
        for (int g = 0; g < G; g++) acc[g] += row[g];
    }
    float best = -INFINITY; int bi = 0;
    for (int g = 0; g < G; g++) if (acc[g] > best) { best = acc[g]; bi = g; }
    out->genus = bi; out->score = best;

    /* A8: bootstrap. */
    int k = n / RDP_WORDSIZE;
    if (k < min_boot_words) k = min_boot_words;
    jrandom rng; jr_set_seed(&rng, RDP_SEED);
    for (int run = 0; run < RDP_NUM_OF_RUNS; run++) {
        for (int g = 0; g < G; g++) acc[g] = 0.0f;
        for (int j = 0; j < k; j++) {
            int r = jr_next_int(&rng, n);
            const float *row = md->logP + (size_t)words[r] * G;
            for (int g = 0; g < G; g++) acc[g] += row[g];
        }
        best = -INFINITY; bi = 0;
        for (int g = 0; g < G; g++) if (acc[g] > best) { best = acc[g]; bi = g; }
        out->boot[run] = bi;
    }
    free(acc); free(words); free(rc);
}

/* reads are independent: OpenMP over reads for the timed CPU baseline. */
void rdp_classify_batch(const rdp_model *md, const char *bytes, const int64_t *off,
                        int64_t nreads, int min_boot_words, rdp_result *out, int nthreads)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int64_t i = 0; i < nreads; i++)
        rdp_classify(md, bytes + off[i], (int)(off[i + 1] - off[i]), min_boot_words, &out[i]);
}

void rdp_train_batch(rdp_model *md, const char *bytes, const int64_t *off, int64_t nseq,
                     const int32_t *genus)
{
    for (int64_t i = 0; i < nseq; i++)
        rdp_train_add(md, bytes + off[i], (int)(off[i + 1] - off[i]), genus[i]);
    rdp_train_finish(md);
}

/* A9: votes.  anc is [G][depth] node ids root-first (-1 padded); the winner
 * of a replicate and every ancestor get +1, so the vote of lineage position d
 * of the determined genus = #replicates whose winner shares that ancestor. */
void rdp_votes(const rdp_result *res, const int32_t *anc, int depth, int32_t *votes)
{
    for (int d = 0; d < depth; d++) votes[d] = 0;
    if (res->genus < 0) return;
    const int32_t *mine = anc + (size_t)res->genus * depth;
    for (int run = 0; run < RDP_NUM_OF_RUNS; run++) {
        const int32_t *his = anc + (size_t)res->boot[run] * depth;
        for (int d = 0; d < depth; d++)
            if (mine[d] >= 0 && mine[d] == his[d]) votes[d]++;
    }
}

int rdp_sizeof_result(void) { return (int)sizeof(rdp_result); }
