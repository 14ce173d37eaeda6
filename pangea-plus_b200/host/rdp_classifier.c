/*
 * rdp_classifier -- drop-in for the RDP stage of the PANGEA+ pipeline,
 *     java -Xmx1g -jar rdp_classifier-2.5.jar -q <in.fa> -o <out.txt>          (README.md:119-122)
 * Same -q/-o surface (plus -f allrank|fixrank|pangea and -t <model>), one output line per
 * classifiable read in input order, short reads reported on stdout and left out of the file
 * (SURVEY.md rows A0, A2, A9).  All arithmetic runs on the GPU through libpangea_b200.
 *
 * --gpus N [--devices a,b,..]: the reads shard across N GPUs.  Device 0 loads the model file, its
 * integer training counts are replicated once with ncclBroadcast (peer copies when a device is listed twice), every
 * device derives its tables locally, and the query file is cut into record-aligned pieces that the devices take in
 * turn; lines are written in input order.  The run is a pipeline: reader threads pread() pieces into pinned buffers,
 * per device an ingest context (highest stream priority) uploads and parses them while a classify context works on the
 * piece before, formatter threads turn records into text, and an in-order thread hands file offsets to a pwrite pool.
 *
 * The jar carries its training data inside; this tool cannot ship RDP's trainset, so the
 * model is a file: `-t model.pgm` (default: $PANGEA_RDP_MODEL, else ./rdp_model.pgm), made by
 *     rdp_classifier --train <training.fa> -t <model.pgm> [--ranks r0,r1,...]
 * Training FASTA headers carry the lineage the way RDP's trainer expects it:
 *     >seqid Root;Bacteria;Firmicutes;...;Genus          (blank or TAB before the lineage)
 * or, with --genus-token N, the genus is the N-th blank-separated token of the header (the
 * layout of the reference's validation_dataset/rdp_download_*.fa: N = 2) under a flat Root.
 *
 * Stock RDP trainset files (layouts per SURVEY.md appendix B, recalled from upstream and
 * UNVERIFIED against a real RDP 2.5 distribution -- no copy of it is reachable from the build environment):
 *     rdp_classifier --export-rdp <dir> -t model.pgm      writes rRNAClassifier.properties, bergeyTrainingTree.xml,
 *                                                          logWordPrior.txt, wordConditionalProbIndexArr.txt,
 *                                                          genus_wordConditionalProbList.txt
 *     rdp_classifier -q in.fa -o out.txt -t <dir>/rRNAClassifier.properties        classifies from such files
 * so that a user with Java can run `java -jar rdp_classifier-2.5.jar -t <dir>/rRNAClassifier.properties` on the same
 * trainset and diff the two outputs -- the one way to pin Stage A to the jar itself.
 * -g 16srrna|fungallsu picks the default model when -t is absent ($PANGEA_RDP_MODEL_16SRRNA / _FUNGALLSU).
 *
 * Output formats:
 *   db       id TAB trainset-no TAB taxid TAB conf, one line per rank of the assignment (UNVERIFIED layout)
 *   allrank  id TAB [-] (TAB name TAB rank TAB conf)*      the jar's default
 *   fixrank  id TAB [-] then domain, phylum, class, order, family, genus triples
 *   pangea   id, five TABs, then the triples below Root -- the layout that
 *            Consensus_BLAST_SOAP_RDP-1.1.pl:126 actually parses
 */
#include <getopt.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include <unistd.h>
#include <cuda_runtime_api.h>
#include <nccl.h>
#include "pangea_b200.h"
#include "pg_host_common.h"

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
    int    nnodes;
    int   *parent, *depth;
    char **name, **rank;
    int    G, maxdepth;
    int   *genus_node;
    int   *taxid;          /* NULL: a node's taxid is its index (models trained here) */
    int    trainset_no;
} taxonomy;

static const char *DEFAULT_RANKS[] = {"rootrank", "domain", "phylum", "class", "order", "family", "genus"};

static int tax_add(taxonomy *t, int parent, int depth, const char *name, size_t nlen, const char *rank)
{
    for (int i = 0; i < t->nnodes; i++)
        if (t->parent[i] == parent && strlen(t->name[i]) == nlen && memcmp(t->name[i], name, nlen) == 0) return i;
    int i = t->nnodes++;
    t->parent = (int *)realloc(t->parent, sizeof(int) * (size_t)t->nnodes);
    t->depth = (int *)realloc(t->depth, sizeof(int) * (size_t)t->nnodes);
    t->name = (char **)realloc(t->name, sizeof(char *) * (size_t)t->nnodes);
    t->rank = (char **)realloc(t->rank, sizeof(char *) * (size_t)t->nnodes);
    t->parent[i] = parent;
    t->depth[i] = depth;
    t->name[i] = (char *)malloc(nlen + 1);
    memcpy(t->name[i], name, nlen);
    t->name[i][nlen] = 0;
    t->rank[i] = strdup(rank);
    if (depth + 1 > t->maxdepth) t->maxdepth = depth + 1;
    return i;
}

/* blob: "PGTAX1\n" nnodes G\n, nnodes lines "parent\tdepth\trank\tname\n", G lines "node\n" */
static char *tax_to_blob(const taxonomy *t, int64_t *len)
{
    size_t cap = 64 + (size_t)t->nnodes * 160 + (size_t)t->G * 12, n = 0;
    char *b = (char *)malloc(cap);
    n += (size_t)snprintf(b + n, cap - n, "PGTAX1\n%d %d\n", t->nnodes, t->G);
    for (int i = 0; i < t->nnodes; i++) {
        if (cap - n < strlen(t->name[i]) + strlen(t->rank[i]) + 64) { cap = cap * 2 + 1024; b = (char *)realloc(b, cap); }
        n += (size_t)snprintf(b + n, cap - n, "%d\t%d\t%s\t%s\n", t->parent[i], t->depth[i], t->rank[i], t->name[i]);
    }
    for (int g = 0; g < t->G; g++) {
        if (cap - n < 32) { cap = cap * 2 + 1024; b = (char *)realloc(b, cap); }
        n += (size_t)snprintf(b + n, cap - n, "%d\n", t->genus_node[g]);
    }
    *len = (int64_t)n;
    return b;
}

static int tax_from_blob(const char *b, int64_t len, taxonomy *t)
{
    memset(t, 0, sizeof *t);
    if (len < 8 || memcmp(b, "PGTAX1\n", 7) != 0) return -1;
    const char *p = b + 7, *end = b + len;
    int nn = 0, G = 0;
    if (sscanf(p, "%d %d", &nn, &G) != 2 || nn <= 0 || G <= 0) return -1;
    p = (const char *)memchr(p, '\n', (size_t)(end - p));
    if (!p) return -1;
    p++;
    t->nnodes = nn;
    t->G = G;
    t->parent = (int *)malloc(sizeof(int) * (size_t)nn);
    t->depth = (int *)malloc(sizeof(int) * (size_t)nn);
    t->name = (char **)malloc(sizeof(char *) * (size_t)nn);
    t->rank = (char **)malloc(sizeof(char *) * (size_t)nn);
    t->genus_node = (int *)malloc(sizeof(int) * (size_t)G);
    for (int i = 0; i < nn; i++) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        if (!nl) return -1;
        const char *t1 = (const char *)memchr(p, '\t', (size_t)(nl - p));
        const char *t2 = t1 ? (const char *)memchr(t1 + 1, '\t', (size_t)(nl - t1 - 1)) : NULL;
        const char *t3 = t2 ? (const char *)memchr(t2 + 1, '\t', (size_t)(nl - t2 - 1)) : NULL;
        if (!t3) return -1;
        t->parent[i] = atoi(p);
        t->depth[i] = atoi(t1 + 1);
        t->rank[i] = strndup(t2 + 1, (size_t)(t3 - t2 - 1));
        t->name[i] = strndup(t3 + 1, (size_t)(nl - t3 - 1));
        if (t->depth[i] + 1 > t->maxdepth) t->maxdepth = t->depth[i] + 1;
        p = nl + 1;
    }
    for (int g = 0; g < G; g++) {
        if (p >= end) return -1;
        t->genus_node[g] = atoi(p);
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        if (!nl) return -1;
        p = nl + 1;
    }
    return 0;
}

static int32_t *tax_lineage_table(const taxonomy *t)          /* anc[g*maxdepth + d], root first, -1 padded */
{
    int D = t->maxdepth;
    int32_t *anc = (int32_t *)malloc(sizeof(int32_t) * (size_t)t->G * (size_t)D);
    for (int g = 0; g < t->G; g++) {
        for (int d = 0; d < D; d++) anc[(size_t)g * D + d] = -1;
        for (int n = t->genus_node[g]; n >= 0; n = t->parent[n]) anc[(size_t)g * D + t->depth[n]] = n;
    }
    return anc;
}

static int do_train(pg_ctx *ctx, const char *fasta_path, const char *model_path, const char *ranks_csv, int genus_token)
{
    pg_fasta fa;
    if (pg_fasta_read(fasta_path, &fa)) { fprintf(stderr, "rdp_classifier: cannot read %s\n", fasta_path); return 1; }
    const char *ranks[64];
    int nranks = 0;
    char *rc = ranks_csv ? strdup(ranks_csv) : NULL;
    if (rc) {
        for (char *tok = strtok(rc, ","); tok && nranks < 64; tok = strtok(NULL, ",")) ranks[nranks++] = tok;
    } else {
        for (; nranks < 7; nranks++) ranks[nranks] = DEFAULT_RANKS[nranks];
    }
    taxonomy t;
    memset(&t, 0, sizeof t);
    int32_t *genus = (int32_t *)malloc(sizeof(int32_t) * (size_t)(fa.count + 1));
    for (int64_t i = 0; i < fa.count; i++) {
        const char *h = fa.header[i];
        int node = -1, depth = 0;
        if (genus_token > 0) {
            const char *p = h;
            for (int k = 1; k < genus_token; k++) {
                while (*p && *p != ' ' && *p != '\t') p++;
                while (*p == ' ' || *p == '\t') p++;
            }
            const char *e = p;
            while (*e && *e != ' ' && *e != '\t') e++;
            int root = tax_add(&t, -1, 0, "Root", 4, "rootrank");
            node = tax_add(&t, root, 1, p, (size_t)(e - p), "genus");
        } else {
            const char *p = h;
            while (*p && *p != ' ' && *p != '\t') p++;
            while (*p == ' ' || *p == '\t') p++;
            if (!*p) { fprintf(stderr, "rdp_classifier: header of %s carries no lineage\n", fa.id[i]); return 1; }
            int parent = -1;
            while (*p) {
                const char *e = strchr(p, ';');
                size_t n = e ? (size_t)(e - p) : strlen(p);
                while (n && (p[n - 1] == ' ' || p[n - 1] == '\t')) n--;
                if (n) {
                    char rk[32];
                    if (depth < nranks) snprintf(rk, sizeof rk, "%s", ranks[depth]);
                    else snprintf(rk, sizeof rk, "rank%d", depth);
                    parent = node = tax_add(&t, parent, depth, p, n, rk);
                    depth++;
                }
                if (!e) break;
                p = e + 1;
            }
        }
        if (t.maxdepth > PG_MAX_DEPTH) { fprintf(stderr, "rdp_classifier: lineage of %s is deeper than %d\n", fa.id[i], PG_MAX_DEPTH); return 1; }
        /* genus index = order of first appearance in the training FASTA */
        int g = -1;
        for (int k = 0; k < t.G; k++)
            if (t.genus_node[k] == node) { g = k; break; }
        if (g < 0) {
            t.genus_node = (int *)realloc(t.genus_node, sizeof(int) * (size_t)(t.G + 1));
            g = t.G;
            t.genus_node[t.G++] = node;
        }
        genus[i] = g;
    }
    pg_seqbatch sb = {fa.bytes, fa.off, fa.count};
    pg_model *md = NULL;
    if (pg_train(ctx, &sb, genus, t.G, &md) != PG_OK) { fprintf(stderr, "rdp_classifier: %s\n", pg_last_error(ctx)); return 1; }
    int32_t *anc = tax_lineage_table(&t);
    if (pg_model_set_lineage(md, anc, t.maxdepth) != PG_OK) { fprintf(stderr, "rdp_classifier: %s\n", pg_last_error(ctx)); return 1; }
    int64_t blen = 0;
    char *blob = tax_to_blob(&t, &blen);
    if (pg_model_save(md, model_path, blob, blen) != PG_OK) { fprintf(stderr, "rdp_classifier: %s\n", pg_last_error(ctx)); return 1; }
    printf("trained %lld sequences, %d genera, %d taxa -> %s\n", (long long)fa.count, t.G, t.nnodes, model_path);
    pg_model_free(md);
    pg_fasta_free(&fa);
    return 0;
}

/* ------------------------------------------------------------------ stock RDP trainset files (UNVERIFIED layouts) */

static const char *RDP_VERSION = "RDP Naive Bayesian rRNA Classifier Version 2.5, May 2012";

static void xml_escape(FILE *f, const char *s)
{
    for (; *s; s++) {
        if (*s == '&') fputs("&amp;", f);
        else if (*s == '<') fputs("&lt;", f);
        else if (*s == '>') fputs("&gt;", f);
        else if (*s == '"') fputs("&quot;", f);
        else fputc(*s, f);
    }
}

static void xml_unescape(char *s)
{
    char *o = s;
    while (*s) {
        if (!strncmp(s, "&amp;", 5)) { *o++ = '&'; s += 5; }
        else if (!strncmp(s, "&lt;", 4)) { *o++ = '<'; s += 4; }
        else if (!strncmp(s, "&gt;", 4)) { *o++ = '>'; s += 4; }
        else if (!strncmp(s, "&quot;", 6)) { *o++ = '"'; s += 6; }
        else if (!strncmp(s, "&apos;", 6)) { *o++ = '\''; s += 6; }
        else *o++ = *s++;
    }
    *o = 0;
}

/* model.pgm -> the five files of an RDP trainset directory.  Floats are written with nine significant digits, which
 * is enough to read every fp32 value back exactly. */
static int export_rdp(pg_ctx *ctx, const pg_model *md, const taxonomy *t, const char *dir)
{
    const int G = t->G;
    float *lp = (float *)malloc(sizeof(float) * 65536), *ll = (float *)malloc(sizeof(float) * (size_t)G);
    float *logp = (float *)malloc(sizeof(float) * 65536 * (size_t)G);
    int32_t *m = (int32_t *)malloc(sizeof(int32_t) * 65536 * (size_t)G), *M = (int32_t *)malloc(sizeof(int32_t) * (size_t)G);
    if (!lp || !ll || !logp || !m || !M) { fprintf(stderr, "rdp_classifier: out of memory\n"); return 1; }
    if (pg_model_tables(md, lp, ll, logp) != PG_OK || pg_model_counts(md, m, NULL, M, NULL) != PG_OK) {
        fprintf(stderr, "rdp_classifier: %s\n", pg_last_error(ctx));
        return 1;
    }
    char path[4096], hdr[256];
#define RDP_OPEN(name) snprintf(path, sizeof path, "%s/%s", dir, name); f = fopen(path, "w"); \
    if (!f) { fprintf(stderr, "rdp_classifier: cannot write %s\n", path); return 1; }
#define RDP_HDR(file) snprintf(hdr, sizeof hdr, "<trainsetNo>%d</trainsetNo><version>pangea_b200</version><modversion>1</modversion><file>%s</file>", t->trainset_no, file)
    FILE *f;
    RDP_OPEN("rRNAClassifier.properties")
    fprintf(f, "# written by pangea_b200 rdp_classifier --export-rdp (layouts per SURVEY.md appendix B, UNVERIFIED)\n"
               "bergeyTree=bergeyTrainingTree.xml\nprobabilityList=genus_wordConditionalProbList.txt\n"
               "probabilityIndex=wordConditionalProbIndexArr.txt\nwordPrior=logWordPrior.txt\nclassifierVersion=%s\n", RDP_VERSION);
    fclose(f);
    /* the tree: leaveCount = training sequences under the node */
    long long *leave = (long long *)calloc((size_t)t->nnodes, sizeof(long long));
    int *gidx = (int *)malloc(sizeof(int) * (size_t)t->nnodes);
    for (int k = 0; k < t->nnodes; k++) gidx[k] = -1;
    for (int g = 0; g < G; g++) {
        gidx[t->genus_node[g]] = g;
        for (int n = t->genus_node[g]; n >= 0; n = t->parent[n]) leave[n] += M[g];
    }
    RDP_OPEN("bergeyTrainingTree.xml")
    RDP_HDR("bergeyTrainingTree");
    fprintf(f, "%s\n", hdr);
    for (int k = 0; k < t->nnodes; k++) {
        fputs("<TreeNode name=\"", f);
        xml_escape(f, t->name[k]);
        fprintf(f, "\" taxid=\"%d\" rank=\"%s\" parentTaxid=\"%d\" leaveCount=\"%lld\" genusIndex=\"%d\"></TreeNode>\n",
                t->taxid ? t->taxid[k] : k, t->rank[k], t->parent[k] < 0 ? -1 : (t->taxid ? t->taxid[t->parent[k]] : t->parent[k]),
                leave[k], gidx[k]);
    }
    fclose(f);
    RDP_OPEN("logWordPrior.txt")
    RDP_HDR("logWordPrior");
    fprintf(f, "%s\n", hdr);
    for (int w = 0; w < 65536; w++) fprintf(f, "%d\t%.9g\n", w, (double)lp[w]);
    fclose(f);
    RDP_OPEN("wordConditionalProbIndexArr.txt")
    RDP_HDR("wordConditionalProbIndexArr");
    fprintf(f, "%s\n", hdr);
    long long at = 0;
    for (int w = 0; w < 65536; w++) {
        fprintf(f, "%d\t%lld\n", w, at);
        for (int g = 0; g < G; g++) at += m[(size_t)w * G + g] > 0;
    }
    fprintf(f, "%d\t%lld\n", 65536, at);
    fclose(f);
    RDP_OPEN("genus_wordConditionalProbList.txt")
    RDP_HDR("genus_wordConditionalProbList");
    fprintf(f, "%s\n", hdr);
    for (int w = 0; w < 65536; w++)
        for (int g = 0; g < G; g++)
            if (m[(size_t)w * G + g] > 0) fprintf(f, "%d\t%.9g\n", g, (double)logp[(size_t)w * G + g]);
    fclose(f);
#undef RDP_OPEN
#undef RDP_HDR
    printf("exported %d genera, %d taxa, %lld conditional probabilities -> %s/rRNAClassifier.properties\n", G, t->nnodes, at, dir);
    free(lp); free(ll); free(logp); free(m); free(M); free(leave); free(gidx);
    return 0;
}

static char *xml_attr(const char *line, const char *key, char *buf, size_t cap)
{
    char pat[64];
    snprintf(pat, sizeof pat, " %s=\"", key);
    const char *p = strstr(line, pat);
    if (!p) return NULL;
    p += strlen(pat);
    const char *e = strchr(p, '"');
    if (!e || (size_t)(e - p) >= cap) return NULL;
    memcpy(buf, p, (size_t)(e - p));
    buf[e - p] = 0;
    xml_unescape(buf);
    return buf;
}

/* pg_lines keeps the newline bytes in place: make every line a C string */
static int read_lines_z(const char *path, pg_lines *pl)
{
    if (pg_lines_read(path, pl)) return -1;
    for (int64_t i = 0; i < pl->count; i++) {
        size_t n = pl->len[i];
        if (n && pl->line[i][n - 1] == '\r') n--;
        pl->line[i][n] = 0;
    }
    return 0;
}

/* rRNAClassifier.properties -> model on the device + taxonomy.  Paths in the properties file are relative to it. */
static int import_rdp(pg_ctx *ctx, const char *props, pg_model **out, taxonomy *t)
{
    memset(t, 0, sizeof *t);
    pg_lines pl;
    if (read_lines_z(props, &pl)) { fprintf(stderr, "rdp_classifier: cannot read %s\n", props); return 1; }
    char dir[4096], tree[256] = "", plist[256] = "", pindex[256] = "", prior[256] = "";
    snprintf(dir, sizeof dir, "%s", props);
    char *slash = strrchr(dir, '/');
    if (slash) *slash = 0; else strcpy(dir, ".");
    for (int64_t i = 0; i < pl.count; i++) {
        char *l = pl.line[i];
        while (*l == ' ' || *l == '\t') l++;
        if (*l == '#' || *l == '!') continue;
        char *eq = strpbrk(l, "=:");
        if (!eq) continue;
        *eq = 0;
        char *v = eq + 1;
        while (*v == ' ' || *v == '\t') v++;
        size_t vl = strlen(v);
        while (vl && (v[vl - 1] == '\r' || v[vl - 1] == ' ' || v[vl - 1] == '\t')) v[--vl] = 0;
        size_t kl = strlen(l);
        while (kl && (l[kl - 1] == ' ' || l[kl - 1] == '\t')) l[--kl] = 0;
        if (!strcmp(l, "bergeyTree")) snprintf(tree, sizeof tree, "%s", v);
        else if (!strcmp(l, "probabilityList")) snprintf(plist, sizeof plist, "%s", v);
        else if (!strcmp(l, "probabilityIndex")) snprintf(pindex, sizeof pindex, "%s", v);
        else if (!strcmp(l, "wordPrior")) snprintf(prior, sizeof prior, "%s", v);
    }
    pg_lines_free(&pl);
    if (!tree[0] || !plist[0] || !pindex[0] || !prior[0]) {
        fprintf(stderr, "rdp_classifier: %s lacks one of bergeyTree / probabilityList / probabilityIndex / wordPrior\n", props);
        return 1;
    }
    char path[4400];
    /* ---- the tree */
    snprintf(path, sizeof path, "%s/%s", dir, tree);
    if (read_lines_z(path, &pl)) { fprintf(stderr, "rdp_classifier: cannot read %s\n", path); return 1; }
    int cap = (int)pl.count + 1, maxg = -1;
    t->parent = (int *)malloc(sizeof(int) * (size_t)cap);
    t->depth = (int *)malloc(sizeof(int) * (size_t)cap);
    t->taxid = (int *)malloc(sizeof(int) * (size_t)cap);
    t->name = (char **)malloc(sizeof(char *) * (size_t)cap);
    t->rank = (char **)malloc(sizeof(char *) * (size_t)cap);
    int *ptax = (int *)malloc(sizeof(int) * (size_t)cap), *gidx = (int *)malloc(sizeof(int) * (size_t)cap);
    long long *leave = (long long *)malloc(sizeof(long long) * (size_t)cap);
    for (int64_t i = 0; i < pl.count; i++) {
        const char *l = pl.line[i];
        char b[1024];
        if (i == 0) {
            const char *q = strstr(l, "<trainsetNo>");
            if (q) t->trainset_no = atoi(q + 12);
        }
        const char *tn = strstr(l, "<TreeNode");
        if (!tn) continue;
        const int k = t->nnodes++;
        t->name[k] = strdup(xml_attr(tn, "name", b, sizeof b) ? b : "");
        t->taxid[k] = xml_attr(tn, "taxid", b, sizeof b) ? atoi(b) : k;
        t->rank[k] = strdup(xml_attr(tn, "rank", b, sizeof b) ? b : "");
        ptax[k] = xml_attr(tn, "parentTaxid", b, sizeof b) ? atoi(b) : -1;
        leave[k] = xml_attr(tn, "leaveCount", b, sizeof b) ? atoll(b) : 0;
        gidx[k] = xml_attr(tn, "genusIndex", b, sizeof b) ? atoi(b) : -1;
        if (gidx[k] > maxg) maxg = gidx[k];
    }
    pg_lines_free(&pl);
    if (t->nnodes == 0 || maxg < 0) { fprintf(stderr, "rdp_classifier: %s holds no genus node\n", path); return 1; }
    for (int k = 0; k < t->nnodes; k++) {                   /* parent taxid -> node index (the root's parent is absent) */
        t->parent[k] = -1;
        if (ptax[k] != t->taxid[k])
            for (int j = 0; j < t->nnodes; j++)
                if (t->taxid[j] == ptax[k] && j != k) { t->parent[k] = j; break; }
    }
    for (int k = 0; k < t->nnodes; k++) {
        int d = 0;
        for (int n = t->parent[k]; n >= 0 && d <= t->nnodes; n = t->parent[n]) d++;
        t->depth[k] = d;
        if (d + 1 > t->maxdepth) t->maxdepth = d + 1;
    }
    if (t->maxdepth > PG_MAX_DEPTH) { fprintf(stderr, "rdp_classifier: the taxonomy of %s is deeper than %d\n", path, PG_MAX_DEPTH); return 1; }
    const int G = t->G = maxg + 1;
    t->genus_node = (int *)malloc(sizeof(int) * (size_t)G);
    int32_t *ll = (int32_t *)malloc(sizeof(int32_t) * (size_t)G);
    for (int g = 0; g < G; g++) t->genus_node[g] = -1;
    for (int k = 0; k < t->nnodes; k++)
        if (gidx[k] >= 0) { t->genus_node[gidx[k]] = k; ll[gidx[k]] = (int32_t)leave[k]; }
    for (int g = 0; g < G; g++)
        if (t->genus_node[g] < 0) { fprintf(stderr, "rdp_classifier: genusIndex %d is missing from %s\n", g, path); return 1; }
    /* ---- word priors */
    float *lp = (float *)malloc(sizeof(float) * 65536);
    for (int w = 0; w < 65536; w++) lp[w] = 0.f;
    snprintf(path, sizeof path, "%s/%s", dir, prior);
    if (read_lines_z(path, &pl)) { fprintf(stderr, "rdp_classifier: cannot read %s\n", path); return 1; }
    for (int64_t i = 1; i < pl.count; i++) {
        char *e;
        long w = strtol(pl.line[i], &e, 10);
        if (e == pl.line[i] || w < 0 || w > 65535) continue;
        lp[w] = strtof(e, NULL);
    }
    pg_lines_free(&pl);
    /* ---- index and list */
    int64_t *idx = (int64_t *)calloc(65537, sizeof(int64_t));
    snprintf(path, sizeof path, "%s/%s", dir, pindex);
    if (read_lines_z(path, &pl)) { fprintf(stderr, "rdp_classifier: cannot read %s\n", path); return 1; }
    int seen_end = 0;
    for (int64_t i = 1; i < pl.count; i++) {
        char *e;
        long w = strtol(pl.line[i], &e, 10);
        if (e == pl.line[i] || w < 0 || w > 65536) continue;
        idx[w] = strtoll(e, NULL, 10);
        if (w == 65536) seen_end = 1;
    }
    pg_lines_free(&pl);
    snprintf(path, sizeof path, "%s/%s", dir, plist);
    if (read_lines_z(path, &pl)) { fprintf(stderr, "rdp_classifier: cannot read %s\n", path); return 1; }
    const int64_t nnz = pl.count > 0 ? pl.count - 1 : 0;
    if (!seen_end) idx[65536] = nnz;
    int32_t *eg = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz + 1));
    float *ep = (float *)malloc(sizeof(float) * (size_t)(nnz + 1));
    for (int64_t i = 0; i < nnz; i++) {
        char *e;
        eg[i] = (int32_t)strtol(pl.line[i + 1], &e, 10);
        ep[i] = strtof(e, NULL);
    }
    pg_lines_free(&pl);
    if (idx[65536] != nnz) { fprintf(stderr, "rdp_classifier: %s lists %lld entries, the index says %lld\n", path, (long long)nnz, (long long)idx[65536]); return 1; }
    if (pg_model_from_tables(ctx, G, lp, ll, idx, eg, ep, out) != PG_OK) { fprintf(stderr, "rdp_classifier: %s\n", pg_last_error(ctx)); return 1; }
    int32_t *anc = tax_lineage_table(t);
    if (pg_model_set_lineage(*out, anc, t->maxdepth) != PG_OK) { fprintf(stderr, "rdp_classifier: %s\n", pg_last_error(ctx)); return 1; }
    free(anc); free(ptax); free(gidx); free(leave); free(ll); free(lp); free(idx); free(eg); free(ep);
    return 0;
}

static int is_properties(const char *path)
{
    const size_t n = strlen(path);
    return n > 11 && strcmp(path + n - 11, ".properties") == 0;
}

/* ------------------------------------------------------------------ pipeline plumbing */

typedef struct {
    int64_t    seq;                 /* position of the piece in the file */
    char      *text;                /* pinned host buffer */
    int64_t    len, textcap;
    int64_t    nrec, cap;
    int64_t   *hdr_off;
    int32_t   *id_len;
    pg_result *res;                 /* pinned */
    pg_reads  *reads;               /* packed reads of the piece, between the ingest and the classify stage */
    int64_t    rescap;
    int64_t    out_at;              /* where the piece's lines go in the output file (set by the in-order writer) */
    char      *out, *msg;           /* formatted lines / stdout messages */
    size_t     out_len, out_cap, msg_len, msg_cap;
} piece_t;

typedef struct {
    piece_t       **item;
    int             cap, head, count, closed;
    pthread_mutex_t mu;
    pthread_cond_t  cv;
} queue_t;

static void q_init(queue_t *q, int cap)
{
    memset(q, 0, sizeof *q);
    q->item = (piece_t **)calloc((size_t)cap, sizeof(piece_t *));
    q->cap = cap;
    pthread_mutex_init(&q->mu, NULL);
    pthread_cond_init(&q->cv, NULL);
}
static void q_push(queue_t *q, piece_t *p)
{
    pthread_mutex_lock(&q->mu);
    q->item[(q->head + q->count) % q->cap] = p;          /* never more pieces in flight than cap */
    q->count++;
    pthread_cond_broadcast(&q->cv);
    pthread_mutex_unlock(&q->mu);
}
static piece_t *q_pop(queue_t *q)                        /* NULL once the queue is closed and drained */
{
    pthread_mutex_lock(&q->mu);
    while (q->count == 0 && !q->closed) pthread_cond_wait(&q->cv, &q->mu);
    piece_t *p = NULL;
    if (q->count) {
        p = q->item[q->head];
        q->head = (q->head + 1) % q->cap;
        q->count--;
    }
    pthread_mutex_unlock(&q->mu);
    return p;
}
static void q_close(queue_t *q)
{
    pthread_mutex_lock(&q->mu);
    q->closed = 1;
    pthread_cond_broadcast(&q->cv);
    pthread_mutex_unlock(&q->mu);
}

typedef struct {
    /* configuration */
    const taxonomy *t;
    int             ifmt;
    pg_classify_opts opts;
    int64_t         piece_bytes;
    FILE           *fq, *fo;
    /* queues */
    queue_t         q_free, q_gpu, q_fmt, q_out, q_write;
    queue_t        *q_mid;          /* per GPU: ingested pieces waiting for the classify context (split mode) */
    int             ngpu, split, ingest_alive;
    pg_ctx        **ictx;           /* per GPU: the ingest context (its stream has the highest priority) */
    int             out_fd, nwriters, write_alive;
    int             npieces;
    /* workers */
    int             nworkers, nformat;
    pg_ctx        **wctx;           /* one context per GPU worker */
    pg_model      **wmodel;         /* the model of the worker's device (shared by the contexts of one device) */
    pthread_mutex_t mu;
    int             gpu_alive, fmt_alive, read_alive, failed;
    int64_t         fsize, next_piece, total_pieces, textcap0;
    char            err[512];
    /* formatter tables */
    char            conf_tab[101][8];
    char          **tpiece;
    size_t         *tpiece_len;
    /* statistics */
    int64_t         reads_total;
    double          busy_read, busy_gpu, busy_ingest, busy_fmt, busy_write;
} pipeline_t;

static void pl_fail(pipeline_t *pl, const char *msg)
{
    pthread_mutex_lock(&pl->mu);
    if (!pl->failed) { pl->failed = 1; snprintf(pl->err, sizeof pl->err, "%s", msg); }
    pthread_mutex_unlock(&pl->mu);
}

static void *pinned_alloc(size_t bytes)
{
    void *p = NULL;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) return NULL;
    return p;
}

/* Readers.  Piece k of the file is the records that START in [k*P, (k+1)*P): its first byte is the first '>' at a line
 * start at or after k*P (byte 0 for k = 0, so text before the first header stays with piece 0), its end is the start
 * of piece k+1.  Every piece is therefore defined by the file alone and several threads can read pieces at once
 * (pread); a thread takes a free buffer FIRST and the next piece index second, so the pieces in flight are always the
 * lowest unwritten ones and the in-order writer never waits for a piece that has no buffer. */
static int64_t record_start_at_or_after(int fd, int64_t pos, int64_t fsize)
{
    if (pos <= 0) return 0;
    char win[65536 + 1];
    int64_t at = pos - 1;                                   /* the byte before pos may be the '\n' of a "\n>" at pos */
    while (at < fsize) {
        const ssize_t got = pread(fd, win, sizeof win - 1, (off_t)at);
        if (got <= 1) break;
        for (ssize_t i = 0; i + 1 < got; i++)
            if (win[i] == '\n' && win[i + 1] == '>') return at + i + 1;
        at += got - 1;                                      /* keep the last byte: it may be the '\n' */
    }
    return fsize;
}

static void *reader_main(void *arg)
{
    pipeline_t *pl = (pipeline_t *)arg;
    const int fd = fileno(pl->fq);
    for (;;) {
        piece_t *p = q_pop(&pl->q_free);
        if (!p) break;
        pthread_mutex_lock(&pl->mu);
        const int64_t k = pl->next_piece < pl->total_pieces ? pl->next_piece++ : -1;
        pthread_mutex_unlock(&pl->mu);
        if (k < 0 || pl->failed) { q_push(&pl->q_free, p); break; }
        const double t0 = now_s();
        const int64_t a = record_start_at_or_after(fd, k * pl->piece_bytes, pl->fsize);
        const int64_t e = k + 1 >= pl->total_pieces ? pl->fsize : record_start_at_or_after(fd, (k + 1) * pl->piece_bytes, pl->fsize);
        const int64_t len = e > a ? e - a : 0;
        if (!p->text || len > p->textcap) {                 /* buffers are pinned on first use (and when a piece outgrows them) */
            if (p->text) cudaFreeHost(p->text);
            p->textcap = len > pl->textcap0 ? len + len / 8 : pl->textcap0;
            p->text = (char *)pinned_alloc((size_t)p->textcap + 1);
            if (!p->text) { pl_fail(pl, "pinned allocation failed"); p->textcap = 0; }
        }
        int64_t have = 0;
        while (p->text && have < len) {
            const ssize_t got = pread(fd, p->text + have, (size_t)(len - have), (off_t)(a + have));
            if (got <= 0) { pl_fail(pl, "read failed"); break; }
            have += got;
        }
        p->len = p->text ? have : 0;
        p->seq = k;
        pthread_mutex_lock(&pl->mu);
        pl->busy_read += now_s() - t0;
        pthread_mutex_unlock(&pl->mu);
        q_push(&pl->q_gpu, p);
    }
    pthread_mutex_lock(&pl->mu);
    const int last = --pl->read_alive == 0;
    pthread_mutex_unlock(&pl->mu);
    if (last) q_close(&pl->q_gpu);
    return NULL;
}

typedef struct { pipeline_t *pl; int idx; } worker_arg;

/* FASTA ingest of one piece (text -> records, ids, packed reads on the device) */
static int ingest_piece(pg_ctx *ctx, piece_t *p)
{
    if (p->cap == 0) p->cap = p->len / 32 + 1024;
    p->reads = NULL;
    for (;;) {
        if (!p->hdr_off) {
            p->hdr_off = (int64_t *)malloc(sizeof(int64_t) * (size_t)p->cap);
            p->id_len = (int32_t *)malloc(sizeof(int32_t) * (size_t)p->cap);
        }
        const int rc = pg_fasta_ingest(ctx, p->text, p->len, p->cap, &p->nrec, p->hdr_off, p->id_len, NULL, &p->reads);
        if (rc == PG_ERANGE && p->nrec > p->cap) {
            p->cap = p->nrec;
            free(p->hdr_off); free(p->id_len);
            p->hdr_off = NULL; p->id_len = NULL;
            continue;
        }
        return rc;
    }
}

static int classify_piece(pipeline_t *pl, pg_ctx *ctx, const pg_model *md, piece_t *p)
{
    int rc = PG_OK;
    if (p->nrec + 1 > p->rescap) {
        if (p->res) cudaFreeHost(p->res);
        p->rescap = p->nrec + p->nrec / 8 + 1024;
        p->res = (pg_result *)pinned_alloc(sizeof(pg_result) * (size_t)p->rescap);
        if (!p->res) { pl_fail(pl, "pinned allocation failed"); return PG_ENOMEM; }
    }
    if (p->nrec > 0) rc = pg_classify_packed_host(ctx, md, p->reads, &pl->opts, p->res, NULL);
    return rc;
}

/* GPU worker, one context for both stages (--contexts-per-gpu 1): FASTA ingest, then classification, one piece at a time */
static void *gpu_main(void *arg)
{
    worker_arg *wa = (worker_arg *)arg;
    pipeline_t *pl = wa->pl;
    pg_ctx *ctx = pl->wctx[wa->idx];
    const pg_model *md = pl->wmodel[wa->idx];
    for (;;) {
        piece_t *p = q_pop(&pl->q_gpu);
        if (!p) break;
        if (pl->failed || p->len == 0) { p->nrec = 0; q_push(&pl->q_fmt, p); continue; }
        double t0 = now_s();
        int rc = ingest_piece(ctx, p);
        const double t_ing = now_s() - t0;
        if (rc == PG_OK) rc = classify_piece(pl, ctx, md, p);
        if (p->reads) { pg_reads_free(p->reads); p->reads = NULL; }
        if (rc != PG_OK) { pl_fail(pl, pg_last_error(ctx)); p->nrec = 0; }
        pthread_mutex_lock(&pl->mu);
        pl->busy_gpu += now_s() - t0;
        pl->busy_ingest += t_ing;
        pl->reads_total += p->nrec;
        pthread_mutex_unlock(&pl->mu);
        q_push(&pl->q_fmt, p);
    }
    pthread_mutex_lock(&pl->mu);
    int last = --pl->gpu_alive == 0;
    pthread_mutex_unlock(&pl->mu);
    if (last) q_close(&pl->q_fmt);
    return NULL;
}

/* Split mode (the default): two threads and two contexts per GPU.  The ingest context runs on a stream of the highest
 * priority, so its short kernels and copies slip in between the thread blocks of the classify context's long kernels:
 * the upload and parsing of piece i+1 hide under the classification of piece i, and neither stage waits behind the
 * other's queue (two equal contexts that each do both stages were measured: 3-11 M reads/s, depending on how their
 * synchronisation points happened to interleave). */
static void *ingest_main(void *arg)
{
    worker_arg *wa = (worker_arg *)arg;
    pipeline_t *pl = wa->pl;
    const int g = wa->idx;
    pg_ctx *ctx = pl->ictx[g];
    for (;;) {
        piece_t *p = q_pop(&pl->q_gpu);
        if (!p) break;
        if (pl->failed || p->len == 0) { p->nrec = 0; p->reads = NULL; q_push(&pl->q_mid[g], p); continue; }
        const double t0 = now_s();
        const int rc = ingest_piece(ctx, p);
        if (rc != PG_OK) { pl_fail(pl, pg_last_error(ctx)); p->nrec = 0; }
        pthread_mutex_lock(&pl->mu);
        pl->busy_ingest += now_s() - t0;
        pthread_mutex_unlock(&pl->mu);
        q_push(&pl->q_mid[g], p);
    }
    q_close(&pl->q_mid[g]);
    return NULL;
}

static void *classify_main(void *arg)
{
    worker_arg *wa = (worker_arg *)arg;
    pipeline_t *pl = wa->pl;
    const int g = wa->idx;
    pg_ctx *ctx = pl->wctx[g];
    const pg_model *md = pl->wmodel[g];
    for (;;) {
        piece_t *p = q_pop(&pl->q_mid[g]);
        if (!p) break;
        const double t0 = now_s();
        if (!pl->failed && p->nrec > 0) {
            const int rc = classify_piece(pl, ctx, md, p);
            if (rc != PG_OK) { pl_fail(pl, pg_last_error(ctx)); p->nrec = 0; }
        }
        if (p->reads) { pg_reads_free(p->reads); p->reads = NULL; }
        pthread_mutex_lock(&pl->mu);
        pl->busy_gpu += now_s() - t0;
        pl->reads_total += p->nrec;
        pthread_mutex_unlock(&pl->mu);
        q_push(&pl->q_fmt, p);
    }
    pthread_mutex_lock(&pl->mu);
    int last = --pl->gpu_alive == 0;
    pthread_mutex_unlock(&pl->mu);
    if (last) q_close(&pl->q_fmt);
    return NULL;
}

static void grow(char **buf, size_t *cap, size_t need)
{
    if (need <= *cap) return;
    size_t n = *cap ? *cap : (size_t)1 << 20;
    while (n < need) n *= 2;
    *buf = (char *)realloc(*buf, n);
    *cap = n;
}

/* formatter: records of one piece -> output lines (and the ShortSequenceException messages for stdout).  Lines are
 * assembled from pieces prepared once per taxon ("\tname\trank\t") and once per vote count, a few memcpy()s each. */
static void format_piece(pipeline_t *pl, piece_t *p)
{
    static const char *FIX[6] = {"domain", "phylum", "class", "order", "family", "genus"};
    const taxonomy *t = pl->t;
    const int ifmt = pl->ifmt;
    size_t on = 0, mn = 0;
#define OUT_STR(s, n) do { memcpy(p->out + on, (s), (n)); on += (n); } while (0)
    for (int64_t i = 0; i < p->nrec; i++) {
        const pg_result *r = &p->res[i];
        const char *rid = p->text + p->hdr_off[i];
        const size_t idlen = (size_t)p->id_len[i];
        if (r->status) {
            grow(&p->msg, &p->msg_cap, mn + idlen + 128);
            mn += (size_t)sprintf(p->msg + mn, "ShortSequenceException: The length of sequence with recordID=%.*s is less than %d\n",
                                  (int)idlen, rid, PG_MIN_SEQ_LEN);
            continue;
        }
        int path[PG_MAX_DEPTH], np = 0;                   /* the genus' lineage, leaf first */
        size_t need = idlen + 16;
        for (int n = t->genus_node[r->genus]; n >= 0 && np < PG_MAX_DEPTH; n = t->parent[n]) {
            path[np++] = n;
            need += pl->tpiece_len[n] + 16;               /* "\tname\trank\t" + confidence; fixrank names a longer rank at most */
        }
        if (ifmt == 3) {                                  /* db: one line per rank of the assignment */
            for (int j = np - 1; j >= 0; j--) {
                grow(&p->out, &p->out_cap, on + idlen + 64);
                OUT_STR(rid, idlen);
                on += (size_t)sprintf(p->out + on, "\t%d\t%d\t%s\n", t->trainset_no, t->taxid ? t->taxid[path[j]] : path[j],
                                      pl->conf_tab[r->votes[t->depth[path[j]]]]);
            }
            continue;
        }
        grow(&p->out, &p->out_cap, on + need);
        OUT_STR(rid, idlen);
        if (ifmt == 2) OUT_STR("\t\t\t\t", 4);           /* with the piece's own leading TAB: five */
        else { OUT_STR("\t", 1); if (r->reversed) OUT_STR("-", 1); }
        if (ifmt == 1) {
            for (int k = 0; k < 6; k++) {
                int pick = -1;
                for (int j = np - 1; j >= 0 && pick < 0; j--)
                    if (strcmp(t->rank[path[j]], FIX[k]) == 0) pick = j;
                for (int kk = k + 1; kk < 6 && pick < 0; kk++)   /* missing rank: back-fill from the next lower one */
                    for (int j = np - 1; j >= 0 && pick < 0; j--)
                        if (strcmp(t->rank[path[j]], FIX[kk]) == 0) pick = j;
                if (pick < 0) continue;
                const char *c = pl->conf_tab[r->votes[t->depth[path[pick]]]];
                const size_t nl = strlen(t->name[path[pick]]), rl = strlen(FIX[k]), cl = strlen(c);
                grow(&p->out, &p->out_cap, on + nl + rl + cl + 8);
                OUT_STR("\t", 1); OUT_STR(t->name[path[pick]], nl);
                OUT_STR("\t", 1); OUT_STR(FIX[k], rl);
                OUT_STR("\t", 1); OUT_STR(c, cl);
            }
        } else {
            for (int j = np - 1; j >= 0; j--) {
                if (ifmt == 2 && t->depth[path[j]] == 0) continue;       /* Root cells are blank in the 5-TAB layout */
                const char *c = pl->conf_tab[r->votes[t->depth[path[j]]]];
                OUT_STR(pl->tpiece[path[j]], pl->tpiece_len[path[j]]);
                OUT_STR(c, strlen(c));
            }
        }
        grow(&p->out, &p->out_cap, on + 2);
        OUT_STR("\n", 1);
    }
#undef OUT_STR
    p->out_len = on;
    p->msg_len = mn;
}

static void *fmt_main(void *arg)
{
    pipeline_t *pl = (pipeline_t *)arg;
    for (;;) {
        piece_t *p = q_pop(&pl->q_fmt);
        if (!p) break;
        double t0 = now_s();
        p->out_len = p->msg_len = 0;
        if (!pl->failed) format_piece(pl, p);
        pthread_mutex_lock(&pl->mu);
        pl->busy_fmt += now_s() - t0;
        pthread_mutex_unlock(&pl->mu);
        q_push(&pl->q_out, p);
    }
    pthread_mutex_lock(&pl->mu);
    int last = --pl->fmt_alive == 0;
    pthread_mutex_unlock(&pl->mu);
    if (last) q_close(&pl->q_out);
    return NULL;
}

/* writer: pieces leave in file order whatever order they were finished in.  The in-order thread only prints the
 * stdout messages and hands out file offsets (a running sum of the pieces' sizes); a small pool of threads does the
 * pwrite()s, so the copy into the page cache is not one thread's job */
static void *writer_main(void *arg)
{
    pipeline_t *pl = (pipeline_t *)arg;
    piece_t **done = (piece_t **)calloc((size_t)pl->npieces, sizeof(piece_t *));
    int64_t next = 0, at = 0;
    for (;;) {
        piece_t *p = q_pop(&pl->q_out);
        if (!p) break;
        done[p->seq % pl->npieces] = p;
        while (done[next % pl->npieces] && done[next % pl->npieces]->seq == next) {
            piece_t *w = done[next % pl->npieces];
            done[next % pl->npieces] = NULL;
            if (!pl->failed && w->msg_len) fwrite(w->msg, 1, w->msg_len, stdout);
            w->out_at = at;
            at += (int64_t)w->out_len;
            next++;
            q_push(&pl->q_write, w);
        }
    }
    free(done);
    q_close(&pl->q_write);
    return NULL;
}

static void *pwrite_main(void *arg)
{
    pipeline_t *pl = (pipeline_t *)arg;
    for (;;) {
        piece_t *w = q_pop(&pl->q_write);
        if (!w) break;
        const double t0 = now_s();
        size_t off = 0;
        while (!pl->failed && off < w->out_len) {
            const ssize_t got = pwrite(pl->out_fd, w->out + off, w->out_len - off, (off_t)(w->out_at + (int64_t)off));
            if (got <= 0) { pl_fail(pl, "write failed"); break; }
            off += (size_t)got;
        }
        pthread_mutex_lock(&pl->mu);
        pl->busy_write += now_s() - t0;
        pthread_mutex_unlock(&pl->mu);
        q_push(&pl->q_free, w);
    }
    return NULL;
}

/* counts of device 0's model -> the models of the other devices: one ncclBroadcast per buffer (NVLink / NVSwitch);
 * a device that is listed twice cannot be two NCCL ranks, so such layouts use peer copies */
static int replicate_counts(int ndev, const int *dev, pg_model **models, char *err, size_t errlen)
{
    void *ptr[16][8];
    size_t nbytes[8];
    int nbuf = 0;
    for (int r = 0; r < ndev; r++)
        if (pg_model_buffers(models[r], ptr[r], nbytes, 8, &nbuf) != PG_OK) { snprintf(err, errlen, "pg_model_buffers failed"); return 1; }
    int distinct = 1;
    for (int a = 0; a < ndev; a++)
        for (int b = a + 1; b < ndev; b++)
            if (dev[a] == dev[b]) distinct = 0;
    if (!distinct) {
        for (int r = 1; r < ndev; r++)
            for (int b = 0; b < nbuf; b++)
                if (cudaMemcpyPeer(ptr[r][b], dev[r], ptr[0][b], dev[0], nbytes[b]) != cudaSuccess) { snprintf(err, errlen, "peer copy failed"); return 1; }
        cudaDeviceSynchronize();
        return 0;
    }
    ncclComm_t comms[16];
    cudaStream_t streams[16];
    /* stdout belongs to the tool's contract (the ShortSequenceException lines): whatever NCCL has to say (its version
     * banner under NCCL_DEBUG=VERSION is a plain printf) goes to stderr -- file descriptor 1 points at 2 while it starts */
    fflush(stdout);
    const int saved_out = dup(1);
    dup2(2, 1);
    ncclResult_t nr = ncclCommInitAll(comms, ndev, dev);
    fflush(stdout);
    if (saved_out >= 0) { dup2(saved_out, 1); close(saved_out); }
    if (nr != ncclSuccess) { snprintf(err, errlen, "ncclCommInitAll: %s", ncclGetErrorString(nr)); return 1; }
    for (int r = 0; r < ndev; r++) { cudaSetDevice(dev[r]); cudaStreamCreateWithFlags(&streams[r], cudaStreamNonBlocking); }
    for (int b = 0; b < nbuf && nr == ncclSuccess; b++) {
        ncclGroupStart();
        for (int r = 0; r < ndev; r++) {
            cudaSetDevice(dev[r]);
            ncclResult_t x = ncclBroadcast(ptr[r][b], ptr[r][b], nbytes[b], ncclUint8, 0, comms[r], streams[r]);
            if (x != ncclSuccess) nr = x;
        }
        ncclResult_t x = ncclGroupEnd();
        if (x != ncclSuccess) nr = x;
    }
    for (int r = 0; r < ndev; r++) { cudaSetDevice(dev[r]); cudaStreamSynchronize(streams[r]); cudaStreamDestroy(streams[r]); ncclCommDestroy(comms[r]); }
    if (nr != ncclSuccess) { snprintf(err, errlen, "ncclBroadcast: %s", ncclGetErrorString(nr)); return 1; }
    return 0;
}

int main(int argc, char **argv)
{
    static struct option lo[] = {{"train", required_argument, 0, 'T'}, {"ranks", required_argument, 0, 'R'},
                                 {"genus-token", required_argument, 0, 'K'}, {"device", required_argument, 0, 'G'},
                                 {"strict", no_argument, 0, 'S'}, {"min-boot-words", required_argument, 0, 'M'},
                                 {"gpus", required_argument, 0, 'N'}, {"devices", required_argument, 0, 'D'},
                                 {"contexts-per-gpu", required_argument, 0, 'C'}, {"format-threads", required_argument, 0, 'F'},
                                 {"export-rdp", required_argument, 0, 'E'}, {0, 0, 0, 0}};
    const char *q = NULL, *o = NULL, *model = NULL, *fmt = "allrank", *train = NULL, *ranks = NULL, *devlist = NULL, *export_dir = NULL;
    const char *gene = "16srrna";
    int device = 0, genus_token = 0, strict = 0, min_boot = 0, ngpu = 1, per_gpu = 2, nformat = 0;
    for (;;) {
        int c = getopt_long(argc, argv, "q:o:t:f:g:", lo, NULL);
        if (c == -1) break;
        switch (c) {
        case 'q': q = optarg; break;
        case 'o': o = optarg; break;
        case 't': model = optarg; break;
        case 'f': fmt = optarg; break;
        case 'g': gene = optarg; break;                   /* -g 16srrna|fungallsu: which default model when -t is absent */
        case 'E': export_dir = optarg; break;
        case 'T': train = optarg; break;
        case 'R': ranks = optarg; break;
        case 'K': genus_token = atoi(optarg); break;
        case 'G': device = atoi(optarg); break;
        case 'S': strict = 1; break;
        case 'M': min_boot = atoi(optarg); break;
        case 'N': ngpu = atoi(optarg); break;
        case 'D': devlist = optarg; break;
        case 'C': per_gpu = atoi(optarg); break;
        case 'F': nformat = atoi(optarg); break;
        default: break;
        }
    }
    if (strcmp(gene, "16srrna") != 0 && strcmp(gene, "fungallsu") != 0) { fprintf(stderr, "rdp_classifier: -g takes 16srrna or fungallsu\n"); return 1; }
    const int lsu = strcmp(gene, "fungallsu") == 0;
    if (!model) model = getenv(lsu ? "PANGEA_RDP_MODEL_FUNGALLSU" : "PANGEA_RDP_MODEL_16SRRNA");
    if (!model && !lsu) model = getenv("PANGEA_RDP_MODEL");
    if (!model) model = lsu ? "rdp_model_fungallsu.pgm" : "rdp_model.pgm";
    if (!train && !export_dir && (!q || !o)) {
        printf("Usage: rdp_classifier -q <query.fa> -o <out.txt> [-t model.pgm | -t rRNAClassifier.properties] [-g 16srrna|fungallsu]\n"
               "                      [-f allrank|fixrank|db|pangea] [--gpus N] [--devices a,b,...]\n"
               "       rdp_classifier --train <training.fa> -t <model.pgm> [--ranks r0,r1,...] [--genus-token N]\n"
               "       rdp_classifier --export-rdp <dir> -t <model.pgm>     (stock RDP trainset files; layouts UNVERIFIED against RDP 2.5)\n");
        return 0;
    }
    int ifmt = strcmp(fmt, "allrank") == 0 ? 0 : strcmp(fmt, "fixrank") == 0 ? 1 : strcmp(fmt, "pangea") == 0 ? 2 : strcmp(fmt, "db") == 0 ? 3 : -1;
    if (ifmt < 0) { fprintf(stderr, "rdp_classifier: unknown format %s\n", fmt); return 1; }
    if (ngpu < 1 || ngpu > 16 || per_gpu < 1 || per_gpu > 2) { fprintf(stderr, "rdp_classifier: --gpus 1..16, --contexts-per-gpu 1 or 2\n"); return 1; }
    int dev[16];
    for (int r = 0; r < ngpu; r++) dev[r] = device + r;
    if (devlist) {
        char *dl = strdup(devlist);
        int r = 0;
        for (char *tok = strtok(dl, ","); tok && r < 16; tok = strtok(NULL, ",")) dev[r++] = atoi(tok);
        if (r != ngpu) { fprintf(stderr, "rdp_classifier: --devices lists %d devices, --gpus says %d\n", r, ngpu); return 1; }
        free(dl);
    }
    const int timing = getenv("PG_TIMING") != NULL;
    double t0 = now_s(), t1, t_begin = t0;
#define LAP(what) do { if (timing) { t1 = now_s(); fprintf(stderr, "[timing] %-22s %.3f s\n", what, t1 - t0); t0 = t1; } } while (0)
    const int split = !(train || export_dir) && per_gpu >= 2;          /* ingest and classify contexts per GPU */
    const int nworkers = (train || export_dir) ? 1 : ngpu;
    pg_ctx **wctx = (pg_ctx **)calloc((size_t)nworkers, sizeof(pg_ctx *));
    for (int w = 0; w < nworkers; w++) {
        wctx[w] = pg_init(dev[w % ngpu]);                /* workers 0..ngpu-1 own the models of their devices */
        if (!wctx[w]) { fprintf(stderr, "rdp_classifier: %s\n", pg_last_error(NULL)); return 1; }
    }
    LAP("pg_init");
    if (train) {
        int rc = do_train(wctx[0], train, model, ranks, genus_token);
        pg_shutdown(wctx[0]);
        return rc;
    }

    /* ---- the model: device 0 reads the file; the other devices get its integer counts and derive their own tables */
    pg_model **models = (pg_model **)calloc((size_t)ngpu, sizeof(pg_model *));
    void *blob = NULL;
    int64_t blen = 0;
    taxonomy t;
    const int from_tables = is_properties(model);
    if (from_tables) {
        if (import_rdp(wctx[0], model, &models[0], &t)) return 1;
    } else {
        if (pg_model_load(wctx[0], model, &models[0], &blob, &blen) != PG_OK) { fprintf(stderr, "rdp_classifier: %s\n", pg_last_error(wctx[0])); return 1; }
        if (!blob || tax_from_blob((const char *)blob, blen, &t)) { fprintf(stderr, "rdp_classifier: %s has no taxonomy section\n", model); return 1; }
    }
    LAP("model load");
    if (export_dir) {
        if (from_tables) { fprintf(stderr, "rdp_classifier: --export-rdp needs a .pgm model (counts), not trainset files\n"); return 1; }
        const int rc = export_rdp(wctx[0], models[0], &t, export_dir);
        pg_model_free(models[0]);
        pg_shutdown(wctx[0]);
        return rc;
    }
    if (ngpu > 1 && from_tables) {
        /* a model given as tables has no counts to broadcast: every device reads the files itself */
        for (int r = 1; r < ngpu; r++) {
            taxonomy t2;
            if (import_rdp(wctx[r], model, &models[r], &t2)) return 1;
        }
    } else if (ngpu > 1) {
        int32_t *anc = tax_lineage_table(&t);
        for (int r = 1; r < ngpu; r++) {
            if (pg_model_create(wctx[r], t.G, &models[r]) != PG_OK || pg_model_set_lineage(models[r], anc, t.maxdepth) != PG_OK) {
                fprintf(stderr, "rdp_classifier: %s\n", pg_last_error(wctx[r]));
                return 1;
            }
        }
        for (int r = 0; r < ngpu; r++) pg_sync(wctx[r]);
        char err[256];
        if (replicate_counts(ngpu, dev, models, err, sizeof err)) { fprintf(stderr, "rdp_classifier: %s\n", err); return 1; }
        for (int r = 1; r < ngpu; r++)
            if (pg_model_commit(models[r]) != PG_OK) { fprintf(stderr, "rdp_classifier: %s\n", pg_last_error(wctx[r])); return 1; }
        free(anc);
        LAP("model replication");
    }

    pipeline_t pl;
    memset(&pl, 0, sizeof pl);
    pl.t = &t;
    pl.ifmt = ifmt;
    pl.opts.min_boot_words = min_boot;
    pl.opts.mode = strict ? 0 : 1;
    /* pieces of about 64 MiB (PG_CLI_PIECE_BYTES overrides, tests): large enough for full chunks of reads on the device,
     * small enough that ten of them -- three being read, one per device context, the rest with the formatters and the
     * writer -- stay pinned without a long start-up */
    pl.piece_bytes = (int64_t)64 << 20;
    if (getenv("PG_CLI_PIECE_BYTES") && atoll(getenv("PG_CLI_PIECE_BYTES")) > 0) pl.piece_bytes = atoll(getenv("PG_CLI_PIECE_BYTES"));
    pl.fq = fopen(q, "rb");
    if (!pl.fq) { fprintf(stderr, "rdp_classifier: cannot read %s\n", q); return 1; }
    pl.fo = fopen(o, "w");
    if (!pl.fo) { fprintf(stderr, "rdp_classifier: cannot write %s\n", o); return 1; }
    setvbuf(pl.fo, NULL, _IONBF, 0);                      /* pieces are written whole */
    if (nformat <= 0) {
        long nc = sysconf(_SC_NPROCESSORS_ONLN);
        nformat = nc > 16 ? 8 : (nc > 4 ? 4 : 2);
        if (nformat < ngpu * 3) nformat = ngpu * 3;
    }
    pl.nworkers = nworkers;
    pl.nformat = nformat;
    pl.ngpu = ngpu;
    pl.split = split;
    if (split) {
        pl.ictx = (pg_ctx **)calloc((size_t)ngpu, sizeof(pg_ctx *));
        pl.q_mid = (queue_t *)calloc((size_t)ngpu, sizeof(queue_t));
        for (int g = 0; g < ngpu; g++) {
            pl.ictx[g] = pg_init(dev[g]);
            if (!pl.ictx[g]) { fprintf(stderr, "rdp_classifier: %s\n", pg_last_error(NULL)); return 1; }
            int lo_p = 0, hi_p = 0;
            cudaStream_t hs = NULL;
            cudaSetDevice(dev[g]);
            if (cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p) == cudaSuccess &&
                cudaStreamCreateWithPriority(&hs, cudaStreamNonBlocking, hi_p) == cudaSuccess)
                pg_set_stream(pl.ictx[g], (void *)hs);
        }
    }
    pl.wctx = wctx;
    pl.wmodel = (pg_model **)calloc((size_t)nworkers, sizeof(pg_model *));
    for (int w = 0; w < nworkers; w++) pl.wmodel[w] = models[w % ngpu];
    pl.gpu_alive = nworkers;
    pl.fmt_alive = nformat;
    pthread_mutex_init(&pl.mu, NULL);
    for (int v = 0; v <= 100; v++) pg_fmt_conf(v, pl.conf_tab[v]);
    pl.tpiece = (char **)malloc(sizeof(char *) * (size_t)t.nnodes);
    pl.tpiece_len = (size_t *)malloc(sizeof(size_t) * (size_t)t.nnodes);
    for (int k = 0; k < t.nnodes; k++) {
        size_t n = strlen(t.name[k]) + strlen(t.rank[k]) + 4;
        pl.tpiece[k] = (char *)malloc(n);
        pl.tpiece_len[k] = (size_t)snprintf(pl.tpiece[k], n, "\t%s\t%s\t", t.name[k], t.rank[k]);
    }
    fseek(pl.fq, 0, SEEK_END);
    pl.fsize = ftell(pl.fq);
    fseek(pl.fq, 0, SEEK_SET);
    pl.total_pieces = pl.fsize > 0 ? (pl.fsize + pl.piece_bytes - 1) / pl.piece_bytes : 0;
    /* a piece holds the records that start inside its window, so it can run a record past it; small files: one piece
     * is all there is -- size the buffers by the file */
    pl.textcap0 = pl.piece_bytes + pl.piece_bytes / 16 + 65536;
    if (pl.fsize + 4096 < pl.textcap0) pl.textcap0 = pl.fsize + 4096;
    const int nreaders = pl.total_pieces > 2 ? (ngpu + 2 > 8 ? 8 : ngpu + 2) : 1;
    pl.nwriters = pl.total_pieces > 2 ? (ngpu > 3 ? 4 : 2) : 1;
    pl.out_fd = fileno(pl.fo);
    pl.read_alive = nreaders;
    pl.npieces = nworkers * (split ? 3 : 1) + nformat + nreaders + pl.nwriters + 1;
    if (pl.total_pieces < pl.npieces) pl.npieces = (int)(pl.total_pieces > 0 ? pl.total_pieces : 1);
    q_init(&pl.q_free, pl.npieces + 1);
    q_init(&pl.q_gpu, pl.npieces + 1);
    q_init(&pl.q_fmt, pl.npieces + 1);
    q_init(&pl.q_out, pl.npieces + 1);
    q_init(&pl.q_write, pl.npieces + 1);
    if (split) for (int g = 0; g < ngpu; g++) q_init(&pl.q_mid[g], pl.npieces + 1);
    piece_t *pieces = (piece_t *)calloc((size_t)pl.npieces, sizeof(piece_t));
    cudaSetDevice(dev[0]);
    for (int i = 0; i < pl.npieces; i++) {
        /* pinned up front: cudaHostAlloc takes the driver's lock and would stall the device contexts mid-run */
        pieces[i].textcap = pl.textcap0;
        pieces[i].text = (char *)pinned_alloc((size_t)pl.textcap0 + 1);
        pieces[i].rescap = pl.textcap0 / 256 + 1024;
        pieces[i].res = (pg_result *)pinned_alloc(sizeof(pg_result) * (size_t)pieces[i].rescap);
        if (!pieces[i].text || !pieces[i].res) { fprintf(stderr, "rdp_classifier: pinned allocation failed\n"); return 1; }
        q_push(&pl.q_free, &pieces[i]);
    }
    LAP("buffers");

    pthread_t th_reader[8], th_pwrite[4], th_writer, *th_gpu = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nworkers),
              *th_fmt = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nformat);
    worker_arg *wa = (worker_arg *)malloc(sizeof(worker_arg) * (size_t)nworkers);
    const double t_pipe = now_s();
    pthread_create(&th_writer, NULL, writer_main, &pl);
    for (int w = 0; w < pl.nwriters; w++) pthread_create(&th_pwrite[w], NULL, pwrite_main, &pl);
    for (int f = 0; f < nformat; f++) pthread_create(&th_fmt[f], NULL, fmt_main, &pl);
    pthread_t *th_ing = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nworkers);
    for (int w = 0; w < nworkers; w++) {
        wa[w].pl = &pl; wa[w].idx = w;
        pthread_create(&th_gpu[w], NULL, split ? classify_main : gpu_main, &wa[w]);
        if (split) pthread_create(&th_ing[w], NULL, ingest_main, &wa[w]);
    }
    for (int r = 0; r < nreaders; r++) pthread_create(&th_reader[r], NULL, reader_main, &pl);
    for (int r = 0; r < nreaders; r++) pthread_join(th_reader[r], NULL);
    if (split) for (int w = 0; w < nworkers; w++) pthread_join(th_ing[w], NULL);
    for (int w = 0; w < nworkers; w++) pthread_join(th_gpu[w], NULL);
    for (int f = 0; f < nformat; f++) pthread_join(th_fmt[f], NULL);
    pthread_join(th_writer, NULL);
    for (int w = 0; w < pl.nwriters; w++) pthread_join(th_pwrite[w], NULL);
    q_close(&pl.q_free);
    const double dt = now_s() - t_pipe;
    fclose(pl.fq);
    fclose(pl.fo);
    if (pl.failed) { fprintf(stderr, "rdp_classifier: %s\n", pl.err); return 1; }
    if (timing)
        fprintf(stderr, "[timing] pipeline %.3f s for %lld reads = %.2f M reads/s (%d GPUs x %d contexts, %d formatter threads; busy: "
                        "read %.3f, gpu %.3f (ingest %.3f), format %.3f, write %.3f s); whole run %.3f s\n",
                dt, (long long)pl.reads_total, 1e-6 * (double)pl.reads_total / dt, ngpu, per_gpu, nformat, pl.busy_read, pl.busy_gpu,
                pl.busy_ingest, pl.busy_fmt, pl.busy_write, now_s() - t_begin);
    for (int i = 0; i < pl.npieces; i++) {
        if (pieces[i].text) cudaFreeHost(pieces[i].text);
        if (pieces[i].res) cudaFreeHost(pieces[i].res);
        free(pieces[i].hdr_off); free(pieces[i].id_len); free(pieces[i].out); free(pieces[i].msg);
    }
    pg_free(blob);
    for (int r = 0; r < ngpu; r++) pg_model_free(models[r]);
    for (int w = 0; w < nworkers; w++) pg_shutdown(wctx[w]);
    if (split) for (int g = 0; g < ngpu; g++) pg_shutdown(pl.ictx[g]);
    return 0;
}
