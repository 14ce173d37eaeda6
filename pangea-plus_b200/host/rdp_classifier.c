/*
 * rdp_classifier -- drop-in for the RDP stage of the PANGEA+ pipeline,
 *     java -Xmx1g -jar rdp_classifier-2.5.jar -q <in.fa> -o <out.txt>          (README.md:119-122)
 * Same -q/-o surface (plus -f allrank|fixrank|pangea and -t <model>), one output line per
 * classifiable read in input order, short reads reported on stdout and left out of the file
 * (SURVEY.md rows A0, A2, A9).  All arithmetic runs on the GPU through libpangea_b200.
 *
 * The jar carries its training data inside; this tool cannot ship RDP's trainset, so the
 * model is a file: `-t model.pgm` (default: $PANGEA_RDP_MODEL, else ./rdp_model.pgm), made by
 *     rdp_classifier --train <training.fa> -t <model.pgm> [--ranks r0,r1,...]
 * Training FASTA headers carry the lineage the way RDP's trainer expects it:
 *     >seqid Root;Bacteria;Firmicutes;...;Genus          (blank or TAB before the lineage)
 * or, with --genus-token N, the genus is the N-th blank-separated token of the header (the
 * layout of the reference's validation_dataset/rdp_download_*.fa: N = 2) under a flat Root.
 *
 * Output formats:
 *   allrank  id TAB [-] (TAB name TAB rank TAB conf)*      the jar's default
 *   fixrank  id TAB [-] then domain, phylum, class, order, family, genus triples
 *   pangea   id, five TABs, then the triples below Root -- the layout that
 *            Consensus_BLAST_SOAP_RDP-1.1.pl:126 actually parses
 */
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pangea_b200.h"
#include "pg_host_common.h"
#include <time.h>

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

int main(int argc, char **argv)
{
    static struct option lo[] = {{"train", required_argument, 0, 'T'}, {"ranks", required_argument, 0, 'R'},
                                 {"genus-token", required_argument, 0, 'K'}, {"device", required_argument, 0, 'G'},
                                 {"strict", no_argument, 0, 'S'}, {"min-boot-words", required_argument, 0, 'M'}, {0, 0, 0, 0}};
    const char *q = NULL, *o = NULL, *model = NULL, *fmt = "allrank", *train = NULL, *ranks = NULL;
    int device = 0, genus_token = 0, strict = 0, min_boot = 0;
    for (;;) {
        int c = getopt_long(argc, argv, "q:o:t:f:g:", lo, NULL);
        if (c == -1) break;
        switch (c) {
        case 'q': q = optarg; break;
        case 'o': o = optarg; break;
        case 't': model = optarg; break;
        case 'f': fmt = optarg; break;
        case 'g': break;                                  /* -g 16srrna|fungallsu: one gene per model file here */
        case 'T': train = optarg; break;
        case 'R': ranks = optarg; break;
        case 'K': genus_token = atoi(optarg); break;
        case 'G': device = atoi(optarg); break;
        case 'S': strict = 1; break;
        case 'M': min_boot = atoi(optarg); break;
        default: break;
        }
    }
    if (!model) model = getenv("PANGEA_RDP_MODEL");
    if (!model) model = "rdp_model.pgm";
    if (!train && (!q || !o)) {
        printf("Usage: rdp_classifier -q <query.fa> -o <out.txt> [-t model.pgm] [-f allrank|fixrank|pangea]\n"
               "       rdp_classifier --train <training.fa> -t <model.pgm> [--ranks r0,r1,...] [--genus-token N]\n");
        return 0;
    }
    int ifmt = strcmp(fmt, "allrank") == 0 ? 0 : strcmp(fmt, "fixrank") == 0 ? 1 : strcmp(fmt, "pangea") == 0 ? 2 : -1;
    if (ifmt < 0) { fprintf(stderr, "rdp_classifier: unknown format %s\n", fmt); return 1; }
    const int timing = getenv("PG_TIMING") != NULL;
    double t0 = now_s(), t1;
    pg_ctx *ctx = pg_init(device);
    if (!ctx) { fprintf(stderr, "rdp_classifier: %s\n", pg_last_error(NULL)); return 1; }
#define LAP(what) do { if (timing) { t1 = now_s(); fprintf(stderr, "[timing] %-22s %.3f s\n", what, t1 - t0); t0 = t1; } } while (0)
    LAP("pg_init");
    if (train) {
        int rc = do_train(ctx, train, model, ranks, genus_token);
        pg_shutdown(ctx);
        return rc;
    }

    pg_model *md = NULL;
    void *blob = NULL;
    int64_t blen = 0;
    if (pg_model_load(ctx, model, &md, &blob, &blen) != PG_OK) { fprintf(stderr, "rdp_classifier: %s\n", pg_last_error(ctx)); return 1; }
    taxonomy t;
    if (!blob || tax_from_blob((const char *)blob, blen, &t)) { fprintf(stderr, "rdp_classifier: %s has no taxonomy section\n", model); return 1; }
    LAP("model load");
    /* the query file goes to the device as text: records, ids and the packed read store come from the GPU
     * (pg_fasta_ingest); the host keeps the text only to print the ids */
    FILE *fq = fopen(q, "rb");
    if (!fq) { fprintf(stderr, "rdp_classifier: cannot read %s\n", q); return 1; }
    FILE *fo = fopen(o, "w");
    if (!fo) { fprintf(stderr, "rdp_classifier: cannot write %s\n", o); return 1; }
    pg_classify_opts opts;
    memset(&opts, 0, sizeof opts);
    opts.min_boot_words = min_boot;
    opts.mode = strict ? 0 : 1;
    /* Files of any size: the text is handed over in pieces of about 1 GiB cut at record boundaries (a line
     * starting with '>'), so the device never holds more than one piece with its word ids and records.
     * PG_CLI_PIECE_BYTES overrides the size (tests). */
    int64_t piece_bytes = (int64_t)1 << 30;
    if (getenv("PG_CLI_PIECE_BYTES") && atoll(getenv("PG_CLI_PIECE_BYTES")) > 0) piece_bytes = atoll(getenv("PG_CLI_PIECE_BYTES"));
    int64_t *hdr_off = NULL;
    int32_t *id_len = NULL;
    pg_result *res = NULL;
    int64_t cap = 0, rescap = 0;
    static const char *FIX[6] = {"domain", "phylum", "class", "order", "family", "genus"};
    /* Output is assembled from pieces prepared once per taxon ("\tname\trank\t") and once per vote
     * count (the 101 possible confidences), so a line costs a few memcpy()s, not a dozen fprintf()s. */
    char conf_tab[101][8];
    for (int v = 0; v <= 100; v++) pg_fmt_conf(v, conf_tab[v]);
    char **piece = (char **)malloc(sizeof(char *) * (size_t)t.nnodes);
    size_t *piece_len = (size_t *)malloc(sizeof(size_t) * (size_t)t.nnodes);
    for (int k = 0; k < t.nnodes; k++) {
        size_t n = strlen(t.name[k]) + strlen(t.rank[k]) + 4;
        piece[k] = (char *)malloc(n);
        piece_len[k] = (size_t)snprintf(piece[k], n, "\t%s\t%s\t", t.name[k], t.rank[k]);
    }
    size_t ocap = (size_t)8 << 20, on = 0;
    char *obuf = (char *)malloc(ocap);
    /* the file is read piece by piece: host memory holds one piece (plus the head of the next record) */
    int64_t bufcap = piece_bytes + 4096, have = 0;
    char *qbuf = (char *)malloc((size_t)bufcap + 1);
    int at_eof = 0;
    for (;;) {
        while (!at_eof && have < bufcap) {
            size_t got = fread(qbuf + have, 1, (size_t)(bufcap - have), fq);
            if (got == 0) at_eof = 1;
            have += (int64_t)got;
        }
        if (have == 0) break;
        int64_t p1 = have;
        if (!at_eof) {                                     /* cut at the start of the last record in the buffer */
            int64_t c = have - 1;
            while (c > 0 && !(qbuf[c] == '>' && qbuf[c - 1] == '\n')) c--;
            if (c > 0) p1 = c;
            else {                                         /* one record longer than the buffer: make room, read on */
                bufcap *= 2;
                qbuf = (char *)realloc(qbuf, (size_t)bufcap + 1);
                continue;
            }
        }
        const char *ptext = qbuf;
        const int64_t plen = p1;
        int64_t nrec = 0;
        pg_reads *reads = NULL;
        if (cap == 0) cap = plen / 32 + 1024;
        for (;;) {
            hdr_off = (int64_t *)realloc(hdr_off, sizeof(int64_t) * (size_t)cap);
            id_len = (int32_t *)realloc(id_len, sizeof(int32_t) * (size_t)cap);
            int rc = pg_fasta_ingest(ctx, ptext, plen, cap, &nrec, hdr_off, id_len, NULL, &reads);
            if (rc == PG_ERANGE && nrec > cap) { cap = nrec; continue; }
            if (rc != PG_OK) { fprintf(stderr, "rdp_classifier: %s\n", pg_last_error(ctx)); return 1; }
            break;
        }
        LAP("FASTA ingest (GPU)");
        if (nrec + 1 > rescap) { rescap = nrec + 1; res = (pg_result *)realloc(res, sizeof(pg_result) * (size_t)rescap); }
        if (nrec > 0 && pg_classify_packed_host(ctx, md, reads, &opts, res, NULL) != PG_OK) {
            fprintf(stderr, "rdp_classifier: %s\n", pg_last_error(ctx));
            return 1;
        }
        pg_reads_free(reads);
        LAP("pg_classify_packed");
#define OUT_ROOM(need) do { if (on + (need) > ocap) { fwrite(obuf, 1, on, fo); on = 0; } } while (0)
#define OUT_STR(s, n) do { memcpy(obuf + on, (s), (n)); on += (n); } while (0)
    for (int64_t i = 0; i < nrec; i++) {
        const pg_result *r = &res[i];
        const char *rid = ptext + hdr_off[i];
        const size_t idlen = (size_t)id_len[i];
        if (r->status) {
            printf("ShortSequenceException: The length of sequence with recordID=%.*s is less than %d\n", (int)idlen, rid, PG_MIN_SEQ_LEN);
            continue;
        }
        int path[PG_MAX_DEPTH], np = 0;                   /* the genus' lineage, leaf first */
        for (int n = t.genus_node[r->genus]; n >= 0 && np < PG_MAX_DEPTH; n = t.parent[n]) path[np++] = n;
        OUT_ROOM(idlen + 16 + (size_t)np * 256);
        OUT_STR(rid, idlen);
        if (ifmt == 2) OUT_STR("\t\t\t\t", 4);           /* with the piece's own leading TAB: five */
        else { OUT_STR("\t", 1); if (r->reversed) OUT_STR("-", 1); }
        if (ifmt == 1) {
            for (int k = 0; k < 6; k++) {
                int pick = -1;
                for (int j = np - 1; j >= 0 && pick < 0; j--)
                    if (strcmp(t.rank[path[j]], FIX[k]) == 0) pick = j;
                for (int kk = k + 1; kk < 6 && pick < 0; kk++)   /* missing rank: back-fill from the next lower one */
                    for (int j = np - 1; j >= 0 && pick < 0; j--)
                        if (strcmp(t.rank[path[j]], FIX[kk]) == 0) pick = j;
                if (pick < 0) continue;
                const char *c = conf_tab[r->votes[t.depth[path[pick]]]];
                on += (size_t)sprintf(obuf + on, "\t%s\t%s\t%s", t.name[path[pick]], FIX[k], c);
            }
        } else {
            for (int j = np - 1; j >= 0; j--) {
                if (ifmt == 2 && t.depth[path[j]] == 0) continue;        /* Root cells are blank in the 5-TAB layout */
                const char *c = conf_tab[r->votes[t.depth[path[j]]]];
                OUT_STR(piece[path[j]], piece_len[path[j]]);
                OUT_STR(c, strlen(c));
            }
        }
        OUT_STR("\n", 1);
    }
        LAP("format + write");
        memmove(qbuf, qbuf + p1, (size_t)(have - p1));    /* the head of the next record moves to the front */
        have -= p1;
    }
    fclose(fq);
    fwrite(obuf, 1, on, fo);
    free(obuf);
    fclose(fo);
    free(res);
    pg_free(blob);
    free(qbuf);
    free(hdr_off);
    free(id_len);
    pg_model_free(md);
    pg_shutdown(ctx);
    return 0;
}
