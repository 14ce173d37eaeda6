/*
 * taxcollector_ref.c -- CPU ORACLE for Stage B (lineage), the part of it that
 * cannot travel to the GPU box: Tax_class/NCBI-taxcollector-0.01.pl (Perl).
 *
 * TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu_baseline).  Restates, string
 * operation by string operation, what the Perl driver does with the answers of the
 * reference's tax_class binary (Tax_class/ncbitc.c), reading the reference's own
 * .bin files directly instead of forking tax_class per lookup.
 *
 * PINNED: tests/test_stage_bc_cpu.py runs the REAL Perl script + the REAL tax_class
 * (compiled from /root/reference/Tax_class/ncbitc.c into oracle/_ref/) on seeded
 * taxonomies in this container and requires byte-identical output files from this
 * restatement; the outputs are also committed under tests/golden/.
 *
 * Reference rows: B1 ncbitc.c:567-599, B2 :601-628, B3 :630-699,
 * B5 taxcollector:166-186, B6 :188-300, B7 :92-155.
 *
 * Inputs outside the reference's defined behaviour (it reads uninitialised memory or
 * recurses on garbage there) are DEFINED here and in the product the same way:
 *   gi missing / <= 0 / past the end of the gi table   -> Unidentified(GI:<text>);
 *   leaf whose parent is 1 (tax_class -s prints nothing) -> the walk starts at the leaf
 *   taxid past the end of the node table, taxid <= 0, or a chain deeper than 128 -> the walk stops
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NODE_REC 28      /* sizeof(struct nodes_dmp), ncbitc.c:98-114 */
#define NAME_REC 196     /* sizeof(struct names_dmp), ncbitc.c:128-133 */

static const char *RANK_STR[29] = {           /* enum ncbitc_rank, ncbitc.c:39-69 */
    "class", "family", "forma", "genus", "infraclass", "infraorder", "kingdom", "no rank", "order",
    "parvorder", "phylum", "species", "species group", "species subgroup", "subclass", "subfamily",
    "subgenus", "subkingdom", "suborder", "subphylum", "subspecies", "subtribe", "superclass",
    "superfamily", "superkingdom", "superorder", "superphylum", "tribe", "varietas"};

typedef struct {
    int32_t *gi2tax; int64_t ngi;
    unsigned char *nodes; int64_t nnodes;
    unsigned char *names; int32_t nnames;
} txc_db;

static unsigned char *slurp(const char *dir, const char *name, int64_t *len)
{
    char path[1024];
    snprintf(path, sizeof path, "%s/%s", dir, name);
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    int64_t n = ftell(f);
    fseek(f, 0, SEEK_SET);
    unsigned char *b = (unsigned char *)malloc((size_t)n + 1);
    if (fread(b, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(b); return NULL; }
    fclose(f);
    *len = n;
    return b;
}

txc_db *txc_load(const char *dir)
{
    txc_db *db = (txc_db *)calloc(1, sizeof *db);
    int64_t n;
    db->gi2tax = (int32_t *)slurp(dir, "gi_taxid_nucl.dmp.bin", &n);
    if (!db->gi2tax) { free(db); return NULL; }
    db->ngi = n / 4;
    db->nodes = slurp(dir, "nodes.dmp.bin", &n);
    if (!db->nodes) { free(db); return NULL; }
    db->nnodes = n / NODE_REC;
    db->names = slurp(dir, "names.dmp.bin", &n);
    if (!db->names) { free(db); return NULL; }
    memcpy(&db->nnames, db->names, 4);
    return db;
}

void txc_free(txc_db *db)
{
    if (!db) return;
    free(db->gi2tax); free(db->nodes); free(db->names); free(db);
}

/* tax_class -n <taxid> followed by get_name (taxcollector:188-224): first record of the
 * run whose class field matches /scientific name/, name trimmed.  The binary search of
 * ncbitc_search_name (:647-699) covers positions 0..num-1 with a 1-based seek, i.e.
 * records 0..num-2: the last record of the file is never seen. */
static int sci_name(const txc_db *db, int taxid, char *out)
{
    int lo = 0, hi = db->nnames - 2, found = -1;
    while (lo <= hi) {
        int mid = (lo + hi) / 2, v;
        memcpy(&v, db->names + 4 + (size_t)mid * NAME_REC, 4);
        if (v == taxid) { found = mid; break; }
        if (v < taxid) lo = mid + 1; else hi = mid - 1;
    }
    if (found < 0) return 0;
    int first = found;
    while (first > 0) {
        int v;
        memcpy(&v, db->names + 4 + (size_t)(first - 1) * NAME_REC, 4);
        if (v != taxid) break;
        first--;
    }
    for (int r = first; r <= db->nnames - 2; r++) {
        const unsigned char *rec = db->names + 4 + (size_t)r * NAME_REC;
        int v;
        memcpy(&v, rec, 4);
        if (v != taxid) break;
        const char *name = (const char *)rec + 4, *cls = (const char *)rec + 4 + 128;
        if (strstr(cls, "scientific name")) {
            /* printed as " <name_txt> " around the stored leading blank, then s/\t//g and trim */
            char tmp[80];
            int n = 0;
            for (const char *p = name; *p && n < 70; p++) if (*p != '\t') tmp[n++] = *p;
            tmp[n] = 0;
            int a = 0, b = n;
            while (a < b && (tmp[a] == ' ' || tmp[a] == '\n' || tmp[a] == '\r' || tmp[a] == '\f')) a++;
            while (b > a && (tmp[b - 1] == ' ' || tmp[b - 1] == '\n' || tmp[b - 1] == '\r' || tmp[b - 1] == '\f')) b--;
            memcpy(out, tmp + a, (size_t)(b - a));
            out[b - a] = 0;
            return 1;
        }
    }
    return 0;
}

static const char *RANKLIST[8] = {"superkingdom", "phylum", "class", "order", "family", "genus", "species", "kingdom"};

/* get_uptaxa (taxcollector:226-300): appends "[idx]" / "Name;|" / "[0]Unclassified;|" pieces */
static void walk(const txc_db *db, int taxid, char *acc, size_t cap)
{
    for (int depth = 0; depth < 128; depth++) {
        if (taxid <= 0 || taxid > db->nnodes) return;
        const unsigned char *rec = db->nodes + (size_t)(taxid - 1) * NODE_REC;
        int parent;
        memcpy(&parent, rec + 4, 4);
        int rk = (signed char)rec[8];
        char rank[32];
        /* printed rank with blanks removed ($nodeline =~ s/\ //g) */
        const char *rs = (rk >= 0 && rk < 29) ? RANK_STR[rk] : "invalid id";
        int n = 0;
        for (const char *p = rs; *p; p++) if (*p != ' ') rank[n++] = *p;
        rank[n] = 0;
        int idx = -1;
        for (int i = 0; i < 8; i++) if (strcmp(rank, RANKLIST[i]) == 0) idx = i;
        if (idx >= 0) {
            char piece[16], nm[80];
            snprintf(piece, sizeof piece, "[%d]", idx);
            strncat(acc, piece, cap - strlen(acc) - 1);
            if (sci_name(db, taxid, nm)) {
                strncat(acc, nm, cap - strlen(acc) - 1);
                strncat(acc, ";|", cap - strlen(acc) - 1);
            }
            if (idx == 0) return;                         /* superkingdom: stop */
            taxid = parent;
        } else if (strcmp(rank, "norank") == 0) {
            if (parent == 1) { strncat(acc, "[0]Unclassified;|", cap - strlen(acc) - 1); return; }
            taxid = parent;
        } else {
            taxid = parent;
        }
    }
}

static int has_char(const char *s, size_t n, char c) { return memchr(s, c, n) != NULL; }

/* lineage text exactly as printed between the first and second TAB (B5-B7) */
int txc_lineage(const txc_db *db, const char *gi_text, char *out, size_t cap)
{
    char acc[4096];
    acc[0] = 0;
    long gi = gi_text ? atol(gi_text) : 0;
    int leaf = 0;
    if (gi_text && gi >= 1 && gi <= db->ngi) leaf = db->gi2tax[gi - 1];
    if (!gi_text || leaf == 0) {
        snprintf(acc, sizeof acc, "Unidentified(GI:%s);|", gi_text ? gi_text : "");
    } else {
        walk(db, leaf, acc, sizeof acc);
    }
    /* split on '|' (trailing empty fields dropped), print last -> first */
    char *el[256];
    size_t len[256];
    int ne = 0;
    char *p = acc;
    while (ne < 256) {
        char *bar = strchr(p, '|');
        el[ne] = p;
        len[ne] = bar ? (size_t)(bar - p) : strlen(p);
        ne++;
        if (!bar) break;
        p = bar + 1;
    }
    while (ne > 0 && len[ne - 1] == 0) ne--;
    size_t o = 0;
    out[0] = 0;
    for (int i = ne - 1; i >= 0; i--) {
        char e[256];
        size_t n = len[i] < 255 ? len[i] : 255;
        memcpy(e, el[i], n);
        e[n] = 0;
        if (has_char(e, n, '6')) {
            for (size_t k = 0; k < n; k++)
                if (e[k] == ' ' || e[k] == '\t' || e[k] == '\n' || e[k] == '\r' || e[k] == '\f') e[k] = '_';
            memcpy(el[i], e, n);                           /* the edit persists in the array */
            int any5 = 0;
            for (int j = 0; j < ne; j++) if (has_char(el[j], len[j], '5')) any5 = 1;
            if (!any5) {
                char *six = (char *)memchr(e, '6', n);
                *six = '5';
                if (o + n < cap) { memcpy(out + o, e, n); o += n; }
                *six = '6';                                /* s/5/6/ hits the same byte: no other '5' exists */
                if (o + n < cap) { memcpy(out + o, e, n); o += n; }
            } else {
                if (o + n < cap) { memcpy(out + o, e, n); o += n; }
            }
        } else {
            char *seven = (char *)memchr(e, '7', n);
            if (seven) *seven = '9';
            if (o + n < cap) { memcpy(out + o, e, n); o += n; }
        }
    }
    out[o] = 0;
    return (int)o;
}

/* Perl split(/ |\t\t|\t/, $line): fields; leading empty field kept, trailing empties dropped */
static int split_fields(const char *line, const char **f, size_t *fl, int maxf)
{
    int n = 0;
    const char *p = line, *start = line;
    while (*p) {
        size_t sep = 0;
        if (*p == ' ') sep = 1;
        else if (*p == '\t') sep = (p[1] == '\t') ? 2 : 1;
        if (sep) {
            if (n < maxf) { f[n] = start; fl[n] = (size_t)(p - start); n++; }
            p += sep;
            start = p;
        } else p++;
    }
    if (n < maxf) { f[n] = start; fl[n] = (size_t)(p - start); n++; }
    while (n > 0 && fl[n - 1] == 0) n--;
    return n;
}

/* one input line -> one output line (without the final newline).  Returns -1 for an empty
 * line: the reference exits there (taxcollector:82-86). */
int txc_line(const txc_db *db, const char *line_in, char *out, size_t cap)
{
    char line[8192];
    size_t L = strlen(line_in);
    if (L >= sizeof line) L = sizeof line - 1;
    memcpy(line, line_in, L);
    line[L] = 0;
    if (L && line[L - 1] == '\n') line[--L] = 0;          /* chomp */
    if (L == 0) return -1;
    /* gi = second field of split(/\|/) */
    char gibuf[256];
    const char *gi_text = NULL;
    const char *b1 = strchr(line, '|');
    if (b1) {
        const char *b2 = strchr(b1 + 1, '|');
        size_t n = b2 ? (size_t)(b2 - b1 - 1) : strlen(b1 + 1);
        if (n > 255) n = 255;
        memcpy(gibuf, b1 + 1, n);
        gibuf[n] = 0;
        gi_text = gibuf;
    }
    const char *f[64];
    size_t fl[64];
    int nf = split_fields(line, f, fl, 64);
    size_t o = 0;
    if (nf > 0) { memcpy(out + o, f[0], fl[0]); o += fl[0]; }
    out[o++] = '\t';
    o += (size_t)txc_lineage(db, gi_text, out + o, cap - o);
    for (int i = 2; i <= 12 && i < nf; i++)
        if (fl[i]) { out[o++] = '\t'; memcpy(out + o, f[i], fl[i]); o += fl[i]; }
    out[o] = 0;
    return (int)o;
}

int txc_file(const char *dir, const char *in_path, const char *out_path)
{
    txc_db *db = txc_load(dir);
    if (!db) return -1;
    FILE *fi = fopen(in_path, "r"), *fo = fopen(out_path, "w");
    if (!fi || !fo) { txc_free(db); return -2; }
    char line[8192], out[16384];
    while (fgets(line, sizeof line, fi)) {
        int n = txc_line(db, line, out, sizeof out);
        if (n < 0) break;
        fwrite(out, 1, (size_t)n, fo);
        fputc('\n', fo);
    }
    fclose(fi); fclose(fo);
    txc_free(db);
    return 0;
}
