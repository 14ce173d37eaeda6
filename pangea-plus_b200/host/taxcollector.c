/*
 * taxcollector -- drop-in for `perl NCBI-taxcollector-0.01.pl -f <hits> -o <out>`
 * (Tax_class/NCBI-taxcollector-0.01.pl; README.md:98-112).  Same flags, same output lines:
 *     <query id> TAB <lineage> (TAB <BLAST column>)*          (rows B5-B7 of SURVEY.md)
 * Like the script it looks for the three .bin tables in ./Tax_class/ relative to the
 * directory it is started from (the script chdir()s there after opening its files,
 * :31-47); --taxdir overrides that.  All lookups of a run are done in batches on the GPU
 * (pg_tax_lineage); the script's per-hit progress chatter on stdout is reduced to one
 * summary line.  An empty input line ends the run, as it does in the script (:82-86).
 */
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pangea_b200.h"
#include "pg_host_common.h"

/* Perl split(/ |\t\t|\t/): leading empty field kept, trailing empties dropped */
static int split_fields(const char *s, size_t n, const char **f, size_t *fl, int maxf)
{
    int nf = 0;
    size_t p = 0, start = 0;
    while (p < n) {
        size_t sep = 0;
        if (s[p] == ' ') sep = 1;
        else if (s[p] == '\t') sep = (p + 1 < n && s[p + 1] == '\t') ? 2 : 1;
        if (sep) {
            if (nf < maxf) { f[nf] = s + start; fl[nf] = p - start; nf++; }
            p += sep;
            start = p;
        } else p++;
    }
    if (nf < maxf) { f[nf] = s + start; fl[nf] = n - start; nf++; }
    while (nf > 0 && fl[nf - 1] == 0) nf--;
    return nf;
}

int main(int argc, char **argv)
{
    static struct option lo[] = {{"taxdir", required_argument, 0, 'T'}, {"device", required_argument, 0, 'G'}, {0, 0, 0, 0}};
    const char *in = NULL, *out = NULL, *taxdir = "./Tax_class";
    int device = 0;
    for (;;) {
        int c = getopt_long(argc, argv, "f:o:", lo, NULL);
        if (c == -1) break;
        if (c == 'f') in = optarg;
        else if (c == 'o') out = optarg;
        else if (c == 'T') taxdir = optarg;
        else if (c == 'G') device = atoi(optarg);
    }
    if (!in || !out) {
        printf("Usage: perl taxcollector_ncbi-0.01.pl \n\t-f Classification results (tabular text file)\n\t-o Output file \n");
        return 0;
    }
    pg_lines L;
    if (pg_lines_read(in, &L)) { printf("Error: Unable to open classification results file %s.\n", in); return 0; }
    FILE *fo = fopen(out, "w");
    if (!fo) { printf("Error: Unable to open output file %s.\n", out); return 0; }

    pg_ctx *ctx = pg_init(device);
    if (!ctx) { fprintf(stderr, "taxcollector: %s\n", pg_last_error(NULL)); return 1; }
    pg_tax *tx = NULL;
    if (pg_tax_load(ctx, taxdir, &tx) != PG_OK) { fprintf(stderr, "taxcollector: %s\n", pg_last_error(ctx)); return 1; }

    int64_t nl = L.count;
    for (int64_t i = 0; i < nl; i++)
        if (L.len[i] == 0) { nl = i; break; }            /* `if (@gi) ... else { exit }` */
    const int64_t CH = 1 << 22;
    int32_t *gi = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nl < CH ? nl + 1 : CH));
    int64_t *off = (int64_t *)malloc(sizeof(int64_t) * (size_t)((nl < CH ? nl : CH) + 1));
    size_t cap = (size_t)64 << 20;
    char *lin = (char *)malloc(cap);
    int64_t unidentified = 0;
    for (int64_t c0 = 0; c0 < nl; c0 += CH) {
        int64_t cn = nl - c0 < CH ? nl - c0 : CH;
        for (int64_t i = 0; i < cn; i++) {
            /* gi = second field of split(/\|/, line), through atoi like tax_class does */
            const char *s = L.line[c0 + i];
            size_t n = L.len[c0 + i];
            const char *b1 = (const char *)memchr(s, '|', n);
            long v = 0;
            if (b1) {
                char tmp[32];
                size_t m = n - (size_t)(b1 + 1 - s);
                if (m > 31) m = 31;
                memcpy(tmp, b1 + 1, m);
                tmp[m] = 0;
                v = atol(tmp);
            }
            gi[i] = (v >= 1 && v <= 0x7fffffffL) ? (int32_t)v : 0;
        }
        int rc;
        while ((rc = pg_tax_lineage(ctx, tx, gi, cn, lin, (int64_t)cap, off)) == PG_ERANGE && (size_t)off[cn] > cap) {
            cap = (size_t)off[cn] + 1024;
            lin = (char *)realloc(lin, cap);
        }
        if (rc != PG_OK) { fprintf(stderr, "taxcollector: %s\n", pg_last_error(ctx)); return 1; }
        for (int64_t i = 0; i < cn; i++) {
            const char *s = L.line[c0 + i];
            size_t n = L.len[c0 + i];
            const char *f[64];
            size_t fl[64];
            int nf = split_fields(s, n, f, fl, 64);
            if (nf > 0) fwrite(f[0], 1, fl[0], fo);
            fputc('\t', fo);
            if (gi[i] == 0) {
                /* no usable gi in the line: outside the script's defined behaviour (it recurses on
                 * tax_class's usage text there); printed as an unidentified hit with the original text */
                const char *b1 = (const char *)memchr(s, '|', n);
                fputs("Unidentified(GI:", fo);
                if (b1) {
                    const char *e = (const char *)memchr(b1 + 1, '|', n - (size_t)(b1 + 1 - s));
                    fwrite(b1 + 1, 1, e ? (size_t)(e - b1 - 1) : n - (size_t)(b1 + 1 - s), fo);
                }
                fputs(");", fo);
                unidentified++;
            } else {
                fwrite(lin + off[i], 1, (size_t)(off[i + 1] - off[i]), fo);
                if (off[i + 1] - off[i] >= 13 && memcmp(lin + off[i], "Unidentified(", 13) == 0) unidentified++;
            }
            for (int k = 2; k <= 12 && k < nf; k++)
                if (fl[k]) { fputc('\t', fo); fwrite(f[k], 1, fl[k], fo); }
            fputc('\n', fo);
        }
    }
    fclose(fo);
    printf("taxcollector: %lld hit lines, %lld without a taxid\n", (long long)nl, (long long)unidentified);
    free(gi); free(off); free(lin);
    pg_lines_free(&L);
    pg_tax_free(tx);
    pg_shutdown(ctx);
    return 0;
}
