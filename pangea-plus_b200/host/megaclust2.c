/*
 * megaclust2 -- drop-in for `perl megaclust2.pl -i <in> -o <out> [-s sim] [-e evalue] [-b bitscore] [-d delim] [-c x]`
 * (README.md:176; Megaclust/megaclust2.pl).  Same options (Getopt::Std 'i:o:s:e:b:c:d:h' -- note that -c
 * takes a value there, and that an option whose value is "" or "0" counts as not given), same usage and
 * run-summary text on stdout, same output file: a header line `OTU<d>times_hit` and one `<subject><d><count>`
 * line per subject that passed the thresholds.  The script prints the subjects in Perl's hash order, which
 * changes from run to run; this tool prints them in order of first appearance.  Thresholding and counting
 * run on the GPU (pg_megaclust).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pangea_b200.h"
#include "../csrc/pg_perlnum.h"

static const char USAGE[] =
    "Usage:\n\t\t   cluster-blast-output.pl -i infile -o outfile [options]\n\t\t   \n\t\t   Required options:\n"
    "\t\t   -i input BLAST tabular results file (megablast or blastall -m 8)\n\t\t   -o output file name\n\n"
    "\t\t   Optional parameters:\n\t\t   -s similarity lower threshold (percent, between 0-100) (default 95)\n"
    "\t\t   -e e-value upper threshold (default 1e-20)\n\t\t   -b bitscore lower threshold (default 200)\n"
    "\t\t   -d delimiter (default to comma)\n\t\t   \n\t\t   Optional switches:\n"
    "\t\t   -c count every query hit (if -c not given, then only count\n\t\t\t\t\t     any query-genome pair as one genome hit)\n"
    "\t\t   -h print usage summary\n\t\t   \n";

static int truthy(const char *v) { return v && v[0] && !(v[0] == '0' && !v[1]); }   /* Perl: "" and "0" are false */

static char *slurp(const char *path, int64_t *len)
{
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *b = (char *)malloc((size_t)n + 1);
    if (n && fread(b, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(b); return NULL; }
    fclose(f);
    b[n] = 0;
    *len = n;
    return b;
}

int main(int argc, char **argv)
{
    const char *in = NULL, *out = NULL, *s = NULL, *e = NULL, *b = NULL, *c = NULL, *d = NULL;
    int help = 0, device = 0;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (strcmp(a, "--device") == 0 && i + 1 < argc) { device = atoi(argv[++i]); continue; }
        if (a[0] != '-' || !a[1]) break;                      /* first non-option word ends the parse */
        if (strcmp(a, "--") == 0) break;
        const char o = a[1];
        if (o == 'h') { help = 1; continue; }
        if (!strchr("iosebcd", o)) continue;                  /* unknown option: Getopt::Std warns and goes on */
        const char *v = a[2] ? a + 2 : (i + 1 < argc ? argv[++i] : "");
        switch (o) {
        case 'i': in = v; break;
        case 'o': out = v; break;
        case 's': s = v; break;
        case 'e': e = v; break;
        case 'b': b = v; break;
        case 'c': c = v; break;
        case 'd': d = v; break;
        }
    }
    if (help) { printf("%s\n", USAGE); return 0; }
    if (!truthy(in) || !truthy(out)) {
        printf("Must specify both an input and output filename\n%s\n", USAGE);
        return 0;
    }
    pg_megaclust_opts opts;
    memset(&opts, 0, sizeof opts);
    opts.sim_threshold = 95.0;
    opts.eval_threshold = pg_perl_number("1e-20", 5);
    opts.bitscore_threshold = 200.0;
    if (truthy(s)) {
        const double v = pg_perl_number(s, (int)strlen(s));
        if (v < 0 || v > 100) {
            printf("similarity threshold must be between 0 and 100\n%s\n", USAGE);
            return 0;
        }
        opts.sim_threshold = v;
    }
    if (truthy(e)) opts.eval_threshold = pg_perl_number(e, (int)strlen(e));
    if (truthy(b)) opts.bitscore_threshold = pg_perl_number(b, (int)strlen(b));
    opts.count_every_hit = truthy(c);
    const char *delim = truthy(d) ? d : ",";

    int64_t len = 0;
    char *text = slurp(in, &len);
    if (!text) { fprintf(stderr, "couldn't open infile at megaclust2 line 72.\n"); return 2; }
    FILE *fo = fopen(out, "wb");
    if (!fo) { fprintf(stderr, "couldn't open outfile at megaclust2 line 73.\n"); return 2; }

    pg_ctx *ctx = pg_init(device);
    if (!ctx) { fprintf(stderr, "megaclust2: %s\n", pg_last_error(NULL)); return 1; }
    int64_t cap = 1 << 16, notu = 0, examined = 0, beyond = 0;
    int64_t *off = NULL, *cnt = NULL;
    int32_t *ln = NULL;
    int rc;
    for (;;) {
        off = (int64_t *)realloc(off, (size_t)cap * 8);
        cnt = (int64_t *)realloc(cnt, (size_t)cap * 8);
        ln = (int32_t *)realloc(ln, (size_t)cap * 4);
        rc = pg_megaclust(ctx, text, len, &opts, cap, &notu, off, ln, cnt, &examined, &beyond);
        if (rc != PG_ERANGE) break;
        cap = notu + 16;
    }
    if (rc != PG_OK) { fprintf(stderr, "megaclust2: %s\n", pg_last_error(ctx)); return 1; }
    fprintf(fo, "OTU%stimes_hit\n", delim);
    for (int64_t i = 0; i < notu; i++) {
        fwrite(text + off[i], 1, (size_t)ln[i], fo);
        fprintf(fo, "%s%lld\n", delim, (long long)cnt[i]);
    }
    fclose(fo);
    printf("Run complete:\n%lld hits examined\n%lld hits beyond thresholds and therefore not counted.\n",
           (long long)examined, (long long)beyond);
    pg_shutdown(ctx);
    return 0;
}
