/*
 * trim2 -- drop-in for `perl trim2.3.pl -a <read 1> [-b <read 2>] [-g gap] [-t truncate]`
 * (README.md:31-33; Trim/trim2.3.pl and trim2.4.pl) on its QSEQ and FASTQ paths.  Like the script it
 * writes output_files/trim2/<basename of -a>_runblast.fasta under the directory it is started from,
 * prints "QSEQ file format found." / the FASTQ records / "Trimming complete." on stdout, and creates
 * <dir of -a>/singletons/<basename>_single.txt (which the script opens and never writes) for QSEQ.
 * -qc and -lc are accepted and have no effect -- the script's getopts string cannot parse them either.
 * The FASTA + quality-file path (-q) and -j are not built: the tool says so and exits.
 * The trim and the join run on the GPU (pg_trim_join).
 */
#include <libgen.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include "pangea_b200.h"

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
    const char *a = NULL, *b = NULL;
    int gap = 189, truncate = 11, device = 0;
    /* Getopt::Std with 'a:b:g:t:q:qc:lc:j': -a -b -g -t -q -c take a value, -l and -j do not; the
     * first non-option word ends the parse.  "-qc 25" is therefore -q with the value "c". */
    for (int i = 1; i < argc; i++) {
        const char *s = argv[i];
        if (s[0] != '-' || !s[1]) break;
        if (strcmp(s, "--device") == 0 && i + 1 < argc) { device = atoi(argv[++i]); continue; }
        char o = s[1];
        if (strchr("abgtqc", o)) {
            const char *v = s[2] ? s + 2 : (i + 1 < argc ? argv[++i] : "");
            if (o == 'a') a = v;
            else if (o == 'b') b = v;
            else if (o == 'g') gap = atoi(v);
            else if (o == 't') truncate = atoi(v);
        } else if (o != 'j' && o != 'l') break;
    }
    if (!a) {
        printf("Usage: perl trim2.pl \n\t-a raw illumina input file read 1\n\t-b raw illumina input file read 2 (if any) \n"
               "\t-g size of GAP between paired-ends (if any) \n\t-t truncate size (if any)\n\t-q quality file (in case of FASTA input)\n"
               "\t-qc quality cutoff value\n\t-j use this option for just joining a and b, without triming\n\t-lc minimum length \n");
        printf("Supported formats: FASTA, FASTQ and QSEQ.\n");
        return 0;
    }
    if (gap == 0) gap = 189;                              /* `if ($parameters{g})`: -g 0 keeps the default */
    if (truncate == 0) truncate = 11;
    int64_t alen = 0, blen = 0;
    char *abuf = slurp(a, &alen), *bbuf = NULL;
    if (!abuf) { printf("Error: Unable to open %s.\n", a); return 0; }
    if (b) {
        bbuf = slurp(b, &blen);
        if (!bbuf) { printf("Error: Unable to open %s.\n", b); return 0; }
    }
    char *a1 = strdup(a), *a2 = strdup(a);
    const char *prefix = basename(a1), *dir = dirname(a2);
    mkdir("output_files", 0777);
    mkdir("output_files/trim2", 0777);
    char outpath[4096];
    snprintf(outpath, sizeof outpath, "output_files/trim2/%s_runblast.fasta", prefix);
    FILE *fo = fopen(outpath, "wb");
    if (!fo) { fprintf(stderr, "trim2: cannot write %s\n", outpath); return 1; }
    if (alen > 0 && abuf[0] == '>') {
        fprintf(stderr, "trim2: the FASTA (+ quality file / -j) paths of trim2.pl are not built in this tool\n");
        return 1;
    }
    int fastq = alen > 0 && abuf[0] == '@';
    if (!fastq) {
        /* the script's QSEQ test: column 8 is 1 or 2 and column 11 is 0 or 1 on the first line */
        const char *p = abuf, *e = (const char *)memchr(abuf, '\n', (size_t)alen);
        if (!e) e = abuf + alen;
        const char *f[12] = {0};
        int nf = 0;
        f[nf++] = p;
        for (; p < e && nf < 12; p++)
            if (*p == '\t') f[nf++] = p + 1;
        int ok = nf >= 11 && (f[7][0] == '1' || f[7][0] == '2') && f[7][1] == '\t' && (f[10][0] == '0' || f[10][0] == '1') &&
                 (f[10] + 1 == e || f[10][1] == '\r');
        if (!ok || !bbuf) {
            printf("Error: file format not recognized.\n");
            printf("Trimming complete.\n");
            return 0;
        }
        printf("QSEQ file format found.\n");
        char sdir[4096], sfile[4400];
        snprintf(sdir, sizeof sdir, "%s/singletons", dir);
        mkdir(sdir, 0777);
        snprintf(sfile, sizeof sfile, "%s/%s_single.txt", sdir, prefix);
        FILE *fs = fopen(sfile, "w");
        if (fs) fclose(fs);
    }
    pg_ctx *ctx = pg_init(device);
    if (!ctx) { fprintf(stderr, "trim2: %s\n", pg_last_error(NULL)); return 1; }
    pg_trim_opts opts;
    memset(&opts, 0, sizeof opts);
    opts.gap = gap;
    opts.truncate = truncate;
    int64_t cap = 2 * (alen + blen) + 4096, n = 0;
    char *out = (char *)malloc((size_t)cap);
    int rc = pg_trim_join(ctx, abuf, alen, bbuf, blen, b != NULL, &opts, out, cap, &n, NULL);
    if (rc == PG_ERANGE) {
        cap = n + 64;
        out = (char *)realloc(out, (size_t)cap);
        rc = pg_trim_join(ctx, abuf, alen, bbuf, blen, b != NULL, &opts, out, cap, &n, NULL);
    }
    if (rc != PG_OK) { fprintf(stderr, "trim2: %s\n", pg_last_error(ctx)); return 1; }
    fwrite(out, 1, (size_t)n, fo);
    fclose(fo);
    if (fastq) fwrite(out, 1, (size_t)n, stdout);          /* parse_fastq prints every record to stdout as well */
    printf("Trimming complete.\n");
    pg_shutdown(ctx);
    return 0;
}
