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
    /* Getopt::Std::getopts('a:b:g:t:q:qc:lc:j'), restated: a word "-Xrest" is taken apart letter by letter.  A letter
     * followed by ':' in the spec (a b g t q c) takes the rest of the word, or else the next word, as its value; l and j
     * are switches and the rest of the word is parsed on ("-lc 80" = -l, then -c 80); an unknown letter is reported
     * on stderr and skipped; "--" or the first word that is not an option ends the parse.  "-qc 25" is therefore -q
     * with the value "c".  (--device N is this tool's own extension.) */
    static const char spec[] = "a:b:g:t:q:qc:lc:j";
    {
        int i = 1;
        char *word = NULL;                                /* the remainder of a bundled word, re-queued as "-rest" */
        while (word || i < argc) {
            const char *s = word ? word : argv[i];
            if (!word && strcmp(s, "--device") == 0 && i + 1 < argc) { device = atoi(argv[i + 1]); i += 2; continue; }
            if (s[0] != '-' || !s[1]) break;
            if (strcmp(s, "--") == 0) break;
            const char first = s[1];
            const char *rest = s + 2;
            const char *pos = strchr(spec, first);
            char *next_word = NULL;
            if (pos && pos[1] == ':') {
                const char *v = rest;
                if (!word) i++;
                if (!*v) v = i < argc ? argv[i++] : "";
                if (first == 'a') a = v;
                else if (first == 'b') b = v;
                else if (first == 'g') gap = atoi(v);
                else if (first == 't') truncate = atoi(v);
            } else {
                if (!pos) fprintf(stderr, "Unknown option: %c\n", first);
                if (*rest) {
                    next_word = (char *)malloc(strlen(rest) + 2);
                    next_word[0] = '-';
                    strcpy(next_word + 1, rest);
                }
                if (!word) i++;
            }
            free(word);
            word = next_word;
        }
        free(word);
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
