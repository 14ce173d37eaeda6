/*
 * trim2 -- drop-in for `perl trim2.3.pl -a <read 1> [-b <read 2>] [-g gap] [-t truncate]`
 * (README.md:31-33; Trim/trim2.3.pl and trim2.4.pl) on its QSEQ and FASTQ paths.  Like the script it
 * writes output_files/trim2/<basename of -a>_runblast.fasta under the directory it is started from,
 * prints "QSEQ file format found." / the FASTQ records / "Trimming complete." on stdout, and creates
 * <dir of -a>/singletons/<basename>_single.txt (which the script opens and never writes) for QSEQ.
 * -qc and -lc are accepted and have no effect -- the script's getopts string cannot parse them either.
 * The trim and the join of QSEQ / FASTQ input run on the GPU (pg_trim_join).
 *
 * FASTA input (Trim/trim2.4.pl join_fasta :300-383, parse_fasta :384-466) is two sequential line cursors over two
 * files, restated here in plain C, accidents included, and pinned to the live script by tests/test_trim_cpu.py:
 *   -j   joins record i of -a with record i of -b (header "a_b", sequence a + N x gap + b); the last line of a
 *        multi-line final record is lost, as in the script (it tests eof() before it appends);
 *   -q   "trims" with the quality file read in lockstep -- with the script's bareword cutoffs (LENGTH_CUTOFF and
 *        QUALITY_CUTOFF without '$' are the strings, numerically 0) nothing is ever cut; records go to STDOUT, 60
 *        "bases" per line counted over the window start..end where end assumes 60 qualities per line, the header
 *        is the text up to the first blank, the last record is never printed and _runblast.fasta stays empty.
 */
#include <libgen.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include "pangea_b200.h"
#include "../csrc/pg_perlnum.h"

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


/* ------------------------------------------------------------------ FASTA input: -j and -q (host-side line cursors) */

typedef struct { FILE *f; char *buf; size_t cap; } liner;

/* Perl's <FH>: the next line with its newline, NULL at end of file */
static char *next_line(liner *l, size_t *len)
{
    if (!l->f) return NULL;
    size_t n = 0;
    int c;
    while ((c = fgetc(l->f)) != EOF) {
        if (n + 2 > l->cap) { l->cap = l->cap * 2 + 256; l->buf = (char *)realloc(l->buf, l->cap); }
        l->buf[n++] = (char)c;
        if (c == '\n') break;
    }
    if (n == 0) return NULL;
    l->buf[n] = 0;
    *len = n;
    return l->buf;
}
static int at_eof(liner *l)
{
    int c = fgetc(l->f);
    if (c == EOF) return 1;
    ungetc(c, l->f);
    return 0;
}
typedef struct { char *p; size_t n, cap; int defined; } str;
static void str_set(str *s, const char *t, size_t n, int defined)
{
    if (n + 1 > s->cap) { s->cap = n * 2 + 64; s->p = (char *)realloc(s->p, s->cap); }
    if (n) memcpy(s->p, t, n);
    s->n = n;
    s->p[n] = 0;
    s->defined = defined;
}
static void str_cat(str *s, const char *t, size_t n)
{
    if (s->n + n + 1 > s->cap) { s->cap = (s->n + n) * 2 + 64; s->p = (char *)realloc(s->p, s->cap); }
    if (n) memcpy(s->p + s->n, t, n);
    s->n += n;
    s->p[s->n] = 0;
}
static void str_chomp(str *s) { if (s->n && s->p[s->n - 1] == '\n') s->p[--s->n] = 0; }
static int has_gt(const str *s) { return s->defined && memchr(s->p, '>', s->n) != NULL; }
static void read_into(liner *l, str *s)
{
    size_t n = 0;
    const char *t = next_line(l, &n);
    if (t) str_set(s, t, n, 1); else str_set(s, "", 0, 0);
}

/* join_fasta (:300-383) */
static void join_fasta(const char *pa, const char *pb, int gap_given, int gap, FILE *out)
{
    liner A = {fopen(pa, "rb"), NULL, 0}, B = {fopen(pb, "rb"), NULL, 0};
    str l1 = {0}, l2 = {0}, h1 = {0}, h2 = {0}, seq = {0};
    str_set(&h1, "", 0, 0); str_set(&h2, "", 0, 0); str_set(&seq, "", 0, 1); str_set(&l2, "", 0, 0);
    int first = 1;
    for (;;) {
        read_into(&A, &l1);
        if (!l1.defined) break;
        str_chomp(&l1);
        if (!first) {
            str_chomp(&h1); str_chomp(&h2);
            size_t o = 0;
            for (size_t k = 0; k < h2.n; k++) if (h2.p[k] != '>') h2.p[o++] = h2.p[k];
            h2.n = o; h2.p[o] = 0;
            fwrite(h1.p, 1, h1.n, out); fputc('_', out); fwrite(h2.p, 1, h2.n, out); fputc('\n', out);
        } else {
            read_into(&B, &l2);
            first = 0;
            fwrite(l1.p, 1, l1.n, out); fputc('_', out);
            for (size_t k = 0; k < l2.n; k++) if (l2.p[k] != '>') fputc(l2.p[k], out);
            read_into(&A, &l1);
        }
        while (!has_gt(&l1)) {                             /* an undefined line has no '>' either */
            str_chomp(&l1);
            str_cat(&seq, l1.p, l1.n);
            read_into(&A, &l1);
            if (at_eof(&A)) break;
        }
        str_set(&h1, l1.p, l1.n, l1.defined);
        if (gap_given && gap > 0)
            for (int k = 0; k < gap; k++) str_cat(&seq, "N", 1);
        read_into(&B, &l2);
        while (!has_gt(&l2)) {
            str_chomp(&l2);
            str_cat(&seq, l2.p, l2.n);
            read_into(&B, &l2);
            if (at_eof(&B)) break;
        }
        str_set(&h2, l2.p, l2.n, l2.defined);
        fwrite(seq.p, 1, seq.n, out); fputc('\n', out);
        str_set(&seq, "", 0, 1);
    }
    if (A.f) fclose(A.f);
    if (B.f) fclose(B.f);
    free(A.buf); free(B.buf); free(l1.p); free(l2.p); free(h1.p); free(h2.p); free(seq.p);
}

/* parse_fasta (:384-466); output on stdout */
static void parse_fasta(const char *pa, const char *pq)
{
    liner A = {fopen(pa, "rb"), NULL, 0}, Q = {fopen(pq, "rb"), NULL, 0};
    const double lengthCut = 0.0, qualityCut = 0.0;         /* the barewords LENGTH_CUTOFF / QUALITY_CUTOFF, as numbers */
    int rejected = -1;
    char *trim = NULL;                                      /* @FinalTrim */
    size_t ntrim = 0, captrim = 0;
    long end = 0, start = 0, first = 0, line_num = 0;
    double max = 0.0, sum = 0.0;
    str ls = {0}, lq = {0}, header = {0};
    str_set(&header, "", 0, 0);
    for (;;) {
        read_into(&A, &ls);
        if (!ls.defined) break;
        read_into(&Q, &lq);
        if (memchr(ls.p, '>', ls.n)) {
            const long length = (end + 1) - start;
            if ((double)length < lengthCut) rejected = 1;
            if (rejected == 0) {
                fwrite(header.p, 1, header.n, stdout); fputc('\n', stdout);
                int count_down = 60;
                for (long a = start; a <= end; a++) {
                    count_down--;
                    if (a >= 0 && (size_t)a < ntrim) fputc(trim[a], stdout);
                    if (count_down == 0) { fputc('\n', stdout); count_down = 60; }
                }
                fputc('\n', stdout);
            }
            rejected = 0;
            const char *sp = (const char *)memchr(ls.p, ' ', ls.n);
            str_set(&header, ls.p, sp ? (size_t)(sp - ls.p) + 1 : 0, 1);
            ntrim = 0;
            max = 0.0; sum = 0.0; first = 0; line_num = 0;
        } else if (rejected == 0) {
            const size_t size = ls.n ? ls.n - 1 : 0;        /* split(//) minus the last element (the newline) */
            if (ntrim + size + 1 > captrim) { captrim = (ntrim + size) * 2 + 256; trim = (char *)realloc(trim, captrim); }
            memcpy(trim + ntrim, ls.p, size);
            ntrim += size;
            str_chomp(&lq);
            /* split(/ /, $lineQual): single blanks separate, trailing empty fields are dropped */
            size_t nf = 0, last_nonempty = 0, st = 0;
            for (size_t p = 0; p <= lq.n; p++)
                if (p == lq.n || lq.p[p] == ' ') { nf++; if (p > st) last_nonempty = nf; st = p + 1; }
            if (!lq.defined) last_nonempty = 0;
            st = 0;
            size_t f = 0;
            for (size_t p = 0; p <= lq.n && f < last_nonempty; p++)
                if (p == lq.n || lq.p[p] == ' ') {
                    sum += pg_perl_number(lq.p + st, (int)(p - st)) - qualityCut;
                    if (sum > max) { max = sum; end = (long)f + 60 * line_num; start = first; }
                    if (sum < 0) { sum = 0; first = (long)f + 60 * line_num; }
                    f++;
                    st = p + 1;
                }
            line_num++;
        }
    }
    if (A.f) fclose(A.f);
    if (Q.f) fclose(Q.f);
    free(A.buf); free(Q.buf); free(trim); free(ls.p); free(lq.p); free(header.p);
}

int main(int argc, char **argv)
{
    const char *a = NULL, *b = NULL, *qfile = NULL;
    int gap = 189, truncate = 11, device = 0, join = 0, gap_arg = 0;
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
                else if (first == 'g') { gap = atoi(v); gap_arg = atoi(v); }
                else if (first == 'q') qfile = v;
                else if (first == 't') truncate = atoi(v);
            } else {
                if (!pos) fprintf(stderr, "Unknown option: %c\n", first);
                if (first == 'j') join = 1;
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
    if (alen > 0 && abuf[0] == '>') {                        /* FASTA: -j joins, -q "trims"; both host-side line cursors */
        if (join) {
            if (b) join_fasta(a, b, gap_arg != 0, gap_arg, fo);
            else printf("Error. Input is -j for joining ends, but you did not provided both sequence a and b with -a and -b options.\n\n");
            fclose(fo);
            return 0;                                         /* the script exits here: no "Trimming complete." */
        }
        if (qfile) {
            printf("%s\n", qfile);
            FILE *fq = fopen(qfile, "rb");
            if (!fq) { printf("Error: Unable to open %s required for FASTA file triming.\n", qfile); return 0; }
            fclose(fq);
            parse_fasta(a, qfile);
        } else {
            printf("Error: Please, specify the FASTA quality file with -q option.\n");
            return 0;
        }
        fclose(fo);
        printf("Trimming complete.\n");
        return 0;
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
