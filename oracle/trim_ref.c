/*
 * trim_ref.c -- CPU ORACLE for the Trim join: Trim/trim2.4.pl
 * (identical to trim2.3.pl on these paths), QSEQ pairs (parse_qseq :169-242, trim_qseq
 * :244-298) and FASTQ (parse_fastq :467-521, trim_fastq :527-578).
 *
 * TEST INFRASTRUCTURE ONLY.  Restates the Perl statement by statement, including its
 * accidents, because they decide the bytes of <prefix>_runblast.fasta:
 *   - getopts('a:b:g:t:q:qc:lc:j') cannot parse -qc / -lc, so the cutoffs are always 20 / 70;
 *   - trim_qseq drops TRUNCATE bases from sequence and quality and then another TRUNCATE-1
 *     from the sequence only; the kept prefix has `end` characters, `end` = index of the
 *     running-sum maximum (sum restarts at 0 whenever it goes negative);
 *   - `$seq[$a] = "N"` writes to an unrelated array: no base is ever masked;
 *   - trim_fastq scans the quality LINE including its newline, returns "SEQ\t" (the global
 *     $qual is unset on this path) and the caller strips the blank from mate 1 only, so a
 *     paired FASTQ record ends in TAB NEWLINE; with -b the second mate is the NEXT record of
 *     the same -a file (READ2 is opened and never read);
 *   - a mate shorter than 70 is "0": QSEQ drops the pair, FASTQ prints the 0.
 * PINNED: byte-identical to the real script on seeded inputs (tests/test_trim_cpu.py) and
 * golden files under tests/golden/trim/.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define QUALITY_CUTOFF 20
#define LENGTH_CUTOFF  70

typedef struct { char *buf; char **line; size_t *len; int64_t n; } lines_t;   /* lines WITHOUT the newline; has_nl says if one followed */

static int read_lines(const char *path, lines_t *L)
{
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    L->buf = (char *)malloc((size_t)sz + 1);
    if (sz && fread(L->buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); return -1; }
    fclose(f);
    int64_t cap = 1024, n = 0;
    L->line = (char **)malloc(sizeof(char *) * (size_t)cap);
    L->len = (size_t *)malloc(sizeof(size_t) * (size_t)cap);
    long p = 0;
    while (p < sz) {
        char *nl = (char *)memchr(L->buf + p, '\n', (size_t)(sz - p));
        long e = nl ? (long)(nl - L->buf) : sz;
        if (n == cap) {
            cap *= 2;
            L->line = (char **)realloc(L->line, sizeof(char *) * (size_t)cap);
            L->len = (size_t *)realloc(L->len, sizeof(size_t) * (size_t)cap);
        }
        L->line[n] = L->buf + p;
        L->len[n] = (size_t)(e - p);
        n++;
        p = e + 1;
    }
    L->n = n;
    return 0;
}

/* index of the running-sum maximum over q[0..n) with per-character score q - base - cutoff;
 * extra = one more character scored after the line (the newline on the FASTQ path), -1 for none */
static long best_end(const char *q, size_t n, int base, int extra)
{
    long sum = 0, max = 0, end = 0;
    for (size_t a = 0; a < n + (extra >= 0 ? 1 : 0); a++) {
        int c = a < n ? (unsigned char)q[a] : extra;
        sum += c - base - QUALITY_CUTOFF;
        if (sum > max) { max = sum; end = (long)a; }
        if (sum < 0) sum = 0;
    }
    return end;
}

/* tab-separated fields of one QSEQ line (trailing empty fields dropped like Perl's split) */
static int qseq_fields(const char *s, size_t n, const char **f, size_t *fl)
{
    int nf = 0;
    size_t start = 0;
    for (size_t p = 0; p <= n && nf < 16; p++)
        if (p == n || s[p] == '\t') { f[nf] = s + start; fl[nf] = p - start; nf++; start = p + 1; }
    while (nf > 0 && fl[nf - 1] == 0) nf--;
    return nf;
}

/* trim_qseq: returns the kept sequence length, or -1 for "0" (too short) */
static long trim_qseq(const char *seq, size_t sl, const char *qual, size_t ql, int truncate, const char **out)
{
    size_t s0 = (size_t)truncate <= sl ? (size_t)truncate : sl;       /* substr($seq, $TRUNCATE) */
    size_t q0 = (size_t)truncate <= ql ? (size_t)truncate : ql;
    size_t s_len = sl - s0, q_len = ql - q0;
    size_t cut = (size_t)(truncate - 1) <= s_len ? (size_t)(truncate - 1) : s_len;   /* substr($seq,0,TRUNCATE-1) = '' */
    if (truncate - 1 < 0) cut = 0;
    s0 += cut;
    s_len -= cut;
    long end = best_end(qual + q0, q_len, 64, -1);
    size_t keep = (size_t)end <= s_len ? (size_t)end : s_len;
    *out = seq + s0;
    if (keep < LENGTH_CUTOFF) return -1;
    return (long)keep;
}

static void put_n(FILE *fo, int gap) { for (int r = 0; r < gap; r++) fputc('N', fo); }

/* format: 0 = QSEQ pairs (a + b), 1 = FASTQ from `a`; paired as given by -b.  Returns 0. */
int trim_run(const char *a_path, const char *b_path, int gap, int truncate, const char *out_path)
{
    lines_t A, B;
    memset(&A, 0, sizeof A);
    memset(&B, 0, sizeof B);
    if (read_lines(a_path, &A)) return -1;
    FILE *fo = fopen(out_path, "wb");
    if (!fo) return -2;
    if (A.n > 0 && A.len[0] > 0 && A.line[0][0] == '@') {
        /* ---- FASTQ */
        int paired = b_path != NULL;
        for (int64_t i = 0; i < A.n; ) {
            const char *h = A.line[i];
            size_t hl = A.len[i];
            const char *sq = i + 1 < A.n ? A.line[i + 1] : "";
            size_t sl = i + 1 < A.n ? A.len[i + 1] : 0;
            const char *q = i + 3 < A.n ? A.line[i + 3] : "";
            size_t ql = i + 3 < A.n ? A.len[i + 3] : 0;
            int q_has_nl = i + 3 < A.n - 1 || (i + 3 == A.n - 1 && A.buf[(A.line[i + 3] - A.buf) + ql] == '\n');
            i += 4;
            /* the sequence line still carries its newline inside trim_fastq: substr may reach it */
            long end = best_end(q, ql, 33, q_has_nl ? '\n' : -1);
            size_t avail = sl + 1;                                   /* sequence + its newline */
            size_t keep = (size_t)end <= avail ? (size_t)end : avail;
            fputc('>', fo);
            for (size_t k = 0; k < hl; k++) if (h[k] != '@') fputc(h[k], fo);
            fputs(":AB\n", fo);
            if (keep < LENGTH_CUTOFF) fputc('0', fo);
            else fwrite(sq, 1, keep <= sl ? keep : sl, fo);          /* s/\s//g removes a captured newline */
            if (paired) {
                const char *sq2 = i + 1 < A.n ? A.line[i + 1] : "";
                size_t sl2 = i + 1 < A.n ? A.len[i + 1] : 0;
                const char *q2 = i + 3 < A.n ? A.line[i + 3] : "";
                size_t ql2 = i + 3 < A.n ? A.len[i + 3] : 0;
                int q2_has_nl = i + 3 < A.n - 1 || (i + 3 == A.n - 1 && A.buf[(A.line[i + 3] - A.buf) + ql2] == '\n');
                int have2 = i < A.n;
                i += 4;
                put_n(fo, gap);
                if (have2) {
                    long end2 = best_end(q2, ql2, 33, q2_has_nl ? '\n' : -1);
                    size_t avail2 = sl2 + 1;
                    size_t keep2 = (size_t)end2 <= avail2 ? (size_t)end2 : avail2;
                    if (keep2 < LENGTH_CUTOFF) fputc('0', fo);
                    else {
                        fwrite(sq2, 1, keep2 <= sl2 ? keep2 : sl2, fo);
                        if (keep2 > sl2) fputc('\n', fo);            /* the newline is inside the kept prefix */
                        fputc('\t', fo);                             /* "SEQ\t": only mate 1 is blank-stripped */
                    }
                } else {
                    fputc('0', fo);                                  /* trim_fastq(undef, undef): empty -> too short */
                }
                fputc('\n', fo);
            } else {
                fputc('\n', fo);
            }
        }
    } else {
        /* ---- QSEQ pairs */
        if (!b_path || read_lines(b_path, &B)) { fclose(fo); return -3; }
        for (int64_t i = 0; i < A.n; i++) {
            const char *f1[16], *f2[16];
            size_t l1[16], l2[16];
            int n1 = qseq_fields(A.line[i], A.len[i], f1, l1);
            int n2 = i < B.n ? qseq_fields(B.line[i], B.len[i], f2, l2) : 0;
            for (int k = n1; k < 16; k++) { f1[k] = ""; l1[k] = 0; }
            for (int k = n2; k < 16; k++) { f2[k] = ""; l2[k] = 0; }
            const char *s1 = f1[8], *s2 = f2[8];
            long k1 = (long)l1[8], k2 = (long)l2[8];
            int trimmed = (l1[7] == 1 && f1[7][0] == '1');
            if (trimmed) {
                k1 = trim_qseq(f1[8], l1[8], f1[9], l1[9], truncate, &s1);
                k2 = trim_qseq(f2[8], l2[8], f2[9], l2[9], truncate, &s2);
            } else {
                /* untrimmed lines: "0" only if the sequence field literally is "0" */
                if (l1[8] == 1 && f1[8][0] == '0') k1 = -1;
                if (l2[8] == 1 && f2[8][0] == '0') k2 = -1;
            }
            if (k1 < 0 || k2 < 0) continue;                         /* one mate too short: the pair is dropped */
            fputc('>', fo);
            for (int k = 0; k < 8; k++) { if (k) fputc(':', fo); fwrite(f1[k], 1, l1[k], fo); }
            fputs(":AB\n", fo);
            for (long k = 0; k < k1; k++) fputc(trimmed && s1[k] == '.' ? 'N' : s1[k], fo);
            put_n(fo, gap);
            for (long k = 0; k < k2; k++) fputc(trimmed && s2[k] == '.' ? 'N' : s2[k], fo);
            fputc('\n', fo);
        }
    }
    fclose(fo);
    return 0;
}
