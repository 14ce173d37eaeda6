/*
 * megaclust_ref.c -- CPU restatement of Megaclust/megaclust2.pl (TEST INFRASTRUCTURE ONLY: tests/,
 * __graft_entry__.smoke() and the CPU-baseline legs of the benches may use it; the product never does).
 *
 * Follows the script line by line:
 *   :80-82    `next if (/^\#/)`; every other line (blank ones too) counts as examined; chomp
 *   :83-97    split /\t\t|\t\s|\s\t|\t/ into query, subject, pident, ..., evalue, bitscore
 *   :126-131  beyond the thresholds when pident < sim || evalue > eval || bitscore < bits, the fields
 *             read as numbers the way Perl reads strings (leading blanks, sign, digits, point, exponent,
 *             trailing garbage ignored, inf/nan, otherwise 0; a missing field is 0)
 *   :133-143  -c: count every passing line per subject; else once per distinct (subject, query) pair
 *   :146-153  header `OTU<d>times_hit`, then `<subject><d><count>` per subject -- in Perl's hash order,
 *             which is unspecified; this restatement (like the product) lists subjects in order of first
 *             appearance, and the tests compare the script's output as a set of lines.
 * Pinned by tests/test_megaclust_cpu.py against the live script on seeded inputs and on hand-made edge lines.
 */
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int is_space(int c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

/* numeric value of s[0..n) as Perl's `<` / `>` see it */
double mc_ref_number(const char *s, int n)
{
    char buf[512];
    int p = 0, o = 0;
    while (p < n && is_space((unsigned char)s[p])) p++;
    if (p < n && (s[p] == '-' || s[p] == '+')) buf[o++] = s[p++];
    if (n - p >= 3 && tolower((unsigned char)s[p]) == 'i' && tolower((unsigned char)s[p + 1]) == 'n' && tolower((unsigned char)s[p + 2]) == 'f')
        return (o && buf[0] == '-') ? -INFINITY : INFINITY;
    if (n - p >= 3 && tolower((unsigned char)s[p]) == 'n' && tolower((unsigned char)s[p + 1]) == 'a' && tolower((unsigned char)s[p + 2]) == 'n')
        return NAN;
    int digits = 0, point = 0;
    while (p < n && o < 480) {
        if (isdigit((unsigned char)s[p])) { buf[o++] = s[p++]; digits++; }
        else if (s[p] == '.' && !point) { buf[o++] = s[p++]; point = 1; }
        else break;
    }
    if (!digits) return 0.0;
    if (p < n && (s[p] == 'e' || s[p] == 'E')) {
        int q = p + 1;
        if (q < n && (s[q] == '-' || s[q] == '+')) q++;
        if (q < n && isdigit((unsigned char)s[q])) {
            while (p < q && o < 500) buf[o++] = s[p++];
            while (p < n && isdigit((unsigned char)s[p]) && o < 510) buf[o++] = s[p++];
        }
    }
    buf[o] = 0;
    return strtod(buf, NULL);
}

typedef struct { const char *s; int n; } str_t;

static uint64_t hash2(str_t a, str_t b)
{
    uint64_t h = 1469598103934665603ULL;
    for (int i = 0; i < a.n; i++) { h ^= (unsigned char)a.s[i]; h *= 1099511628211ULL; }
    h ^= 0xFF; h *= 1099511628211ULL;
    for (int i = 0; i < b.n; i++) { h ^= (unsigned char)b.s[i]; h *= 1099511628211ULL; }
    return h ^ (h >> 29);
}
static int same(str_t a, str_t b) { return a.n == b.n && memcmp(a.s, b.s, (size_t)a.n) == 0; }

/* whole file in `text`; returns the number of subjects, -1 on allocation failure.  subj_off/len/count
 * (capacity cap) in order of first appearance. */
long long mc_ref_megaclust(const char *text, long long len, double sim, double ev, double bs, int every,
                           long long cap, long long *subj_off, int *subj_len, long long *subj_count,
                           long long *examined, long long *beyond)
{
    long long nl = 0;
    for (long long i = 0; i < len; i++) nl += text[i] == '\n';
    if (len && text[len - 1] != '\n') nl++;
    uint64_t size = 16;
    while (size < (uint64_t)nl * 2 + 16) size <<= 1;
    long long *pair_tab = (long long *)malloc(size * sizeof(long long));      /* line index + 1 */
    long long *subj_tab = (long long *)malloc(size * sizeof(long long));      /* subject index + 1 */
    str_t *pq = (str_t *)malloc(sizeof(str_t) * (size_t)(nl + 1)), *ps = (str_t *)malloc(sizeof(str_t) * (size_t)(nl + 1));
    str_t *subjects = (str_t *)malloc(sizeof(str_t) * (size_t)(nl + 1));
    long long *counts = (long long *)calloc((size_t)(nl + 1), sizeof(long long));
    if (!pair_tab || !subj_tab || !pq || !ps || !subjects || !counts) return -1;
    memset(pair_tab, 0, size * sizeof(long long));
    memset(subj_tab, 0, size * sizeof(long long));
    long long nsub = 0, line = 0, ex = 0, by = 0;
    for (long long a = 0; a < len; line++) {
        long long b = a;
        while (b < len && text[b] != '\n') b++;
        const char *s = text + a;
        int n = (int)(b - a);
        a = b + 1;
        if (n > 0 && s[0] == '#') continue;
        ex++;
        str_t f[13];
        int nf = 0, p = 0, f0 = 0;
        while (nf < 13) {
            int dl = 0;
            if (p < n) {
                if (s[p] == '\t') dl = (p + 1 < n && is_space((unsigned char)s[p + 1])) ? 2 : 1;
                else if (is_space((unsigned char)s[p]) && p + 1 < n && s[p + 1] == '\t') dl = 2;
            }
            if (p >= n || dl) {
                f[nf].s = s + f0; f[nf].n = p - f0; nf++;
                if (p >= n) break;
                p += dl; f0 = p;
            } else p++;
        }
        for (int k = nf; k < 13; k++) { f[k].s = s; f[k].n = 0; }
        if (mc_ref_number(f[2].s, f[2].n) < sim || mc_ref_number(f[10].s, f[10].n) > ev || mc_ref_number(f[11].s, f[11].n) < bs) { by++; continue; }
        pq[line] = f[0];
        ps[line] = f[1];
        int count_it = 1;
        if (!every) {
            uint64_t slot = hash2(f[1], f[0]) & (size - 1);
            for (;;) {
                if (!pair_tab[slot]) { pair_tab[slot] = line + 1; break; }
                const long long o = pair_tab[slot] - 1;
                if (same(ps[o], f[1]) && same(pq[o], f[0])) { count_it = 0; break; }
                slot = (slot + 1) & (size - 1);
            }
        }
        if (!count_it) continue;
        str_t none = {"", 0};
        uint64_t slot = hash2(f[1], none) & (size - 1);
        for (;;) {
            if (!subj_tab[slot]) { subjects[nsub] = f[1]; subj_tab[slot] = ++nsub; }
            const long long o = subj_tab[slot] - 1;
            if (same(subjects[o], f[1])) { counts[o]++; break; }
            slot = (slot + 1) & (size - 1);
        }
    }
    if (examined) *examined = ex;
    if (beyond) *beyond = by;
    for (long long i = 0; i < nsub && i < cap; i++) {
        subj_off[i] = subjects[i].s - text;
        subj_len[i] = subjects[i].n;
        subj_count[i] = counts[i];
    }
    free(pair_tab); free(subj_tab); free(pq); free(ps); free(subjects); free(counts);
    return nsub;
}
