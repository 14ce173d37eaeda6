/*
 * uniq_ref.c -- CPU restatement of Scripts/get_uniq.pl (TEST INFRASTRUCTURE ONLY).
 *   :34-39  for every line: `split('\t', $line)`, print the line unless its first column was seen before.
 *           The `chomp` of the script works on $_, not on $line: a line without a TAB keys on its whole text
 *           INCLUDING the newline, and every kept line is printed with its own newline (or without, for a last
 *           line that has none).
 * Pinned by tests/test_uniq_cpu.py against the live script.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* out must have room for len bytes; returns the number of bytes written, -1 on allocation failure */
long long uq_ref_first_hits(const char *text, long long len, char *out, long long *kept_lines, long long *n_kept)
{
    long long nl = 0;
    for (long long i = 0; i < len; i++) nl += text[i] == '\n';
    if (len && text[len - 1] != '\n') nl++;
    uint64_t size = 16;
    while (size < (uint64_t)nl * 2 + 16) size <<= 1;
    long long *tab = (long long *)malloc(size * sizeof(long long));      /* key start + 1 */
    int *klen = (int *)malloc(size * sizeof(int));
    if (!tab || !klen) return -1;
    memset(tab, 0, size * sizeof(long long));
    long long o = 0, kept = 0, line = 0;
    for (long long a = 0; a < len; line++) {
        long long b = a;
        while (b < len && text[b] != '\n') b++;
        if (b < len) b++;                                                 /* the line keeps its newline */
        long long ke = a;
        while (ke < b && text[ke] != '\t') ke++;
        const int kl = (int)(ke - a);
        uint64_t h = 1469598103934665603ULL;
        for (int i = 0; i < kl; i++) { h ^= (unsigned char)text[a + i]; h *= 1099511628211ULL; }
        uint64_t slot = (h ^ (h >> 31)) & (size - 1);
        int seen = 0;
        for (;;) {
            if (!tab[slot]) { tab[slot] = a + 1; klen[slot] = kl; break; }
            if (klen[slot] == kl && memcmp(text + tab[slot] - 1, text + a, (size_t)kl) == 0) { seen = 1; break; }
            slot = (slot + 1) & (size - 1);
        }
        if (!seen) {
            memcpy(out + o, text + a, (size_t)(b - a));
            o += b - a;
            if (kept_lines) kept_lines[kept] = line;
            kept++;
        }
        a = b;
    }
    if (n_kept) *n_kept = kept;
    free(tab);
    free(klen);
    return o;
}
