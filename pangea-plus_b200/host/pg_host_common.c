/* pg_host_common.c -- FASTA / line readers for the drop-in executables. */
#include "pg_host_common.h"
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

static char *dup_n(const char *s, size_t n)
{
    char *r = (char *)malloc(n + 1);
    memcpy(r, s, n);
    r[n] = 0;
    return r;
}

int pg_fasta_read(const char *path, pg_fasta *out)
{
    memset(out, 0, sizeof *out);
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)sz + 1);
    if (sz && fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(buf); return -1; }
    fclose(f);
    buf[sz] = 0;
    int64_t cap = 1024, n = 0;
    out->bytes = (char *)malloc((size_t)sz + 1);
    out->off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cap + 1));
    out->id = (char **)malloc(sizeof(char *) * (size_t)cap);
    out->header = (char **)malloc(sizeof(char *) * (size_t)cap);
    int64_t nb = 0;
    long p = 0;
    while (p < sz) {
        char *nl = (char *)memchr(buf + p, '\n', (size_t)(sz - p));
        long e = nl ? (long)(nl - buf) : sz;
        long le = e;
        while (le > p && (buf[le - 1] == '\r')) le--;
        if (buf[p] == '>') {
            if (n == cap) {
                cap *= 2;
                out->off = (int64_t *)realloc(out->off, sizeof(int64_t) * (size_t)(cap + 1));
                out->id = (char **)realloc(out->id, sizeof(char *) * (size_t)cap);
                out->header = (char **)realloc(out->header, sizeof(char *) * (size_t)cap);
            }
            out->off[n] = nb;
            out->header[n] = dup_n(buf + p + 1, (size_t)(le - p - 1));
            long q = p + 1;
            while (q < le && !isspace((unsigned char)buf[q])) q++;
            out->id[n] = dup_n(buf + p + 1, (size_t)(q - p - 1));
            n++;
        } else if (n > 0) {
            /* residue lines almost never hold blanks: two vectorised scans, then one memcpy */
            size_t ll = (size_t)(le - p);
            if (!memchr(buf + p, ' ', ll) && !memchr(buf + p, '\t', ll)) {
                memcpy(out->bytes + nb, buf + p, ll);
                nb += (int64_t)ll;
            } else {
                for (long k = p; k < le; k++)
                    if (!isspace((unsigned char)buf[k])) out->bytes[nb++] = buf[k];
            }
        }
        p = e + 1;
    }
    out->off[n] = nb;
    out->count = n;
    free(buf);
    return 0;
}

void pg_fasta_free(pg_fasta *f)
{
    if (!f) return;
    for (int64_t i = 0; i < f->count; i++) { free(f->id[i]); free(f->header[i]); }
    free(f->bytes); free(f->off); free(f->id); free(f->header);
    memset(f, 0, sizeof *f);
}

int pg_lines_read(const char *path, pg_lines *out)
{
    memset(out, 0, sizeof *out);
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    out->buf = (char *)malloc((size_t)sz + 1);
    if (sz && fread(out->buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(out->buf); return -1; }
    fclose(f);
    out->buf[sz] = 0;
    int64_t cap = 1024, n = 0;
    out->line = (char **)malloc(sizeof(char *) * (size_t)cap);
    out->len = (size_t *)malloc(sizeof(size_t) * (size_t)cap);
    long p = 0;
    while (p < sz) {
        char *nl = (char *)memchr(out->buf + p, '\n', (size_t)(sz - p));
        long e = nl ? (long)(nl - out->buf) : sz;
        if (n == cap) {
            cap *= 2;
            out->line = (char **)realloc(out->line, sizeof(char *) * (size_t)cap);
            out->len = (size_t *)realloc(out->len, sizeof(size_t) * (size_t)cap);
        }
        out->line[n] = out->buf + p;
        out->len[n] = (size_t)(e - p);
        n++;
        p = e + 1;
    }
    out->count = n;
    return 0;
}

void pg_lines_free(pg_lines *l)
{
    if (!l) return;
    free(l->buf); free(l->line); free(l->len);
    memset(l, 0, sizeof *l);
}

void pg_fmt_conf(int votes, char out[8])
{
    if (votes >= 100) { strcpy(out, "1.0"); return; }
    if (votes <= 0) { strcpy(out, "0.0"); return; }
    if (votes % 10 == 0) snprintf(out, 8, "0.%d", votes / 10);
    else snprintf(out, 8, "0.%02d", votes);
}
