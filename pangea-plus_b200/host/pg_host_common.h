/* pg_host_common.h -- file parsing shared by the drop-in executables (plain C host code). */
#ifndef PG_HOST_COMMON_H
#define PG_HOST_COMMON_H
#include <stdint.h>
#include <stdio.h>
#include <stddef.h>

typedef struct {
    char    *bytes;      /* residues of all records, concatenated, no newlines */
    int64_t *off;        /* count+1 */
    char   **id;         /* first token of the header */
    char   **header;     /* whole header line without '>' */
    int64_t  count;
} pg_fasta;

int  pg_fasta_read(const char *path, pg_fasta *out);
void pg_fasta_free(pg_fasta *f);

/* whole file as lines (chomp'ed, '\r' kept), one malloc'ed block */
typedef struct {
    char   *buf;
    char  **line;
    size_t *len;
    int64_t count;
} pg_lines;

int  pg_lines_read(const char *path, pg_lines *out);
void pg_lines_free(pg_lines *l);

/* votes/100f printed like java.lang.Float.toString: 1.0, 0.98, 0.5, 0.07, 0.0 */
void pg_fmt_conf(int votes, char out[8]);

#endif
