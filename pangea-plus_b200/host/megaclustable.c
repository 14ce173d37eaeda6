/*
 * megaclustable -- drop-in for `perl megaclustable.pl -m <megaclust file>... -t <level> -o <table>`
 * (README.md:185; Megaclustable/megaclustable.pl): pivots the OTU counts of several megaclust outputs
 * into one table per taxonomic level.  For every line holding the token "[<level>]" the name runs from
 * after the token to the next ';' (or ','), the count is the text after the next ',' behind it; rows are
 * names in order of first appearance, columns the files in command-line order.  The inputs are a few
 * thousand short lines, so this stage is plain C on the host -- nothing here is worth a kernel launch.
 *
 * Kept quirks (megaclustable.pl :18-52, :79-106, :113-129): fewer than six arguments -> one line and exit;
 * `-t` outside 0..6 -> message and exit; the first row is "\t1\t2..." without a trailing newline, every
 * later row starts with "\n" and every cell ends with a TAB; a count that is not a number adds 0.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../csrc/pg_perlnum.h"

typedef struct { char *name; double *val; char **raw; } row_t;

static void fmt_num(double v, char *out, size_t cap)
{
    /* Perl prints an NV with %.15g */
    if (v == (double)(long long)v && v > -1e15 && v < 1e15) snprintf(out, cap, "%lld", (long long)v);
    else snprintf(out, cap, "%.15g", v);
}

int main(int argc, char **argv)
{
    if (argc - 1 < 6) { printf("Please enter the correct parameters.\n"); return 0; }
    char **files = (char **)malloc(sizeof(char *) * (size_t)argc);
    int nfiles = 0, in_m = 0;
    const char *output = NULL;
    char token[64] = "";
    for (int a = 1; a < argc; a++) {
        if (strcmp(argv[a], "-m") == 0) in_m = 1;
        else if (strcmp(argv[a], "-o") == 0) { in_m = 0; a++; output = a < argc ? argv[a] : NULL; }
        else if (strcmp(argv[a], "-t") == 0) {
            a++;                                            /* (the script's typo leaves the -m list open here) */
            const char *lv = a < argc ? argv[a] : "";
            const double v = pg_perl_number(lv, (int)strlen(lv));
            if (v > 6 || v < 0) {
                printf("You must enter a number between 0 and 6 for taxonomy level where 0 = domain and 6 = species.\n");
                return 0;
            }
            snprintf(token, sizeof token, "[%s]", lv);
        } else if (in_m) files[nfiles++] = argv[a];
    }
    row_t *rows = NULL;
    int nrows = 0;
    for (int b = 0; b < nfiles; b++) {
        FILE *f = fopen(files[b], "rb");
        if (!f) {
            printf("Unable to open %s\nMake sure you entered the extension when entering the file name.\n", files[b]);
            return 0;
        }
        char *line = NULL;
        size_t cap = 0;
        ssize_t n;
        while ((n = getline(&line, &cap, f)) >= 0) {
            if (n > 0 && line[n - 1] == '\n') line[--n] = 0;
            const char *hit = strstr(line, token);
            if (!hit) continue;
            long loc = (hit - line) + 3;                    /* the script adds 3, whatever the token's length */
            if (loc > n) loc = n;
            const char *endp = strchr(line + loc, ';');
            if (!endp) endp = strchr(line + loc, ',');
            long end = endp ? endp - line : -1;
            long nlen = end - loc;                          /* substr with a negative length drops from the end */
            char *name;
            if (nlen >= 0) {
                name = strndup(line + loc, (size_t)nlen);
            } else {
                long keep = (n - loc) + nlen;
                name = strndup(line + loc, keep > 0 ? (size_t)keep : 0);
            }
            const char *comma = strchr(line + (end >= 0 ? end : 0), ',');      /* index(...) == -1 -> start 0 */
            long nstart = comma ? (comma - line) + 1 : 0;
            const char *num = line + nstart;
            int found = 0;
            for (int r = 0; r < nrows; r++)
                if (strcmp(rows[r].name, name) == 0) {
                    found = 1;
                    const double base = rows[r].raw[b] ? pg_perl_number(rows[r].raw[b], (int)strlen(rows[r].raw[b])) : rows[r].val[b];
                    free(rows[r].raw[b]);
                    rows[r].raw[b] = NULL;
                    rows[r].val[b] = base + pg_perl_number(num, (int)strlen(num));
                }
            if (!found) {
                rows = (row_t *)realloc(rows, sizeof(row_t) * (size_t)(nrows + 1));
                rows[nrows].name = name;
                rows[nrows].val = (double *)calloc((size_t)nfiles, sizeof(double));
                rows[nrows].raw = (char **)calloc((size_t)nfiles, sizeof(char *));
                rows[nrows].raw[b] = strdup(num);           /* pushed as text: printed as it stands unless added to later */
                nrows++;
            } else {
                free(name);
            }
        }
        free(line);
        fclose(f);
    }
    if (!output) { fprintf(stderr, "No such file or directory at megaclustable line 112.\n"); return 2; }
    FILE *fo = fopen(output, "wb");
    if (!fo) { fprintf(stderr, "%s at megaclustable line 112.\n", "No such file or directory"); return 2; }
    for (int a = 1; a <= nfiles; a++) fprintf(fo, "\t%d", a);
    for (int r = 0; r < nrows; r++) {
        fprintf(fo, "\n%s\t", rows[r].name);
        for (int b = 0; b < nfiles; b++) {
            if (rows[r].raw[b]) {
                fprintf(fo, "%s\t", rows[r].raw[b][0] ? rows[r].raw[b] : "0");
            } else {
                char buf[64];
                fmt_num(rows[r].val[b], buf, sizeof buf);
                fprintf(fo, "%s\t", buf);
            }
        }
    }
    fclose(fo);
    return 0;
}
