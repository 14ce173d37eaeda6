/*
 * get_uniq -- drop-in for `perl get_uniq.pl -f <hits>` (Scripts/get_uniq.pl): writes <hits>.unique holding the
 * first line of every distinct first column (the best BLAST hit per read), input order kept.  Same messages on
 * stdout.  The de-duplication runs on the GPU (pg_first_hits).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pangea_b200.h"

int main(int argc, char **argv)
{
    const char *f = NULL;
    int device = 0;
    for (int i = 1; i < argc; i++) {
        if (strcmp(argv[i], "--device") == 0 && i + 1 < argc) { device = atoi(argv[++i]); continue; }
        if (argv[i][0] != '-' || !argv[i][1]) break;
        if (argv[i][1] == 'f') f = argv[i][2] ? argv[i] + 2 : (i + 1 < argc ? argv[++i] : "");
    }
    if (!f || !f[0] || (f[0] == '0' && !f[1])) {
        printf("Usage: perl taxcollector_ncbi-0.01.pl \n\t-f Classification results (tabular text file)\n");
        return 0;
    }
    printf("\nLoading input file...\n");
    FILE *in = fopen(f, "rb");
    if (!in) { printf("Error: Unable to open database file %s.\n", f); return 0; }
    fseek(in, 0, SEEK_END);
    long n = ftell(in);
    fseek(in, 0, SEEK_SET);
    char *text = (char *)malloc((size_t)n + 1), *out = (char *)malloc((size_t)n + 1);
    if (n && fread(text, 1, (size_t)n, in) != (size_t)n) { printf("Error: Unable to open database file %s.\n", f); return 0; }
    fclose(in);
    char path[4096];
    snprintf(path, sizeof path, "%s.unique", f);
    FILE *fo = fopen(path, "wb");
    if (!fo) { printf("Error: Unable to open output file %s.\n", path); return 0; }
    pg_ctx *ctx = pg_init(device);
    if (!ctx) { fprintf(stderr, "get_uniq: %s\n", pg_last_error(NULL)); return 1; }
    int64_t olen = 0;
    if (pg_first_hits(ctx, text, n, out, n, &olen, NULL, 0, NULL) != PG_OK) { fprintf(stderr, "get_uniq: %s\n", pg_last_error(ctx)); return 1; }
    fwrite(out, 1, (size_t)olen, fo);
    fclose(fo);
    pg_shutdown(ctx);
    return 0;
}
