/*
 * tax_class -- drop-in for the reference's Tax_class/tax_class (ncbitc.c), same flags, same
 * cwd-relative file names, same stdout text (ncbitc.c:841-1004):
 *   -c  build gi_taxid_nucl.dmp.bin, nodes.dmp.bin, names.dmp.bin from the NCBI dumps
 *   -s gi   every node up the chain whose parent is not 1      -g gi   the leaf node line
 *   -t taxid  that node line                                    -n taxid  its name lines
 *   -v verbose, -h help.   Failure: "Error." on stdout, exit status 255.
 * The parent chain is walked on the GPU (libpangea_b200: pg_tax_leaf / pg_tax_chain); there is
 * no CPU lookup path.  Several ids may follow one another (-s 5 -s 7 ...): they are answered
 * in order from one load of the tables, which is what makes the tool usable from scripts.
 */
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pangea_b200.h"

static const char *RANK_STR[29] = {
    "class", "family", "forma", "genus", "infraclass", "infraorder", "kingdom", "no rank", "order",
    "parvorder", "phylum", "species", "species group", "species subgroup", "subclass", "subfamily",
    "subgenus", "subkingdom", "suborder", "subphylum", "subspecies", "subtribe", "superclass",
    "superfamily", "superkingdom", "superorder", "superphylum", "tribe", "varietas"};

static void print_help(void)
{
    printf("Usage: ncbitc [options]\n");
    printf("Options:\n");
    printf("   -s --search id         search all tree using a gi index\n");
    printf("   -g --search-gi id      search a tax id using a gi index\n");
    printf("   -t --search-node id    search a node entry using a tax id\n");
    printf("   -n --search-name id    search a name entry using a tax id\n");
    printf("   -v --verbose           turn on verbose output\n");
    printf("   -h --help              print this help message\n");
    exit(0);
}

/* ncbitc_print_node_entry (:467-493) on a raw 28-byte record */
static void print_node(const unsigned char *r)
{
    int32_t tax, parent, mgc;
    int16_t div, gc;
    memcpy(&tax, r, 4); memcpy(&parent, r + 4, 4); memcpy(&div, r + 12, 2); memcpy(&gc, r + 16, 2); memcpy(&mgc, r + 20, 4);
    int rk = (signed char)r[8];
    char embl[4] = {(char)r[9], (char)r[10], (char)r[11], 0};
    printf("%d | %d | %s | %s | %d | %d | %d | %d | %d | %d | %d | %d | %s |\n", tax, parent,
           (rk >= 0 && rk < 29) ? RANK_STR[rk] : "invalid id", embl, div, (signed char)r[14], gc, (signed char)r[18], mgc,
           (signed char)r[24], (signed char)r[25], (signed char)r[26], "");
}

typedef struct { int op; int id; } request;

int main(int argc, char **argv)
{
    static struct option long_options[] = {
        {"help", no_argument, 0, 'h'}, {"verbose", no_argument, 0, 'v'}, {"create", no_argument, 0, 'c'},
        {"search", required_argument, 0, 's'}, {"search-gi", required_argument, 0, 'g'},
        {"search-name", required_argument, 0, 'n'}, {"search-node", required_argument, 0, 't'},
        {"dir", required_argument, 0, 'D'}, {"device", required_argument, 0, 'G'}, {0, 0, 0, 0}};
    request req[256];
    int nreq = 0, verbose = 0, create = 0, device = 0;
    const char *dir = ".";
    for (;;) {
        int c = getopt_long(argc, argv, "hcs:n:vt:g:", long_options, NULL);
        if (c == -1) break;
        switch (c) {
        case 'c': create = 1; break;
        case 's': case 'g': case 'n': case 't':
            if (nreq < 256) { req[nreq].op = c; req[nreq].id = atoi(optarg); nreq++; }
            break;
        case 'v': verbose = 1; break;
        case 'D': dir = optarg; break;
        case 'G': device = atoi(optarg); break;
        default: print_help();
        }
    }
    if (verbose) printf("verbose flag is set\n");
    if (create && nreq == 0) {
        if (pg_tax_build(dir) != PG_OK) { fprintf(stderr, "%s\n", pg_last_error(NULL)); printf("Error.\n"); return 255; }
        return 0;
    }
    if (nreq == 0) print_help();

    pg_ctx *ctx = pg_init(device);
    if (!ctx) { fprintf(stderr, "tax_class: %s\n", pg_last_error(NULL)); printf("Error.\n"); return 255; }
    pg_tax *tx = NULL;
    if (pg_tax_load(ctx, dir, &tx) != PG_OK) { fprintf(stderr, "tax_class: %s\n", pg_last_error(ctx)); printf("Error.\n"); return 255; }
    enum { MAXCHAIN = 256 };
    int32_t *chain = (int32_t *)malloc(sizeof(int32_t) * MAXCHAIN);
    int rc = 0;
    for (int q = 0; q < nreq && rc == 0; q++) {
        int op = req[q].op;
        int32_t id = req[q].id, leaf = 0, len = 0;
        unsigned char rec[28];
        if (op == 's' || op == 'g') {
            if (id < 1) { fprintf(stderr, "fseek: Invalid argument\n"); printf("Error.\n"); rc = 255; break; }
            if (pg_tax_leaf(ctx, tx, &id, 1, &leaf) != PG_OK) { printf("Error.\n"); rc = 255; break; }
            if (verbose) printf("%d\t%d\n", id, leaf);
            if (leaf == 0) { printf("0\n"); continue; }
            if (op == 'g') {
                if (pg_tax_node_record(tx, leaf, rec) != PG_OK) { printf("Error.\n"); rc = 255; break; }
                if (verbose) { int32_t t; memcpy(&t, rec, 4); printf("%d\n", t); }
                print_node(rec);
                continue;
            }
            if (pg_tax_chain(ctx, tx, &leaf, 1, MAXCHAIN, chain, &len) != PG_OK) { printf("Error.\n"); rc = 255; break; }
            int n = len < 0 ? -1 - len : len;
            for (int k = 0; k < n; k++) {
                pg_tax_node_record(tx, chain[k], rec);
                if (verbose) { int32_t t; memcpy(&t, rec, 4); printf("%d\n", t); }
                print_node(rec);
            }
            if (len < 0) { printf("Error.\n"); rc = 255; }        /* the chain left the node table */
        } else if (op == 't') {
            if (pg_tax_node_record(tx, id, rec) != PG_OK) { fprintf(stderr, "fseek: Invalid argument\n"); printf("Error.\n"); rc = 255; break; }
            if (verbose) { int32_t t; memcpy(&t, rec, 4); printf("%d\n", t); }
            print_node(rec);
        } else {
            int cnt = pg_tax_name_records(tx, id, NULL, 0);
            if (cnt <= 0) { printf("0\n"); continue; }
            unsigned char *nr = (unsigned char *)malloc((size_t)cnt * 196);
            pg_tax_name_records(tx, id, nr, cnt);
            for (int k = 0; k < cnt; k++) {
                int32_t t;
                memcpy(&t, nr + (size_t)k * 196, 4);
                printf("%d | %s | %s | %s |\n", t, (char *)nr + (size_t)k * 196 + 4, (char *)nr + (size_t)k * 196 + 68,
                       (char *)nr + (size_t)k * 196 + 132);
            }
            free(nr);
        }
    }
    free(chain);
    pg_tax_free(tx);
    pg_shutdown(ctx);
    return rc;
}
