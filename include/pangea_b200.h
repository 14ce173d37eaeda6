/*
 * pangea_b200.h -- C ABI of libpangea_b200.so: the B200-native (sm_100a CUDA)
 * classification hot path of PANGEA+.
 *
 * The reference has no in-process plugin API: its boundary is argv + text
 * files + stdout + exit status.  The three drop-in
 * executables (rdp_classifier, tax_class/taxcollector, consensus -- see
 * pangea-plus_b200/host/) are thin main()s over the entry points below, and a
 * maintainer binding the path from Perl/Python/cgo binds exactly these
 * (INTEGRATION.md shows the stubs).  Each entry point cites the reference
 * interface it replaces (paths relative to the reference root).
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function
 * returns 0 (PG_OK) or a negative PG_E* code, with text from pg_last_error();
 * the library never writes to stdout; handles are opaque and freed by the
 * matching *_free; one host thread per pg_ctx.  There is NO CPU fallback:
 * without a usable CUDA device pg_init() fails (PG_ENODEV) and nothing else
 * can be called.
 *
 * "host" pointers are ordinary process memory; "dev" pointers are CUDA device
 * memory on the context's device (e.g. torch tensor .data_ptr()).
 */
#ifndef PANGEA_B200_H
#define PANGEA_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_OK        0
#define PG_EINVAL   -1   /* bad argument */
#define PG_ECUDA    -2   /* CUDA runtime / launch failure */
#define PG_ENOMEM   -3   /* host or device allocation failed */
#define PG_EIO      -4   /* file could not be opened / read / written */
#define PG_ENODEV   -5   /* no CUDA device: the library has no CPU path */
#define PG_ERANGE   -6   /* input outside the supported range (e.g. read too long) */
#define PG_EFORMAT  -7   /* malformed file */

#define PG_WORDSIZE        8        /* RDP 8-mers                                  */
#define PG_NWORDS          65536    /* 4^8                                         */
#define PG_NUM_BOOT        100      /* NUM_OF_RUNS                                 */
#define PG_MIN_SEQ_LEN     50       /* ShortSequenceException gate     */
#define PG_MAX_DEPTH       32       /* lineage levels root..genus kept per result  */
#define PG_GENUS_TILE      32       /* genera per 128-byte table row segment       */
#define PG_MAX_WORDS       7000     /* longest read (good words) the kernels stage */

typedef struct pg_ctx   pg_ctx;
typedef struct pg_model pg_model;
typedef struct pg_reads pg_reads;
typedef struct pg_tax   pg_tax;

/* ------------------------------------------------------------------ context */

/* One context per GPU (one process per GPU under torchrun; or one per device
 * inside a multi-device host program).  Returns NULL on failure; the reason is
 * then available from pg_last_error(NULL). */
pg_ctx     *pg_init(int device);
void        pg_shutdown(pg_ctx *ctx);
const char *pg_last_error(const pg_ctx *ctx);
/* Launch everything on the caller's CUDA stream (cudaStream_t / CUstream, e.g.
 * torch.cuda.current_stream().cuda_stream).  NULL = the context's own
 * non-blocking stream; pass cudaStreamLegacy ((void*)1) for the legacy default
 * stream, whose handle is also 0. */
int         pg_set_stream(pg_ctx *ctx, void *cuda_stream);
int         pg_sync(pg_ctx *ctx);
/* Number of kernels this context has launched so far (bench "gpu_launches"). */
int64_t     pg_launch_count(const pg_ctx *ctx);
/* Milliseconds spent inside the dominant classify kernel since the last call to
 * pg_kernel_time_reset(), measured with CUDA events on the launching stream, and
 * how many launches that covers. */
int         pg_kernel_time(pg_ctx *ctx, double *classify_ms, int64_t *classify_launches);
int         pg_kernel_time_reset(pg_ctx *ctx);

/* ------------------------------------------------------------------ Stage A
 * Replaces: `java -Xmx1g -jar rdp_classifier-2.5.jar -q <in.fa> -o <out.txt>`
 * (README.md:119-122; installer Classify/RunRDP/install_RDPClassifier.sh:60-66).
 * Semantics: SURVEY.md section 8(a) rows A1-A9.
 */

/* A batch of sequences: `count` records, record i = bytes[off[i] .. off[i+1]),
 * raw FASTA residue characters (any case, IUPAC allowed, no newlines). */
typedef struct {
    const char    *bytes;
    const int64_t *off;      /* count+1 entries */
    int64_t        count;
} pg_seqbatch;

/* One classified read, 64 bytes (A7-A9). */
typedef struct {
    int32_t genus;                    /* genus index of the assignment; -1 if status!=0 */
    int32_t n_words;                  /* good 8-mers used (after orientation)           */
    float   score;                    /* A7: sequential fp32 sum of the winner          */
    uint8_t reversed;                 /* A3: 1 if the read was reverse-complemented     */
    uint8_t status;                   /* 0 ok; 1 = shorter than PG_MIN_SEQ_LEN (A2)     */
    uint8_t depth;                    /* lineage levels of the genus (root first)       */
    uint8_t _pad0;
    uint8_t votes[PG_MAX_DEPTH];      /* A9: bootstrap votes (0..100) per lineage level */
    uint8_t _pad1[16];
} pg_result;

/* A5+A6 training (upstream RawHierarchyTree.initWordOccurrence /
 * TreeFactory.createGenusWordConditionalProb): integer-atomic word x genus
 * counts, then the genus-tiled fp32 log table.  genus_of_seq[i] in [0,G).
 * Host buffers in; the model lives on the context's device. */
int pg_train(pg_ctx *ctx, const pg_seqbatch *seqs_host, const int32_t *genus_of_seq_host,
             int G, pg_model **out);
/* Same with all three arrays already resident in device memory. */
int pg_train_dev(pg_ctx *ctx, const pg_seqbatch *seqs_dev, const int32_t *genus_of_seq_dev,
                 int G, pg_model **out);
void pg_model_free(pg_model *m);
/* Sharded / streamed training (BASELINE configs[3]): add the counts of one more batch to a model made
 * by pg_model_create() -- or trained before -- without deriving tables.  The counts are integers added atomically, so
 * after an integer sum all-reduce of pg_model_buffers() over the ranks (ncclAllReduce / torch.distributed) every rank
 * holds exactly the counts of one pg_train() over the whole set; pg_model_commit() then derives the tables.
 * genus_of_seq[i] belongs to record i of the batch (seqs->off may be a slice of a larger offset array). */
int pg_train_accumulate(pg_ctx *ctx, pg_model *m, const pg_seqbatch *seqs_host, const int32_t *genus_of_seq_host);
int pg_train_accumulate_dev(pg_ctx *ctx, pg_model *m, const pg_seqbatch *seqs_dev, const int32_t *genus_of_seq_dev);

/* Lineage of every genus for the A9 vote: anc[g*depth + d] = taxonomy node id
 * at level d (root first), -1 beyond the genus' own level.  depth <= PG_MAX_DEPTH.
 * Without it every genus has the one-level lineage {g}. */
int pg_model_set_lineage(pg_model *m, const int32_t *anc_host, int depth);

int     pg_model_genera(const pg_model *m);
int64_t pg_model_sequences(const pg_model *m);

/* Parity hooks (host outputs, any may be NULL): m_wg[w*G+g], n_w[65536], M_g[G], N. */
int pg_model_counts(const pg_model *m, int32_t *m_wg, int32_t *n_w, int32_t *M_g, int64_t *N);
/* logPrior[65536], logLeave[G], logP[w*G+g] (dense, word-major). */
int pg_model_tables(const pg_model *m, float *logPrior, float *logLeave, float *logP);

/* Model persistence (sparse counts; tables are re-derived on load, on the GPU).
 * `blob` is an opaque caller section (the CLI keeps the taxonomy names there). */
int pg_model_save(const pg_model *m, const char *path, const void *blob, int64_t blob_len);
int pg_model_load(pg_ctx *ctx, const char *path, pg_model **out, void **blob, int64_t *blob_len);
void pg_free(void *p);

/* A model given as TABLES instead of counts -- the form RDP's own trainset files hold (upstream
 * files logWordPrior.txt, wordConditionalProbIndexArr.txt, genus_wordConditionalProbList.txt and the leaveCount
 * attributes of bergeyTrainingTree.xml, loaded through rRNAClassifier.properties; call site README.md:119 `-t`):
 * logPrior[65536], leave_count[G] (the genus nodes' leaveCount; logLeave = ln(leaveCount + 1) is derived on the device by
 * the expression training uses), and the listed cells in word-major order -- word w owns entries
 * idx[w] .. idx[w+1] (idx has 65537 entries), entry e is logp_of_entry[e] for genus genus_of_entry[e].  Every cell
 * that is not listed is fp32(logPrior[w] - logLeave[g]) (row A4).  Such a model classifies like any other; it has no
 * counts, so pg_model_save / pg_model_commit / pg_model_counts refuse it.  Host arrays in. */
int pg_model_from_tables(pg_ctx *ctx, int G, const float *logPrior, const int32_t *leave_count, const int64_t *idx,
                         const int32_t *genus_of_entry, const float *logp_of_entry, pg_model **out);

/* Multi-GPU replication: the device buffers that define a model, for an
 * external broadcast (torch.distributed / ncclBroadcast).  A receiving rank
 * creates an empty model of the same G, broadcasts into its buffers, then
 * calls pg_model_commit().  names: "counts_m","counts_n","counts_M","N". */
int pg_model_create(pg_ctx *ctx, int G, pg_model **out);
int pg_model_buffers(pg_model *m, void **dev_ptrs, size_t *nbytes, int max, int *n);
int pg_model_commit(pg_model *m);     /* re-derive tables from the counts */

/* K3: 2-bit packing.  Packs a batch into the compact device read store. */
int  pg_reads_pack(pg_ctx *ctx, const pg_seqbatch *seqs_host, pg_reads **out);
int  pg_reads_pack_dev(pg_ctx *ctx, const pg_seqbatch *seqs_dev, int64_t total_bytes, pg_reads **out);
void pg_reads_free(pg_reads *r);
int64_t pg_reads_count(const pg_reads *r);
/* Parity hook: 2-bit codes (16 bases per uint32, base i at bits 2*(i%16)) and the
 * validity mask (32 bases per uint32) of read i; returns its length. */
int64_t pg_reads_unpack(const pg_reads *r, int64_t i, uint32_t *codes, uint32_t *mask, int64_t cap_words);

/* Options for classification. */
typedef struct {
    int32_t min_boot_words;   /* 0 = RDP 2.5 (k = n/8); later releases use 5            */
    int32_t mode;             /* 0 = strict (reference order);
                                 1 = certified (quantised pre-filter, strict re-check)  */
    int32_t cert_plan;        /* certified mode only, results never depend on it:
                                 0 = default: best 16-position part of the best block + lower bounds of every other
                                     block as one integer matrix product per read on the tensor cores (tcgen05,
                                     reads of up to 640 good words; longer reads as 3), then the open pairs exactly,
                                 1 = every genus block with partial-sum pruning,
                                 2 = the whole best block + lower bounds + items,
                                 3 = as 0 with the shared-memory kernels (the default of round 1/2) */
    int32_t light_max;        /* cert_plan 0: open (task, block) pairs per read above which the
                                 read is redone under cert_plan 1; 0 = default, -1 = none */
    int32_t bound_level;      /* cert_plan 0 / 3, results never depend on it: which kernel bounds the blocks the
                                 best part leaves: 0 = the plan's own (tensor cores under cert_plan 0); any other value
                                 selects cert_plan 3 with: 1 = the 16-bit bounds, one CTA per (read, group of 28 blocks);
                                 2 = a coarse 8-bit first level over all blocks + an exact second level (measured
                                 no faster on 10 000 genera, kept for models with many more groups); 3 = the 16-bit
                                 bounds over the table's 16-position PARTS instead of its 64-position blocks (four times
                                 the columns, far stronger bounds when genera have hundreds of training members) */
    int32_t reserved[3];
} pg_classify_opts;

/* 1 if the model's quantised table certifies every deficit (mode 1 is then the
 * certified path; otherwise mode 1 silently runs the strict kernels). */
int pg_model_certifiable(const pg_model *m);
/* Which columns the certified path's lower bounds use for this model: 0 = the 64-position blocks, 1 = the 16-position
 * parts, -1 = not decided yet.  Models with more than one group of blocks are timed both ways on the head of the first
 * batch they classify (bound_level 0); results never depend on it. */
int pg_model_bound_columns(const pg_model *m);
/* How the reads of the last pg_classify*() call were routed: through the certified
 * kernels, through the strict kernels, and handed back from certified to strict. */
int pg_classify_stats(const pg_ctx *ctx, int64_t *certified_reads, int64_t *strict_reads, int64_t *handed_back);
/* Certified-path detail of the last call: reads redone by the all-block kernel because the block
 * lower bounds left too many (task, block) pairs open, and the number of such pairs ("items")
 * evaluated exactly for the other reads. */
int pg_classify_stats2(const pg_ctx *ctx, int64_t *heavy_reads, int64_t *items);
/* Reads of the last call whose best part and block bounds were computed by the tensor-core kernel (certified mode's
 * default plan, reads of up to 640 good words); the others went through the shared-memory kernels of plan 3. */
int pg_classify_stats3(const pg_ctx *ctx, int64_t *tensor_core_reads);

/* K3-K5: word extraction + orientation, gather-sum + 100 bootstraps, argmax,
 * vote.  results: nreads records (host for pg_classify, device for *_dev).
 * boot_winners (optional, may be NULL): nreads*100 int32 genus index per
 * replicate, parity hook for row A8. */
int pg_classify(pg_ctx *ctx, const pg_model *m, const pg_seqbatch *reads_host,
                const pg_classify_opts *opts, pg_result *results_host, int32_t *boot_winners_host);
int pg_classify_packed(pg_ctx *ctx, const pg_model *m, const pg_reads *reads,
                       const pg_classify_opts *opts, pg_result *results_dev, int32_t *boot_winners_dev);
/* the same with the records (and optional replicate winners) delivered to host memory: what a C host program
 * calls after pg_fasta_ingest() / pg_trim_join(..., reads_out) */
int pg_classify_packed_host(pg_ctx *ctx, const pg_model *m, const pg_reads *reads,
                            const pg_classify_opts *opts, pg_result *results_host, int32_t *boot_winners_host);

/* Parity hook for A1/A3: the word list the classifier uses for each read
 * (after orientation).  words_host[off[i] + j], n_words_host[i], reversed_host[i]. */
int pg_extract_words(pg_ctx *ctx, const pg_model *m, const pg_seqbatch *reads_host,
                     uint16_t *words_host, int32_t *n_words_host, uint8_t *reversed_host);
/* Parity hook for A8: the java.util.Random(1) sample list for a read of n words:
 * out[run*k + j], k = max(n/8, min_boot_words).  Computed on the device. */
int pg_boot_indices(pg_ctx *ctx, int32_t n, int32_t min_boot_words, uint16_t *out_host);

/* ------------------------------------------------------------------ Stage B
 * Replaces: Tax_class/ncbitc.c (`tax_class -c|-s|-g|-t|-n`, main at :841-1004)
 * and the per-hit walk of Tax_class/NCBI-taxcollector-0.01.pl:58-300.
 */

/* == `tax_class -c` (ncbitc.c:995-998): builds gi_taxid_nucl.dmp.bin,
 * nodes.dmp.bin, names.dmp.bin in `dir` with the reference's record layouts
 * (ncbitc.c:98-140), so the files are interchangeable. */
int pg_tax_build(const char *dir);
/* Loads the three .bin files and uploads the lookup arrays. */
int pg_tax_load(pg_ctx *ctx, const char *dir, pg_tax **out);
void pg_tax_free(pg_tax *t);

/* Batched gi -> leaf taxid (ncbitc_search_tax_id :567-599; 0 = unknown). */
int pg_tax_leaf(pg_ctx *ctx, const pg_tax *t, const int32_t *gi_host, int64_t n, int32_t *taxid_host);
/* Batched lineage strings exactly as taxcollector prints them between the first
 * and second TAB (rows B5-B7): out_bytes is a caller buffer of out_cap bytes,
 * out_off gets n+1 offsets.  Returns PG_ERANGE if out_cap is too small (needed
 * size in out_off[n]). */
int pg_tax_lineage(pg_ctx *ctx, const pg_tax *t, const int32_t *gi_host, int64_t n,
                   char *out_bytes, int64_t out_cap, int64_t *out_off);

/* Views for the drop-in `tax_class -s|-g|-t|-n` executable.  pg_tax_chain walks the parent
 * chain on the device exactly like ncbitc.c:941-952 (the nodes whose parent is not 1, leaf
 * first; chain_len < 0 = the walk left the node table after -1-chain_len nodes).  The two
 * record getters copy the raw 28-/196-byte records (ncbitc.c:98-133) of the loaded files;
 * pg_tax_name_records returns how many records `tax_class -n` would print. */
int pg_tax_chain(pg_ctx *ctx, const pg_tax *t, const int32_t *taxid_host, int64_t n, int32_t maxlen,
                 int32_t *chain_host, int32_t *chain_len_host);
int pg_tax_node_record(const pg_tax *t, int32_t taxid, void *rec28);
int pg_tax_name_records(const pg_tax *t, int32_t taxid, void *rec196, int32_t max_records);
int64_t pg_tax_max_gi(const pg_tax *t);
int64_t pg_tax_max_taxid(const pg_tax *t);

/* ------------------------------------------------------------------ Stage C
 * Replaces: Consensus/Consensus_BLAST_SOAP_RDP-1.1.pl:96-237 (rows C2-C8).
 */

/* hits: BLAST-class lines grouped by read (CSR hit_off over reads);
 * per hit the lineage field and the pident field as text; per read the RDP
 * triples as text.  winner[i] = index (into the hit arrays) of the line the
 * reference would print for read i, nmatch[i] = its "#Matches found". */
typedef struct {
    int64_t        nreads;
    const int64_t *hit_off;        /* nreads+1 */
    const char    *lineage_bytes;  /* concatenated lineage fields           */
    const int64_t *lineage_off;    /* nhits+1                               */
    const char    *pident_bytes;   /* concatenated third-column fields      */
    const int64_t *pident_off;     /* nhits+1                               */
    const char    *rdp_bytes;      /* per read: text after the five TABs    */
    const int64_t *rdp_off;        /* nreads+1                              */
    int32_t        first_is_fresh; /* 1: read 0 is the first group the script ever flushes
                                      ($blastsim still undef, :186-204); 0: it is "0" */
    int32_t        reserved;
} pg_consensus_in;

/* winner_host[i] = -1 when no hit of read i updates the script's $tempresult (the script then
 * prints the previous read's line again; the caller carries that over). */
int pg_consensus(pg_ctx *ctx, const pg_consensus_in *in_host, int64_t *winner_host, int32_t *nmatch_host);

/* ------------------------------------------------------------------ Trim join (widening: SURVEY.md 8(f) next-1)
 * Replaces: `perl trim2.3.pl -a <read 1> [-b <read 2>] [-g gap] [-t truncate]`  (README.md:31-33;
 * Trim/trim2.4.pl parse_qseq :169-242, trim_qseq :244-298, parse_fastq :467-521, trim_fastq :527-578):
 * quality trim of Illumina QSEQ pairs or FASTQ records and the mateA + N x gap + mateB join that
 * defines the reads entering Stage A.
 */
typedef struct {
    int32_t gap;        /* -g, the script's default is 189                                        */
    int32_t truncate;   /* -t, default 11 (QSEQ only)                                             */
    int32_t reserved[6];/* the script's -qc / -lc cannot be parsed by its own getopts string:
                           quality cutoff 20 and length cutoff 70 are constants of the contract  */
} pg_trim_opts;

/* a_host / b_host: the whole input files.  FASTQ ('@' first): b_host is ignored and `paired` says
 * whether -b was given (the script then takes the second mate from the NEXT record of the same
 * file); otherwise QSEQ pairs, line i of a with line i of b.  out_host receives the text of
 * <prefix>_runblast.fasta; PG_ERANGE (needed size in *out_len) when out_cap is too small.
 * reads_out (optional): the joined sequences, same order, already in the packed device read store
 * for pg_classify_packed -- records the trim dropped are simply absent from both outputs. */
int pg_trim_join(pg_ctx *ctx, const char *a_host, int64_t a_len, const char *b_host, int64_t b_len, int paired,
                 const pg_trim_opts *opts, char *out_host, int64_t out_cap, int64_t *out_len, pg_reads **reads_out);

/* ------------------------------------------------------------------ FASTA ingest (widening: SURVEY.md 8(f) next-1)
 * Replaces the file reading of `java -jar rdp_classifier-2.5.jar -q <in.fa>` (README.md:119): the whole query
 * file goes to the device as text; out come the packed read store for pg_classify_packed (records in file
 * order) and, per record, where its header sits in text_host: the header is
 * text_host[hdr_off[i] .. hdr_off[i]+hdr_len[i]) (without '>'), the id its first id_len[i] bytes.
 * Lines starting with '>' open a record; other lines are sequence ('\r' and blanks dropped; lines before the
 * first header ignored).  PG_ERANGE (count in *nrec) when cap is too small.  reads_out may be NULL. */
int pg_fasta_ingest(pg_ctx *ctx, const char *text_host, int64_t len, int64_t cap, int64_t *nrec, int64_t *hdr_off,
                    int32_t *id_len, int32_t *hdr_len, pg_reads **reads_out);
/* Parity hook: sequence bytes (concatenated) and offsets [nrec+1] produced by the last pg_fasta_ingest(). */
int pg_fasta_last_bytes(pg_ctx *ctx, int64_t nrec, char *bytes_host, int64_t cap, int64_t *off_host);

/* ------------------------------------------------------------------ Megaclust (widening: SURVEY.md 8(f) next-2)
 * Replaces: `perl megaclust2.pl -i <consensus or BLAST tabular> -o <out> [-s sim] [-e evalue] [-b bitscore] [-c x]`
 * (README.md:176; Megaclust/megaclust2.pl:80-153): lines not starting with '#' are split on the script's
 * delimiter pattern, kept when pident >= sim, evalue <= eval and bitscore >= bits (Perl's numeric reading of
 * the text), and counted per subject (field 2): once per distinct (subject, query) pair, or every line when
 * count_every_hit is set.
 */
typedef struct {
    double  sim_threshold;      /* -s, the script's default is 95   */
    double  eval_threshold;     /* -e, default 1e-20                */
    double  bitscore_threshold; /* -b, default 200                  */
    int32_t count_every_hit;    /* -c <true value>                  */
    int32_t reserved[5];
} pg_megaclust_opts;

/* text_host: the whole input file.  OTUs come back in order of first appearance (the script prints them in
 * Perl's hash order, i.e. unspecified): subject i is text_host[otu_off[i] .. otu_off[i]+otu_len[i]) and was hit
 * otu_count[i] times.  PG_ERANGE (count in *n_otus) when cap is too small.  lines_examined / lines_beyond are
 * the two numbers of the script's run summary. */
int pg_megaclust(pg_ctx *ctx, const char *text_host, int64_t len, const pg_megaclust_opts *opts, int64_t cap,
                 int64_t *n_otus, int64_t *otu_off, int32_t *otu_len, int64_t *otu_count, int64_t *lines_examined,
                 int64_t *lines_beyond);

/* ------------------------------------------------------------------ first hit per read (widening: SURVEY.md 8(f) next-4)
 * Replaces: `perl get_uniq.pl -f <tabular hits>` (Scripts/get_uniq.pl:34-40), which writes <file>.unique: every
 * line whose first TAB-separated column has not occurred on an earlier line, in input order (a line without a TAB
 * keys on its whole text, newline included, as the script's misplaced chomp makes it).
 * out_host receives the kept lines; kept_lines_host (optional) their 0-based line numbers.  PG_ERANGE (needed
 * sizes in *out_len / *n_kept) when a buffer is too small. */
int pg_first_hits(pg_ctx *ctx, const char *text_host, int64_t len, char *out_host, int64_t out_cap, int64_t *out_len,
                  int64_t *kept_lines_host, int64_t lines_cap, int64_t *n_kept);

#ifdef __cplusplus
}
#endif
#endif /* PANGEA_B200_H */
