/*
 * pangea_b200_synth.h -- seeded synthetic workloads generated ON THE DEVICE (bench and test support).
 *
 * Not part of the reference-facing ABI (that is pangea_b200.h): nothing here replaces a reference interface.
 * BASELINE.json configs[3] asks for 100 M reads against a ~3 M-sequence / ~10 000-genus training set; neither
 * fits through host text in a bench's time budget, so members and reads are produced by kernels from a seed
 * ("generated on device from the seed, never through text").  Every byte is a pure function of
 * (seed, record index, position) through splitmix64, so any slice of the set can be produced on any rank, and
 * pangea_b200/synth.py holds the same functions in numpy for the CPU oracle (tests/test_synth_gpu.py compares them).
 */
#ifndef PANGEA_B200_SYNTH_H
#define PANGEA_B200_SYNTH_H

#include "pangea_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Training members [first, first+count) of a synthetic 16S set: member i copies the centroid of its genus
 * (centroids_dev[genus * length ..], lower-case acgt), truncated to off[i+1]-off[i] bases, with 1 % substitutions
 * and 0.1 % 'n'.  off_dev: count+1 offsets into bytes_dev, relative to the slice (off_dev[0] == 0);
 * genus_dev: count entries. */
int pg_synth_members(pg_ctx *ctx, uint64_t seed, const uint8_t *centroids_dev, int length, const int32_t *genus_dev,
                     const int64_t *off_dev, int64_t first, int64_t count, char *bytes_dev);

/* Illumina-like reads [first, first+count) drawn from resident members: a window of
 * `span` = paired ? 2*read_len+gap : read_len bases of a member chosen by hash (members shorter than the span are
 * re-drawn deterministically), 0.5 % substitution errors, the gap filled with 'N' (the Trim join,
 * Trim/trim2.4.pl:228-244), every second record (by hash) reverse-complemented whole.  Fixed record length, so
 * record j is out_dev[j*span ..]; src_genus_dev (optional) gets the member's genus. */
int pg_synth_reads(pg_ctx *ctx, uint64_t seed, const char *members_dev, const int64_t *member_off_dev,
                   const int32_t *member_genus_dev, int64_t nmembers, int64_t first, int64_t count, int read_len,
                   int gap, int paired, char *out_dev, int32_t *src_genus_dev);

#ifdef __cplusplus
}
#endif
#endif
