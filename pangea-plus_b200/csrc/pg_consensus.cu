// pg_consensus.cu -- Stage C: the per-read BLAST/RDP consensus vote
// (Consensus/Consensus_BLAST_SOAP_RDP-1.1.pl:96-237).
//
// The script's hot loops are, per read, (#BLAST rank/name pairs) x (#RDP triples) string
// comparisons per hit followed by a running "best hit" update whose comparisons are all
// STRING comparisons (`gt`, `lt`, `eq` on numbers).  Here:
//   k_rdp_prepare  one thread per read: split the RDP text into triples, sanitise the names
//                  in place (only ASCII letters survive :159-160), resolve the rank index;
//   k_hit_matches  one thread per hit: tokenise the lineage field (:116-122), count matches
//                  (:154-184) against the read's triples -> (rankmatches, blastcount);
//   k_read_vote    one thread per read: the update rules of :186-204 over the read's hits
//                  in file order, with the script's string semantics.
// Grouping BLAST lines under RDP lines (the global cursor, "not found" skips, C2/C8) is
// sequential file parsing and stays in the host program (host/consensus.c).
#include "pg_internal.cuh"

#define PG_MAX_TRIPLES 42       // RDP (name, rank, conf) triples examined per read
#define PG_MAX_TOKENS  128      // lineage tokens examined per hit

struct Triple { int32_t off; int16_t len; int8_t rank; int8_t pad; };   // name bytes inside rdp text

__device__ __forceinline__ bool pg_lin_sep(char c)
{
    return c == '[' || c == ']' || c == ';' || c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v';
}

__device__ int pg_rdp_rank(const char *s, int n)
{
    // position in (domain, phylum, class, order, family, genus, species), else -1 (undef)
    const char *names[7] = {"domain", "phylum", "class", "order", "family", "genus", "species"};
    const int lens[7] = {6, 6, 5, 5, 6, 5, 7};
    for (int q = 0; q < 7; q++) {
        if (n != lens[q]) continue;
        bool eq = true;
        for (int k = 0; k < n; k++) if (s[k] != names[q][k]) { eq = false; break; }
        if (eq) return q;
    }
    return -1;
}

__global__ void k_rdp_prepare(char *__restrict__ rdp, const int64_t *__restrict__ rdp_off, int64_t nreads,
                              Triple *__restrict__ triples, int32_t *__restrict__ ntriples, int *__restrict__ overflow)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreads) return;
    char *s = rdp + rdp_off[r];
    const int n = (int)(rdp_off[r + 1] - rdp_off[r]);
    // split(/\t/): fields; trailing empty fields dropped
    int nf = 0, start = 0, last_nonempty = 0;
    for (int p = 0; p <= n; p++)
        if (p == n || s[p] == '\t') { nf++; if (p > start) last_nonempty = nf; start = p + 1; }
    nf = last_nonempty;
    Triple *tr = triples + r * PG_MAX_TRIPLES;
    int nt = 0, f = 0;
    start = 0;
    int name_off = 0, name_len = 0;
    for (int p = 0; p <= n && f < nf; p++) {
        if (p == n || s[p] == '\t') {
            const int flen = p - start;
            if (f % 3 == 0) {
                // s/"|\\//g ; s/[\W\d_]//g, in place
                int o = start;
                for (int k = start; k < p; k++) {
                    const char c = s[k];
                    if ((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z')) s[o++] = c;
                }
                name_off = start;
                name_len = o - start;
                if (f + 1 >= nf) {                          // name without a rank field: rank undef
                    if (nt < PG_MAX_TRIPLES) { tr[nt].off = name_off; tr[nt].len = (int16_t)name_len; tr[nt].rank = -1; nt++; }
                    else atomicExch(overflow, 1);
                }
            } else if (f % 3 == 1) {
                if (nt < PG_MAX_TRIPLES) { tr[nt].off = name_off; tr[nt].len = (int16_t)name_len; tr[nt].rank = (int8_t)pg_rdp_rank(s + start, flen); nt++; }
                else atomicExch(overflow, 1);
            }
            f++;
            start = p + 1;
        }
    }
    ntriples[r] = nt;
}

// One WARP per hit.  The lineage field is read once, 32 bytes per step, coalesced; separators become a ballot mask,
// token starts and ends fall out of the mask and its shift by one (carried across steps), and their byte offsets go
// to two small per-warp arrays in shared memory.  The (BLAST pair, RDP triple) comparisons are then spread over the
// lanes.  (The first version gave every hit to one thread: two byte-by-byte passes over an uncoalesced line and 512
// bytes of local arrays per thread.)
#define PG_HM_WARPS 8
__global__ void __launch_bounds__(32 * PG_HM_WARPS)
k_hit_matches(const char *__restrict__ lin, const int64_t *__restrict__ lin_off, int64_t nhits,
              const int32_t *__restrict__ read_of_hit, const char *__restrict__ rdp,
              const int64_t *__restrict__ rdp_off, const Triple *__restrict__ triples,
              const int32_t *__restrict__ ntriples, int32_t *__restrict__ rankmatches,
              int32_t *__restrict__ blastcount, int *__restrict__ overflow)
{
    __shared__ short s_toff[PG_HM_WARPS][PG_MAX_TOKENS], s_tend[PG_HM_WARPS][PG_MAX_TOKENS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t h = (int64_t)blockIdx.x * PG_HM_WARPS + warp;
    if (h >= nhits) return;
    const char *s = lin + lin_off[h];
    const int n = (int)(lin_off[h + 1] - lin_off[h]);
    short *toff = s_toff[warp], *tend = s_tend[warp];
    // tokens = maximal runs of bytes that are neither [ ] ; nor whitespace
    int ntok = 0, nend = 0;
    unsigned prev_sep = 1u;                                 // "the byte before the line" counts as a separator
    const unsigned lt = (1u << lane) - 1u;
    for (int p0 = 0; p0 < n; p0 += 32) {
        const int p = p0 + lane;
        const bool sep = p >= n || pg_lin_sep(s[p]);
        const unsigned sm = __ballot_sync(0xffffffffu, sep);
        const unsigned before = (sm << 1) | prev_sep;       // bit i: byte p0+i-1 is a separator
        const unsigned starts = ~sm & before;               // a token starts here
        const unsigned ends = sm & ~before;                 // the first separator after a token: its end (exclusive)
        if ((starts >> lane) & 1u) {
            const int k = ntok + __popc(starts & lt);
            if (k < PG_MAX_TOKENS) toff[k] = (short)p;
        }
        if ((ends >> lane) & 1u && p <= n) {
            const int k = nend + __popc(ends & lt);
            if (k < PG_MAX_TOKENS) tend[k] = (short)(p < n ? p : n);
        }
        ntok += __popc(starts);
        nend += __popc(ends);
        prev_sep = sm >> 31;
    }
    if (nend < ntok && lane == 0 && nend < PG_MAX_TOKENS) tend[nend] = (short)n;     // the line ends inside a token
    if (ntok > PG_MAX_TOKENS && lane == 0) atomicExch(overflow, 1);
    __syncwarp();
    const int usable = ntok < PG_MAX_TOKENS ? ntok : PG_MAX_TOKENS;
    const int64_t r = read_of_hit[h];
    const Triple *tr = triples + r * PG_MAX_TRIPLES;
    const int nt = ntriples[r];
    const char *rs = rdp + rdp_off[r];
    const int npair = (usable + 1) >> 1;
    int matches = 0;
    for (int c = lane; c < npair * nt; c += 32) {
        const int a = 2 * (c / nt), b = c % nt;
        const int alen = tend[a] - toff[a];
        int idx1 = -1;                                      // position in ("0".."6"), else undef
        if (alen == 1 && s[toff[a]] >= '0' && s[toff[a]] <= '6') idx1 = s[toff[a]] - '0';
        const bool has_name = a + 1 < usable;
        const int nl = has_name ? tend[a + 1] - toff[a + 1] : 0;     // undef compares as ""
        const char *nm = has_name ? s + toff[a + 1] : s;
        if (tr[b].rank != idx1 || tr[b].len != nl) continue;
        bool eq = true;
        for (int k = 0; k < nl; k++) if (nm[k] != rs[tr[b].off + k]) { eq = false; break; }
        if (eq) matches++;
    }
    matches = __reduce_add_sync(0xffffffffu, matches);
    if (lane == 0) {
        rankmatches[h] = matches;
        blastcount[h] = ntok;
    }
}

// Perl `$a gt $b` on non-negative integers: compare their decimal text
__device__ int pg_int_str_cmp(int a, int b)
{
    char sa[12], sb[12];
    int na = 0, nb = 0;
    { char t[12]; int k = 0; int v = a; do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v); while (k) sa[na++] = t[--k]; }
    { char t[12]; int k = 0; int v = b; do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v); while (k) sb[nb++] = t[--k]; }
    const int n = na < nb ? na : nb;
    for (int k = 0; k < n; k++) if (sa[k] != sb[k]) return sa[k] < sb[k] ? -1 : 1;
    return na < nb ? -1 : (na > nb ? 1 : 0);
}

__device__ int pg_bytes_cmp(const char *a, int na, const char *b, int nb)
{
    const int n = na < nb ? na : nb;
    for (int k = 0; k < n; k++)
        if (a[k] != b[k]) return (unsigned char)a[k] < (unsigned char)b[k] ? -1 : 1;
    return na < nb ? -1 : (na > nb ? 1 : 0);
}

__global__ void k_read_vote(const int64_t *__restrict__ hit_off, int64_t nreads, const int32_t *__restrict__ rankmatches,
                            const int32_t *__restrict__ blastcount, const char *__restrict__ pid,
                            const int64_t *__restrict__ pid_off, int first_is_fresh, int64_t *__restrict__ winner,
                            int32_t *__restrict__ nmatch)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreads) return;
    int maxblastcount = 0, maxrankmatches = 0;
    int64_t best = -1;
    // $blastsim: undef before the first flush of the run, "0" after every flush
    const char zero = '0';
    const char *sim = &zero;
    int simlen = (r == 0 && first_is_fresh) ? 0 : 1;
    for (int64_t h = hit_off[r]; h < hit_off[r + 1]; h++) {
        const int rm = rankmatches[h], bc = blastcount[h];
        const char *p = pid + pid_off[h];
        const int pl = (int)(pid_off[h + 1] - pid_off[h]);
        if (pg_int_str_cmp(rm, maxrankmatches) > 0) {
            maxrankmatches = rm;
            best = h;
            sim = p; simlen = pl;
        }
        if ((pg_int_str_cmp(bc, maxblastcount) > 0 || pg_bytes_cmp(sim, simlen, p, pl) < 0) && rm == maxrankmatches) {
            maxblastcount = bc;
            best = h;
            sim = p; simlen = pl;
        }
    }
    winner[r] = best;
    nmatch[r] = maxrankmatches;
}

__global__ void k_read_of_hit(const int64_t *__restrict__ hit_off, int64_t nreads, int32_t *__restrict__ read_of_hit)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreads) return;
    for (int64_t h = hit_off[r]; h < hit_off[r + 1]; h++) read_of_hit[h] = (int32_t)r;
}

extern "C" int pg_consensus(pg_ctx *ctx, const pg_consensus_in *in, int64_t *winner_host, int32_t *nmatch_host)
{
    if (!ctx || !in || !winner_host || !nmatch_host || in->nreads < 0 || !in->hit_off)
        return pg_fail(ctx, PG_EINVAL, "pg_consensus: bad arguments");
    const int64_t R = in->nreads;
    if (R == 0) return PG_OK;
    if (R > 0x7fffffffLL) return pg_fail(ctx, PG_ERANGE, "pg_consensus: more than 2^31-1 reads in one batch");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t H = in->hit_off[R];
    const int64_t lb = H ? in->lineage_off[H] : 0, pb = H ? in->pident_off[H] : 0, rb = in->rdp_off[R];
    // one device block: [hit_off | lin_off | pid_off | rdp_off | lin | pid | rdp | triples | ntr | read_of_hit | rm | bc | winner | nmatch | flag]
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
    const size_t o_hoff = take((size_t)(R + 1) * 8), o_loff = take((size_t)(H + 1) * 8), o_poff = take((size_t)(H + 1) * 8),
                 o_roff = take((size_t)(R + 1) * 8), o_lin = take((size_t)lb + 1), o_pid = take((size_t)pb + 1),
                 o_rdp = take((size_t)rb + 1), o_tr = take((size_t)R * PG_MAX_TRIPLES * sizeof(Triple)),
                 o_ntr = take((size_t)R * 4), o_roh = take((size_t)H * 4 + 4), o_rm = take((size_t)H * 4 + 4),
                 o_bc = take((size_t)H * 4 + 4), o_win = take((size_t)R * 8), o_nm = take((size_t)R * 4), o_flag = take(4);
    PG_TRY(pg_scratch(ctx, &ctx->s_candl, o));
    char *base = (char *)ctx->s_candl.p;
    cudaStream_t st = ctx->stream;
    PG_CUDA(ctx, cudaMemcpyAsync(base + o_hoff, in->hit_off, (size_t)(R + 1) * 8, cudaMemcpyHostToDevice, st));
    if (H) {
        PG_CUDA(ctx, cudaMemcpyAsync(base + o_loff, in->lineage_off, (size_t)(H + 1) * 8, cudaMemcpyHostToDevice, st));
        PG_CUDA(ctx, cudaMemcpyAsync(base + o_poff, in->pident_off, (size_t)(H + 1) * 8, cudaMemcpyHostToDevice, st));
        if (lb) PG_CUDA(ctx, cudaMemcpyAsync(base + o_lin, in->lineage_bytes, (size_t)lb, cudaMemcpyHostToDevice, st));
        if (pb) PG_CUDA(ctx, cudaMemcpyAsync(base + o_pid, in->pident_bytes, (size_t)pb, cudaMemcpyHostToDevice, st));
    }
    PG_CUDA(ctx, cudaMemcpyAsync(base + o_roff, in->rdp_off, (size_t)(R + 1) * 8, cudaMemcpyHostToDevice, st));
    if (rb) PG_CUDA(ctx, cudaMemcpyAsync(base + o_rdp, in->rdp_bytes, (size_t)rb, cudaMemcpyHostToDevice, st));
    PG_CUDA(ctx, cudaMemsetAsync(base + o_flag, 0, 4, st));
    const unsigned rb_blocks = (unsigned)((R + 127) / 128);
    k_rdp_prepare<<<rb_blocks, 128, 0, st>>>(base + o_rdp, (const int64_t *)(base + o_roff), R, (Triple *)(base + o_tr),
                                            (int32_t *)(base + o_ntr), (int *)(base + o_flag));
    PG_LAUNCHED(ctx);
    k_read_of_hit<<<rb_blocks, 128, 0, st>>>((const int64_t *)(base + o_hoff), R, (int32_t *)(base + o_roh));
    PG_LAUNCHED(ctx);
    if (H) {
        k_hit_matches<<<(unsigned)((H + PG_HM_WARPS - 1) / PG_HM_WARPS), 32 * PG_HM_WARPS, 0, st>>>(base + o_lin, (const int64_t *)(base + o_loff), H, (const int32_t *)(base + o_roh),
                                                base + o_rdp, (const int64_t *)(base + o_roff), (const Triple *)(base + o_tr),
                                                (const int32_t *)(base + o_ntr), (int32_t *)(base + o_rm), (int32_t *)(base + o_bc),
                                                (int *)(base + o_flag));
        PG_LAUNCHED(ctx);
    }
    k_read_vote<<<rb_blocks, 128, 0, st>>>((const int64_t *)(base + o_hoff), R, (const int32_t *)(base + o_rm),
                                          (const int32_t *)(base + o_bc), base + o_pid, (const int64_t *)(base + o_poff),
                                          in->first_is_fresh, (int64_t *)(base + o_win), (int32_t *)(base + o_nm));
    PG_LAUNCHED(ctx);
    int flag = 0;
    PG_CUDA(ctx, cudaMemcpyAsync(winner_host, base + o_win, (size_t)R * 8, cudaMemcpyDeviceToHost, st));
    PG_CUDA(ctx, cudaMemcpyAsync(nmatch_host, base + o_nm, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
    PG_CUDA(ctx, cudaMemcpyAsync(&flag, base + o_flag, 4, cudaMemcpyDeviceToHost, st));
    PG_CUDA(ctx, cudaStreamSynchronize(st));
    if (flag)
        return pg_fail(ctx, PG_ERANGE, "pg_consensus: a read has more than %d RDP triples or a hit more than %d lineage tokens",
                       PG_MAX_TRIPLES, PG_MAX_TOKENS);
    return PG_OK;
}
