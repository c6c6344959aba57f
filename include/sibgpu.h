/* sibgpu.h -- C ABI of the B200-native de Bruijn-graph hot path of Sibelia.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every entry point names the
 * reference interface it replaces (paths relative to the reference tree, /root/reference).  The reference binds through
 * sibelia_b200/csrc/facade/: translation units that DEFINE the reference's own members (IndexedSequence::
 * EnumerateBifurcationsSArray{,InRAM}, BlockFinder::PerformGraphSimplifications, FASTAReader::GetSequences) on top of
 * this header and are linked instead of the reference's, plus generated edits of the index-building sites of
 * synteny.cpp / serialization.cpp; INTEGRATION.md shows the binding a maintainer of the reference would add.  (No
 * array-backed re-implementation of the IndexedSequence / BifurcationStorage / DNASequence classes exists: every
 * production site that built those objects around the hot path calls this ABI instead.)
 *
 * Conventions
 *   - every function returns 0 on success, a sibgpu_status otherwise; sibgpu_last_error() gives the message
 *     (the reference throws std::runtime_error, src/sibelia.cpp:351-365; the facade rethrows).
 *   - input chromosomes must already be sanitised to upper-case ACGT: the reference replaces every other
 *     character with DEFINITE_BASE[rand() % 4] on the HOST in row-major order (src/indexedsequence.cpp:31-37);
 *     that step stays on the host (facade) because it consumes the process-wide glibc rand() stream.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with SIBGPU_ERR_CUDA.
 *   - one context per device; calls on one context are serialised by the caller (the reference is single-threaded).
 */
#ifndef SIBGPU_H_
#define SIBGPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sibgpu_ctx sibgpu_ctx;

typedef enum sibgpu_status {
	SIBGPU_OK = 0,
	SIBGPU_ERR_CUDA = 1,          /* no device / CUDA runtime error / out of device memory */
	SIBGPU_ERR_INVALID = 2,       /* bad argument (k == 0, NULL pointer, input too large for 32-bit positions, ...) */
	SIBGPU_ERR_INPUT = 3,         /* a character outside ACGT reached the device (caller skipped sanitising) */
	SIBGPU_ERR_INTERNAL = 4,      /* capacity fallback exhausted, fingerprint verification failed twice, ... */
	SIBGPU_ERR_STATE = 5          /* staged call made out of order (e.g. download before enumerate) */
} sibgpu_status;

/* == IndexedSequence::BifurcationInstance (src/indexedsequence.h:57-68): vertex id, chromosome, position in that
 *    strand's own coordinates (negative strand: position in the reverse complement of the chromosome). */
typedef struct sibgpu_inst { uint32_t bifId, chr, pos; } sibgpu_inst;

/* per-kernel device timing of the last enumerate on a context (CUDA events on the launching stream) */
typedef struct sibgpu_kernel_stat {
	const char *name;             /* kernel name, static storage */
	uint32_t launches;            /* launches in the last run */
	float ms;                     /* summed device time of those launches */
	uint64_t algo_bytes;          /* algorithmic bytes those launches had to move */
} sibgpu_kernel_stat;

const char *sibgpu_last_error(void);
const char *sibgpu_version(void);

/* number of visible CUDA devices (0 if none / no driver); never fails */
int sibgpu_device_count(void);

int sibgpu_create(int device, sibgpu_ctx **out);
void sibgpu_destroy(sibgpu_ctx *ctx);

/* Frees any buffer handed out by this library (instance arrays, sequences, position maps). */
void sibgpu_free(void *p);

/* ---------------------------------------------------------------------------------------------------------------
 * sibgpu_enumerate: replaces
 *     size_t IndexedSequence::EnumerateBifurcationsSArrayInRAM(const std::vector<std::string>& data,
 *            std::vector<BifurcationInstance>& positiveBif, std::vector<BifurcationInstance>& negativeBif)
 *     src/indexedsequence.h:73, src/vertexenumeration.cpp:263-364   (and its file-backed twin :160-261, which returns
 *     the same tables), called from IndexedSequence::Init, src/indexedsequence.cpp:40-47.
 * Host buffers in, host buffers out; the host<->device copies are inside the call.
 *   chr[i], len[i]   the nchr sanitised chromosomes (not NUL-terminated)
 *   k                vertex size, k >= 1
 *   *pos, *neg       library-allocated arrays sorted by (chr, pos) exactly like the reference's two vectors after
 *                    vertexenumeration.cpp:361-362; release with sibgpu_free
 *   *count           the reference's return value `bifurcationCount` (ids are 0 .. count-1; a vertex id is the
 *                    lexicographic rank of its k-mer among all vertex k-mers)
 */
int sibgpu_enumerate(sibgpu_ctx *ctx, const char *const *chr, const uint64_t *len, uint32_t nchr, uint32_t k,
	sibgpu_inst **pos, uint64_t *npos, sibgpu_inst **neg, uint64_t *nneg, uint32_t *count);

/* ---------------------------------------------------------------------------------------------------------------
 * Staged form of the same operation, for callers that keep the genome resident in HBM across calls (one upload,
 * several k: the reference builds five indexes per `-s loose` run, src/sibelia.cpp:242-289) and for bench.py's
 * device-resident timing.  sibgpu_enumerate == upload + enumerate_resident + download.
 */
int sibgpu_upload(sibgpu_ctx *ctx, const char *const *chr, const uint64_t *len, uint32_t nchr);
int sibgpu_enumerate_resident(sibgpu_ctx *ctx, uint32_t k, uint64_t *ninst_per_strand, uint32_t *count);
int sibgpu_download(sibgpu_ctx *ctx, sibgpu_inst **pos, uint64_t *npos, sibgpu_inst **neg, uint64_t *nneg);

/* Kernel statistics of the last enumerate on this context: fills at most cap entries, returns how many exist.
 * Timing is only recorded when enabled with sibgpu_set_profiling(ctx, 1) (it inserts events between kernels). */
int sibgpu_set_profiling(sibgpu_ctx *ctx, int enabled);
int sibgpu_kernel_stats(sibgpu_ctx *ctx, sibgpu_kernel_stat *out, int cap);
/* number of kernels this library launched during the last enumerate / simplify on this context */
uint64_t sibgpu_last_launches(sibgpu_ctx *ctx);
/* how often, over the life of the context, a fixed-capacity hash partition overflowed (a k-mer repeated millions of
 * times) and the enumeration fell back to exactly sized partitions (histogram pass); diagnostic */
uint64_t sibgpu_partition_fallbacks(sibgpu_ctx *ctx);
/* how often a shared-memory bucket (~1 Ki records, fixed capacity) overflowed -- one k-mer repeated hundreds of times --
 * and the enumeration regrouped the partitions through the L2-resident tables instead; diagnostic */
uint64_t sibgpu_bucket_fallbacks(sibgpu_ctx *ctx);
/* device time of the last sibgpu_enumerate_resident / device part of sibgpu_simplify on this context: milliseconds
 * between two CUDA events recorded on the library's stream around the whole operation */
float sibgpu_last_device_ms(sibgpu_ctx *ctx);

/* ---------------------------------------------------------------------------------------------------------------
 * sibgpu_simplify: replaces one stage of
 *     size_t BlockFinder::PerformGraphSimplifications(size_t k, size_t minBranchSize, size_t maxIterations,
 *            ProgressCallBack f)                              src/blockfinder.h:45, src/blockfinder.cpp:78-98
 * i.e. IndexedSequence(rawSeq_, originalPos_, k, tempDir_, true) + SimplifyGraph (blockfinder.cpp:16-51, which
 * calls RemoveBulges, bulgeremoval.cpp:330-430, for every vertex id and sweep) + the copy-back of :85-95.
 * State in/out is the reference's inter-stage state (src/blockfinder.h:52-54):
 *   seq[i] / origpos[i] / len[i]   rawSeq_[i], originalPos_[i] and their common length.  On success the three
 *                                  arrays are overwritten with library-allocated buffers of the new lengths
 *                                  (release with sibgpu_free); the caller keeps ownership of the old ones.
 *                                  EXCEPT when the stage finds no bulge at all (*bulges == 0, decided on the device
 *                                  from the resident enumeration: no host index is built): then seq / origpos / len
 *                                  are left exactly as passed -- compare seq[i] with the pointer handed in before
 *                                  adopting or freeing it.
 *   progress(done, state, user)    optional; same protocol as BlockFinder::ProgressCallBack (blockfinder.h:39):
 *                                  state 0 = start, 1 = run (<= 50 ticks), 2 = end
 *   *bulges                        the return value (cumulative number of collapsed bulges)
 */
typedef void (*sibgpu_progress_fn)(size_t done, int state, void *user);
int sibgpu_simplify(sibgpu_ctx *ctx, char **seq, uint32_t **origpos, uint64_t *len, uint32_t nchr,
	uint32_t k, uint32_t min_branch_size, uint32_t max_iterations,
	sibgpu_progress_fn progress, void *user, uint64_t *bulges);

/* ---------------------------------------------------------------------------------------------------------------
 * sibgpu_list_edges: replaces the pair
 *     IndexedSequence iseq(rawSeq_, originalPos_, k, tempDir_);
 *     ListEdges(iseq.Sequence(), iseq.BifStorage(), k, edge);
 * of BlockFinder::GenerateSyntenyBlocks (src/synteny.cpp:238-241) and SerializeCondensedGraph
 * (src/serialization.cpp:90-93): the edges of the condensed de Bruijn graph, i.e. one BlockFinder::Edge
 * (src/blockfinder.h:59-85, built by BlockFinder::ListEdges, src/serialization.cpp:56-86) per pair of consecutive
 * vertex marks on each strand of each chromosome, in the reference's order (positive strand: chromosomes in order,
 * positions ascending; then the negative strand likewise in its own coordinates).  No host-side index is built.
 *   seq / origpos / len   rawSeq_, originalPos_ (NULL = identity) and their lengths; seq must be sanitised ACGT
 *   *edges, *nedges       library-allocated array (release with sibgpu_free)
 */
typedef struct sibgpu_edge {
	uint32_t chr;
	uint32_t direction;           /* 0 = DNASequence::positive, 1 = DNASequence::negative (src/dnasequence.h:24-28) */
	uint32_t start_vertex, end_vertex;
	uint32_t actual_position, actual_length;
	uint32_t original_position, original_length;
	uint32_t first_char;          /* ASCII, read along the strand */
} sibgpu_edge;
int sibgpu_list_edges(sibgpu_ctx *ctx, const char *const *seq, const uint32_t *const *origpos, const uint64_t *len,
	uint32_t nchr, uint32_t k, sibgpu_edge **edges, uint64_t *nedges);

/* ---------------------------------------------------------------------------------------------------------------
 * sibgpu_trim_blocks: replaces the index-and-search part of
 *     bool BlockFinder::TrimBlocks(std::vector<Edge> & block, size_t trimK, size_t minSize)   src/synteny.cpp:31-122
 * (IndexedSequence iseq(blockSeq, trimK, "") + ConstructChrIndex + the walk over every vertex mark and all of its
 * instances; GenerateSyntenyBlocks calls it in a loop for every block, src/synteny.cpp:263).
 *   seq[i], len[i]     blockSeq[i] (sanitised ACGT), the original sequence spelled by block[i]
 *   direction[i]       block[i].GetDirection(): 0 = positive, 1 = negative
 *   out[i]             found = 0: no vertex of sequence i (read along its direction) occurs on another sequence
 *                      (the reference sets drop = true); otherwise start / end = element index (0-based position in
 *                      seq[i], positive-strand coordinates) of the reference's trimStart / trimEnd iterators.
 * The caller finishes like synteny.cpp:105-116 (size test against minSize, std::advance(trimEnd, trimK - 1), Edge).
 */
typedef struct sibgpu_trim { uint32_t found, start, end; } sibgpu_trim;
int sibgpu_trim_blocks(sibgpu_ctx *ctx, const char *const *seq, const uint64_t *len, const uint8_t *direction,
	uint32_t nchr, uint32_t trim_k, sibgpu_trim *out);

/* ---------------------------------------------------------------------------------------------------------------
 * sibgpu_fasta_parse: replaces
 *     size_t FASTAReader::GetSequences(std::vector<FASTARecord> & record)            src/fasta.cpp:22-73
 * (with ValidateHeader :75-90 and ValidateSequence :92-106) on the bytes of a whole FASTA file: line splitting, trimming,
 * '>' records (description = text up to the first blank), upper-casing and validation against "ACGTURYKMSWBDHWNX-",
 * all on the GPU, independent of how the sequence is wrapped.  Quirks kept: sequence lines before the first header are
 * glued to the first record; no header at all = one record with an empty description.
 *   data, nbytes     the file as read from disk (any line ending)
 *   out              nrec records {name (NUL-terminated), name_len, seq (NOT NUL-terminated), len}; the sequences are
 *                    slices of one pinned block; release everything with sibgpu_fasta_free
 *   *err_line        on SIBGPU_ERR_INPUT: the reference's line counter (non-empty lines, 1-based); sibgpu_last_error()
 *                    is the reference's `what` ("empty sequence", "empty header", "illegal character: x"), so that
 *                    "parse error in <file> on line <err_line>: <what>" is the reference's exception text
 */
typedef struct sibgpu_fasta_record { const char *name; uint64_t name_len; const char *seq; uint64_t len; } sibgpu_fasta_record;
typedef struct sibgpu_fasta {
	uint32_t nrec;
	sibgpu_fasta_record *rec;
	void *text_block, *name_block;                     /* owned storage behind the records */
	uint64_t total;                                    /* sum of the sequence lengths */
} sibgpu_fasta;
int sibgpu_fasta_parse(sibgpu_ctx *ctx, const char *data, uint64_t nbytes, sibgpu_fasta *out, uint64_t *err_line);
void sibgpu_fasta_free(sibgpu_fasta *f);

/* ---------------------------------------------------------------------------------------------------------------
 * Sharded enumeration over `world` GPUs of one box, ONE PROCESS PER GPU (k <= 32).  The concatenated genome is split
 * into `world` contiguous text ranges; records are bucketed by hash prefix so that partition p belongs to rank
 * p / (nparts_total / world); the caller moves them with one all-to-all (NCCL) between two device buffers it owns,
 * and all-gathers the (few) vertex keys so every rank can compute the global lexicographic ids.
 *
 *   sibgpu_dist_upload   every rank passes the whole input, only the bytes of its own range (+ halo) are copied
 *   sibgpu_dist_scan     -> nparts_total, hist[nparts_total] (records of this rank per partition), nrec_local
 *   sibgpu_dist_scatter  fills send_dev (nrec_local records of sibgpu_dist_record_bytes, ordered by partition)
 *   [all_gather hist -> counts[world][nparts_total]; all_to_all records: to rank r go this rank's partitions of r]
 *   sibgpu_dist_group    recv_dev = for each source rank in order, its records of my partitions in partition order
 *                        -> nkeys_local canonical vertex keys; sibgpu_dist_keys copies them to a device buffer
 *   [all_gather keys -> allkeys_dev, nkeys_total]
 *   sibgpu_dist_finish   global ids, instances of the own text range; sibgpu_download then returns the LOCAL tables,
 *                        both in text order (the caller concatenates ranks in order and reverses the negative table
 *                        inside each chromosome, vertexenumeration.cpp:361-362 -- see binding.assemble_tables).
 */
int sibgpu_dist_upload(sibgpu_ctx *ctx, const char *const *chr, const uint64_t *len, uint32_t nchr, uint32_t rank, uint32_t world);
int sibgpu_dist_scan(sibgpu_ctx *ctx, uint32_t k, uint32_t *nparts_total, uint32_t *hist, uint64_t *nrec_local);
uint32_t sibgpu_dist_record_bytes(sibgpu_ctx *ctx);
int sibgpu_dist_scatter(sibgpu_ctx *ctx, void *send_dev);
int sibgpu_dist_group(sibgpu_ctx *ctx, const void *recv_dev, const uint32_t *counts, uint64_t *nkeys_local);
int sibgpu_dist_keys(sibgpu_ctx *ctx, void *keys_dev);
int sibgpu_dist_finish(sibgpu_ctx *ctx, const void *allkeys_dev, uint64_t nkeys_total, uint64_t *ninst_local, uint32_t *count);

/* Peer variant of the same sharded enumeration: no histogram pass and no separate exchange step.
 *
 *   sibgpu_dist_scatter_local  pack + scatter the own text range into the context's own send buffer: one segment of
 *                              *seg_cap records per GLOBAL partition p (at record index p * seg_cap); counts[p] =
 *                              records written to segment p (counts must hold 1024 entries).  *overflow != 0: a
 *                              segment did not fit (a k-mer repeated very often) -- use the histogram path above.
 *   sibgpu_dist_export_send    CUDA IPC handle (64 bytes) of the send buffer
 *   [all_gather: counts, seg_cap, overflow, handle -- this is also the barrier after which every send buffer is final]
 *   sibgpu_dist_import_peers   handles[world][64]: maps the other ranks' send buffers (cached while unchanged)
 *   sibgpu_dist_group_peer     counts[world][nparts_total], seg_caps[world]: for every owned partition the insert kernel
 *                              reads its `world` segments directly from the peers' send buffers over NVLink (the
 *                              exchange is fused into the grouping kernel; nothing is received into a buffer)
 *   [all_gather keys (sibgpu_dist_keys) -- also the barrier before any send buffer may be rewritten]
 *   sibgpu_dist_finish         as above
 */
int sibgpu_dist_scatter_local(sibgpu_ctx *ctx, uint32_t k, uint32_t *nparts_total, uint64_t *counts, uint64_t *seg_cap, int *overflow);
/* sibgpu_dist_upload + sibgpu_dist_scatter_local in one call with the host-to-device copy of the own text range
 * pipelined against the pack and scatter kernels (pieces of 8 Mi positions, as in sibgpu_enumerate) */
int sibgpu_dist_upload_scatter(sibgpu_ctx *ctx, const char *const *chr, const uint64_t *len, uint32_t nchr, uint32_t rank,
	uint32_t world, uint32_t k, uint32_t *nparts_total, uint64_t *counts, uint64_t *seg_cap, int *overflow);
int sibgpu_dist_export_send(sibgpu_ctx *ctx, void *handle64);
int sibgpu_dist_import_peers(sibgpu_ctx *ctx, const void *handles);
int sibgpu_dist_group_peer(sibgpu_ctx *ctx, const uint64_t *counts, const uint64_t *seg_caps, uint64_t *nkeys_local);

/* Fused variant (k <= 28, and k > 32 through sibgpu_fused_run_fp): no collective and no host round trip inside a step.  Every rank owns one exported buffer
 * [header | fill cursors | vertex keys | one fixed-capacity segment per global partition]; the peers map it once
 * (CUDA IPC).  In a step a rank scatters its text range into its own segments and publishes a step counter in its
 * header; the level-2 split kernel of the partition owner waits for the peers' counters on the device and pulls the
 * segments with TMA bulk copies straight out of the peers' memory (the all-to-all, fused into the kernel); the vertex
 * keys are published the same way and concatenated by a pull kernel (the all-gather, fused).
 *
 *   sibgpu_fused_plan           layout of this call (every rank passes the whole input).  *need_alloc: 0 = the live
 *                               buffers fit, 1 = (re)allocation needed -- the SAME answer on every rank, the caller
 *                               then runs, collectively: barrier, sibgpu_fused_release_peers, barrier,
 *                               sibgpu_fused_alloc, all-gather of the 64-byte handles, sibgpu_fused_import;
 *                               -1 = not applicable (k = 29..32, input beyond the bucket fan-out): use the phased API above.
 *                               resident != 0: the text range is already in HBM (sibgpu_dist_upload).
 *   sibgpu_fused_run            one step.  *status: 0 = done (sibgpu_download returns the LOCAL tables, text order);
 *                               1 = a segment or bucket overflowed on some rank: every rank gets 1 and uses the phased
 *                               API; 2 = a key region was too small: plan again (need_alloc will be 1) and rerun.
 *                               A peer that never publishes makes the kernels give up after 4 s (SIBGPU_ERR_INTERNAL).
 */
int sibgpu_fused_plan(sibgpu_ctx *ctx, const char *const *chr, const uint64_t *len, uint32_t nchr, uint32_t rank, uint32_t world,
	uint32_t k, int resident, int *need_alloc);
int sibgpu_fused_release_peers(sibgpu_ctx *ctx);
int sibgpu_fused_alloc(sibgpu_ctx *ctx, void *handle64);
int sibgpu_fused_import(sibgpu_ctx *ctx, const void *handles);
int sibgpu_fused_run(sibgpu_ctx *ctx, const char *const *chr, const uint64_t *len, uint32_t nchr, int resident, uint32_t *count,
	uint64_t *ninst_local, int *status);

/* k > 32 (fingerprint classes, the -s loose stages k = 100, 1000, 5000 of src/util.cpp:50-61) on the same exported buffers.
 * The packed text is made complete on every rank by peer pulls (the rolling hashes, the string ranking of the vertex
 * classes and the verification of the instances read it beyond the own range); a class is {fingerprint, partition}.
 * The step has two halves because the representative occurrence of a class is its smallest text position over ALL ranks:
 *
 *   sibgpu_fused_run_fp         scatter, fused exchange, grouping, key pull, map, marking of the own range.
 *                               *status as sibgpu_fused_run (1: no phased API exists for k > 32 -- the caller lets every
 *                               rank run sibgpu_enumerate on the whole input and keeps its own range).
 *                               *rep_dev = device array of *nclasses 64-bit words: the caller min-reduces it over the
 *                               ranks in place (signed or unsigned: all values are < 2^63), e.g. ncclAllReduce(ncclMin)
 *   sibgpu_fused_finish_fp      string ranking (every rank ranks all classes), instance tables of the own range with
 *                               verification.  *collision != 0 on any rank (the caller max-reduces it): two different
 *                               k-mers shared a fingerprint inside a vertex class -- run both halves again with attempt + 1.
 */
int sibgpu_fused_run_fp(sibgpu_ctx *ctx, const char *const *chr, const uint64_t *len, uint32_t nchr, int resident, uint32_t attempt,
	uint64_t *nclasses, void **rep_dev, int *status);
int sibgpu_fused_finish_fp(sibgpu_ctx *ctx, uint32_t *count, uint64_t *ninst_local, int *collision);

/* Test hook (host only, no GPU needed): iteration order of the reference's boost::unordered_map<size_t, BranchData>
 * (Boost 1.54, src/bulgeremoval.cpp:168,203-215) after inserting n distinct keys in the given order, as restated in
 * sibelia_b200/csrc/boost_order.h.  out receives the n keys in begin()..end() order. */
void sibgpu_debug_unordered_order(const uint64_t *keys, uint64_t n, uint64_t *out);

/* Test hook (host only, no GPU needed): the O(instances) part of sibgpu_trim_blocks, fed with the two instance tables of
 * an enumeration of the block's sequences at trim_k (any source: the CPU tests pass the oracle's). */
void sibgpu_debug_trim_from_tables(const sibgpu_inst *pos, uint64_t npos, const sibgpu_inst *neg, uint64_t nneg,
	uint32_t count, const uint64_t *len, const uint8_t *direction, uint32_t nchr, sibgpu_trim *out);

#ifdef __cplusplus
}
#endif
#endif /* SIBGPU_H_ */
