// Internal launch interfaces between the C-ABI layer (vq_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define VQ_DTYPE_F32 0
#define VQ_DTYPE_BF16 1
#define VQ_DTYPE_F16 2
#define VQ_DTYPE_F32X3 3

namespace vq {

// ---- K0: ingest / query preparation (vq_prep.cu)
// src [n, d] (src_dtype, row stride ld_src) -> dst [n, ld_dst] (dst_dtype); optional row L2-normalisation;
// columns d..ld_dst-1 are zero-filled.
// rows [n, n_pad) of dst (if n_pad > n) are written as zero rows (the padded query operand of the tensor-core path)
int prep_rows_launch(const void* src, int src_dtype, long long ld_src, void* dst, int dst_dtype, long long ld_dst,
                     long long n, int d, int l2norm, cudaStream_t stream, long long n_pad = 0);
// VQ_F32X3 storage: [n, 6*seg] bf16 per row, six segments of seg columns.  Corpus side: lo hi mid mid hi hi;
// query side: hi lo mid hi mid hi (segment s of a query row times segment s of a corpus row are the six product terms,
// smallest first).  split: src rows -> segments (optional L2 normalisation first, same arithmetic as prep_rows);
// join: corpus-side segments -> hi + (mid + lo), exactly the fp32 value that was split.
int split_rows_launch(const void* src, int src_dtype, long long ld_src, void* dst, int seg, long long n, int d, int l2norm,
                      int query_side, cudaStream_t stream, long long n_pad = 0);
int join_rows_launch(const void* src, int seg, void* dst, int dst_dtype, long long ld_dst, long long n, int d, cudaStream_t stream);


// ---- K3: tensor-core contraction + fused top-k (vq_tc.cu)
struct TcPlan {
    int cg;               // CTAs per cluster (1 or 2)
    int m_tiles;          // query tiles of 128*cg
    int n_tiles;          // corpus tiles of 256 rows
    int n_chunks;         // residue classes G of corpus tiles; work item = (m_tile, g), tiles g, g+G, ...
    int tiles_per_chunk;
    int list_sets;        // candidate-list sets written: n_chunks * column halves
    int classes_per_slab; // classes that share one contiguous slab of tiles (== n_chunks when there is one round)
    int nq_pad;           // m_tiles * 128 * cg
    int n_clusters;       // persistent clusters launched
};
struct TcParams {
    uint64_t* lists;      // [list_sets][nq_pad][VQ_CAP] candidate (score, row) pairs
    int* counts;          // [list_sets][nq_pad]
    long long n_rows;
    int nq, nq_pad, k, num_kb, m_tiles, n_tiles, n_chunks, tiles_per_chunk;
    int tile_begin;       // first VIRTUAL tile index of this pass (n_tiles is the virtual end); see map_stride
    // virtual -> physical tile: j = t - tile_begin; if (map_ratio) j = (j / (map_ratio-1)) * map_ratio + j % (map_ratio-1) + 1;
    // physical = j * map_stride.  Pass p of the priming schedule takes the multiples of map_stride that the earlier
    // (sparser) passes did not take, so every pass samples the WHOLE corpus, not a prefix.
    int map_stride, map_ratio;
    int classes_per_slab;
    int* progress;        // [m_tiles * n_chunks] tile counters of the L2 sharing window (may be null)
    int sync_every, sync_window;
    int list_set_base;    // first candidate-list set this pass writes
    const float* init_thr;  // optional primed per-query thresholds: init_thr[q * init_thr_stride]
    int init_thr_stride;
    // optional exclusive UPPER bound per query (banded rounds for k > 128): only candidates that rank strictly
    // after (ub_score[q], ub_row[q]) are taken
    const float* ub_score;
    const uint32_t* ub_row;
    uint32_t idesc_base;
    int pdl;              // programmatic dependent launch: trigger the dependents at the start; only the EPILOGUE waits for the
                          // preceding kernel (the merge that produces init_thr) — producer and MMA warps run ahead
};
// one pass over corpus tiles [tile_begin, tile_end): plan must come from tc_plan(nq, (tile_end - tile_begin) * 256 rows)
struct TcPass {
    int tile_begin, tile_end, list_set_base;
    int map_stride, map_ratio;
    const float* init_thr;
    int init_thr_stride;
    const float* ub_score;
    const uint32_t* ub_row;
    int pdl;   // 0 = plain launch; 1 = kernel triggers its dependents early, launched plainly itself (first pass);
               // 2 = also launched with programmatic stream serialization (its epilogue waits for the preceding merge)
};
void tc_plan(int nq, long long n_rows, int n_sms, int force_cg, TcPlan* plan);
int tc_topk_launch(const void* q16, int nq, int ld_q, const void* x16, long long n_rows, int ld_x, int d, int dtype, int k,
                   const TcPlan& plan, const TcPass& pass, uint64_t* lists, int* counts, int* progress, cudaStream_t stream);

// ---- K1: fp32 exact SIMT contraction + fused top-k (vq_fp32.cu)
struct F32Plan {
    int m_tiles, n_tiles, n_chunks, tiles_per_chunk, nq_pad, n_ctas;
};
void f32_plan(int nq, long long n_rows, int n_sms, F32Plan* plan);
// K1's operands are K-MAJOR tiles of 128 rows: [tile][kpad][128] floats, kpad = f32_kpad(d) (d rounded up to 32, zero columns)
int f32_kpad(int d);
// [n][d] rows of src_dtype -> rows row0 .. row0+n-1 of a tiled matrix (+ zero rows up to n_pad); cast/L2norm as prep_rows_launch
int tile_rows_launch(const void* src, int src_dtype, long long ld_src, float* dst, long long row0, int kpad, long long n, int d, int l2norm,
                     cudaStream_t stream, long long n_pad = 0);
// rows row0 .. row0+n-1 of a tiled matrix -> [n][ld_dst] rows of dst_dtype
int untile_rows_launch(const float* src, long long row0, int kpad, void* dst, int dst_dtype, long long ld_dst, long long n, int d, cudaStream_t stream);
// one pass over corpus tiles (128 rows each) [tile_begin, tile_end); plan from f32_plan(nq, (tile_end - tile_begin) * 128 rows);
// qt: the prepared queries as plan.m_tiles tiles, xt: the corpus tiles
int f32_topk_launch(const float* qt, int nq, const float* xt, long long n_rows, int d, int k,
                    const F32Plan& plan, int tile_begin, int tile_end, int map_stride, int map_ratio, const float* init_thr,
                    const float* ub_score, const uint32_t* ub_row, uint64_t* lists, int* counts, cudaStream_t stream);

// ---- K4: merges (vq_merge.cu)
// Merge step of one pass.  Candidate lists are [set][nq_pad][VQ_CAP] raw pairs + counts[set][nq_pad]; set 0 is
// the RUNNING merged top-k of all passes so far.  Gathers set 0 (if use_set0) + sets [set_begin, set_begin+n_sets),
// keeps the best k in set 0, writes every query's k-th score to thr_out (may be null) and, if out_scores is
// given, the final [nq][k] scores (f32) / ids (i64, + id_offset).  expected_total sizes the shared-memory sort.
// pdl != 0: launched with programmatic stream serialization (the kernel triggers ITS dependents at once and waits for the
// preceding contraction kernel itself) — only for grids that are fully resident at once (few queries)
int merge_lists_launch(uint64_t* lists, int* counts, int use_set0, int set_begin, int n_sets, int nq_pad, int nq, int k, long long id_offset,
                       float* thr_out, float* out_scores, long long* out_ids, int expected_total, cudaStream_t stream, int pdl = 0);
// banded rounds (k > 128): copies one round's [nq][kr] result into columns [col0, col0+kr) of the final [nq][k]
// outputs (+ id_offset) and sets the next round's exclusive upper bound to the round's last entry
int band_scatter_launch(const float* tmp_s, const long long* tmp_i, int nq, int kr, int k, int col0, long long id_offset,
                        float* out_scores, long long* out_ids, float* ub_score, uint32_t* ub_row, cudaStream_t stream);
// per-shard results [n_lists][nq][k_in] (scores f32, ids i64; id < 0 = padding) -> [nq][k_out]
// list c starts at scores + c * stride_s and ids + c * stride_i (elements)
int merge_topk_launch(const float* scores, const long long* ids, long long stride_s, long long stride_i, int n_lists, int nq, int k_in,
                      int k_out, float* out_scores, long long* out_ids, cudaStream_t stream);

// ---- cross-GPU exchange over NVLink peer memory + fused pull-merge (vq_xchg.cu)
#define VQ_XCHG_MAX_RANKS 16
struct XchgControl {          // head of every rank's exchange window (device memory, mapped by all peers)
    unsigned seq;             // this rank's sequence number (local; bumped by publish)
    unsigned done;            // CTA counter of the publish kernel in flight
    unsigned pad[14];
    unsigned flag[VQ_XCHG_MAX_RANKS];  // flag[r] = last sequence number rank r published (written REMOTELY by rank r)
    unsigned pad2[16];
};
struct XchgPeers {            // peer-mapped pointers, indexed by rank (own rank = local pointers)
    XchgControl* ctl[VQ_XCHG_MAX_RANKS];
    uint8_t* slots[VQ_XCHG_MAX_RANKS];
};
int xchg_publish_launch(XchgControl* ctl, uint8_t* slots, size_t slot_bytes, const void* src, size_t bytes, const XchgPeers& peers, int rank,
                        int world, cudaStream_t stream);
int xchg_merge_launch(const XchgControl* ctl, const XchgPeers& peers, size_t slot_bytes, int world, int nq, int k_in, int k_out,
                      float* out_scores, long long* out_ids, cudaStream_t stream);
int xchg_gather_launch(const XchgControl* ctl, const XchgPeers& peers, size_t slot_bytes, int world, size_t bytes, void* dst, cudaStream_t stream);

// ---- K6: run post-processing (vq_post.cu)
int expand_hits_launch(const float* scores, const long long* ids, int nq, int k_in, const long long* map_offsets,
                       const long long* map_targets, int many2one_max, int k, int out_cap, double* out_scores, long long* out_ids,
                       int* out_counts, int* overflow, cudaStream_t stream);

// ---- adjacent rows (vq_inbatch.cu): in-batch similarity of the bi-encoder evaluation step; hard-negative marking
int inbatch_launch(const void* q, const void* c, int dtype, int n_q, int n_c, int d, const long long* labels, long long ignore_index,
                   float* log_probs, float* nll, int* rank, cudaStream_t stream);
int mark_relevant_launch(const long long* ids, int nq, int k, const long long* rel_off, const long long* rel_ids, unsigned char* is_rel,
                         cudaStream_t stream);

// ---- K5: score fusion (vq_fuse.cu)
struct FuseParams {
    int n_runs, nq, k_in, u_max;
    int defmin;       // fill missing docs with the run's per-query minimum (meerqat/ir/fuse.py:132-149)
    int norm;         // 0 none, 1 zmuv (per query), 2 gzmuv (global, fuse.py:86-129), 3 min-max
};
size_t fuse_workspace_bytes(int n_runs, int nq, int k_in);
int fuse_launch(const double* scores, const long long* ids, const double* weights, const FuseParams& fp,
                double* out_scores, long long* out_ids, int* out_counts, void* workspace, size_t ws_bytes, cudaStream_t stream);
// grid search of Fusion.fit: every weight vector of weights[n_weights][n_runs] evaluated against the relevant documents
// rel[nq][max_rel] (doc ids, < 0 padding); metric 0 mrr, 1 precision, 2 hit_rate, 3 recall, all @cut (0 = no cut);
// out_metric[n_weights] = mean over n_judged queries, out_best = index of the first best weight vector
size_t fuse_grid_workspace_bytes(int n_runs, int nq, int k_in, int n_weights);
int fuse_grid_launch(const double* scores, const long long* ids, const FuseParams& fp, const double* weights, int n_weights, const long long* rel,
                     int max_rel, const int* n_rel_total, int n_judged, int metric, int cut, double* out_metric, int* out_best, void* workspace,
                     size_t ws_bytes, cudaStream_t stream);

}  // namespace vq
