/* viquae_b200 — C ABI of the B200-native exact inner-product top-k + score-fusion engine.
 *
 * This is the drop-in boundary for the one hot path of PaulLerner/ViQuAE (meerqat) that the engine
 * replaces: the FAISS IndexFlatIP object that meerqat.ir.search reaches through HF datasets.
 * Each entry point names the reference interface it stands in for (paths relative to the reference
 * checkout; `datasets/` is the HF datasets dependency, faiss is un-vendored third-party C++).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * error code (never throws); vq_last_error() gives the message for the calling thread.  An index is
 * bound to one CUDA device and is thread-compatible, not thread-safe (the reference has a single
 * caller thread: Dataset.map without num_proc, meerqat/ir/search.py:482).  Functions taking a
 * `stream` enqueue asynchronously on it and do not synchronise; the *_host variants synchronise.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef VIQUAE_B200_H
#define VIQUAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element types of vectors crossing the ABI */
#define VQ_F32 0
#define VQ_BF16 1
#define VQ_F16 2
/* storage type only (vq_index_create): every fp32 value kept as three bf16 terms hi + mid + lo (exactly the fp32
 * value), contracted on the tensor cores as hi*hi + hi*mid + mid*hi + mid*mid + hi*lo + lo*hi with fp32
 * accumulation: fp32-faithful scores (the dropped terms are below 2^-24 of a product) at tensor-core speed.
 * 12 bytes per element on the device.  Stands in for the fp32 sgemm FAISS runs (datasets/search.py:384). */
#define VQ_F32X3 3

/* error codes */
#define VQ_OK 0
#define VQ_ERR_CUDA (-1)        /* a CUDA runtime/driver call failed (see vq_last_error) */
#define VQ_ERR_INVALID (-2)     /* bad argument (shape, dtype, k, null pointer) */
#define VQ_ERR_UNSUPPORTED (-3) /* valid request the engine does not serve (e.g. k > 2048) */
#define VQ_ERR_NOMEM (-4)       /* workspace too small / allocation failed */

/* kernel variants (vq_index_set_variant): 0 lets the engine choose from (dtype, nq) */
#define VQ_VARIANT_AUTO 0
#define VQ_VARIANT_TC_CG1 1   /* tcgen05 128x256 tiles, one CTA per SM */
#define VQ_VARIANT_TC_CG2 2   /* tcgen05 cta_group::2, 256x256 tiles per CTA pair */

typedef struct vq_index vq_index;

/* Library / device probe.  Returns the number of CUDA devices (0 = none) or a negative error. */
int vq_device_count(void);
const char* vq_last_error(void);
const char* vq_version(void);

/* faiss.IndexFlat(d, faiss.METRIC_INNER_PRODUCT)            [datasets/search.py:276-288]
 * `dtype` is the STORAGE type of the corpus on the device (VQ_F32 = exact fixed-order mode, VQ_BF16 / VQ_F16 =
 * tensor-core mode on rounded embeddings, VQ_F32X3 = tensor-core mode on split fp32 embeddings); `capacity` pre-sizes the device buffer in rows (it grows on demand).
 * The device layout is the engine's own (VQ_F32: k-major tiles of 128 rows, d padded to a multiple of 32; VQ_F32X3:
 * six bf16 segments per row); rows always cross this interface as plain [n, d] row-major arrays. */
int vq_index_create(vq_index** out, int d, int dtype, int device, uint64_t capacity);
void vq_index_destroy(vq_index* idx);

/* faiss Index::add(n, x)                                    [datasets/search.py:311-313]
 * x: [n, d] row-major, host or device pointer, element type src_dtype.  The engine copies, converts
 * to the storage dtype and (l2norm != 0) L2-normalises rows on ingest, which is what the
 * "L2norm,Flat" factory string does on add (meerqat/ir/search.py:230-245).  At most 2^31-256 rows per index
 * (shard): rows travel as uint32 in the candidate keys and as int32 TMA coordinates. */
int vq_index_add(vq_index* idx, const void* x, uint64_t n, int src_dtype, int l2norm, void* stream);

/* faiss Index::ntotal / Index::d */
uint64_t vq_index_ntotal(const vq_index* idx);
int vq_index_dim(const vq_index* idx);
int vq_index_dtype(const vq_index* idx);

/* Row-sharding support: ids returned by search are local row + id_offset. */
int vq_index_set_id_offset(vq_index* idx, int64_t id_offset);
int vq_index_set_variant(vq_index* idx, int variant);

/* Measurement support (bench.py): when enabled, every vq_index_search records a CUDA event pair around
 * its dominant kernel (the contraction + fused top-k) on the launch stream; read_timing waits for the
 * recorded pairs, returns their summed duration and count, and resets.  At most 256 searches are kept. */
int vq_index_set_timing(vq_index* idx, int enable);
int vq_index_read_timing(vq_index* idx, double* total_ms, int* n_launches);
/* Cumulative number of CUDA kernels vq_index_search has launched on this index (bench.py: gpu_launches). */
uint64_t vq_index_launch_count(const vq_index* idx);

/* faiss Index::reconstruct_n — copy rows [start, start+n) out as dst_dtype, [n, d] row-major,
 * host or device destination.  Used by save/load (meerqat/ir/search.py:234-248). */
int vq_index_copy_rows(const vq_index* idx, uint64_t start, uint64_t n, void* dst, int dst_dtype, void* stream);

/* Bytes of device workspace vq_index_search needs for (nq, k) on this index. */
size_t vq_index_workspace_bytes(const vq_index* idx, int nq, int k);

/* faiss Index::search(n, x, k, D, I)                        [datasets/search.py:384; meerqat/ir/search.py:146]
 * All pointers are DEVICE pointers.  q: [nq, d] of q_dtype; l2norm_q != 0 applies the reference's
 * L2norm(queries) (meerqat/ir/search.py:43-46,144-145).  Outputs: out_scores float32 [nq, k] sorted
 * descending, out_ids int64 [nq, k]; rows with fewer than k hits are padded with (-FLT_MAX, -1);
 * equal scores are ordered by lower id.  k <= 128 is one fused search; 128 < k <= 2048 runs ceil(k/128) banded
 * rounds (each round takes what ranks strictly after the previous round's last hit), still exact.
 * Asynchronous on `stream`. */
int vq_index_search(vq_index* idx, const void* q, int nq, int q_dtype, int l2norm_q, int k, float* out_scores,
                    int64_t* out_ids, void* workspace, size_t workspace_bytes, void* stream);

/* Same call with HOST buffers — exactly the shape of faiss_index.search(queries, k) as HF datasets
 * calls it with numpy arrays: float32 queries in, float32 scores / int64 ids out.  Host<->device
 * copies and the engine-owned workspace are inside the call; returns after the results are on the host. */
int vq_index_search_host(vq_index* idx, const float* q_host, int nq, int l2norm_q, int k, float* out_scores_host,
                         int64_t* out_ids_host);

/* meerqat.ir.search.L2norm (search.py:43-46) / the dtype cast of np.array(queries, dtype=float32)
 * (search.py:143) as a standalone device operation: src [n, d] -> dst [n, d], optional row-wise
 * x / ||x||_2 (no epsilon: a zero row becomes NaN like the reference).  Device pointers. */
int vq_prep_rows(const void* src, int src_dtype, void* dst, int dst_dtype, uint64_t n, int d, int l2norm, void* stream);

/* Cross-GPU k-way merge after the all-gather of per-shard results.  Device pointers:
 * scores float32 / ids int64 [n_lists, nq, k_in] (id < 0 = padding) -> [nq, k_out]. */
int vq_merge_topk(const float* scores, const int64_t* ids, int n_lists, int nq, int k_in, int k_out, float* out_scores,
                  int64_t* out_ids, void* stream);
/* Same merge when every rank's (scores, ids) travel in ONE packed all-gather buffer: list c starts
 * list_stride_scores floats / list_stride_ids int64s after list c-1. */
int vq_merge_topk_strided(const float* scores, const int64_t* ids, int64_t list_stride_scores, int64_t list_stride_ids, int n_lists,
                          int nq, int k_in, int k_out, float* out_scores, int64_t* out_ids, void* stream);

/* ---- Cross-GPU exchange over NVLink peer memory (one process per GPU on one node), fused with the k-way merge.
 * Stands where the reference has nothing (FAISS replicates the index per GPU, datasets/search.py:329-340) and where
 * an ncclAllGather + vq_merge_topk pair would go.  Every rank creates a window, exports its 64-byte IPC handle,
 * the host exchanges the handles over any side channel (the Python layer uses its process group's object all-gather) and
 * attaches every peer.  Per search: vq_xchg_publish(packed local result) then vq_xchg_merge_topk, both asynchronous
 * on `stream`, no host synchronisation, CUDA-graph capturable (sequence numbers live in device memory).  Collective:
 * every rank makes the same calls in the same order. */
typedef struct vq_xchg vq_xchg;
#define VQ_XCHG_HANDLE_BYTES 64
int vq_xchg_create(vq_xchg** out, int device, int rank, int world_size, size_t slot_bytes);
int vq_xchg_export(vq_xchg* x, void* handle_out /* VQ_XCHG_HANDLE_BYTES */);
int vq_xchg_attach(vq_xchg* x, int peer_rank, const void* handle /* VQ_XCHG_HANDLE_BYTES, from that rank's export */);
size_t vq_xchg_slot_bytes(const vq_xchg* x);
/* src: device pointer to this rank's contribution (packed result: scores f32 [nq, k] | pad to 8 | ids i64 [nq, k]) */
int vq_xchg_publish(vq_xchg* x, const void* src, size_t bytes, void* stream);
/* waits (on the device) until every rank published, pulls each query's k_in candidates of every rank straight from the
 * peers' windows and writes the global top-k_out: out_scores f32 / out_ids i64 [nq, k_out] device pointers */
int vq_xchg_merge_topk(vq_xchg* x, int nq, int k_in, int k_out, float* out_scores, int64_t* out_ids, void* stream);
/* generic consumer: dst[r * bytes, (r+1) * bytes) = what rank r published (bytes % 16 == 0); used to rebuild a query batch
 * from per-rank slices */
int vq_xchg_gather(vq_xchg* x, size_t bytes, void* dst, void* stream);
void vq_xchg_destroy(vq_xchg* x);

/* Searcher.__call__'s per-hit post-processing (meerqat/ir/search.py:413-440) for one index and one query
 * batch: article -> passage expansion through a CSR mapping (map_offsets int64 [n_rows+1], map_targets int64
 * [nnz]; both NULL = identity), many2one (0 = overwrite like a dict, 1 = 'max'), stop after the hit that
 * brings the run to >= k documents.  Device pointers.  scores f32 / ids i64 [nq, k_in] as returned by
 * vq_index_search.  Outputs in dict INSERTION order: out_ids i64 / out_scores f64 [nq, out_cap] (-1 padding),
 * out_counts i32 [nq]; out_cap must be >= k + (largest fan-out of the mapping) - 1, else *overflow_flag = 1. */
int vq_expand_hits(const float* scores, const int64_t* ids, int nq, int k_in, const int64_t* map_offsets, const int64_t* map_targets,
                   int many2one_max, int k, int out_cap, double* out_scores, int64_t* out_ids, int32_t* out_counts, int32_t* overflow_flag,
                   void* stream);

/* meerqat.ir.fuse: default_minimum (fuse.py:132-149) -> normalisation (gzmuv fuse.py:86-129; ranx
 * zmuv / min-max) -> ranx wsum -> descending sort, for n_runs dense runs of one query batch.
 * Device pointers.  scores float64 / ids int64 [n_runs, nq, k_in] (id < 0 = absent entry);
 * weights float64 [n_runs].  norm: 0 none, 1 zmuv, 2 gzmuv, 3 min-max.  Outputs: the fused union per
 * query, out_scores float64 / out_ids int64 [nq, n_runs*k_in] sorted descending (padding -1),
 * out_counts int32 [nq] = size of the union.  The union is NOT truncated to k, like ranx.fuse. */
size_t vq_fuse_workspace_bytes(int n_runs, int nq, int k_in);
int vq_fuse(const double* scores, const int64_t* ids, const double* weights, int n_runs, int nq, int k_in, int defmin,
            int norm, double* out_scores, int64_t* out_ids, int32_t* out_counts, void* workspace, size_t workspace_bytes,
            void* stream);

/* Fusion.fit (meerqat/ir/fuse.py:193-217 -> ranx.optimize_fusion): the grid search over weight vectors as one
 * device pass.  Same inputs as vq_fuse plus weights float64 [n_weights, n_runs] (the reference's grid: step 0.1,
 * summing to 1), the relevant documents of every query rel int64 [nq, max_rel] (doc ids, < 0 = padding; at most 256
 * of them can lie in a query's union), n_rel_total int32 [nq] (relevant documents in the qrels, retrieved or not:
 * recall's denominator) and n_judged (queries in the qrels: the mean's denominator, like ranx).  metric: 0 mrr,
 * 1 precision, 2 hit_rate, 3 recall, each @cut (cut = 0: no cut).  Union + default-minimum + normalisation happen
 * once; per weight vector only the RANKS of the relevant documents are computed (a counting pass, no sort).
 * Outputs: out_metric float64 [n_weights], out_best int32 [1] = index of the first best weight vector. */
size_t vq_fuse_grid_workspace_bytes(int n_runs, int nq, int k_in, int n_weights);
int vq_fuse_grid(const double* scores, const int64_t* ids, int n_runs, int nq, int k_in, int defmin, int norm, const double* weights,
                 int n_weights, const int64_t* rel, int max_rel, const int32_t* n_rel_total, int n_judged, int metric, int cut,
                 double* out_metric, int32_t* out_best, void* workspace, size_t workspace_bytes, void* stream);

/* In-batch similarity of the bi-encoder's evaluation step (meerqat/train/trainee.py:454-458 + train/metrics.py:22-37),
 * forward only: similarities = Q . C^T (fp32 accumulation), log_softmax over each row, NLL of the labelled context and
 * its 1-based rank in the row (ties: lower index first).  Device pointers.  q [n_q, d], c [n_c, d] of `dtype`; labels
 * int64 [n_q] (ignore_index or out-of-range = ignored: nll 0, rank 0); log_probs float32 [n_q, n_c] or NULL;
 * nll float32 [n_q]; rank int32 [n_q].  The caller averages nll over the non-ignored rows (the NLL loss's 'mean' reduction). */
int vq_inbatch_logprobs(const void* q, const void* c, int dtype, int n_q, int n_c, int d, const int64_t* labels, int64_t ignore_index,
                        float* log_probs, float* nll, int32_t* rank, void* stream);

/* Hard-negative mining after a search (EXPERIMENTS.rst:78-103: retrieved ids split into relevant ones and
 * `search_irrelevant_indices`): is_relevant[q, j] = 1 iff ids[q, j] is in the query's ASCENDING relevant-id list
 * rel_ids[rel_offsets[q] .. rel_offsets[q+1]).  Device pointers; ids int64 [nq, k] as returned by vq_index_search. */
int vq_mark_relevant(const int64_t* ids, int nq, int k, const int64_t* rel_offsets, const int64_t* rel_ids, uint8_t* is_relevant, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIQUAE_B200_H */
