/*
 * crx.h -- C ABI of the B200-native engine for crypto-recommendation's data-parallel hot path.
 *
 * The reference (YannisLamp/crypto-recommendation) has no FFI: its boundary is the set of header
 * templates main.cpp calls.  Every entry point below replaces one of those templates on flat,
 * contiguous buffers; the drop-in C++ headers under include/crx/ re-create the reference's own
 * signatures (CustVector / CustHashtable / std::vector arguments) on top of this ABI.
 * Citations are file:line in the reference tree.
 *
 * Conventions
 *   - plain pointers and sizes only; opaque handles own device memory;
 *   - every function returns CRX_OK (0) or a negative status; crx_last_error() gives the text
 *     (the reference has no error channel at all: cust_vector.hpp:111-115 prints and returns -1);
 *   - `mem` says where a caller buffer lives: CRX_HOST (library copies through pinned staging)
 *     or CRX_DEVICE (pointer is device memory on the context's GPU, used in place, stream-ordered);
 *   - rows are 0-based indices into a crx_points object -- they stand for the CustVector*
 *     pointers of the reference (ids are assumed unique, as its string-keyed caches require);
 *   - metric: CRX_EUCLIDEAN / CRX_COSINE  <->  metric_type "euclidean" / "cosine";
 *   - `seed` is the value the reference reads from system_clock::now() (lsh_cube.hpp:49,112;
 *     initialization.hpp:42,75); RNG consumption order is the reference's (libstdc++ <random>);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef CRX_H
#define CRX_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRX_OK 0
#define CRX_ERR_INVALID -1
#define CRX_ERR_CUDA -2
#define CRX_ERR_NOMEM -3
#define CRX_ERR_UNSUPPORTED -4

enum { CRX_F32 = 0, CRX_F64 = 1 };
enum { CRX_HOST = 0, CRX_DEVICE = 1 };
enum { CRX_EUCLIDEAN = 0, CRX_COSINE = 1 };

typedef struct crx_ctx crx_ctx;       /* one GPU + one stream; calls on a context are serialised */
typedef struct crx_points crx_points; /* std::vector<CustVector<T>> resident in HBM (cust_vector.hpp:23-36) */
typedef struct crx_lsh crx_lsh;       /* std::vector<CustHashtable<T>*> of create_LSH_hashtables */
typedef struct crx_cube crx_cube;     /* CustHashtable<T>* of create_hypercube */

/* ---- context ---- */
int crx_version(void);
const char* crx_last_error(void);
/* cuda_stream: a cudaStream_t to launch on (e.g. torch's current stream) or NULL for an own stream */
int crx_ctx_create(int device, void* cuda_stream, crx_ctx** out);
int crx_ctx_destroy(crx_ctx* ctx);
int crx_ctx_synchronize(crx_ctx* ctx);
/* number of this library's own kernels launched so far on the context */
int64_t crx_ctx_launch_count(const crx_ctx* ctx);
/* synchronises the stream and returns the memory held by the stream-ordered pool to the driver (freed buffers are kept by
 * default so that rebuilding tables every step costs no driver call) */
int crx_ctx_trim(crx_ctx* ctx);
/* per-kernel CUDA-event timing (on the context's stream).  enable!=0 starts recording every
 * launch; crx_ctx_kernel_time sums the launches whose name starts with `prefix`. */
int crx_ctx_profile(crx_ctx* ctx, int enable);
int crx_ctx_profile_reset(crx_ctx* ctx);
int crx_ctx_kernel_time(crx_ctx* ctx, const char* prefix, double* total_ms, int64_t* launches);
/* near-boundary / fallback counters since the last reset (see DESIGN.md "exactness"):
 * [0] hash projections recomputed in double-double, [1] top-P queries that needed the exhaustive
 * re-scan, [2] k-means++ draws within tolerance of a prefix boundary, [3] PAM rows re-summed exactly,
 * [4] Lloyd points re-evaluated exactly, [5] top-P queries with equal similarities among the P best whose
 * reference order (quicksort partition history) could not be reconstructed from the listed candidates: they are
 * returned in descending similarity, ties by row, [6] top-P queries with equal similarities among the P best (all),
 * [7] top-P queries decided by the second, targeted pass (threshold re-scan + exact similarities + literal sort).
 * With the second pass on (the default; CRX_TOPP_EXACT=0 switches it off) [1] and [5] stay 0: every query carries
 * the reference's list. */
int crx_ctx_counters(crx_ctx* ctx, int64_t out[8], int reset);

/* ---- points: vector<CustVector<T>> (cust_vector.hpp:23-72) ---- */
/* data: row-major [n][d] of dtype; copied.  d <= 128 in this round. */
int crx_points_create(crx_ctx* ctx, const void* data, int dtype, int64_t n, int32_t d, int mem, crx_points** out);
/* unknown_indexes / known_mean of every row (cust_vector.hpp:30-32): unknown[n][d] (1 = unknown), mean[n] */
int crx_points_set_ratings(crx_points* p, const uint8_t* unknown, const double* known_mean, int mem);
int crx_points_destroy(crx_points* p);
int64_t crx_points_n(const crx_points* p);
int32_t crx_points_d(const crx_points* p);

/* ---- vector math on rows (cust_vector.hpp:107-174); out[m] for row pairs (a[i], b[i]) ---- */
/* op: 0 inner_product, 1 euclideanDistance, 2 cosineDistance, 3 cosineSimilarity */
int crx_pair_op(crx_ctx* ctx, const crx_points* pa, const int32_t* a, const crx_points* pb, const int32_t* b, int64_t m,
                int op, double* out /* host */);

/* ---- LSH tables (lsh_cube.hpp:45-74; generators/*.hpp; cust_hashtable.hpp:51-70) ---- */
int crx_create_LSH_hashtables(crx_ctx* ctx, const crx_points* input_vectors, int metric, int k, int L,
                              int lsh_bucket_div, double euclidean_h_w, uint64_t seed, crx_lsh** out);
int crx_lsh_destroy(crx_lsh* t);
/* CustHashtable::getHash of every stored row: out[L][N] (cust_hashtable.hpp:123) */
int crx_lsh_bucket_ids(const crx_lsh* t, int32_t* out, int mem);
/* EuclideanPhiGen::getDetailedHashes: out[L][N][k] (euclidean_phi_gen.hpp:94); CRX_ERR_INVALID for cosine */
int crx_lsh_detailed_hashes(const crx_lsh* t, int32_t* out, int mem);
/* get_LSH_combined_buckets (filtered=0, lsh_cube.hpp:78) / get_LSH_filtered_combined_buckets
 * (filtered=1, lsh_cube.hpp:94) for a stored row; ascending row order; *count may exceed cap */
int crx_get_LSH_combined_buckets(const crx_lsh* t, int64_t query_row, int filtered, int32_t* out /* host */,
                                 int64_t cap, int64_t* count);
/* CustHashtable::getHash of a vector that is NOT stored in the tables (e.g. a query user, main.cpp:207):
 * bucket_ids[L]; detailed[L][k] (euclidean, may be NULL).  x[D] doubles on the host. */
int crx_lsh_hash_vector(const crx_lsh* t, const double* x, int32_t* bucket_ids, int32_t* detailed);
/* The same for every row of a point set that is not the stored one (all the query users of main.cpp:205-216 at once):
 * bucket_ids[L][n]; detailed[L][n][k] (euclidean, may be NULL); buffers in `mem`. */
int crx_lsh_hash_points(const crx_lsh* t, const crx_points* queries, int32_t* bucket_ids, int32_t* detailed, int mem);
/* hash parameters actually drawn (for inspection / parity): cosine r[L][k][D] doubles,
 * euclidean v[L][k][D] floats, t[L][k] floats, r_i[L][k] ints.  Any pointer may be NULL.  Host. */
int crx_lsh_params(const crx_lsh* t, double* cos_r, float* euc_v, float* euc_t, int32_t* euc_r);

/* ---- hypercube (lsh_cube.hpp:109-177; euclidean_f_gen.hpp; hypercube_gen.hpp) ---- */
int crx_create_hypercube(crx_ctx* ctx, const crx_points* input_vectors, int metric, int k, double euclidean_h_w,
                         uint64_t seed, crx_cube** out);
int crx_cube_destroy(crx_cube* c);
int crx_cube_vertex_ids(const crx_cube* c, int32_t* out /* [N] */, int mem);
/* get_hypercube_combined_buckets for a stored row: vertex visit order x insertion order */
int crx_get_hypercube_combined_buckets(const crx_cube* c, int64_t query_row, int probes, int32_t* out /* host */,
                                       int64_t cap, int64_t* count);
/* utils.cpp:22-50 (host helper, exposed for known-answer tests); returns count */
int crx_get_num_hamming_dist_from(int num, int dist, int min_bit, int bits, int32_t* out, int cap);

/* ---- clustering: initialisation (initialization.hpp:40-156) ---- */
int crx_rand_selection(crx_ctx* ctx, const crx_points* input_vectors, int cluster_num, uint64_t seed,
                       int32_t* centroid_rows /* host [K] */);
int crx_k_means_pp(crx_ctx* ctx, const crx_points* input_vectors, int cluster_num, int metric, uint64_t seed,
                   int32_t* centroid_rows /* host [K] */);

/* ---- clustering: assignment (assignment.hpp:55-217) ---- */
/* centroids: [K][D] doubles in `cmem`; centroid_rows[K] (host, may be NULL): row the centroid pointer
 * aliases or -1 -- reproduces `centroids[c]->setCluster(c,0)` (assignment.hpp:77-78).
 * labels[N] int32, dists[N] double in `mem`.  dists may be NULL (labels only: the k-means loop of main.cpp:96-103
 * never reads the stored distance; at K >= 32 this skips the exact-distance pass over the rows). */
int crx_lloyds_assignment(crx_ctx* ctx, const crx_points* input_vectors, const double* centroids, int cmem, int K,
                          const int32_t* centroid_rows, int metric, int32_t* labels, double* dists, int mem);
/* assignment.hpp:84-105: only rows with labels[v] == -1 are (re)assigned; labels/dists are in/out */
int crx_lloyds_for_remaining(crx_ctx* ctx, const crx_points* input_vectors, const double* centroids, int cmem, int K,
                             int metric, int32_t* labels, double* dists, int mem);
/* assignment.hpp:109-129 / 132-152: centroids are stored rows (after init or PAM).
 * labels_before_lloyd (nullable): labels after range_assignment, before lloyds_for_remaining. */
int crx_lsh_range_assignment(crx_ctx* ctx, const crx_points* input_vectors, const crx_lsh* lsh_hashtables,
                             const int32_t* centroid_rows, int K, int metric, int32_t* labels, double* dists, int mem,
                             int32_t* labels_before_lloyd);
/* lsh_range_assignment (assignment.hpp:109-129) for centroids that are ANY vectors -- the heap centres k_means leaves behind
 * (update.hpp:48) from the second iteration of {range assignment, k_means} on: centroids[K][D] doubles on the host, hashed
 * like a query (cust_hashtable.hpp:123); centroid_rows (nullable) = the stored row a centroid aliases, or -1.
 * shared_ids != 0 reproduces SURVEY App. A-2: the reference keys its distance cache by "<centroid id>to<vector id>"
 * (assignment.hpp:183-194) and every k_means centre is called "k_means_center", so all centroids share ONE cached distance
 * per vector -- the one of the lowest centroid whose bucket list holds it.  Pass it when all K centroid ids are equal
 * (the drop-in header does, from the ids themselves); with unique ids (shared_ids = 0) the cache is value-transparent. */
int crx_lsh_range_assignment_vectors(crx_ctx* ctx, const crx_points* input_vectors, const crx_lsh* lsh_hashtables,
                                     const double* centroids /* host [K][D] */, const int32_t* centroid_rows, int K, int metric,
                                     int shared_ids, int32_t* labels, double* dists, int mem, int32_t* labels_before_lloyd);
int crx_cube_range_assignment(crx_ctx* ctx, const crx_points* input_vectors, const crx_cube* hypercube,
                              const int32_t* centroid_rows, int K, int metric, int probes, int32_t* labels,
                              double* dists, int mem, int32_t* labels_before_lloyd);

/* ---- clustering: update (update.hpp:38-142) ---- */
/* per-cluster coordinate sums [K][D] and member counts [K] (device or host); the data-parallel
 * half of k_means -- the only part that needs an all-reduce when rows are sharded across GPUs */
int crx_cluster_sums(crx_ctx* ctx, const crx_points* input_vectors, const int32_t* labels, int lmem, int K,
                     double* sums, int64_t* counts, int mem);
/* means (empty cluster => zeros, cust_vector.hpp:189), convergence test of update.hpp:63-80;
 * new_centroids = centres after the call (old ones when *continue_clustering == 0). All host or all device. */
int crx_k_means_finish(crx_ctx* ctx, const double* sums, const int64_t* counts, const double* old_centroids, int K,
                       int D, int metric, double min_dist, double* new_centroids, int mem, int* continue_clustering);
/* k_means (update.hpp:38) = crx_cluster_sums + crx_k_means_finish on one GPU */
int crx_k_means(crx_ctx* ctx, const crx_points* input_vectors, const int32_t* labels, int lmem,
                const double* old_centroids, int K, int metric, double min_dist, double* new_centroids, int cmem,
                int* continue_clustering);
/* pam_lloyds (update.hpp:90): new_centroid_rows[K] host */
int crx_pam_lloyds(crx_ctx* ctx, const crx_points* input_vectors, const int32_t* labels, int lmem,
                   const int32_t* centroid_rows, int K, int metric, int32_t* new_centroid_rows, int* median_swapped);
/* silhouette_cluster (silhouette.hpp:32): sils[K+1] host */
int crx_silhouette_cluster(crx_ctx* ctx, const crx_points* input_vectors, const int32_t* labels, int lmem,
                           const double* centroids, int cmem, int K, int metric, double* sils);

/* ---- the same phases with the work split over several GPUs ---------------------------
 * One process per GPU.  The engine drives the algorithm and calls back into the host for the (few) exchange steps:
 * the caller supplies the three collectives (torch.distributed over NCCL in crypto_recommendation_b200/dist.py, or
 * ncclAllReduce & co. directly from C++).  A callback returns 0 on success; buffers are host memory (mem ==
 * CRX_HOST) or device memory of the context's GPU (CRX_DEVICE; the context's stream is idle when the callback runs
 * and the callback must have finished the transfer when it returns).  comm == NULL or world == 1: no exchange. */
enum { CRX_I32 = 2, CRX_I64 = 3 };               /* with CRX_F32 / CRX_F64: element types of a collective */
enum { CRX_SUM = 0, CRX_MAX = 1, CRX_MIN = 2 };
typedef struct crx_comm {
    int rank, world;
    void* user;
    int (*allreduce)(void* user, void* buf, int64_t count, int dtype, int op, int mem);                 /* in place */
    int (*allgather)(void* user, const void* send, void* recv, int64_t count_per_rank, int dtype, int mem);
    int (*broadcast)(void* user, void* buf, int64_t count, int dtype, int root, int mem);
    int stream_ordered;   /* 1: collectives on DEVICE buffers are enqueued on the context's stream (no synchronisation on
                           * either side; crx_comm_nccl_create); 0: the callback runs elsewhere and returns when done */
} crx_comm;
/* The same three collectives over NCCL inside libcrx.so (csrc/comm_nccl.cu), enqueued on the context's stream.  Rank 0 makes
 * the 128-byte id and hands it to the other ranks through the caller's bootstrap (torch.distributed in dist.py, a file, MPI);
 * every rank then calls crx_comm_nccl_create (collective).  NCCL is bound at run time (dlopen), preferring the copy the
 * process already holds.  crx_comm_nccl_calls: number of all-reduce / all-gather / broadcast calls issued so far. */
#define CRX_NCCL_ID_BYTES 128
int crx_comm_nccl_version(void);   /* ncclGetVersion, 0 when NCCL cannot be loaded */
int crx_comm_nccl_unique_id(uint8_t id[CRX_NCCL_ID_BYTES]);
int crx_comm_nccl_create(crx_ctx* ctx, const uint8_t id[CRX_NCCL_ID_BYTES], int rank, int world, crx_comm** out);
int crx_comm_nccl_calls(const crx_comm* comm, int64_t out[3]);
int crx_comm_nccl_destroy(crx_comm* comm);
#define CRX_ERR_COMM -5
/* k_means_pp with the points sharded by contiguous row range: this rank holds global rows
 * [row_offset, row_offset + n_local).  Per round: all-reduce(max) of the normaliser, all-gather of the per-shard
 * totals, the owner shard searches its prefix sums, broadcast of the chosen row and of its vector.
 * centroid_rows[K]: GLOBAL rows, identical on every rank; centroid_vectors[K][D] (host, nullable): their coordinates. */
int crx_k_means_pp_sharded(crx_ctx* ctx, const crx_points* local_vectors, int64_t row_offset, int64_t n_global,
                           int cluster_num, int metric, uint64_t seed, const crx_comm* comm, int64_t* centroid_rows,
                           double* centroid_vectors);
/* k_means over sharded rows: local cluster sums, all-reduce(sum) of K*D sums + K counts, redundant finish */
int crx_k_means_sharded(crx_ctx* ctx, const crx_points* local_vectors, const int32_t* labels, int lmem,
                        const double* old_centroids, int K, int metric, double min_dist, const crx_comm* comm,
                        double* new_centroids, int cmem, int* continue_clustering);
/* pam_lloyds / silhouette_cluster / range assignment with the points REPLICATED and the work split: candidate
 * medoid rows (PAM), silhouette rows, centroids (range search; the Lloyd pass over the remainder is split by
 * rows).  One all-reduce of an N-long device array per call; results identical on every rank and bit-identical
 * to the single-GPU call. */
int crx_pam_lloyds_sharded(crx_ctx* ctx, const crx_points* input_vectors, const int32_t* labels, int lmem,
                           const int32_t* centroid_rows, int K, int metric, const crx_comm* comm,
                           int32_t* new_centroid_rows, int* median_swapped);
int crx_silhouette_cluster_sharded(crx_ctx* ctx, const crx_points* input_vectors, const int32_t* labels, int lmem,
                                   const double* centroids, int cmem, int K, int metric, const crx_comm* comm, double* sils);
int crx_lsh_range_assignment_sharded(crx_ctx* ctx, const crx_points* input_vectors, const crx_lsh* lsh_hashtables,
                                     const int32_t* centroid_rows, int K, int metric, const crx_comm* comm,
                                     int32_t* labels, double* dists, int mem, int32_t* labels_before_lloyd);
int crx_cube_range_assignment_sharded(crx_ctx* ctx, const crx_points* input_vectors, const crx_cube* hypercube,
                                      const int32_t* centroid_rows, int K, int metric, int probes, const crx_comm* comm,
                                      int32_t* labels, double* dists, int mem, int32_t* labels_before_lloyd);

/* ---- recommendation (crypto_rec.hpp:214-345; loops of main.cpp:159-170, 205-216, 260-269, 353-373) ---- */
/* For queries [q_begin, q_end) of `queries` (NULL => the table's own rows, rec A):
 *   neighbours = get_LSH_filtered_combined_buckets; sims = get_P_closest(neighbours, user, P);
 *   recs = get_top_N_recom(neighbours, user, Nrec, sims).
 * Outputs (rows relative to q_begin; any may be NULL): recs[nq][Nrec] (-1 when the reference prints
 * nothing), nbr_rows[nq][P] (-1 padded), nbr_sims[nq][P], ncand[nq].  P <= 64 for cosine tables whose L k-bit bucket
 * ids pack into 32 bits (the tensor path), P <= 32 otherwise; every query carries the reference's own list (rows, order
 * among equal similarities, similarity doubles). */
int crx_recommend_lsh(crx_ctx* ctx, const crx_lsh* lsh_hashtables, const crx_points* queries, int64_t q_begin,
                      int64_t q_end, int P, int Nrec, int32_t* recs, int32_t* nbr_rows, double* nbr_sims,
                      int32_t* ncand, int mem);
/* The same call with a per-query status (the per-query form of counters [1] and [5]): status[q] = CRX_Q_EXACT when the
 * neighbour list is the reference's, order of equal similarities included; CRX_Q_PLATEAU when more equal similarities
 * reach the P-th place than the candidate list of the query holds; CRX_Q_TIE_ORDER when equal similarities among the P
 * best come back ordered by row because the reference's order could not be rebuilt.  Both only occur with the second
 * pass switched off (CRX_TOPP_EXACT=0): by default every query is CRX_Q_EXACT. */
enum { CRX_Q_EXACT = 0, CRX_Q_PLATEAU = 1, CRX_Q_TIE_ORDER = 2 };
int crx_recommend_lsh_status(crx_ctx* ctx, const crx_lsh* lsh_hashtables, const crx_points* queries, int64_t q_begin,
                             int64_t q_end, int P, int Nrec, int32_t* recs, int32_t* nbr_rows, double* nbr_sims,
                             int32_t* ncand, int32_t* status /* [nq] */, int mem);
/* neighbours = all rows of `users` with labels == qlabels[q]; get_top_N_recom without similarities
 * (crypto_rec.hpp:328).  queries NULL => users themselves with qlabels = labels. */
int crx_recommend_cluster(crx_ctx* ctx, const crx_points* users, const int32_t* labels, int lmem, int K,
                          const crx_points* queries, const int32_t* qlabels, int Nrec, int32_t* recs, int mem);
/* Per-user forms of the same functions, for the drop-in headers (include/crx/lib/crypto_rec.hpp):
 * get_P_closest (crypto_rec.hpp:214): similarities of the listed neighbours to the user, literal co-sort,
 *   first min(n, P) kept.  neighbor_rows[n] is reordered in place (host): its first min(n, P) entries are the reference's,
 *   the rest is some permutation of the remaining neighbours; similarities[min(n,P)] (host). */
int crx_get_P_closest(crx_ctx* ctx, const crx_points* users, int32_t* neighbor_rows, int64_t n, const crx_points* query_set,
                      int64_t query_row, int P, double* similarities, int64_t* kept);
/* get_predicted_user_sim (:281) and get_top_N_recom (:310 with similarities, :328 with similarities == NULL) for
 *   one user and an explicit neighbour list.  predicted[D] and recs[N] are host buffers; either may be NULL. */
int crx_get_top_N_recom(crx_ctx* ctx, const crx_points* users, const int32_t* neighbor_rows, const double* similarities,
                        int64_t n, const crx_points* query_set, int64_t query_row, int N, double* predicted, int32_t* recs);
/* parallel_quickSort (crypto_rec.hpp:269) on the device, one thread: known-answer tests only */
int crx_parallel_quickSort(crx_ctx* ctx, double* sims /* host, in/out */, int32_t* ids /* host, in/out */, int n);
/* The same sort when only the first `need` (<= 126) positions are consumed, as get_P_closest does (crypto_rec.hpp:225-228):
 * the warp-parallel closed form of the Lomuto partition that the second pass of crx_recommend_lsh runs on candidate sets
 * of any size.  The first min(n, need) entries of sims / ids are the literal sort's; the rest is unspecified. */
int crx_parallel_quickSort_topn(crx_ctx* ctx, double* sims /* host, in/out */, int32_t* ids /* host, in/out */, int n, int need);

/* ---- input ingest (vector_reader.hpp:55-85): binary columnar files in place of line-by-line getline + stod -----------
 * A file converted once (crx_columnar_write, tools/csv_to_columnar.py; values = strtod of every token = what the
 * reference's stod lambda returns, main.cpp:82) holds the ids and the coordinates as float64 columns:
 *   "CRXCOL1\0" | int64 n | int32 d | int32 0 | int64 ids_bytes | n NUL-terminated ids | zero padding to 8 bytes |
 *   float64 [d][n]
 * Loading = one mapped read, one upload, one transpose on the GPU into the engine's row-major layout. */
typedef struct crx_columnar crx_columnar;
int crx_columnar_write(const char* path, const char* const* ids, const double* rows /* host [n][d] */, int64_t n, int d);
int crx_columnar_open(const char* path, crx_columnar** out);
int64_t crx_columnar_n(const crx_columnar* f);
int32_t crx_columnar_d(const crx_columnar* f);
const char* crx_columnar_id(const crx_columnar* f, int64_t i);   /* valid until crx_columnar_close */
/* the vectors as a point set on the GPU (VectorReader::read + the packing every clustering call would do) */
int crx_columnar_points(crx_ctx* ctx, const crx_columnar* f, crx_points** out);
/* the same rows on the host, [n][d] doubles, transposed by the GPU (for callers that keep CustVector objects) */
int crx_columnar_rows(crx_ctx* ctx, const crx_columnar* f, double* rows /* host */);
int crx_columnar_close(crx_columnar* f);

/* ---- the step in front of the path: user rating vectors from tweet mentions ------------------------------
 * tweets_to_user_vectors (crypto_rec.hpp:79-140) and clusters_to_user_vectors (:143-210) after the strings are
 * resolved to indices: mention i says "tweet with sentiment mention_score[i] names coin mention_coin[i], posted
 * by user (or assigned to cluster) mention_user[i]", in the order the reference walks its tweets.  Per user:
 * X[u][c] = sum of the positive scores of its mentions of c (mention order), known coins = mentioned ones,
 * known_mean[u] = sum of the known coordinates (coin order) / their count, unknown coordinates := known_mean,
 * unknown[u][c] = 1 for those, keep[u] = 0 when every coordinate is 0 (the reference drops such users, :127;
 * their row is left as accumulated).  All buffers live in `mem`. */
int crx_user_vectors_build(crx_ctx* ctx, const int32_t* mention_user, const int32_t* mention_coin, const double* mention_score,
                           int64_t n_mentions, int64_t n_users, int n_coins, double* X, uint8_t* unknown, double* known_mean,
                           uint8_t* keep, int mem);

#ifdef __cplusplus
}
#endif
#endif /* CRX_H */
