/*
 * geoa3_b200.h — C ABI of libgeoa3_b200.so: the B200 (sm_100a) kernels behind GeoA3's
 * geometry-aware loss path and the pointnet2_ops sampling/grouping ops.
 *
 * Conventions (all entry points):
 *   - plain device pointers + int sizes + a CUDA stream handle (cudaStream_t passed as void*);
 *     no torch types.  The caller allocates every input, output and workspace buffer and keeps
 *     it alive until the stream has passed the call; the library never allocates user-visible
 *     memory, never synchronises, and launches on the stream it is given (the reference launches
 *     on at::cuda::getCurrentCUDAStream(): _ext-src/src/sampling_gpu.cu:25-27,180).
 *   - return value: 0 = success, >0 = cudaError_t of the launch, <0 = GEOA3_E* argument error.
 *     Nothing ever calls exit() (the reference does: _ext-src/include/cuda_utils.h:30-39).
 *   - loss-side clouds are channel-first float32 [b][3][n] (what Lib/loss_utils.py receives);
 *     pointnet2-side coordinates are AoS float32 [b][n][3] and features [b][c][n]
 *     (the pointnet2_ops convention); every index tensor is int32.
 *   - index outputs are bit-exact w.r.t. TWO pinned squared-distance chains:
 *       loss side (geoa3_nn_pair, geoa3_knn):  t = dx*dx; t = fma(dy,dy,t); t = fma(dz,dz,t)
 *         (pytorch3d's `dist += diff*diff` loop under nvcc's default -fmad=true);
 *       pointnet2 side (FPS, ball_query, three_nn):  t = dy*dy; t = fma(dx,dx,t); t = fma(dz,dz,t)
 *         (what nvcc makes of the reference's (dx*dx)+(dy*dy)+(dz*dz): read off the SASS of the
 *         reference sources built for sm_100, and checked against that binary);
 *     ties -> lowest index (FPS: the reference's tournament order); reductions are fixed-order
 *     (run-to-run deterministic, no float atomics).
 *
 * Paths are relative to the reference checkout (Gorilla-Lab-SCUT/GeoA3).
 */
#ifndef GEOA3_B200_H_
#define GEOA3_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GEOA3_OK 0
#define GEOA3_EINVAL (-1)       /* null pointer / non-positive size */
#define GEOA3_EUNSUPPORTED (-2) /* size outside what the kernels are built for (e.g. K > 33) */
#define GEOA3_EWORKSPACE (-3)   /* workspace too small */

#define GEOA3_KNN_MAX_K 33 /* K = k+1 <= 33 (curv_loss_knn <= 32, BASELINE config 4) */

typedef void *geoa3_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define GEOA3_API __attribute__((visibility("default")))
#else
#define GEOA3_API
#endif

/* ABI version (major*1000+minor) and error text for any return code. */
GEOA3_API int geoa3_version(void);
GEOA3_API const char *geoa3_error_string(int code);

/* ----------------------------------------------------------------------------------------------
 * Loss path  (replaces pytorch3d.ops.knn_points/knn_gather as used by Lib/loss_utils.py:28-97)
 * ---------------------------------------------------------------------------------------------- */

/* Fused bidirectional 1-NN between adv [b][3][n] and ori [b][3][m]: the N x M distance matrix is
 * tiled through shared memory and never written.  d_a2o[b][n]/jstar[b][n] = min_j / argmin_j
 * d(adv_i, ori_j); d_o2a[b][m]/istar[b][m] the other direction (pass NULL for both to skip it:
 * pseudo_chamfer_loss, hausdorff_loss, find_offset).
 * hint_a2o [b][n] / hint_o2a [b][m] (nullable) seed each query with a candidate index (default: its own
 * index); results are exact for any seed, a good one (e.g. last step's argmin) makes the search ~2x faster.
 * A hint may alias the corresponding output (in-place update of a persistent buffer).
 * perm_a [b][n] / perm_o [b][m] (both or none): the clouds are passed ALREADY ARRANGED in a visiting order —
 * position t of adv holds original point perm_a[t] (likewise ori/perm_o), e.g. the Morton order of the original
 * cloud computed once per attack.  With a spatially coherent order whole groups of 32 candidates are skipped by
 * a bounding-box test.  Outputs and hints use ORIGINAL numbering (d_a2o[i], jstar[i] belong to original adv point
 * i), ties resolve on original indices: results are identical for ANY permutation.  iperm_* (inverse
 * permutations, nullable) are only needed to honour hints; without them the seed is the same-position point.
 * Replaces: knn_points(adv, ori, K=1) + knn_points(ori, adv, K=1), Lib/loss_utils.py:32-33,41,48,70,92;
 *           Attacker/geoA3_attack.py:65,80. */
GEOA3_API int geoa3_nn_pair(const float *adv, const float *ori, int b, int n, int m, const int32_t *perm_a,
                            const int32_t *perm_o, const int32_t *iperm_a, const int32_t *iperm_o,
                            const int32_t *hint_a2o, const int32_t *hint_o2a, float *d_a2o, int32_t *jstar,
                            float *d_o2a, int32_t *istar, geoa3_stream_t stream);

/* K nearest of every query [b][3][n] among ref [b][3][m], ascending (dist, idx); the first `drop`
 * columns are discarded (drop=1 removes the self match exactly like "[:,:,:,1:]").
 * idx [b][n][K-drop] int32, dist (nullable) [b][n][K-drop].  K <= GEOA3_KNN_MAX_K, K <= m.
 * hint [b][n][hint_k] (nullable): candidate indices per query used only to start the selection threshold
 * (self-verifying, results are exact for any hint); may alias idx when hint_k == K-drop.
 * perm_q [b][n] / perm_c [b][m] (nullable): query / ref are passed already arranged in a visiting order (see
 * geoa3_nn_pair); iperm_c [b][m] = inverse of perm_c (needed to use hints with a permuted ref).  Results are in
 * original numbering and identical for any permutation; a spatially coherent one enables bounding-box pruning.
 * Replaces: knn_points(pc, pc, K=k+1).idx[..., 1:], Lib/loss_utils.py:57-58,77-78,139,174. */
GEOA3_API int geoa3_knn(const float *query, const float *ref, int b, int n, int m, int K, int drop,
                        const int32_t *perm_q, const int32_t *perm_c, const int32_t *iperm_c, const float *bb_c,
                        const int32_t *hint, int hint_k, int32_t *idx, float *dist, geoa3_stream_t stream);

/* The same K - drop members as geoa3_knn (bit-exact membership: the K lexicographically smallest
 * (pinned distance, index) pairs minus the `drop` smallest), written in ascending VISITING order of the candidates
 * (ascending index without perm_c) instead of by distance — for consumers that sum over the neighbourhood (kappa and
 * its gradient, Lib/loss_utils.py:59-62,79-82) and never look at the order.  Not sorting makes it ~2x cheaper.
 * Arguments as geoa3_knn; `dist` (optional) receives the members' distances in the same order.
 * Replaces: knn_points(pc, pc, K=k+1).idx[..., 1:] inside _get_kappa_adv / _get_kappa_ori, Lib/loss_utils.py:57-58,77-78. */
GEOA3_API int geoa3_knn_set(const float *query, const float *ref, int b, int n, int m, int K, int drop,
                            const int32_t *perm_q, const int32_t *perm_c, const int32_t *iperm_c, const float *bb_c,
                            const int32_t *hint, int hint_k, int32_t *idx, float *dist, geoa3_stream_t stream);

/* Cell-grid searches (the attack step's 1-NN and kNN searches on clouds of up to 65535 points).
 *
 * geoa3_cell_sort arranges each cloud [b][3][n] into a uniform grid of CUBIC cells over its bounding box (one CTA per
 * cloud, counting sort: cell (cz*gy + cy)*gx + cx major, ascending original index inside a cell — so the cells of one
 * (cz, cy) row are one contiguous range of positions) and writes one self-contained blob per cloud
 * (geoa3_cell_blob_bytes(n, ncap) bytes, 16-byte aligned: grid parameters, the cloud as float4 (x, y, z, original
 * index) in cell order, the cell start table, the inverse permutation).  ncap is the capacity of the cell table (it
 * fixes the blob layout; <= geoa3_cell_grid_max(n): the sorting pass keeps its tables in one CTA's shared memory).
 * Grid: gx, gy, gz > 0 -> that many cells per axis (gx*gy*gz <= ncap, <= 64 per axis), cell edge = longest side of the
 * box / largest g.  gx = gy = gz = 0 -> chosen PER CLOUD: cell edge ~ radius of a ball holding `kref` points, from the
 * cloud's surface density (occupied cells of a 16^3 probe grid), as many cells per axis as the box needs within ncap —
 * a thin rod and a sphere both get cells that match their own neighbourhood size.  The grid only affects speed.
 *
 * geoa3_knn_cells (SELF queries: the curvature term's neighbour lists) visits, per query, only the cell rows its search
 * ball can reach (radius^2 = the largest pinned distance to the hinted candidates; without a usable hint the search
 * starts from +inf and is still exact) and keeps every candidate whose PINNED distance is inside the bound; the K
 * smallest (distance, original index) keys minus the `drop` smallest are the members.  Same members as geoa3_knn /
 * geoa3_knn_set for ANY hint and ANY grid (bit-exact membership and distances); they are written in ascending
 * position of the cell arrangement — a function of the cloud and its grid only, never of the hint.
 * idx [b][n][K-drop] (rows by ORIGINAL query index), dist nullable; hint [b][n][hint_k] may alias idx.
 * Replaces: knn_points(pc, pc, K=k+1).idx[..., 1:] inside _get_kappa_adv, Lib/loss_utils.py:77-78 (no counterpart
 * for the grid in the reference: its search is pytorch3d's unordered brute force). */
GEOA3_API int geoa3_cell_grid_max(int n);
GEOA3_API size_t geoa3_cell_blob_bytes(int n, int ncap);
GEOA3_API int geoa3_cell_sort(const float *pc, int b, int n, int ncap, float kref, int gx, int gy, int gz, void *blobs,
                              geoa3_stream_t stream);
GEOA3_API int geoa3_knn_cells(const void *blobs, int b, int n, int ncap, int K, int drop, const int32_t *hint, int hint_k,
                              int32_t *idx, float *dist, geoa3_stream_t stream);

/* geoa3_nn_pair on the cell-grid blobs of geoa3_cell_sort (blobs_adv: n points, table capacity ncap_adv; blobs_ori: m,
 * ncap_ori): the same outputs, bit for bit — d_a2o[i], jstar[i] = (min, lowest argmin) over ori of the pinned distance
 * to adv point i, d_o2a / istar the other direction (both NULL to skip it) — for ANY seed and ANY grids.  A query walks
 * only the cell rows within sqrt(best) of itself, best starting at the pinned distance to its seed (hint_*, default:
 * the same index; may alias the index output).  In an attack step that is ~10 candidates per query.  The candidate
 * blob must fit one CTA's shared memory (clouds up to ~10 000 points); otherwise GEOA3_EUNSUPPORTED (use
 * geoa3_nn_pair).  The original cloud's blob is built once per attack, the adversarial one every step (it also serves
 * geoa3_knn_cells).
 * Replaces: knn_points(adv, ori, K=1) + knn_points(ori, adv, K=1), Lib/loss_utils.py:32-33,41,48,70,92. */
GEOA3_API int geoa3_nn_pair_cells(const void *blobs_adv, const void *blobs_ori, int b, int n, int m, int ncap_adv,
                                  int ncap_ori, const int32_t *hint_a2o, const int32_t *hint_o2a, float *d_a2o,
                                  int32_t *jstar, float *d_o2a, int32_t *istar, geoa3_stream_t stream);

/* Bounding boxes of a cloud that is already arranged in visiting order: per cloud geoa3_group_bbox_floats(n)
 * floats = [G0 + G1][8] (lo xyz, hi xyz, max |p|^2, pad), G0 = ceil(n/32) boxes of 32 consecutive positions
 * followed by G1 = ceil(n/1024) boxes of 1024.  Input to geoa3_knn's `bb_c` (required with perm_c): whole
 * 1024-candidate chunks are skipped before they are staged, then 32-candidate groups inside a staged chunk. */
GEOA3_API size_t geoa3_group_bbox_floats(int n);
GEOA3_API int geoa3_group_bbox(const float *pc_arranged, int b, int n, float *bb, geoa3_stream_t stream);

/* Arranges a cloud in visiting order and (bb != NULL) boxes it in the same launch:
 * arranged[b][3][n] with position t = original point perm[t]; bb as for geoa3_group_bbox.  Replaces the
 * host-side index_select the reference would use for a permuted cloud (no counterpart in the reference: its
 * searches are unordered brute force, Lib/loss_utils.py:32-33,57,77). */
GEOA3_API int geoa3_arrange(const float *pc, const int32_t *perm, int b, int n, float *arranged, float *bb,
                            geoa3_stream_t stream);

/* Local curvature + per-cloud loss reductions, one CTA per cloud, fixed-order reductions.
 *   kappa_i = (1/k) sum_m |<nrm_i, v_im/max(|v_im|,1e-12)>|,  v_im = pc[nbr[i][m]] - pc[i]
 *   nrm_i   = normal[:, jstar[i]]  (jstar != NULL: _get_kappa_adv)  or normal[:, i] (jstar NULL: _get_kappa_ori)
 * pc [b][3][n]; normal [b][3][m], kappa_ori [b][m], d_o2a [b][m] live on the m original points
 * (jstar NULL requires m == n).  Optional outputs (NULL to skip): kappa [b][n], nrm_out [b][3][n],
 *   cd[b] = mean d_a2o + mean d_o2a (d_o2a NULL -> one-sided), hd[b] = max d_a2o, hd_arg[b] (lowest i),
 *   curv[b] = mean (kappa_i - kappa_ori[jstar[i]])^2.   nbr NULL / k 0 skips the curvature part.
 * Replaces: Lib/loss_utils.py:34,42,49,59-62,71,79-82,93-95; Lib/utility.py:30-31. */
GEOA3_API int geoa3_kappa_loss_fwd(const float *pc, const float *normal, const int32_t *jstar, const int32_t *nbr, int k,
                         const float *d_a2o, const float *d_o2a, const float *kappa_ori, int b, int n, int m,
                         float *kappa, float *nrm_out, float *cd, float *hd, int32_t *hd_arg, float *curv,
                         geoa3_stream_t stream);

/* Backward of  sum_b (g_cd*CD + g_hd*HD + g_cu*CUR) + sum_{b,i} g_kappa*kappa  w.r.t. adv, in one
 * kernel: contributions are gathered per target point through CSR lists built in shared memory from
 * istar / nbr (ascending source order) — deterministic, no float atomics.
 * Any of g_cd/g_hd/g_cu [b] and g_kappa [b][n] may be NULL (term skipped); istar NULL = one-sided CD.
 * grad_adv [b][3][n] is fully overwritten.
 * Clouds whose lists do not fit shared memory (n*k > 65535 edges, e.g. n >= 4096 at k = 16) take the same
 * gather with the lists in a caller-provided device workspace of geoa3_loss_bwd_workspace_bytes(b,n,m,k)
 * bytes (16-byte aligned; 0 = not needed, workspace may be NULL): identical summation order, identical bits.
 * Replaces: autograd through pytorch3d _knn_points.backward + knn_gather scatter_add (float atomics),
 *           reached from Attacker/geoA3_attack.py:326. */
GEOA3_API size_t geoa3_loss_bwd_workspace_bytes(int b, int n, int m, int k);
GEOA3_API int geoa3_loss_bwd(const float *adv, const float *ori, const float *nrm_adv, const float *kappa_adv,
                   const float *kappa_ori, const int32_t *jstar, const int32_t *istar, const int32_t *nbr,
                   const int32_t *hd_arg, const float *g_cd, const float *g_hd, const float *g_cu,
                   const float *g_kappa, int b, int n, int m, int k, float *grad_adv, void *workspace,
                   size_t workspace_bytes, geoa3_stream_t stream);

/* geoa3_kappa_loss_fwd + geoa3_loss_bwd in ONE launch, for the attack step (forward and backward always run
 * together there): kappa with normals borrowed through jstar, cd / hd / curv per cloud (same definitions as
 * geoa3_kappa_loss_fwd; outputs nullable, kappa [b][n], nrm_out [b][3][n], hd_arg [b] optional) AND
 * grad_adv [b][3][n] = d(w_cd*CD + w_hd*HD + w_cu*CUR)/d adv for a UNIT upstream gradient per cloud — the loss is linear in
 * its upstream gradient, so the caller's backward is a scale by g[b].  normal [b][3][m] / kappa_ori [b][m] live on the
 * original points; istar / d_o2a both NULL = one-sided CD; k = 0 or w_cu = 0 skips the curvature term.
 * Same arithmetic and summation orders as the two separate kernels (deterministic, no float atomics); the cloud, the
 * normals and the neighbour rows are staged once.  Needs the single-kernel shared-memory plan:
 * geoa3_geo_fwd_bwd_supported(n, m, k) != 0 (n*k <= 65535 edges, n <= ~3 900 at k = 16), else GEOA3_EUNSUPPORTED.
 * Replaces: Lib/loss_utils.py:28-97 + autograd backward, as assembled by Attacker/geoA3_attack.py:131-162,326. */
GEOA3_API int geoa3_geo_fwd_bwd_supported(int n, int m, int k);
GEOA3_API int geoa3_geo_fwd_bwd(const float *adv, const float *ori, const float *normal, const float *kappa_ori,
                                const int32_t *jstar, const int32_t *istar, const int32_t *nbr, int k, const float *d_a2o,
                                const float *d_o2a, float w_cd, float w_hd, float w_cu, int b, int n, int m, float *cd,
                                float *hd, float *curv, float *kappa, float *nrm_out, int32_t *hd_arg, float *grad_adv,
                                geoa3_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * pointnet2_ops  (replaces the 9 functions exported by _ext-src/src/bindings.cpp:6-19)
 * ---------------------------------------------------------------------------------------------- */

/* xyz [b][n][3] -> idx [b][m]; same selection (origin skip, 1e10 init, tournament tie order) as
 * furthest_point_sampling (sampling.cpp:66-87, sampling_gpu.cu:69-229).  No temp buffer needed. */
GEOA3_API int geoa3_furthest_point_sampling(const float *xyz, int b, int n, int m, int32_t *idx, geoa3_stream_t stream);

/* Plain farthest point sampling from a given first pick per cloud: start[b] (int32, clamped to [0,n)), no frozen
 * points, running minimum from +inf, arg-max ties -> lowest index; idx[b][m] int32 with idx[.][0] = start.
 * Replaces: farthest_points_sample, Lib/utility.py:175-187 (a Python loop of m-1 torch.min / argmax passes), the
 * per-iteration subsampling of --is_subsample_opt (Attacker/geoA3_attack.py:283-292).  Distances are Euclidean
 * norms like the reference's (`torch.norm`, :183): sqrt.rn of (dx*dx + dy*dy) + dz*dz with separately rounded
 * products, the arithmetic of torch's CUDA norm reduction, so rounding ties fall the same way. */
GEOA3_API int geoa3_farthest_points_sample(const float *xyz, int b, int n, int m, const int32_t *start, int32_t *idx,
                                           geoa3_stream_t stream);

/* points [b][c][n], idx [b][m] -> out [b][c][m]        (gather_points, sampling.cpp:15-39) */
GEOA3_API int geoa3_gather_points(const float *points, const int32_t *idx, int b, int c, int n, int m, float *out,
                        geoa3_stream_t stream);
/* grad_out [b][c][m], idx [b][m] -> grad_points [b][c][n] (overwritten; duplicates summed in
 * ascending j; workspace of geoa3_group_points_grad_workspace_bytes(b, n, m, 1) bytes)
 *                                                       (gather_points_grad, sampling.cpp:41-65) */
GEOA3_API int geoa3_gather_points_grad(const float *grad_out, const int32_t *idx, int b, int c, int n, int m,
                             float *grad_points, void *workspace, size_t workspace_bytes,
                             geoa3_stream_t stream);

/* new_xyz [b][m][3], xyz [b][n][3] -> idx [b][m][nsample] (ball_query, ball_query.cpp:8-32,
 * ball_query_gpu.cu:9-54: first nsample hits in index order, d2 < r*r strict, first-hit fill,
 * no hit -> zeros). */
GEOA3_API int geoa3_ball_query(const float *new_xyz, const float *xyz, int b, int n, int m, float radius, int nsample,
                     int32_t *idx, geoa3_stream_t stream);

/* points [b][c][n], idx [b][npoints][nsample] -> out [b][c][npoints][nsample]
 *                                                       (group_points, group_points.cpp:12-36) */
GEOA3_API int geoa3_group_points(const float *points, const int32_t *idx, int b, int c, int n, int npoints, int nsample,
                       float *out, geoa3_stream_t stream);
/* Deterministic scatter: a CSR-by-target of idx is built once into `workspace` (size from
 * geoa3_group_points_grad_workspace_bytes) and shared by all c channels; sums in ascending (j,k).
 * grad_points [b][c][n] is fully overwritten.           (group_points_grad, group_points.cpp:38-62) */
GEOA3_API size_t geoa3_group_points_grad_workspace_bytes(int b, int n, int npoints, int nsample);
GEOA3_API int geoa3_group_points_grad(const float *grad_out, const int32_t *idx, int b, int c, int n, int npoints,
                            int nsample, float *grad_points, void *workspace, size_t workspace_bytes,
                            geoa3_stream_t stream);

/* unknown [b][n][3], known [b][m][3] -> dist2 [b][n][3], idx [b][n][3]   (three_nn, interpolate.cpp) */
GEOA3_API int geoa3_three_nn(const float *unknown, const float *known, int b, int n, int m, float *dist2, int32_t *idx,
                   geoa3_stream_t stream);
/* points [b][c][m], idx/weight [b][n][3] -> out [b][c][n]               (three_interpolate) */
GEOA3_API int geoa3_three_interpolate(const float *points, const int32_t *idx, const float *weight, int b, int c, int m,
                            int n, float *out, geoa3_stream_t stream);
/* grad_out [b][c][n] -> grad_points [b][c][m] (overwritten, deterministic; workspace of
 * geoa3_group_points_grad_workspace_bytes(b, m, n, 3) bytes)             (three_interpolate_grad) */
GEOA3_API int geoa3_three_interpolate_grad(const float *grad_out, const int32_t *idx, const float *weight, int b, int c,
                                 int n, int m, float *grad_points, void *workspace, size_t workspace_bytes,
                                 geoa3_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GEOA3_B200_H_ */
