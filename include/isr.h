/* libisr -- C ABI of the B200-native registration-and-verification hot path.
 *
 * Drop-in boundary for the arithmetic that the reference
 * (Kudo510/ImageSequenceRegistrationfor6DPoseEstimationLabeling) delegates to Open3D
 * and scikit-learn.  The reference has no native code and no FFI of its own; the
 * calls below are what a binding for this path would replace, cited per entry point
 * as <reference file>:<line>.
 *
 * Conventions
 *  - Every function returns an int status: ISR_OK (0) or a negative ISR_E_* code.
 *    No C++ exception crosses the boundary.  isr_last_error() returns a thread-local
 *    message for the most recent failure on the calling thread.
 *  - All data pointers are DEVICE pointers owned by the caller (PyTorch's allocator in
 *    this repo) unless the parameter name ends in _host.  Scratch is passed in as
 *    `workspace` (size from the matching *_workspace_bytes function, 256-byte aligned).
 *    The library itself allocates only two small persistent objects: a 64-byte profiling
 *    counter per device (first profiled search) and the exchange buffer of isr_peer_create.
 *  - Every call works on the CUDA device that is current in the calling thread; all
 *    pointers of a call must belong to it.
 *  - `stream` is a cudaStream_t passed as void*.  Calls only enqueue work; none
 *    synchronises unless documented.
 *  - Points: row-major float32 [n][3] ("AoS", the reference's numpy N x 3 layout,
 *    genFeat.py:223-228).  Internally clouds are repacked to padded planes
 *    [3][npad] ("SoA", npad = isr_soa_padded_len(n)); padded slots hold ISR_PAD_COORD.
 *  - Poses: row-major float64 [16] 4x4, column-vector convention p' = T[:3,:3] p + T[:3,3]
 *    (Open3D PointCloud.transform, icp.py:22,110).  The reference's right-multiplication
 *    `pc.dot(M)` (verfication.py:83-85) is the pose with rotation M^T.
 *  - Indices int32 (n < 2^31 - 1024), counts int64.
 *  - Arithmetic of the nearest-neighbour search (isr_nn2, nn2.cu): clouds are centred and
 *    kept as float32 hi/lo pairs of their FP64 coordinates (isr_prepare_cloud); an FP32
 *    3-FMA filter a_j = |p_j|^2 - 2 q.p_j over the tiles that can hold the neighbour, then an
 *    exact FP64 distance for the few targets inside the filter's proven error window; the
 *    index is the float64 argmin, lowest ORIGINAL target index on exact ties.  (isr_nn_soa is
 *    the round-1 direct-difference FP32 kernel, kept for bit-exact A/B.)  Transforms,
 *    distances fed to ICP, all reductions, Kabsch: FP64, fixed summation order.
 */
#ifndef ISR_H_
#define ISR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISR_VERSION 100 /* 0.1.0 */

#define ISR_OK 0
#define ISR_E_INVALID_ARG (-1)
#define ISR_E_SHAPE (-2)
#define ISR_E_ALIGN (-3)
#define ISR_E_CUDA (-4)
#define ISR_E_WORKSPACE (-5)

#define ISR_SOA_TILE 1024      /* SoA planes are padded to a multiple of this        */
#define ISR_SUB_TILE 64        /* pruning granularity inside a 1024-point stage      */
#define ISR_PAD_COORD 1.0e18f  /* coordinate stored in padded slots (d2 ~ 3e36, finite) */

/* ---- library / device --------------------------------------------------------------- */
int isr_version(void);
const char *isr_last_error(void);
/* SM count, max SM clock (kHz), shared memory per SM (bytes) of the current device. */
int isr_device_info(int *sm_count, int *sm_clock_khz, int *smem_per_sm);
/* Number of CUDA kernels this library has launched since the last reset (process-wide). */
uint64_t isr_launch_count(void);
void isr_reset_launch_count(void);

/* Per-kernel device timing for bench.py's roofline: while enabled, every launch of a
 * profiled kernel kind is bracketed by CUDA events on its stream.  collect() waits for
 * the recorded events, writes total milliseconds and launch counts per kind
 * (0 transform, 1 nearest-neighbour, 2 mean-sqrt reduce, 3 ICP accumulate, 4 ICP solve;
 * arrays of ISR_PROFILE_KINDS) and clears the records. */
#define ISR_PROFILE_KINDS 5
int isr_profile_enable(int on);
int isr_profile_collect(double *ms_by_kind_host, uint64_t *launches_by_kind_host);

/* ---- K1: batched rigid transform ---------------------------------------------------- */
int64_t isr_soa_padded_len(int64_t n);

/* out[b][i][:] = R_b pts[i] + t_b, float32 [b][n][3].  Replaces `pc.dot(R.T) + t`
 * (verfication.py:83-85, icp.py:68, choosePose.py:21-22) and PointCloud.transform
 * (icp.py:22,110).  FP64 math, one rounding to float32. */
int isr_transform_points(const float *pts, int64_t n, const double *poses, int64_t b,
                         float *out, void *stream);

/* out[i][:] = R pts[i] + t in float64 for ONE pose (out may alias pts): Open3D's
 * PointCloud.transform keeps double coordinates (icp.py:22,110). */
int isr_transform_points_f64(const double *pts, int64_t n, const double *pose, double *out,
                             void *stream);

/* Same transform written as padded planes out[b][3][npad] (the layout K2 consumes).
 * poses == NULL copies the cloud unchanged (b must be 1).  `pose_stride` is in doubles
 * (16 for a dense [b][16] array; larger when poses live inside IsrIcpState records).
 * `skip`, if not NULL, is an int32 per batch at stride `skip_stride` ints: non-zero
 * -> that batch is left untouched (finished ICP starts). */
int isr_transform_points_soa(const float *pts, int64_t n, const double *poses,
                             int64_t pose_stride, int64_t b, float *out_soa, int64_t npad,
                             const int32_t *skip, int64_t skip_stride, void *stream);

/* ---- K2 (direct-difference form, kept for A/B and bit-exact FP32 checks) ---------------- */
/* For every query of every batch: squared distance to, and index of, its nearest target.
 * q_soa [batch][3][nq_pad] (q_bstride floats between batches, 0 = shared),
 * t_soa likewise.  out_d2 float32 [batch][nq]; out_idx int32 [batch][nq] or NULL.
 * Replaces PointCloud.compute_point_cloud_distance (verfication.py:97,99; icp.py:113,115),
 * the KD-tree 1-NN inside registration_icp / evaluate_registration (icp.py:97-103) and
 * sklearn KDTree.query(k=1) (choosePose.py:21-22). */
size_t isr_nn_workspace_bytes(int64_t nq, int64_t nt, int64_t batch);
int isr_nn_soa(const float *q_soa, int64_t nq, int64_t nq_pad, int64_t q_bstride,
               const float *t_soa, int64_t nt, int64_t nt_pad, int64_t t_bstride,
               int64_t batch, float *out_d2, int32_t *out_idx, const int32_t *skip,
               int64_t skip_stride, void *workspace, size_t workspace_bytes, void *stream);

/* ---- K1' + K2 (production): ordered, centred hi/lo planes and the filtered exact search -- */
/* out3[0..2] = FP64 centroid of pts (single CTA, fixed summation order). */
int isr_centroid(const float *pts, int64_t n, double *out3, void *stream);

/* perm[i] = original index of the i-th point along a 30-bit Hilbert curve through the cloud's
 * bounding box (the curve never jumps: any run of consecutive points is one connected patch).  NeRF surface clouds
 * arrive in farthest-point-sampling order (genFeat.py:199-202), i.e. spatially random; the
 * scan kernel is fastest when consecutive stored points are neighbours.  Deterministic. */
size_t isr_spatial_order_workspace_bytes(int64_t n);
int isr_spatial_order(const float *pts, int64_t n, int32_t *perm, void *workspace,
                      size_t workspace_bytes, void *stream);

/* K1 with centring and FP64-accurate output, "SoA7" planes out[b][7][npad]:
 *   0-2 hi = float32(R_b p + t_b - c_b), 3 = fl32 |hi|^2, 4-6 lo = float32(exact - hi).
 * c_b = C_b . centroid (centre_poses NULL: c_b = centroid; centroid NULL: c_b = 0).
 * Both clouds of a pair must be prepared with the same c_b.  poses NULL = identity (b == 1).
 * pts_lo (float32 [n][3], may be NULL): low part of a float64 input cloud, p = pts + pts_lo
 * (icp.py:68 hands Open3D a float64 camera-frame source).  perm (may be NULL): stored
 * position i holds original point perm[i].  Strides are in doubles.
 * Replaces the same reference lines as isr_transform_points_soa. */
int isr_prepare_cloud(const float *pts, const float *pts_lo, const int32_t *perm, int64_t n,
                      const double *poses, int64_t pose_stride, const double *centre_poses,
                      int64_t centre_pose_stride, const double *centroid, int64_t b,
                      float *out_soa7, int64_t npad, const int32_t *skip, int64_t skip_stride,
                      void *stream);

/* Bounding spheres of the stored tiles of a SoA7 cloud, (cx, cy, cz, r) float32 each:
 * out_stage [batch][isr_stage_sphere_count(npad)]: first the npad/1024 spheres of the
 * 1024-point stages, then one sphere per chunk of 32 consecutive stages (it bounds the 32
 * stage spheres; the search tests a chunk before it looks at its stages -- 977 stage spheres
 * per query block at 1 M target points otherwise); out_sub [batch][npad/64] for the
 * 64-point sub-tiles.  r bounds |p - c| for every real point of the tile, inflated to hold
 * for the FP64 (hi + lo) coordinates; r = -1 marks a tile of padding only.  The search uses
 * them to scan the nearest stage first and to skip tiles that provably cannot hold a
 * query's nearest neighbour (what the KD-tree of Open3D / sklearn does by construction).
 * out_box (uint32 [batch][npad/64], may be NULL) receives each sub-tile's axis-aligned bounding
 * box about the same centre: three 10-bit fields k_x | k_y << 10 | k_z << 20, half-extent
 * h = r * k / 1023 (rounded up, inflated like r); 0x3FFFFFFF for a tile of padding only.  The
 * patches of a surface cloud are thin sheets: the box follows them where the sphere is mostly
 * empty, and a tile is skipped when either volume is out of reach. */
int64_t isr_stage_sphere_count(int64_t npad);
int isr_tile_spheres(const float *soa7, int64_t n, int64_t npad, int64_t bstride, int64_t batch,
                     float *out_stage, float *out_sub, uint32_t *out_box, void *stream);

/* A prepared cloud (or batch of clouds) as the search kernel sees it. */
typedef struct IsrCloud {
    const float *soa7;    /* [batch][7][npad] from isr_prepare_cloud                          */
    int64_t n;            /* real points                                                      */
    int64_t npad;         /* padded plane length (multiple of ISR_SOA_TILE)                   */
    int64_t bstride;      /* floats between batch items; 0 = one cloud shared by the batch    */
    const float *stage_c; /* isr_tile_spheres out_stage (stage + chunk spheres, batch stride
                             isr_stage_sphere_count(npad)), or NULL (scan in storage order)   */
    const int32_t *perm;  /* the perm it was prepared with, or NULL; results are reported in
                             original indices either way                                      */
    const float *sub_c;   /* isr_tile_spheres out_sub, or NULL (no tile pruning: every pair is
                             evaluated); only read when this cloud is the target             */
    int32_t *hint;        /* only read when this cloud is the QUERY of a pruned search; NULL or
                             int32 [batch][npad], by stored query position: stored position in
                             the target of a point near the query (-1 = none).  It seeds the
                             query's bound -- any value gives the same result, a good one less
                             work -- and is overwritten with the neighbour found, so that the
                             next search of the same clouds (the next ICP iteration) starts
                             from it                                                          */
    const uint32_t *sub_box; /* isr_tile_spheres out_box, or NULL (sphere tests only); only read
                             when this cloud is the target                                    */
} IsrCloud;

/* Process-wide switch for the tile pruning of isr_nn2 (default on).  Off = exhaustive brute
 * force: every (query, target) pair goes through the FP32 scan.  Results are identical either
 * way; bench.py uses it to time the exhaustive kernel against its FP32 roofline. */
int isr_set_nn_pruning(int on);
int isr_get_nn_pruning(void);

/* While isr_profile_enable(1): pairs the isr_nn2 launches actually evaluated (scanned
 * sub-tiles x queries, padded lanes included) and the pairs they answered for (nq x nt x
 * batch).  Synchronises the device, then clears both counters. */
int isr_profile_nn_pairs(uint64_t *evaluated_host, uint64_t *answered_host);
/* Raw per-warp event counters behind the figure above, uint64 [8], not cleared: scanned
 * (warp, sub-tile) units, stages walked, stages skipped by their sphere, sub-tile sphere
 * tests, warps, warps whose stage list overflowed, 2 reserved.  Synchronises the device. */
int isr_profile_nn_counters(uint64_t *out8_host);

/* Exact 1-NN of every query in its target by tiled brute force: FP32 3-FMA filter over the
 * pairs of every tile that can hold the neighbour (all tiles when pruning is off or the
 * target has no sub_c), FP64 resolve of the few targets inside the proven error window
 * (nn2.cu).  out_idx
 * [batch][nq] equals the float64 brute-force argmin of the prepared coordinates (lowest
 * original index on exact ties); out_d2 is that FP64 squared distance rounded to float32.
 * use_lo == 0 ignores the lo planes (distances between the float32 hi coordinates).
 * Same reference call sites as isr_nn_soa. */
size_t isr_nn2_workspace_bytes(int64_t nq, int64_t nt, int64_t batch);
int isr_nn2(const IsrCloud *q, const IsrCloud *t, int64_t batch, int use_lo, float *out_d2,
            int32_t *out_idx, const int32_t *skip, int64_t skip_stride, void *workspace,
            size_t workspace_bytes, void *stream);

/* out_mean[b] = mean_i sqrt(d2[b][i]) in FP64, fixed summation order (deterministic).
 * np.mean(np.asarray(compute_point_cloud_distance(..))) -- verfication.py:98,100. */
int isr_mean_sqrt(const float *d2, int64_t n, int64_t batch, double *out_mean, void *stream);

/* ---- batched candidate verification ------------------------------------------------- */
/* Candidate k compares X_k = Pq[k] . cloud_q with Y_k = Pt[k] . cloud_t:
 *   bidirectional != 0: loss = (mean d(X->Y) + mean d(Y->X)) / 2  -- the Chamfer loop of
 *                       verfication.py:61-102 (icp.py:113-117);
 *   bidirectional == 0: loss = mean d(X->Y)                      -- ADDS, choosePose.py:20-22.
 * valid (uint8 [b], may be NULL): 0 marks a failed PnP candidate (choosePose.py:29-33);
 * its loss is +inf.  out_loss float64 [b].  out_best: int64 [2] = {argmin index with the
 * first minimum winning (verfication.py:105-106), bit pattern of its float64 loss}. */
size_t isr_verify_workspace_bytes(int64_t nq, int64_t nt, int64_t b, int bidirectional);
int isr_verify_poses(const float *cloud_q, int64_t nq, const float *cloud_t, int64_t nt,
                     const double *poses_q, const double *poses_t, const uint8_t *valid,
                     int64_t b, int bidirectional, double *out_loss, int64_t *out_best,
                     void *workspace, size_t workspace_bytes, void *stream);

/* ---- the callers of the batched ADD-S in choosePose.py, on the device --- */
/* out[k - pair0] (float64 [count][16]) = relative_poses[i][j] for the flat pair index
 * k = i * n + j in [pair0, pair0 + count): the 4x4 of compute_rel_poses(R_i, t_i, R_j, t_j) =
 * (R_i^T R_j, t_j - t_i) -- NOT an SE(3) composition, kept as the reference has it
 * (choosePose.py:43-51, 98-107).  R float64 [n][9] row-major, t float64 [n][3]. */
int isr_rel_pose_table(const double *R, const double *t, int64_t n, int64_t pair0, int64_t count,
                       double *out, void *stream);
/* out[k] = poses_t[k]^-1 . poses_q[k] for RIGID poses_t (rotation inverse = transpose). */
int isr_rigid_relative(const double *poses_q, const double *poses_t, int64_t b, double *out,
                       void *stream);
/* ADD-S of candidate k = mean 1-NN distance from poses_q[k] . cloud_q into cloud_t ITSELF
 * (the target is sorted / prepared / given its tile spheres once for the whole batch).  With
 * poses_q = isr_rigid_relative(Pq, Pt) this is ADDS(verts, gtR, gtT, R, T) of choosePose.py:20-22
 * for rotation matrices R (the error is |R^T R - I| x the cloud diameter otherwise).
 * valid / out_loss / out_best as in isr_verify_poses. */
size_t isr_adds_fixed_target_workspace_bytes(int64_t nq, int64_t nt, int64_t b);
int isr_adds_fixed_target(const float *cloud_q, int64_t nq, const float *cloud_t, int64_t nt,
                          const double *poses_q, const uint8_t *valid, int64_t b, double *out_loss,
                          int64_t *out_best, void *workspace, size_t workspace_bytes, void *stream);
/* Rigorous bounds on the ADD-S of candidate k (poses_q[k] . cloud_q against a prepared, FIXED
 * target) from the target's tile spheres alone: out_lower[k] <= ADDS_k <= out_upper[k].
 * centroid / stage_c / sub_c: the centre the target was prepared with and its isr_tile_spheres
 * outputs (single cloud; `stages` = npad / 1024 stage spheres are read, then 16 sub-tile spheres per
 * stage).  The vote of choosePose.py:135 only asks whether ADDS < 0.1 x diameter: pairs with
 * upper < threshold or lower >= threshold are decided without touching a point, the rest go
 * through isr_adds_fixed_target.  One CTA per pose pair, spheres in shared memory. */
int isr_adds_bounds(const float *cloud_q, int64_t nq, const double *poses_q, int64_t b,
                    const double *centroid, const float *stage_c, int64_t stages, const float *sub_c,
                    double *out_lower, double *out_upper, void *stream);
/* The vote of choosePose.py:135-151 over a loss table float64 [rows][cols]:
 * out_error[i][j] (uint8, may be NULL) = loss[i][j] < threshold (NaN / inf: 0),
 * out_votes[i] (int32) = sum_j error[i][j], out_best (int64 [2], may be NULL) = {index of the
 * FIRST maximum of the votes (np.argmax), its vote count}. */
int isr_vote(const double *loss, int64_t rows, int64_t cols, double threshold, uint8_t *out_error,
             int32_t *out_votes, int64_t *out_best, void *stream);

/* ---- K3: point-to-point ICP --------------------------------------------------------- */
/* Device-resident state of one ICP start (one per symmetry-seeded start when batched). */
typedef struct IsrIcpState {
    double T[16];          /* current transformation, row-major 4x4                   */
    double fitness;        /* #correspondences / n_source of the last evaluation       */
    double inlier_rmse;    /* sqrt(sum d2 / #correspondences) of the last evaluation   */
    double prev_fitness;
    double prev_rmse;
    int64_t n_corr;        /* correspondences of the last evaluation                   */
    int32_t iters;         /* updates applied so far                                   */
    int32_t evals;         /* evaluations done so far                                  */
    int32_t done;          /* 1 once the loop has finished (criteria or max_iteration) */
    int32_t reserved;
} IsrIcpState;

#define ISR_ICP_NSUMS 17 /* sum s[3], sum t[3], sum t s^T [9], sum d2, count */

/* One evaluation pass for `starts` states: transform src by state.T, 1-NN into the
 * target, and reduce the 17 FP64 sums over correspondences with d2 < max_dist^2 (strict)
 * into sums[starts][17].  corr_idx int32 [starts][ns] receives the NN index of every
 * source point, inlier uint8 [starts][ns] its correspondence flag.  States whose `done`
 * is set are skipped.  Prepared once per problem: `centroid` (device double[3], normally
 * isr_centroid(tgt)), tgt_cloud = the target through isr_spatial_order / isr_prepare_cloud
 * (identity pose, that centroid) / isr_tile_spheres, src_perm = isr_spatial_order(src)
 * (may be NULL).  src_lo (may be NULL) is the float32 low part of a float64 source.  This is GetRegistrationResultAndCorrespondences of Open3D's
 * RegistrationICP (icp.py:97-103, upstream). */
size_t isr_icp_workspace_bytes(int64_t ns, int64_t nt, int64_t starts);
int isr_icp_accumulate(IsrIcpState *states, int64_t starts, const float *src,
                       const float *src_lo, const int32_t *src_perm, int64_t ns, const float *tgt,
                       const IsrCloud *tgt_cloud, const double *centroid, double max_dist,
                       double *sums, int32_t *corr_idx, uint8_t *inlier, void *workspace,
                       size_t workspace_bytes, void *stream);

/* The two halves of isr_icp_accumulate, for callers that exchange correspondences between
 * them (target-sharded ICP: every rank searches its own slice of the target, the ranks
 * agree on the nearest one with two MIN all-reduces, and each rank accumulates the
 * correspondences that landed in its slice):
 *   isr_icp_search          transform by state.T + 1-NN -> corr_idx [starts][ns] (target index)
 *   isr_icp_corr_dist       out_D [starts][ns] float64 = |T src[i] - tgt[corr_idx]|^2, +inf where
 *                           corr_idx < 0 (the accumulate kernel's arithmetic: what the ranks compare)
 *   isr_icp_accumulate_corr the 17 sums over corr_idx; an index < 0 means "no correspondence
 *                           on this rank".  nt = rows of tgt.  The sums are formed in a fixed
 *                           order over the STORED source order (src_perm, as in isr_icp_search;
 *                           NULL = file order): 32-point rows, 256-point blocks, groups of 64
 *                           blocks -- the order of the fused iteration of isr_icp_run, so both
 *                           give the same bits. */
int isr_icp_search(IsrIcpState *states, int64_t starts, const float *src, const float *src_lo,
                   const int32_t *src_perm, int64_t ns, const IsrCloud *tgt_cloud,
                   const double *centroid, int32_t *corr_idx, void *workspace,
                   size_t workspace_bytes, void *stream);
int isr_icp_corr_dist(const IsrIcpState *states, int64_t starts, const float *src,
                      const float *src_lo, int64_t ns, const float *tgt, const int32_t *corr_idx,
                      double *out_D, void *stream);
int isr_icp_accumulate_corr(IsrIcpState *states, int64_t starts, const float *src,
                            const float *src_lo, const int32_t *src_perm, int64_t ns,
                            const float *tgt, int64_t nt, const int32_t *corr_idx, double max_dist,
                            double *sums, uint8_t *inlier, void *workspace, size_t workspace_bytes,
                            void *stream);

/* Consume sums[starts][17] (already reduced over all source shards): set fitness / rmse
 * (ns_total = global source count), apply Open3D's break test against the previous
 * evaluation, and otherwise solve Kabsch (Eigen::umeyama without scaling, 3x3 Jacobi SVD,
 * det-sign fix) and update T <- U T.  `final_eval` != 0 marks the evaluation after the
 * last allowed update (sets done).  TransformationEstimationPointToPoint, icp.py:101-103. */
int isr_icp_solve(IsrIcpState *states, int64_t starts, const double *sums, int64_t ns_total,
                  double rel_fitness, double rel_rmse, int final_eval, void *stream);

/* Full loop: states[i].T must hold init_i (other fields zero).  Enqueues
 * max_iteration + 1 evaluation passes; converged starts skip the rest on the device.
 * With the pruned search (the default) one pass is ONE kernel launch: the search kernel
 * transforms the source itself, gathers each neighbour's original coordinates, reduces the
 * 17 sums in a fixed order across the grid and solves Kabsch in the last warp to arrive
 * (nn2.cu / icp_device.cuh); otherwise K1' + K2 + accumulate + solve per pass.  A single start
 * also measures the cycles of every CTA and, once or twice early in the run, cuts its launch
 * list again from them (block_rebalance_kernel: heavy query blocks split, light ones merge);
 * the sums are reduced in an order that does not depend on that cut, so results are
 * bit-identical with and without.  Launches after the first are programmatic dependent
 * launches of the one before.  corr_idx and inlier describe the LAST evaluation of every start.
 * o3d.pipelines.registration.registration_icp, icp.py:101-103; with max_iteration == 0
 * it is evaluate_registration, icp.py:97-98. */
int isr_icp_run(IsrIcpState *states, int64_t starts, const float *src, const float *src_lo,
                const int32_t *src_perm, int64_t ns, const float *tgt, const IsrCloud *tgt_cloud,
                const double *centroid, double max_dist, int max_iteration, double rel_fitness,
                double rel_rmse, double *sums, int32_t *corr_idx, uint8_t *inlier,
                void *workspace, size_t workspace_bytes, void *stream);

/* ---- source-sharded ICP over the GPUs of one box (collective C2) ------ */
/* One process per GPU.  Instead of a collective-library call, the exchange of the 17 sums is
 * FUSED into the iteration's one kernel: the warp that finishes this rank's reduction stores
 * the rank's sums straight into every peer's exchange buffer (CUDA IPC mapping of peer HBM,
 * carried by NVLink / NVSwitch), waits on the flags in its own buffer and adds the ranks'
 * vectors in rank order, so that all ranks solve bit-identical 3x3 problems.  No extra
 * launch, no host synchronisation.
 *   isr_peer_create   allocates this rank's buffer; handle_out receives ISR_PEER_HANDLE_BYTES
 *                     bytes that the caller passes to all other ranks by any means
 *                     (torch.distributed all_gather in dist.py)
 *   isr_peer_connect  handles = world x ISR_PEER_HANDLE_BYTES bytes, in rank order (the own
 *                     entry is ignored); maps every peer's buffer
 *   isr_peer_destroy  unmaps and frees; all ranks must have finished their loops (barrier)
 * world == 1 is allowed (the rank exchanges with itself; used by single-GPU tests). */
#define ISR_PEER_MAX_RANKS 8
#define ISR_PEER_MAX_STARTS 64
#define ISR_PEER_HANDLE_BYTES 64
typedef struct IsrPeer IsrPeer;
int isr_peer_create(int rank, int world, IsrPeer **out, unsigned char *handle_out);
int isr_peer_connect(IsrPeer *peer, const unsigned char *handles);
/* How long a rank waits (inside the kernel) for a peer's message before it gives up;
 * default about 10 s of SM clock. */
int isr_peer_set_timeout(IsrPeer *peer, double seconds);
int isr_peer_destroy(IsrPeer *peer);

/* isr_icp_run for one SOURCE shard: src / src_lo / src_perm / corr_idx / inlier describe this
 * rank's ns rows, ns_total is the global source count, the target is replicated.  Every
 * rank of `peer` must make the same sequence of calls with the same starts, criteria and
 * max_iteration.  starts <= ISR_PEER_MAX_STARTS.  A peer that never arrives makes the waiting
 * warp give up after the time-out: the state gets done = 1, reserved = 1 and NaN fitness.
 * registration_icp, icp.py:101-103. */
int isr_icp_run_sharded(IsrIcpState *states, int64_t starts, const float *src, const float *src_lo,
                        const int32_t *src_perm, int64_t ns, int64_t ns_total, const float *tgt,
                        const IsrCloud *tgt_cloud, const double *centroid, double max_dist,
                        int max_iteration, double rel_fitness, double rel_rmse, double *sums,
                        int32_t *corr_idx, uint8_t *inlier, void *workspace, size_t workspace_bytes,
                        IsrPeer *peer, void *stream);

/* ---- radius neighbour count ----------------------------------- */
/* out_count[i] (int32 [nq], original query indexing) = number of target points with
 * d^2 < radius^2 (strict, float64 decision, a point coinciding with the query included):
 * the count behind Open3D's PointCloud.remove_radius_outlier(nb_points, radius), which keeps
 * point i iff count > nb_points (generateCors.py:254-258, trainPose.py:343-347).  q and t are
 * prepared clouds (isr_prepare_cloud with a common centre; single clouds, bstride 0); t needs
 * its tile spheres.  Pass the same cloud twice for the self-count of the reference. */
int isr_radius_count(const IsrCloud *q, const IsrCloud *t, double radius, int32_t *out_count,
                     void *stream);

/* out_normals float32 [n][3]: pytorch3d.ops.estimate_pointcloud_normals(points,
 * neighborhood_size=k, disambiguate_directions=disambiguate) as generateCors.py:200-215 calls it
 * (k = 400 on the 1000 farthest-point samples of the NeRF cloud; the reference negates the
 * result): the k nearest points of the same cloud (the point itself included; exact distance
 * ties go to the lower index), covariance about their mean, eigenvector of the smallest
 * eigenvalue, flipped when fewer than k / 2 neighbours lie on its positive side.  One CTA per
 * point, all n distances in shared memory: n <= 49152. */
int isr_knn_normals(const float *pts, int64_t n, int64_t k, int disambiguate, float *out_normals,
                    void *stream);

/* ---- PnP hypothesis scoring ------------------------------------ */
/* out_count[j] (int32 [b]) = number of the n 2-D/3-D correspondences that hypothesis j
 * (poses float64 [b][16], object -> camera) reprojects within `reperr` pixels: the consensus
 * test of cv2.solvePnPRansac as choosePose.py:23-33,280-300 calls it (camera matrix `cam`,
 * device double[9] row-major, no distortion; error = float32(du^2 + dv^2) <= reperr^2; a point
 * with z == 0 is projected with z = 1, as OpenCV's projectPoints does).  p3d float32 [n][3],
 * p2d float32 [n][2] (u, v).  out_inlier (uint8 [b][n], may be NULL) receives the flags
 * (the `in1` array of choosePose.py:24,31 for the winning hypothesis). */
int isr_pnp_score(const float *p3d, const float *p2d, int64_t n, const double *cam,
                  const double *poses, int64_t b, double reperr, int32_t *out_count,
                  uint8_t *out_inlier, void *stream);

/* out2 (int64 [2]) = {index of the FIRST maximum of v[0..n) (np.argmax), that maximum}. */
int isr_first_max(const int32_t *v, int64_t n, int64_t *out2, void *stream);

/* ---- PnP-RANSAC hypothesis generation (second half) ---------------- */
/* All solutions of the perspective three-point problem for b explicit samples (Grunert's quartic,
 * FP64): pts float64 [b][3][3] object points, uv float64 [b][3][2] pixels, cam double[9].
 * out_poses float64 [b][4][16] (NaN-filled beyond the out_n[b] real solutions), camera = R object
 * + t.  The minimal solver behind cv2.solveP3P / SOLVEPNP_P3P (choosePose.py:23). */
int isr_p3p_solve(const double *pts, const double *uv, const double *cam, int64_t b, double *out_poses,
                  int32_t *out_n, void *stream);
/* cv2.solvePnPRansac(p3d, p2d, cam, None, iterationsCount=iterations, reprojectionError=reperr,
 * flags=SOLVEPNP_P3P) as choosePose.py:23-33 calls it, on the device: `iterations` hypotheses
 * (4 correspondences each from a counter-based generator seeded with `seed`; P3P on three, the
 * fourth picks the solution), the consensus test of isr_pnp_score for all of them, the first
 * hypothesis with the largest consensus, and `refine_rounds` rounds of a Gauss-Newton refit on
 * the inliers (consensus re-evaluated between rounds; OpenCV refits with EPnP).  All
 * `iterations` are evaluated (no early exit: they run in parallel).
 * out_pose float64 [16]; out_counts int32 [2] = {consensus of the winning hypothesis, consensus
 * of out_pose}; out_inlier uint8 [n] = the winning hypothesis' consensus set (cv2's `inliers`).
 * out_counts[0] == 0: no pose (the reference's `pnp` then returns (1, 1, 1)). */
size_t isr_pnp_ransac_workspace_bytes(int64_t n, int64_t iterations);
int isr_pnp_ransac(const float *p3d, const float *p2d, int64_t n, const double *cam, int64_t iterations,
                   uint64_t seed, double reperr, int refine_rounds, double *out_pose,
                   int32_t *out_counts, uint8_t *out_inlier, void *workspace, size_t workspace_bytes,
                   void *stream);

/* ---- measurement helpers ------------------------------------------------------------ */
/* Developer probe of the pruned search: while dev_buf (device, 4 x uint64 per record) is set,
 * every CTA of every isr_nn2 / ICP launch writes record [batch * gridDim.x + blockIdx.x] =
 * {SM cycles it ran, scanned sub-tiles << 32 | exact sub-tile tests, scanned quarter units << 32
 * | candidate stages, query block << 32 | row code << 24 | resolve passes}.  NULL switches it off. */
int isr_debug_cta_log(uint64_t *dev_buf, int64_t capacity_records);

/* FFMA-chain microbenchmark: launches `blocks` x 256 threads, each running `iters`
 * rounds of 16 independent FMAs (packed != 0: fma.rn.f32x2).  flops_out_host receives the
 * number of FP32 flops executed.  Used by bench.py to report the measured FP32 peak. */
int isr_bench_ffma(int blocks, int iters, int packed, float *sink, double *flops_out_host,
                   void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ISR_H_ */
