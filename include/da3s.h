/*
 * da3s.h — C ABI of libda3s.so: the B200 (sm_100a) implementation of DA3-SLAM's
 * submap-alignment hot path.
 *
 * The reference (joey674/DA3-SLAM) is pure Python and has no FFI of its own; its
 * boundary for this path is a set of Python functions taking numpy arrays.  Each
 * entry point below names the reference function(s) whose arithmetic it replaces
 * (paths relative to the reference tree).  INTEGRATION.md shows the ctypes stub a
 * maintainer adds on the reference side.
 *
 * Rules for every function unless stated otherwise
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void*;
 *   - all array pointers are DEVICE pointers owned by the caller; *_host functions
 *     take HOST pointers and do the copies themselves;
 *   - asynchronous on `stream`; nothing is allocated after da3s_create();
 *   - returns 0 (DA3S_OK) or a negative DA3S_E* code, never throws;
 *   - one da3s_ctx per (device, stream); not thread-safe across threads sharing it;
 *   - there is no CPU fallback: a missing device or driver is DA3S_ECUDA.
 */
#ifndef DA3S_H
#define DA3S_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DA3S_VERSION 100

/* ---- error codes ------------------------------------------------------------ */
#define DA3S_OK       0
#define DA3S_EINVAL  (-1)   /* bad argument (null pointer, non-positive size, unknown flag)   */
#define DA3S_EALIGN  (-2)   /* a flat depth/conf/xyz base pointer is not 16-byte aligned     */
#define DA3S_ENOMEM  (-3)   /* the context workspace is too small for this call              */
#define DA3S_ECUDA   (-4)   /* a CUDA runtime call failed (see da3s_last_cuda_error)        */
#define DA3S_ETOOFEW (-5)   /* fewer usable points than the algorithm needs                  */

typedef struct da3s_ctx da3s_ctx;

/* Creates the per-device context and its workspace (scratch partial sums, selection
 * histograms, hypothesis tables, the voxel hash table).  The only allocating call. */
int da3s_create(int device, size_t workspace_bytes, da3s_ctx** out);
int da3s_destroy(da3s_ctx* ctx);
const char* da3s_strerror(int code);
int da3s_version(void);
/* cudaError_t of the last failing runtime call seen by this context (0 if none). */
int da3s_last_cuda_error(const da3s_ctx* ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
unsigned long long da3s_launch_count(const da3s_ctx* ctx);
/* lets kernels of this context's device address memory of `peer_device` (NVLink peer access); needed once
 * per peer before da3s_voxel_send is given pointers into that peer's memory */
int da3s_enable_peer_access(da3s_ctx* ctx, int peer_device);

/* Measures this device's float32 FMA throughput with the library's own microbenchmark kernel (dependent FFMA chains on
 * every SM, `iters` x 4096 FMAs per thread) and returns TFLOP/s (2 flop per FMA) in *tflops_out (HOST pointer).
 * Synchronises `stream`.  bench.py uses it as the denominator for the ALU-bound RANSAC scoring kernel. */
int da3s_measure_fp32_peak(da3s_ctx* ctx, int iters, double* tflops_out, void* stream);

/* Per-kernel device timers: with timers on, the library records a CUDA event pair on the caller's stream around every
 * launch of the kernels below (nothing else changes; the records cost no device time).  da3s_kernel_time synchronises
 * with the last recorded pair and returns the SUM of the durations of the `timed` most recent launches (all launches
 * since the previous read, at most DA3S_TIMER_RING), `timed`, the number of launches since the previous read and, where the
 * kernel counts it, the work ALL those launches executed (ransac_score_kernel: (hypothesis, correspondence) evaluations,
 * 27 flop each — invalid hypotheses and masked correspondences are skipped, so this is below n_hyp x pixels; else 0);
 * then it resets the ring.  work_out may be NULL.  bench.py's roofline lines divide
 * algorithmic bytes / flops by these. */
#define DA3S_TIMED_RANSAC_SCORE 0   /* ransac_score_kernel          */
#define DA3S_TIMED_IRLS         1   /* pair_moments_mixed_kernel    */
#define DA3S_TIMED_EXPORT_VOXEL 2   /* export_voxel_kernel          */
#define DA3S_TIMED_VOXEL_EMIT   3   /* voxel_emit_kernel            */
#define DA3S_TIMED_KERNELS      4
#define DA3S_TIMER_RING         256
int da3s_kernel_timers(da3s_ctx* ctx, int on);
int da3s_kernel_time(da3s_ctx* ctx, int which, double* sum_ms_out, int* timed_out, int* launches_out, double* work_out);

/* ---- camera table ------------------------------------------------------------- */
/* One entry per frame, device resident.  Built by da3s_build_cams from the network's
 * float32 intrinsics [n,3,3] and world-to-camera extrinsics [n,3,4]
 * (solver.py:168-176 lists the Prediction fields). */
typedef struct da3s_cam {
    float  fu, fv, cu, cv;      /* pinhole intrinsics, zero skew assumed by the closed form  */
    float  inv_fu, inv_fv;      /* float32(1)/fu, float32(1)/fv (fast float32 unprojection)  */
    float  skew_flag;           /* != 0 when K[0][1] or K[1][0] is non-zero                  */
    float  reserved;
    double kinv[9];             /* K^-1 row-major, float64 (general unprojection)            */
    double c2w[12];             /* camera-to-world [R|t] row-major, float64                  */
} da3s_cam;

#define DA3S_CAM_CLOSED_FORM 0  /* c2w = [R^T | -R^T t]  (src/vggt/utils/geometry.py:119-168) */
#define DA3S_CAM_GENERAL_INV 1  /* c2w = inverse of the 4x4 (utils/geometry.py:32-36)          */
int da3s_build_cams(da3s_ctx* ctx, const float* K9, const float* E12, int n_frames, int inverse_mode,
                    da3s_cam* cams_out, void* stream);

/* ---- K1: depth -> xyz, fused with confidence / validity filtering -------------------
 * Replaces  align_geometry.py:192-256, utils/align_geometry_single.py:52-102 (float32),
 *           utils/geometry.py:4-40 (float64, world),
 *           src/vggt/utils/geometry.py:14-116 (closed form),
 *           and, with `sim3`, utils/geometry.py:43-70 fused behind it
 *           (utils/da3_streaming.py:639-644 does the two back to back);
 *           the mask replaces utils/align.py:145-146 ('>'), viewer.py:333-336 ('>=' and
 *           conf > 0), utils/viser_server.py:108,186, align_geometry.py:325 (validity).
 * xyz_out: [n_frames,H,W,3] float32 (float64 with DA3S_UNPROJ_OUT_F64); mask_out:
 * [n_frames,H,W] uint8 (nullable); n_kept: one counter, ADDED to (nullable). */
#define DA3S_UNPROJ_CLOSED   0x0   /* x=(u-cu)*d/fu in float64 then float32: bit-exact vs VGGT   */
#define DA3S_UNPROJ_KINV     0x1   /* cam = K^-1 [u,v,1] d in float64 (skew allowed)              */
#define DA3S_UNPROJ_FAST     0x2   /* float32: ((u-cu)*d)*inv_fu  (oracle/SPEC.md section 1)      */
#define DA3S_UNPROJ_MODEMASK 0x3
#define DA3S_UNPROJ_WORLD    0x4   /* apply c2w (else camera coordinates)                         */
#define DA3S_UNPROJ_OUT_F64  0x8   /* write float64 xyz                                           */
#define DA3S_MASK_CONF_GT    0x10  /* keep conf >  thr   (utils/align.py:145)                     */
#define DA3S_MASK_CONF_GE    0x20  /* keep conf >= thr   (viewer.py:336)                          */
#define DA3S_MASK_CONF_FLOOR 0x40  /* and conf > conf_floor (viewer.py:334: 0; viser: 0.1, 1e-5)  */
#define DA3S_MASK_DEPTH      0x80  /* and depth > depth_eps and finite (align_geometry.py:325)    */
#define DA3S_MASK_WORLD_Z    0x100 /* and 0.1 < z_out < 50 and finite xyz (viewer.py:214-218)     */
#define DA3S_SIM3_PER_FRAME  0x200 /* sim3 holds one [13] row per frame instead of one row        */
int da3s_unproject_filter(da3s_ctx* ctx, const float* depth, const float* conf, const da3s_cam* cams,
                          int n_frames, int H, int W, int flags,
                          float conf_thr, const float* conf_thr_dev /* overrides conf_thr if non-null */,
                          float conf_floor, float depth_eps,
                          const double* sim3 /* nullable: s, R[9] row-major, t[3] */,
                          void* xyz_out, uint8_t* mask_out, unsigned long long* n_kept, void* stream);

/* Batched form: one launch over frames that live anywhere (e.g. every frame of every submap of a
 * sequence, each with its own cumulative Sim(3) and its own device-resident threshold).  `jobs`
 * is a DEVICE array; every depth/conf/xyz pointer must be 16-byte aligned, every mask pointer
 * 4-byte aligned, and H*W a multiple of 4.  conf_thr is used where a job's conf_thr is null. */
typedef struct da3s_frame_job {
    const float*    depth;      /* [H,W] */
    const float*    conf;       /* [H,W] or null */
    const da3s_cam* cam;        /* this frame's camera record */
    const double*   sim3;       /* [13] or null */
    const float*    conf_thr;   /* device scalar or null */
    void*           xyz;        /* [H,W,3] float32 (float64 with DA3S_UNPROJ_OUT_F64) */
    uint8_t*        mask;       /* [H,W] or null */
} da3s_frame_job;
int da3s_unproject_filter_jobs(da3s_ctx* ctx, const da3s_frame_job* jobs, int n_frames, int H, int W, int flags,
                               float conf_thr, float conf_floor, float depth_eps,
                               unsigned long long* n_kept, void* stream);

/* ---- K5: apply Sim(3) to a cloud -----------------------------------------------------
 * Replaces utils/geometry.py:43-70 (apply_sim3_transform): out = s * (p R^T) + t.
 * in_f64/out_f64 select float64 arrays (the reference returns float64 for float32 input). */
int da3s_apply_sim3(da3s_ctx* ctx, const void* xyz_in, int in_f64, long long n_points,
                    const double* sim3 /* device [13] */, void* xyz_out, int out_f64, void* stream);

/* ---- ordered filter + compaction of a resident cloud --------------------------------------
 * Replaces the map push of viewer.py:333-355 (`points[conf >= thr]`): keeps point i iff valid[i] != 0 (nullable = all) and,
 * when use_thr, conf[i] >= thr; writes the kept points (and colours, nullable pair) densely in INPUT order and their number
 * to *n_out (device).  At most max_out points are written (n_out still reports the full count). */
int da3s_filter_points(da3s_ctx* ctx, const float* xyz /* [n,3] */, const uint8_t* rgb /* [n,3] or null */,
                       const float* conf /* [n] */, const uint8_t* valid /* [n] or null */, long long n, int use_thr, float thr,
                       long long max_out, float* xyz_out, uint8_t* rgb_out, unsigned long long* n_out, void* stream);

/* ---- exact selection: medians and percentiles --------------------------------------------
 * Replaces np.median (utils/align.py:140-141, align_geometry.py:329,
 * utils/align_geometry_single.py:46) and np.percentile (viewer.py:334-335,
 * utils/viser_server.py:107,182) on float32 data, bit for bit: the order statistics are
 * selected exactly (radix select) and the float32 finishing arithmetic of numpy >= 2 is
 * reproduced on the device. */
#define DA3S_SEL_VALUES   0     /* keys = a[i]                                                   */
#define DA3S_SEL_POSITIVE 1     /* keys = a[i] where a[i] > 0            (viewer.py:334)          */
#define DA3S_SEL_RATIO    2     /* keys = a[i]/b[i] over the depth-scale mask (align_geometry.py:325-329) */
#define DA3S_SEL_MEDIAN     0
#define DA3S_SEL_PERCENTILE 1
typedef struct da3s_select_seg {
    const float* a;             /* values, or d_prev for RATIO                                    */
    const float* b;             /* d_cur for RATIO, else null                                     */
    const float* ca;            /* conf_prev for RATIO (nullable), else null                      */
    const float* cb;            /* conf_cur  for RATIO (nullable), else null                      */
    long long    n;             /* number of elements                                             */
    int          kind;          /* DA3S_SEL_VALUES / POSITIVE / RATIO                             */
    int          stat;          /* DA3S_SEL_MEDIAN / PERCENTILE                                   */
    float        percent;       /* for PERCENTILE, e.g. 65.0                                      */
    float        conf_th;       /* RATIO: conf > conf_th on both                                  */
    float        eps;           /* RATIO: depth > eps on both                                     */
    float        reserved;
} da3s_select_seg;
typedef struct da3s_select_out {
    long long n_valid;          /* elements that entered the selection                            */
    float     lo, hi;           /* the two order statistics used                                  */
    float     value;            /* median / percentile, numpy float32 arithmetic; NaN if n_valid==0 */
    float     gamma;            /* interpolation weight (percentile)                              */
} da3s_select_out;
int da3s_select(da3s_ctx* ctx, const da3s_select_seg* segs /* device */, int n_segs,
                long long max_n /* >= every segs[i].n */, da3s_select_out* out /* device */, void* stream);

/* ---- pair alignment --------------------------------------------------------------------
 * One "pair" = the `overlap` shared frames of two submaps: A = the previous submap's last
 * frames (target), B = the current submap's first frames (source); pixel i of A's frame k
 * corresponds to pixel i of B's frame k.  Result rows: target ~= s R source + t. */
typedef struct da3s_pair {
    const float*    depth_a;    /* [overlap,H,W] */
    const float*    conf_a;
    const float*    depth_b;
    const float*    conf_b;
    const da3s_cam* cam_a;      /* [overlap] */
    const da3s_cam* cam_b;
} da3s_pair;

typedef struct da3s_align_opts {
    int    world;               /* 1: points in each submap's world frame (utils/align.py:324-335,
                                      utils/da3_streaming.py:552-562); 0: camera frame of the overlap
                                      frame (align_geometry.py:272-285)                               */
    int    depth_scale_mode;    /* 0 off; 1 guarded median (utils/align_geometry_single.py:31-49);
                                      2 plain median (align_geometry.py:307-330).  Multiplies B's depth
                                      (solver.py:125-126, main_align.py:33-34)                        */
    float  depth_conf_th;       /* 0.2 in the reference                                            */
    float  depth_eps;           /* 1e-6 in the reference                                           */
    int    valid_depth;         /* also require depth > depth_eps and finite on both sides         */
    float  conf_thr_override;   /* if not NaN: use instead of min(median,median)*0.1               */
    int    huber;               /* 1: IRLS with Huber weights (utils/align.py:174-211); 0: one
                                      confidence-weighted Umeyama solve (utils/align.py:14-40)          */
    double huber_delta;         /* 1.0 in the reference (utils/align.py:94)                        */
    int    max_iterations;      /* 20 (utils/align.py:114)                                         */
    double tol;                 /* 1e-6 (utils/align.py:115)                                       */
    int    min_points;          /* 100 (utils/align.py:113): fewer -> identity, status 1           */
    int    n_hyp;               /* RANSAC hypotheses per pair, 0 = no RANSAC (oracle/SPEC.md 4)    */
    float  ransac_thr;          /* inlier iff r^2 < thr^2 (strict, align_geometry.py:123)          */
    int    ransac_min_inliers;  /* 20 (align_geometry.py:124): fewer -> identity, status 2          */
    int    precise;             /* 0: float32 micro-batches flushed into float64 accumulators (default,
                                      <= 1e-8 from the float64 oracle); 1: float64 per correspondence      */
} da3s_align_opts;
void da3s_align_opts_default(da3s_align_opts* opts);

/* Row layout of the Sim(3) table, float64[16] per pair — the only data that crosses
 * NVLink in the multi-GPU path (one all_gather of these rows). */
#define DA3S_ROW_S        0
#define DA3S_ROW_R        1     /* 9 entries row-major */
#define DA3S_ROW_T        10    /* 3 entries */
#define DA3S_ROW_NVALID   13    /* correspondences that entered the last solve */
#define DA3S_ROW_ITERS    14    /* IRLS iterations executed */
#define DA3S_ROW_STATUS   15    /* 0 ok, 1 too few points, 2 RANSAC found no model */
#define DA3S_ROW_LEN      16

typedef struct da3s_pair_aux {  /* optional per-pair diagnostics, float64[8] per pair */
    double conf_thr;            /* threshold actually used (float32 value)       */
    double depth_scale;         /* depth scale applied to B (float32 value)      */
    double median_a, median_b;  /* confidence medians                            */
    double best_hyp;            /* index of the winning hypothesis or -1         */
    double best_count;          /* its inlier count                              */
    double mean_residual;       /* residual of the last IRLS pass: rms (default kernel) or mean |r| (precise) */
    double last_change;         /* |ds| + ||dR||_F + ||dt|| of the last update   */
} da3s_pair_aux;

/* The whole hot path for a batch of pairs, no host synchronisation inside:
 * exact medians -> thresholds/depth scale -> [3-point hypotheses -> inlier counts -> winner]
 * -> IRLS (fused unproject + mask + weight + float64 moments, closed-form solve) -> rows.
 * Replaces utils/align.py:111-218 (+ :307-343), align_geometry.py:259-330,
 * utils/align_geometry_single.py:31-49,105-122 and the per-pair body of
 * utils/da3_streaming.py:322-363.
 * sample_idx: device int32 [n_pairs, n_hyp, 3] pixel indices in [0, overlap*H*W) (null if n_hyp==0).
 * hyp_counts_out: device int32 [n_pairs, n_hyp] (nullable).  Without it the hypotheses are scored in rounds and those
 * that can no longer reach the leader's count are dropped (exact: same winner, same count, same rows). */
int da3s_align_pairs(da3s_ctx* ctx, const da3s_pair* pairs /* device */, int n_pairs,
                     int overlap, int H, int W, const da3s_align_opts* opts,
                     const int32_t* sample_idx, double* sim3_rows /* [n_pairs,16] */,
                     da3s_pair_aux* aux /* nullable */, int32_t* hyp_counts_out /* nullable */,
                     void* stream);

/* Sim(3) chain accumulation on the device (utils/geometry.py:73-119): rows[k] maps submap k+1
 * into submap k; cum[0] = identity, cum[k+1] = cum[k] o rows[k].  cum: [n_rows+1, 13] float64
 * (s, R row-major, t) — directly usable as the per-frame/per-submap `sim3` of
 * da3s_unproject_filter, so the map export needs no host round trip. */
int da3s_accumulate_sim3(da3s_ctx* ctx, const double* rows /* [n_rows,16] */, int n_rows,
                         double* cum /* [n_rows+1,13] */, void* stream);

/* Stage-level entry points (the same kernels; used by the parity tests and by callers
 * that already hold thresholds / hypotheses). */
int da3s_pair_thresholds(da3s_ctx* ctx, const da3s_pair* pairs, int n_pairs, int overlap, int H, int W,
                         const da3s_align_opts* opts, float* conf_thr /* [n_pairs] */,
                         float* depth_scale /* [n_pairs] */, da3s_pair_aux* aux, void* stream);
/* RANSAC scoring given hypotheses: hyp_A [n_pairs,n_hyp,9] float32 (= s*R), hyp_t [n_pairs,n_hyp,3]
 * float32, hyp_ok [n_pairs,n_hyp] uint8 -> counts [n_pairs,n_hyp] int32 (overwritten). */
int da3s_ransac_score(da3s_ctx* ctx, const da3s_pair* pairs, int n_pairs, int overlap, int H, int W,
                      int world, int valid_depth, float depth_eps, const float* conf_thr, const float* depth_scale,
                      const float* hyp_A, const float* hyp_t, const uint8_t* hyp_ok, int n_hyp,
                      float ransac_thr, int32_t* counts, void* stream);
int da3s_ransac_hypotheses(da3s_ctx* ctx, const da3s_pair* pairs, int n_pairs, int overlap, int H, int W,
                           int world, int valid_depth, float depth_eps, const float* conf_thr, const float* depth_scale,
                           const int32_t* sample_idx, int n_hyp,
                           float* hyp_A, float* hyp_t, uint8_t* hyp_ok, double* hyp_sim3 /* nullable [.,.,13] */,
                           void* stream);
/* Inlier mask of one hypothesis per pair: mask_out [n_pairs, overlap*H*W] uint8. */
int da3s_ransac_inlier_mask(da3s_ctx* ctx, const da3s_pair* pairs, int n_pairs, int overlap, int H, int W,
                            int world, int valid_depth, float depth_eps, const float* conf_thr, const float* depth_scale,
                            const float* best_A /* [n_pairs,9] */, const float* best_t /* [n_pairs,3] */,
                            const uint8_t* best_ok /* [n_pairs] */, float ransac_thr, uint8_t* mask_out, void* stream);

/* How da3s_align_pairs scores when no count table is asked for (no reference counterpart; oracle/SPEC.md 4): the
 * correspondence tiles of a frame (DA3S_RANSAC_TILE pixels each, row-major) are dealt into DA3S_RANSAC_ROUNDS rounds;
 * after round 0 the hypothesis with the most inliers so far is counted on every other tile, and before each later round
 * a hypothesis is dropped when its count so far plus ALL correspondences of the tiles it has not seen is below that
 * complete count.  da3s_ransac_round_of is the dealing rule (host function, no device needed): the round of `tile`
 * in a frame of `tiles_per_frame` tiles. */
#define DA3S_RANSAC_ROUNDS 3
#define DA3S_RANSAC_TILE   16384
int da3s_ransac_round_of(int tile, int tiles_per_frame);

/* ---- Umeyama on materialised correspondences ----------------------------------------------
 * Replaces utils/align.py:14-40 (weighted_umeyama_alignment; variant 0),
 * align_geometry.py:59-82 (_umeyama_sim3; variant 1) and utils/align.py:224-276
 * (norm-ratio scale + Kabsch; variant 2) and utils/align.py:42-92 (variant 3).  src/dst: [n,3] float32 or float64; weights:
 * [n] float32/float64 or null (=1); idx_src/idx_dst: optional int64 gather lists of length
 * n_idx (the reference's independent-mask subsample, utils/align.py:145-165). */
#define DA3S_UMEYAMA_WEIGHTED  0
#define DA3S_UMEYAMA_MEAN      1
#define DA3S_UMEYAMA_NORMRATIO 2
#define DA3S_UMEYAMA_LEGACY_TRACE 3  /* utils/align.py:42-92 (weighted_umeyama_alignment0): scale = trace(Sigma)/var */
int da3s_umeyama_points(da3s_ctx* ctx, const void* src, const void* dst, int points_f64,
                        const void* weights, int weights_f64, long long n,
                        const long long* idx_src, const long long* idx_dst, long long n_idx,
                        int variant, double* sim3_row /* device [16] */, void* stream);
/* IRLS on materialised correspondences (utils/align.py:169-211) with an explicit gather. */
int da3s_irls_points(da3s_ctx* ctx, const void* src, const void* dst, int points_f64,
                     const float* conf_src, const float* conf_dst, long long n,
                     const long long* idx_src, const long long* idx_dst, long long n_idx,
                     double huber_delta, int max_iterations, double tol,
                     double* sim3_row /* device [16] */, void* stream);

/* ---- K6: voxel-grid downsample (hash grid) -------------------------------------------------
 * No counterpart in the reference (main_3dgs.py:1-3 is a stub; viewer.py:205-206 strides;
 * utils/da3_streaming.py:665-673 random-samples).  Spec: oracle/SPEC.md section 5.
 * xyz: [n,3] float32; rgb: [n,3] uint8 (nullable); mask: [n] uint8 (nullable).
 * Outputs sized for `max_voxels`: xyz_out [.,3] float32, rgb_out [.,3] uint8 (nullable),
 * count_out [.] int32, key_out [.] int64 (nullable); *n_voxels receives the number written
 * (device scalar); order is unspecified (sort by key_out for a canonical order).
 * The hash table lives in the context workspace and is cleared by da3s_voxel_begin; several
 * da3s_voxel_insert calls may accumulate into it before da3s_voxel_finish compacts it. */
int da3s_voxel_begin(da3s_ctx* ctx, long long table_slots /* power of two */, void* stream);
int da3s_voxel_insert(da3s_ctx* ctx, const float* xyz, const uint8_t* rgb, const uint8_t* mask,
                      long long n, float voxel, void* stream);
/* Same, for a DEVICE table of clouds in one launch (e.g. every submap of a sequence after
 * da3s_unproject_filter_jobs); max_n = the largest job's n.  width > 0 declares the clouds to be
 * image sequences with rows of `width` points ([frames,H,W,3]): warps then walk 8x16-pixel patches
 * instead of runs of 128 points, which lets them merge the points of a voxel before they reach the
 * table (same result — the sums are integers — fewer atomics); width = 0 for unstructured clouds. */
typedef struct da3s_voxel_job {
    const float*   xyz;         /* [n,3] float32                                  */
    const uint8_t* rgb;         /* [n,3] uint8, nullable (all jobs alike)         */
    const uint8_t* mask;        /* [n] uint8, nullable = every point              */
    long long      n;
} da3s_voxel_job;
int da3s_voxel_insert_jobs(da3s_ctx* ctx, const da3s_voxel_job* jobs_dev, int n_jobs, long long max_n,
                           int width, float voxel, void* stream);
/* Fused export of image frames straight into the grid: what da3s_unproject_filter_jobs (fast float32
 * mode, same flags) followed by da3s_voxel_insert_jobs would insert — bit for bit — without writing
 * the points (utils/da3_streaming.py:639-644 + the map export that follows it).  Only
 * DA3S_UNPROJ_FAST (+ DA3S_UNPROJ_WORLD, DA3S_MASK_CONF_*, DA3S_MASK_DEPTH) is accepted. */
typedef struct da3s_export_job {
    const float*    depth;      /* [H,W], 16-byte aligned          */
    const float*    conf;       /* [H,W] or null                   */
    const da3s_cam* cam;
    const double*   sim3;       /* [13] or null                    */
    const float*    conf_thr;   /* device scalar or null           */
    const uint8_t*  rgb;        /* [H,W,3] or null (all jobs alike) */
} da3s_export_job;
int da3s_unproject_voxel_jobs(da3s_ctx* ctx, const da3s_export_job* jobs_dev, int n_frames, int H, int W, int flags,
                              float conf_thr, float conf_floor, float depth_eps, float voxel, void* stream);
/* ---- nearest-neighbour registration of two unordered clouds -----------
 * DA3S_ICP_SIM3  replaces align_geometry.py:84-140 (align_two_point_clouds_umeyama: KD-tree nearest
 *                neighbour, d^2 < thr^2, fewer than 20 inliers stops, _umeyama_sim3 per iteration, composed);
 * DA3S_ICP_RIGID replaces the Open3D registration_icp calls at align_geometry.py:29-45 and
 *                utils/align_geometry_single.py:146-160 (point-to-point, s == 1, relative fitness / rmse 1e-6).
 * src [n_src,3], dst [n_dst,3] float32 or float64 device arrays; rows with a non-finite coordinate are
 * ignored (:94-95).  Result row as DA3S_ROW_* (NVALID = inliers of the last evaluation, ITERS = updates
 * applied, STATUS 1 = the first evaluation already had too few inliers: identity returned).
 * cell_size: edge of the search grid's cells, 0 = threshold; pass about twice the point spacing of `dst`
 * when the threshold is much larger than that (the search walks shells of cells and stops early). */
#define DA3S_ICP_SIM3  0
#define DA3S_ICP_RIGID 1
int da3s_icp_points(da3s_ctx* ctx, const void* src, long long n_src, const void* dst, long long n_dst, int points_f64,
                    int mode, double threshold, double cell_size, int max_iterations, double* sim3_row, void* stream);

/* ---- multi-GPU merge of voxel grids (global map export) -------------------
 * One rank per GPU fills its own grid (inserts above), then
 *   da3s_voxel_send         compacts the local table (leaving it clean and still active) and writes every
 *                           48-byte record {key, sum_qx, sum_qy, sum_qz, (n,sum_r), (sum_g,sum_b)} into the
 *                           inbox of the rank that owns its key: inbox_ptrs[d] is rank d's inbox
 *                           [world][cap][6] u64 mapped into this process (peer memory over NVLink, e.g. a
 *                           CUDA-IPC mapping), count_ptrs[d] its [world] u64 record counts, flag_ptrs[d] its
 *                           [world] u64 arrival flags; rank r writes segment r, counts[r] and — after a
 *                           system-scope fence over all its stores — flags[r] = step.  Records beyond `cap`
 *                           are dropped and reported by finish.
 *   da3s_voxel_merge_inbox  waits ON THE DEVICE until flags[s] >= step for every source rank s (no host
 *                           synchronisation, no process-group barrier; flags = null skips the wait when the
 *                           caller has synchronised otherwise), then folds this rank's own inbox into its table;
 *                           da3s_voxel_finish emits the rank's share of the global map.
 * `step` must increase by one per merge; callers alternate between two inbox/counts/flags buffers by step parity so
 * that a peer's stores of step k+1 never land in the buffer being merged for step k (sharding.VoxelExchange).
 * Integer sums => bit-identical to one GPU. */
int da3s_voxel_send(da3s_ctx* ctx, int world, int rank, void* const* inbox_ptrs /* host array [world] */,
                    void* const* count_ptrs /* host array [world] */, void* const* flag_ptrs /* host array [world] */,
                    unsigned long long step, long long cap, void* stream);
int da3s_voxel_merge_inbox(da3s_ctx* ctx, const void* inbox, const void* counts, const void* flags /* nullable */,
                           unsigned long long step, int world, long long cap, void* stream);
int da3s_voxel_finish(da3s_ctx* ctx, float voxel, long long max_voxels, float* xyz_out, uint8_t* rgb_out,
                      int32_t* count_out, long long* key_out, unsigned long long* n_voxels,
                      unsigned long long* n_dropped /* nullable: points lost to a full table */, void* stream);

/* ---- host-buffer entry point (end-to-end path) ---------------------------------------------
 * Same as da3s_align_pairs but every array is a HOST pointer (pinned memory recommended):
 * depth/conf are [n_pairs, overlap, H, W] per side, K [n_pairs,overlap,3,3] and
 * E [n_pairs,overlap,3,4] per side.  Copies in, runs, copies the rows out, synchronises. */
int da3s_align_pairs_host(da3s_ctx* ctx, int n_pairs, int overlap, int H, int W,
                          const float* depth_a, const float* conf_a, const float* K_a, const float* E_a,
                          const float* depth_b, const float* conf_b, const float* K_b, const float* E_b,
                          const da3s_align_opts* opts, const int32_t* sample_idx,
                          double* sim3_rows, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DA3S_H */
