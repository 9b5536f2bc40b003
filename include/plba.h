/*
 * plba.h — C ABI of the B200-native local bundle adjustment (LBA) for point + Plücker-line SLAM.
 *
 * This is the drop-in boundary for the reference's LBA hot path (kongan/PL-SLAM-plucker):
 *   - MapHandler::levMarquardtOptimizationLBA            include/mapHandler.h:128, src/mapHandler.cpp:2334-3016  (profile H_END)
 *   - MapHandler::levMarquardtOptimizationLBAForPluker   include/mapHandler.h:132, src/mapHandler.cpp:1618-2332  (profile H_PLK)
 *   - MapHandler::localBundleAdjustmentForPlukerWithG2O  include/mapHandler.h:134, src/mapHandler.cpp:5851-6323  (profile G)
 *     together with g2o_types/g2o_types.h:1-506 (vertex / edge arithmetic).
 *
 * Plain pointers and sizes only; no torch / Eigen / STL types cross this boundary.  All floating point is FP64,
 * all indices int32.  Host pointers in / out for plba_solve*(); the library stages H2D / D2H itself.
 * A handle owns one CUDA device, one stream and its workspace; it is re-entrant per handle and NOT
 * thread-safe per handle (the reference calls LBA from a single local-mapping thread, src/mapHandler.cpp:1251-1300).
 *
 * The same plba_problem / plba_options / plba_result layouts are consumed by the CPU oracle (oracle/), which is
 * test infrastructure only.
 */
#ifndef PLBA_H
#define PLBA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLBA_VERSION 1

/* Which reference implementation is reproduced ("Read this first"). */
enum {
    PLBA_PROFILE_G     = 0, /* g2o graph LBA: Huber, additive lambda, Schur, 5+10 iterations (src/mapHandler.cpp:5851)  */
    PLBA_PROFILE_H_END = 1, /* hand LM, endpoint lines, Cauchy, multiplicative lambda        (src/mapHandler.cpp:2334)  */
    PLBA_PROFILE_H_PLK = 2  /* hand LM, orthonormal Plücker lines (dead + buggy in reference) (src/mapHandler.cpp:1618)  */
};

/* Profile H_END only: the LM shell around the per-observation arithmetic.  GBA = MapHandler::levMarquardtOptimizationGBA
 * (include/mapHandler.h:137, src/mapHandler.cpp:3128-3728; the "next" row 1 of SURVEY.md §8f): identical observation terms (incl.
 * Q3 / Q4, :3547-3550), every keyframe but KF 0 free, and a shell that differs in: lambda0 scaled by an `int Hmax` (truncated,
 * :3386-3392), all three stop tests against numeric_limits<double>::epsilon() (:3664, :3694), err /= 0 in EVERY iteration
 * (Npt_obs, Nls_obs are never counted, :3383, :3662), no `inlier` rule on write-back (:3705-3726), void return. */
enum { PLBA_SHELL_LBA = 0, PLBA_SHELL_GBA = 1 };

/* Bug-for-bug (FAITHFUL) or intended maths (FIXED); every difference is listed in SURVEY.md §8.Q. */
enum { PLBA_QUIRKS_FAITHFUL = 0, PLBA_QUIRKS_FIXED = 1 };

/* Return codes.  0 / -1 follow the reference (src/mapHandler.cpp:3011-3012, 1499-1500). */
enum {
    PLBA_OK          = 0,   /* optimisation applied                                   */
    PLBA_DISCARDED   = -1,  /* nothing to do (no observations) / results discarded    */
    PLBA_E_ARG       = -2,  /* malformed problem (index out of range, NULL pointer …) */
    PLBA_E_CUDA      = -3,  /* CUDA runtime failure, see plba_last_error()            */
    PLBA_E_NUMERIC   = -4,  /* the reduced camera system was not positive definite in EVERY trial of an LM iteration: outputs are
                               written (the rejected trials left the state untouched) but no longer follow the reference, whose
                               SimplicialLDLT has no positivity requirement */
    PLBA_E_UNSUPPORTED = -5
};

/*
 * One LBA window, flattened.  Mirrors what MapHandler::localBundleAdjustment() (src/mapHandler.cpp:1392-1502)
 * gathers into X_aux / kf_list / pt_list / ls_list / pt_obs_list / ls_obs_list plus the map values the LM
 * functions read through `this` (T_kf_w, obs_list, sigma_list, cam).
 *
 * Keyframes: n_kf rows = every KF referenced by an observation, free or fixed.  kf_slot[i] = position in the
 * reference's kf_list (0..n_free-1, strictly ascending with i among free KFs) or -1 for a fixed observer
 * (Vector6i(4) == -1, src/mapHandler.cpp:1434; g2o setFixed(true), :5962).
 * Observations are landmark-major, i.e. po_lm / lo_lm are non-decreasing (reference order, :1424-1447).
 */
typedef struct plba_problem {
    int32_t n_kf;            /* rows of kf_T_wc / kf_slot                                          */
    int32_t n_free;          /* == kf_list.size()                                                  */
    int32_t n_pt;            /* point landmarks  (== pt_list.size())                               */
    int32_t n_ls;            /* line landmarks   (== ls_list.size())                               */
    int32_t n_pobs;          /* point observations                                                 */
    int32_t n_lobs;          /* line observations                                                  */
    double  cam[4];          /* fx, fy, cx, cy   (PinholeStereoCamera getters)                     */
    const double  *kf_T_wc;  /* [n_kf][12] rows of the 3x4 [R|t] of KeyFrame::T_kf_w (camera->world) */
    const int32_t *kf_slot;  /* [n_kf]                                                             */
    const double  *x_pose;   /* [n_free][6] pose part of X_aux (= KeyFrame::x_kf_w, translation first); profiles H only;
                                NULL => logmap_se3(T_kf_w)                                          */
    const double  *pt_xyz;   /* [n_pt][3]  MapPoint::point3D                                       */
    const double  *ls_plk;   /* [n_ls][6]  MapLine::NDw = [n; d]            (profiles G, H_PLK)    */
    const double  *ls_end;   /* [n_ls][6]  MapLine::line3D = [P; Q]         (profile H_END)        */
    const int32_t *po_lm;    /* [n_pobs] local point index 0..n_pt-1, non-decreasing               */
    const int32_t *po_kf;    /* [n_pobs] row of kf_T_wc                                            */
    const double  *po_uv;    /* [n_pobs][2] MapPoint::obs_list[j]                                  */
    const double  *po_sig2;  /* [n_pobs] MapPoint::sigma_list[j]; NULL => 1                        */
    const int32_t *lo_lm;    /* [n_lobs] local line index 0..n_ls-1, non-decreasing                */
    const int32_t *lo_kf;    /* [n_lobs]                                                           */
    const double  *lo_ab;    /* [n_lobs][4] G / H_PLK: NDw_obs_list[j] = (spl.x, spl.y, epl.x, epl.y);
                                H_END: obs_list[j] = (lx, ly, lz, unused)                          */
    const double  *lo_sig2;  /* [n_lobs]; NULL => 1                                                */
} plba_problem;

/* LM schedule + robust weighting.  Defaults (plba_default_options) are the reference's:
 * src/slamConfig.cpp:65-67, src2/config.cpp:80-85, src/mapHandler.cpp:5978,6122,6129,6152. */
typedef struct plba_options {
    int32_t profile;          /* PLBA_PROFILE_*                                                    */
    int32_t quirks;           /* PLBA_QUIRKS_*                                                     */
    /* profile H */
    double  lambda_lba_lm;    /* 1e-5  SlamConfig::lambdaLbaLM                                     */
    double  lambda_lba_k;     /* 10    SlamConfig::lambdaLbaK                                      */
    int32_t max_iters_lba;    /* 15    SlamConfig::maxItersLba                                     */
    int32_t shell;            /* PLBA_SHELL_LBA (0) or PLBA_SHELL_GBA (1): which hand-LM shell wraps profile H_END */
    double  homog_th;         /* 1e-7  Config::homogTh                                             */
    double  min_error;        /* 1e-7  Config::minError                                            */
    double  min_error_change; /* 1e-7  Config::minErrorChange                                      */
    /* profile G */
    double  huber_delta;      /* (double)(float)sqrt(5.991)  (Q13)                                 */
    double  chi2_gate;        /* 5.991                                                             */
    int32_t iters_stage1;     /* 5                                                                 */
    int32_t iters_stage2;     /* 10                                                                */
    double  lm_tau;           /* 1e-5  g2o OptimizationAlgorithmLevenberg _tau                     */
    int32_t lm_max_trials;    /* 10    g2o _maxTrialsAfterFailure                                  */
    int32_t reserved1;
} plba_options;

/* One record per LM trial (profile G) or per LM iteration (profile H): the observable trace used for parity. */
typedef struct plba_trace_rec {
    int32_t window;    /* batch element                                                            */
    int32_t stage;     /* G: 0 = Huber stage, 1 = gated stage; H: 0                                */
    int32_t iter;      /* outer iteration                                                          */
    int32_t trial;     /* G: trial inside the iteration; H: 0                                      */
    int32_t accepted;  /* G: rho > 0 && finite; H: step applied                                    */
    int32_t stop;      /* 0 continue; 1 stopped here (H: which test: 1 err, 2 dx; G: 1 terminate)  */
    double  chi;       /* G: robust chi2 at linearisation point; H: err (normalised as reference)  */
    double  chi_new;   /* G: robust chi2 after the trial step;    H: unused (0)                    */
    double  rho;       /* G: gain ratio                                                            */
    double  lambda;    /* damping used for this solve                                              */
    double  scale;     /* G: computeScale() + 1e-3                                                 */
    double  dx_norm;   /* H: ||DX||                                                                */
    double  err_pt;    /* H: un-normalised point_error_lm (printed at src/mapHandler.cpp:2794)     */
    double  err_ls;    /* H: un-normalised line_error_lm  (:2795)                                  */
} plba_trace_rec;

/* Per-observation flags written by profile G (src/mapHandler.cpp:6125-6147, 6156-6293). */
enum {
    PLBA_OBS_LEVEL1 = 1,  /* gated out after stage 1 (setLevel(1))                                 */
    PLBA_OBS_BAD    = 2,  /* final test: chi2 > gate (or depth <= 0 for points) => observation to be erased */
    PLBA_OBS_NEGDEPTH = 4 /* point edge with z_c <= 0 at the final estimate                        */
};

/*
 * Outputs.  Every array pointer may be NULL (then it is not written).  Caller-owned buffers.
 */
typedef struct plba_result {
    double  *kf_T_wc;     /* [n_kf][12]  updated T_kf_w rows (fixed KFs copied through)            */
    double  *x_pose;      /* [n_free][6] final X pose part (profiles H)                            */
    double  *pt_xyz;      /* [n_pt][3]                                                             */
    double  *ls_plk;      /* [n_ls][6]   G / H_PLK: changeOrthToPluker(final orth)                 */
    double  *ls_orth;     /* [n_ls][4]   G / H_PLK: final orthonormal coordinates                  */
    double  *ls_end;      /* [n_ls][6]   H_END                                                     */
    uint8_t *pt_inlier;   /* [n_pt]  H: 0 if the landmark moved > 0.01 (src/mapHandler.cpp:2858-2860), else 1 */
    uint8_t *ls_inlier;   /* [n_ls]                                                                */
    double  *po_chi2;     /* [n_pobs] final chi2 per point observation (profile G, as e->chi2())   */
    double  *lo_chi2;     /* [n_lobs]                                                              */
    uint8_t *po_flags;    /* [n_pobs] PLBA_OBS_*                                                   */
    uint8_t *lo_flags;    /* [n_lobs]                                                              */
    plba_trace_rec *trace;/* [trace_cap]                                                           */
    int32_t  trace_cap;
    int32_t  n_trace;     /* out: records written                                                  */
    int32_t  status;      /* out: PLBA_OK / PLBA_DISCARDED / error                                 */
    int32_t  n_trials;    /* out: number of linear solves performed                                */
} plba_result;

typedef struct plba_handle_s *plba_handle;

/* ---- life cycle ---- */
void plba_default_options(int32_t profile, plba_options *opt);
/* stream: a cudaStream_t (as void*) the library launches on, or NULL to create its own. */
int  plba_create(int32_t device, void *stream, plba_handle *out);
void plba_destroy(plba_handle h);
const char *plba_last_error(plba_handle h);
int  plba_version(void);

/* ---- the drop-in call: replaces the three reference LBA functions named above ---- */
int  plba_solve(plba_handle h, const plba_problem *prob, const plba_options *opt, plba_result *res);
/* n independent windows (BASELINE config 3); identical result to n plba_solve() calls. */
int  plba_solve_batch(plba_handle h, int32_t n, const plba_problem *probs, const plba_options *opt, plba_result *res);

/* ---- staged interface: keeps the problem resident in HBM (benchmarks, multi-GPU sharding) ---- */
int  plba_upload(plba_handle h, int32_t n, const plba_problem *probs, const plba_options *opt);
int  plba_reset_state(plba_handle h);                       /* state <- uploaded initial values   */
int  plba_run(plba_handle h);                               /* full LM schedule on the resident problem */
int  plba_download(plba_handle h, int32_t n, plba_result *res);
/* One LM trial at a given damping on the resident problem, without the controller:
 * linearise + robust weights + block accumulation + Schur  -> reduced camera system (stage A),
 * solve + back-substitution + retraction + new cost          (stage B).
 * Used by bench.py for per-kernel roofline timing and by the sharded multi-GPU driver. */
int  plba_trial_assemble(plba_handle h, double lambda);
int  plba_trial_finish(plba_handle h, double lambda, double *chi_new, double *scale);
/* Reduced camera system exposed for the cross-GPU reduction: device pointer + length in doubles of
 * [S blocks | g_red | cost scalars]; the caller all-reduces it in place between assemble and finish. */
int  plba_reduced_system(plba_handle h, void **dev_ptr, int64_t *n_doubles);
/* Host copy of one window's reduced camera system as left by plba_trial_assemble(): S_out[6 n_free][6 n_free] row-major
 * (upper triangle of 6x6 blocks filled, undamped) and g_out[6 n_free].  Either pointer may be NULL. */
int  plba_copy_reduced_system(plba_handle h, int32_t window, double *S_out, double *g_out);
/* Average device time (ms) of one launch of a stage kernel on the resident problem, `reps` back-to-back launches between
 * two CUDA events on the handle's stream.  which: 0 = assembly (linearise + blocks + Schur), 1 = reduced-system solve,
 * 2 = update (back-substitution + retraction + new cost).  The LM state is not advanced. */
int  plba_time_kernel(plba_handle h, int32_t which, int32_t reps, double lambda, double *ms_avg);
/* Measured FP64 roof of the device (the denominators of an FP64-bound roofline are measured, not nominal): which 0 = DFMA on the vector
 * pipe, 1 = mma.sync.m8n8k4.f64 on the tensor pipe; best of `reps` launches, CUDA events; TFLOP/s. */
int  plba_measure_fp64_peak(plba_handle h, int32_t which, int32_t reps, double *tflops);
/* Large windows (6 n_free > 144): the reduced camera system is solved by a banded Cholesky when no landmark track spans more
 * than 15 free keyframes (sliding-window shape), by the dense tensor-core (FP64 DMMA) Cholesky otherwise.  on != 0 forces the
 * dense path (benchmarks, tests). */
int  plba_set_force_dense(plba_handle h, int on);
/* Assembly / update have two implementations: warp-autonomous kernels (one warp per run of landmarks with identical keyframe
 * sequence; large uploads whose longest track has at most 32 observations) and CTA-chunk kernels (everything else).
 * mode 0 = route by size (default), 1 = always the CTA-chunk kernels, 2 = the warp kernels whenever the tracks allow
 * (tests, A/B measurements; also the PLBA_FORCE_CHUNK environment variable). */
int  plba_set_force_chunk(plba_handle h, int mode);
/* Layout of the resident problem, for roofline accounting: out8 = { point chunks, line chunks, point segments, line segments,
 * off-diagonal Schur tasks, diagonal Schur tasks, structurally non-zero upper 6x6 blocks of S (nnzb), device arena bytes }. */
int  plba_layout_stats(plba_handle h, int64_t *out8);
/* Which kernels the resident upload runs on: out4 = { assembly / update: 0 = CTA-chunk kernels, 1 = warp-autonomous kernels;
 * reduced-system solver: 0 = one CTA per window in shared memory (6 n_free <= 144), 1 = block cyclic reduction (block-banded
 * large windows), 2 = single-CTA banded Cholesky, 3 = dense tiled FP64-tensor-core Cholesky; half bandwidth of S in pose
 * blocks; number of work units (chunks or warp items) }. */
int  plba_kernel_path(plba_handle h, int32_t *out4);
/* Optional all-reduce hook called on the handle's stream wherever the path has its exchange step
 * (SURVEY.md §8e).  fn(dev_ptr, n_doubles, stream, user) must sum the buffer in place over all ranks. */
typedef void (*plba_allreduce_fn)(void *dev_ptr, int64_t n_doubles, void *stream, void *user);
int  plba_set_allreduce(plba_handle h, plba_allreduce_fn fn, void *user);
/* with a caller-supplied all-reduce the library also needs the rank layout (the lambda-init maximum travels as one slot per rank
 * inside a SUM exchange): nranks <= 16 */
int  plba_set_allreduce_ranks(plba_handle h, int32_t nranks, int32_t rank);

/* ---- landmark-sharded LBA over several GPUs (the partition notion is the reference's base keyframe of a landmark,
 * include/mapHandler.h:148-149).  Every rank uploads ITS landmarks (with all their observations) and ALL keyframes; the library then
 * sums the reduced camera system over the ranks once per LM trial with ncclAllReduce on the handle's stream (plus one scalar
 * exchange after the update kernel), and every rank takes the same LM decisions.  NCCL is resolved at run time (libnccl.so.2 already
 * mapped into the process, or PLBA_NCCL_LIB, or the loader path): the library has no link-time dependency on it.
 *   one process per GPU : rank 0 calls plba_comm_unique_id(), ships the 128 bytes to the other ranks (MPI, sockets, torch.distributed,
 *                         a file ...), every rank calls plba_comm_init_rank() on its handle;
 *   one process, N GPUs : plba_create_group() = plba_create per device + ncclCommInitAll; drive each handle from its own host thread.
 * plba_set_allreduce() (a caller-supplied all-reduce) remains for hosts without NCCL and for the CPU tests. */
#define PLBA_COMM_ID_BYTES 128
int  plba_comm_unique_id(void *id_out);
int  plba_comm_init_rank(plba_handle h, int32_t nranks, int32_t rank, const void *id);
int  plba_comm_destroy(plba_handle h);
int  plba_create_group(int32_t ndev, const int32_t *devs, plba_handle *out_handles);
void plba_destroy_group(int32_t ndev, plba_handle *handles);
/* out2 = { ranks of the handle's communicator (1 = not sharded), this rank } */
int  plba_comm_info(plba_handle h, int32_t *out2);

/* Timing of the last plba_run(): CUDA-event milliseconds per kernel family and launch counts. */
typedef struct plba_timing {
    double  ms_total, ms_assemble, ms_solve, ms_update, ms_other;
    int64_t n_launches;      /* kernels launched by the library in the last plba_run / plba_solve  */
    int64_t n_assemble;      /* launches of the assembly kernel                                    */
    int64_t h2d_bytes, d2h_bytes;  /* host<->device bytes moved by the last plba_upload / plba_download (plba_solve = both) */
    int64_t n_launches_run;  /* kernels launched by the last plba_run                              */
    int64_t n_assemble_run;  /* assembly-kernel launches of the last plba_run                      */
    int64_t n_trials_run;    /* LM trials (linear solves) of the last plba_run, summed over windows */
    double  ms_host_prep;    /* host wall time of the last plba_upload before its H2D copy (validation, signature sort, flattening) */
    double  ms_host_unpack;  /* host wall time of the last plba_download after its D2H copy (un-permutation into caller buffers)  */
} plba_timing;
int  plba_get_timing(plba_handle h, plba_timing *t);
/* on: record CUDA events around the assembly / solve / update kernels of every LM round (fills ms_assemble ...). */
int  plba_set_detail_timing(plba_handle h, int on);

/* ---- synthetic scenes: deterministic generator shared by tests, bench and oracle ---- */
typedef struct plba_scene_spec {
    int32_t n_kf_free, n_kf_fixed;   /* free KFs; fixed observers (KF 0 + non-local)               */
    int32_t n_pt, n_ls;
    double  mean_track;              /* k                                                          */
    int32_t width, height;
    double  fx, fy, cx, cy;
    double  kf_spacing, depth_min, depth_max;
    double  pixel_noise, outlier_frac;
    double  pose_rot_noise, pose_trans_noise, pt_noise, ls_noise;
    uint64_t seed;
    int32_t line_mode;               /* 0: Plücker (ls_plk + endpoint-pixel obs); 1: endpoint (ls_end + normalised 2-D line obs) */
    int32_t loop_every;              /* >0: KF i also sees landmarks of KF i-loop_every (non-banded S) */
} plba_scene_spec;
void plba_scene_preset(int32_t config_id, plba_scene_spec *spec);   /* 1..5 = BASELINE configs (one window) */
/* The scene owns its arrays; plba_scene_problem() hangs them into *out (valid until plba_scene_destroy). */
typedef struct plba_scene_s *plba_scene;
int  plba_scene_create(const plba_scene_spec *spec, plba_scene *out);
int  plba_scene_problem(plba_scene s, plba_problem *out);
/* ground truth of the same scene (for convergence checks): kf_T_wc[n_kf][12], pt_xyz[n_pt][3], ls_plk[n_ls][6] */
int  plba_scene_truth(plba_scene s, const double **kf_T_wc, const double **pt_xyz, const double **ls_plk);
void plba_scene_destroy(plba_scene s);

/* ------------------------------------------------------------------------------------------------------------------
 * Frame-to-frame pose tracking in Plücker mode (the step BEFORE the LBA path, it produces the keyframe
 * poses): StereoFrameHandler::gaussNewtonOptimizationforPluker + optimizeFunctionsUsingPluker
 * (src2/stereoFrameHandler.cpp:563-853).  Pose-only Gauss-Newton on point re-projection errors and endpoint-to-projected-
 * Plücker-line distances, Cauchy weights scaled by the MAD of the residuals (src2/auxiliar.cpp:444-460), line weights
 * multiplied by the segment overlap (src2/stereoFrame.cpp:547-660).  Frames are independent: a batch runs one CTA per frame.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct plba_track_frame {
    int32_t n_pt, n_ls;
    double  DT[12];                 /* initial guess, rows of [R|t] of the 4x4 DT (previous frame -> current frame)                 */
    const double  *pt_P;            /* [n_pt][3] PointFeature::P       3-D point in the previous frame                               */
    const double  *pt_obs;          /* [n_pt][2] PointFeature::pl_obs  pixel in the current frame                                    */
    const uint8_t *pt_inlier;       /* [n_pt]    PointFeature::inlier  (NULL = all inliers)                                          */
    const double  *ls_sP, *ls_eP;   /* [n_ls][3] LineFeature::sP / eP  3-D endpoints in the previous frame (overlap weight only)     */
    const double  *ls_NDc;          /* [n_ls][6] LineFeature::NDc      Plücker line [n; d] in the previous frame                     */
    const double  *ls_obs;          /* [n_ls][4] spl_obs.xy, epl_obs.xy  endpoint pixels in the current frame                        */
    const double  *ls_seg;          /* [n_ls][4] spl.xy, epl.xy          endpoint pixels in the previous frame (overlap weight)      */
    const double  *ls_sigma2;       /* [n_ls]    LineFeature::sigma2   (NULL = 1)                                                    */
    const uint8_t *ls_inlier;       /* [n_ls]    LineFeature::inlier   (NULL = all inliers)                                          */
} plba_track_frame;
typedef struct plba_track_options {
    double  cam[4];                 /* fx, fy, cx, cy                                                                                */
    double  homog_th;               /* 1e-7  Config::homogTh                                                                         */
    double  min_error;              /* 1e-7  Config::minError                                                                        */
    double  min_error_change;       /* 1e-7  Config::minErrorChange                                                                  */
    int32_t max_iters;              /* 5     Config::maxIters (10 = maxItersRef for the refinement call)                             */
    int32_t reserved;
} plba_track_options;
typedef struct plba_track_result {
    double  DT[12];                 /* optimised DT; the initial guess if the solve failed (:846-851)                                */
    double  DT_cov[36];             /* H^-1 of the last linearisation; identity if the solve failed                                  */
    double  err;                    /* normalised robust cost of the last linearisation; -1 if the solve failed                      */
    int32_t iters;                  /* linearisations performed                                                                      */
    int32_t good;                   /* solution_is_good (:809)                                                                       */
} plba_track_result;
void plba_track_default_options(plba_track_options *opt);
/* n_frames independent frames, host buffers in and out.  At most 1024 point and 1024 line matches per frame. */
int  plba_track_solve(plba_handle h, int32_t n_frames, const plba_track_frame *frames, const plba_track_options *opt, plba_track_result *results);

/* ------------------------------------------------------------------------------------------------------------------
 * Creation of Plücker line landmarks (it defines the on-wire meaning of NDw and the sqrt(5.991) gate):
 * stereo triangulation of a segment as the intersection of the two viewing planes (StereoFrame::pi_from_ppp / pipi_plk,
 * src2/stereoFrame.cpp:381-397, 870-883), transformation into the world with the |n| / |d| renormalisation and the
 * re-projection gate against the matched segment of the current keyframe (src/mapHandler.cpp:449-492).  One thread per candidate.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct plba_newline_batch {
    int32_t n, n_kf;
    double  cam[5];                 /* fx, fy, cx, cy, stereo baseline                                                              */
    const double  *seg_l;           /* [n][4] sp_l.xy, ep_l.xy: the segment in the left image of the keyframe that creates the line  */
    const double  *seg_r;           /* [n][4] sp_r.xy, ep_r.xy: its match in the right image, after the row interpolation (:367-368) */
    const double  *seg_curr;        /* [n][4] spl.xy, epl.xy of the matched segment in the current keyframe (the gate)               */
    const int32_t *kf_prev, *kf_curr; /* [n] rows of kf_T_wc: creating keyframe, current keyframe                                    */
    const double  *kf_T_wc;         /* [n_kf][12] KeyFrame::T_kf_w                                                                   */
} plba_newline_batch;
/* NDc [n][6] (camera frame of the creating keyframe), NDw [n][6] (world, MapLine::NDw), err_first / err_curr [n] = norm of the two
 * endpoint-to-line distances in the creating / the current keyframe, accept [n] = !(err_curr > sqrt(5.991)) (:487).  Any output may be NULL. */
int  plba_create_lines(plba_handle h, const plba_newline_batch *batch, double *NDc, double *NDw, double *err_first, double *err_curr, uint8_t *accept);

#ifdef __cplusplus
}
#endif
#endif /* PLBA_H */
