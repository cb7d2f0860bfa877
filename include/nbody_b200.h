/* nbody_b200.h -- C ABI of the B200-native NBodySimProject hot path.
 *
 * The reference (calkan27/NBodySimProject, pure Python/NumPy) has no FFI of its own: its boundary
 * is the Python call surface.  Each entry point below replaces the reference
 * function(s) cited next to it; `nbodysimproject_b200/` binds them with ctypes and INTEGRATION.md
 * shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types.  `*_f64` / `*_f32` entry points take DEVICE pointers
 *     (caller-owned, e.g. torch tensors) and a cudaStream_t passed as `void*`; they never synchronise, never
 *     allocate device memory and are safe to call concurrently on distinct streams.  (nb_ensemble_run_f64 creates a
 *     per-device pool of internal priority streams and events on first use: it launches the sub-step-heavy head of
 *     a large batch at higher priority than the bulk and joins both back into the caller's stream before returning
 *     control of it, so stream-ordered semantics are unchanged.)
 *   - `*_host` entry points take HOST pointers and do H2D -> kernels -> D2H themselves (they own a
 *     cached per-device workspace); they return after the result is in the host buffers.
 *   - arrays use the reference's own layouts: m[B][N], q[B][N][2], v[B][N][2] row-major fp64
 *     (simulation_state.py:28-31), one batch = B systems with the same body count N (2..NB_MAX_N_MID).
 *   - every function returns NB_OK (0) or a negative error code and never throws.  Numerical failure
 *     of one system is reported per system in `status[B]`, never as a hang or exception.
 */
#ifndef NBODY_B200_H
#define NBODY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB_OK 0
#define NB_ERR_ARG (-1)        /* bad argument (null pointer, N out of range, ...) */
#define NB_ERR_CUDA (-2)       /* a CUDA runtime call failed; see nb_last_error() */
#define NB_ERR_UNSUPPORTED (-3)

#define NB_MIN_N 2
#define NB_MAX_N 8             /* register-resident ensemble kernels are templated on N = 2..8 (all modes) */
#define NB_MAX_N_MID 64        /* 9..64 bodies: one CTA per system (verlet / yoshida4 / whfast / ham_soft incl. setup and
                                  probe, pair + variational calls, prepare, host entry); classic adaptive softening: one
                                  thread per system */

/* integrator_mode (sim_config.py:19-24) */
#define NB_MODE_VERLET 0
#define NB_MODE_YOSHIDA4 1
#define NB_MODE_WHFAST 2
#define NB_MODE_HAMSOFT 3

/* per-system status word written by the ensemble kernels */
#define NB_STATUS_OK 0
#define NB_STATUS_NONFINITE 1      /* state became non-finite (reference: inf/NaN features, OverflowError) */
#define NB_STATUS_EPS_OOB 2        /* ham_soft: epsilon left [eps_min, eps_max] by more than their span */
#define NB_STATUS_KEPLER_NOCONV 4  /* whfast: Newton loop hit the 64-iteration cap */

/* flags for nb_ensemble_prepare_f64 */
#define NB_PREP_REMOVE_COM 1u      /* v -= sum(m v)/sum(m)            simulation.py:85-86, physics_utils.py:16-26 */
#define NB_PREP_CTOR_KICK 2u       /* v += 0.5*kick_dt*a(q)           simulation.py:150-157, integration_scheme_base.py:154-192 */
#define NB_PREP_SNAPSHOT_KICK 4u   /* a second, identical half kick   simulation.py:319-326 (snapshot()/copy()) */
#define NB_PREP_STATIC_FEATURES 8u /* fill static_features[B][NB_N_STATIC] from the (kicked) state */

/* flags for nb_ensemble_run_f64 */
#define NB_RUN_ENERGY 1u           /* E0/E1, L0/L1 + drifts           stability_analyzer.py:106-131 */
#define NB_RUN_WRITE_STATE 2u      /* write final q, v (and eps, pi) back */
#define NB_RUN_KEPLER_EXACT 4u     /* whfast: physically correct Kepler solver instead of the bug-compatible one */

/* columns of dyn_features[B][NB_N_DYN] (stability_analyzer.py:226-252, in the reference's order) */
enum {
  NB_F_IS_STABLE = 0, NB_F_ENERGY_DRIFT, NB_F_ANGMOM_DRIFT, NB_F_COM_MEAN, NB_F_COM_MAX,
  NB_F_JEPS_MEAN, NB_F_JEPS_STD, NB_F_THETA_MEAN, NB_F_THETA_STD, NB_F_COS_MEAN, NB_F_COS_MIN,
  NB_F_VARL_MEAN, NB_F_VARL_MAX, NB_F_TIDAL_MEAN, NB_F_TIDAL_MAX, NB_F_MEGNO, NB_F_LYAP_TIME,
  NB_F_E0, NB_F_E1, NB_F_L0, NB_F_L1, NB_F_T_END, NB_N_DYN
};
#define NB_N_DYN_USER 17   /* the columns a reference user sees; E0/E1/L0/L1/t_end behind them are parity taps */

/* columns of static_features[B][NB_N_STATIC] (dynamical_features.py:27-155, in the reference's order) */
enum {
  NB_S_TOTAL_MASS = 0, NB_S_MASS_VAR, NB_S_MASS_RATIO_MAX, NB_S_MASS_CENTER_OFFSET,
  NB_S_MEAN_SEP, NB_S_STD_SEP, NB_S_MIN_SEP, NB_S_MAX_SEP, NB_S_SEP_RATIO,
  NB_S_MEAN_SPEED, NB_S_STD_SPEED, NB_S_MAX_SPEED, NB_S_MEAN_RELVEL, NB_S_MAX_RELVEL,
  NB_S_KINETIC, NB_S_POTENTIAL, NB_S_TOTAL_ENERGY, NB_S_VIRIAL, NB_S_ENERGY_PER_MASS, NB_S_IS_BOUND,
  NB_S_TOTAL_ANGMOM, NB_S_MEAN_SPEC_ANGMOM, NB_S_ANGMOM_VAR, NB_S_SOFT_MEAN, NB_S_SOFT_STD,
  NB_N_STATIC
};

/* ham_soft per-system parameters (hamiltonian_softening_integrator.py:47-141, hamsoft_params.py:31-76);
 * one row of hs_params[B][NB_HS_NPARAM] */
enum {
  NB_HS_K_SOFT = 0, NB_HS_MU_SOFT, NB_HS_EPS_MIN, NB_HS_EPS_MAX, NB_HS_ALPHA_RUN, NB_HS_K_WALL,
  NB_HS_BARRIER_N, NB_HS_ETA, NB_HS_J_MAX_CAP, NB_HS_LAMBDA, NB_HS_POLICY /*0 soft barrier, 1 reflection (fold), 2 none*/,
  NB_HS_THETA_IMP, NB_HS_THETA_CAP, NB_HS_CHI_PI, NB_HS_OMEGA_SPR0, NB_HS_S0,
  NB_HS_FLAGS /*bit mask NB_HS_FLAG_*, stored as a double*/,
  NB_HS_NPARAM
};
/* reference test hooks: SimConfig.freeze_s_subsystem (hamsoft_stepper.py:119-124, 592-600: no S half-flow, no pi half
 * kick -- epsilon and pi stay frozen) and cfg._validate_S_only (hamsoft_stepper.py:270-284,
 * hamiltonian_softening_integrator.py:804-835: a macro step is ONE sub-step S S between the folds, no V, no T) */
#define NB_HS_FLAG_FREEZE_S 1
#define NB_HS_FLAG_S_ONLY 2

const char* nb_last_error(void);
int nb_version(void);

/* ---- a1-a5: forces.py:63-75 gravitational_force (+ simulation.py:551-552 division by m_i),
 *      forces.py:77-112 dV_d_epsilon, potential.py:23-64 softened_potential, geometry_cache.py:24-39.
 *      acc[B][N][2], U[B], dVdeps[B]; any output pointer may be NULL.  eps[B], G scalar. */
int nb_pair_batched_f64(const double* q, const double* m, const double* eps, double G, int B, int N,
                        double* acc, double* U, double* dVdeps, void* stream);

/* ---- a6: tangent_map.py:21-59 TangentMap.variational_accel; s2[B] = manager.step_s2 */
int nb_variational_batched_f64(const double* q, const double* m, const double* s2, const double* dr,
                               double G, int B, int N, double* da, void* stream);

/* ---- a3/a8/a9 construction-time work, one thread per system:
 *      COM removal, corrector half kicks (ctor and snapshot), the frozen sub-step schedule
 *      h_sub_ref (timestep_manager.py:139-253) and the 25 static features.
 *      v is updated in place; h_sub_ref[B] and n_sub[B] (for step size dt) are outputs. */
int nb_ensemble_prepare_f64(const double* m, const double* q, double* v, const double* eps, double G,
                            int B, int N, int mode, unsigned flags, double kick_dt, double sched_dt,
                            double dt, int split_n_max, double* h_sub_ref, int32_t* n_sub,
                            double* static_features, void* stream);

/* ---- a7/a8/a10/a17: the persistent ensemble integrator, one thread per system, state in registers.
 *      Runs n_steps macro steps of size dt (each n_sub[s] sub-steps; integrator.py:78-104), sampling
 *      step_metrics every sample_interval steps (0 = never; stability_analyzer.py:115-127), then n_megno
 *      further steps with the tangent map (evolution_features.py:34-66; raw_dr/raw_dv are the two
 *      randn(N,2) draws per system).  perm[B] (optional) maps thread -> system so that callers can sort
 *      by n_sub; n_heavy (optional, device int32 = workspace[64] of nb_sort_by_nsub) says how many leading
 *      entries of perm are sub-step-heavy: those run on the latency-optimised mappings (one body per lane;
 *      one pair per lane for N >= 5).
 *      q, v are advanced in place when NB_RUN_WRITE_STATE, NB_RUN_ENERGY or n_megno > 0 is requested.
 *      ham_soft additionally takes eps_pi[B][2] (epsilon, pi; in/out) and hs_params. */
int nb_ensemble_run_f64(const double* m, double* q, double* v, const double* eps, double G, int B, int N,
                        int mode, unsigned flags, double dt, int n_steps, int sample_interval, int n_megno,
                        const int32_t* n_sub, const int32_t* perm, const int32_t* n_heavy,
                        const double* raw_dr, const double* raw_dv,
                        double* eps_pi, const double* hs_params,
                        double* dyn_features, int32_t* status, void* stream);

/* the same call, instrumented for the roofline figures.  t_main (device uint64[2], may be NULL; the
 *      caller initialises it to {UINT64_MAX, 0}) receives {earliest start, latest end} of the MAIN-phase kernels' CTAs in
 *      %globaltimer nanoseconds, published by the kernels themselves (one atomic per CTA / warp), so a benchmark can
 *      time the dominant kernels inside its timed loop without adding stream operations.  work[B][2] (may be NULL)
 *      receives, per system,
 *      whfast  : {Newton iterations summed over all Kepler solves, number of Kepler solves}   (kepler_solver.py:48-91)
 *      ham_soft: {Jacobi sweeps of _solve_hi summed over the 4N+1 evaluations of every S half-flow, number of S half-flows}
 *                (hamsoft_eps_model.py:316-400); verlet / yoshida4 leave it untouched (their work is n_sub x n_steps) */
int nb_ensemble_run_counted_f64(const double* m, double* q, double* v, const double* eps, double G, int B, int N,
                                int mode, unsigned flags, double dt, int n_steps, int sample_interval, int n_megno,
                                const int32_t* n_sub, const int32_t* perm, const int32_t* n_heavy,
                                const double* raw_dr, const double* raw_dv,
                                double* eps_pi, const double* hs_params,
                                double* dyn_features, int32_t* status, double* work, uint64_t* t_main,
                                void* stream);

/* ---- classic ADAPTIVE softening (adaptive_softening=True with verlet / yoshida4): n_steps macro steps in which the
 *      softening is re-derived from the minimum separation after every sub-step (integrator.py:126-136, 204-225;
 *      SofteningManager.softening_from_min_sep / refresh_softening / _compute_energy_correction,
 *      softening_manager.py:298-336, 423-471, 541-547).  eps[B] = manager.s (in/out); soft_par[B][3] = {s0,
 *      min_softening, softening_scale}; energy_delta[B] = sim.softening_energy_delta (in/out, may be NULL);
 *      eps_hist[B][n_steps] = softening after each macro step (may be NULL).  q, v advance in place. */
int nb_ensemble_run_adaptive_f64(const double* m, double* q, double* v, double* eps, const double* soft_par, double G,
                                 int B, int N, int mode, double dt, int n_steps, const int32_t* n_sub, double k_wall,
                                 int barrier_exponent, double* energy_delta, double* eps_hist, int32_t* status,
                                 void* stream);

/* ---- StabilityAnalyzer.run_stability_analysis (stability_analyzer.py:69-259) for adaptive-softening copies: E0,
 *      n_steps adaptive macro steps with step_metrics sampling, E1, n_megno tangent-map steps, drifts and is_stable.
 *      eps[B] = the softening the restored copy starts from (the original's sim._epsilon, simulation.py:473-482; in/out);
 *      eps_energy[B] = the CONSTANT epsilon of the reference's energy diagnostics (diagnostics.py:474);
 *      the variational equations use the current manager.step_s2 (tangent_map.py:21-59). */
int nb_ensemble_analyze_adaptive_f64(const double* m, double* q, double* v, double* eps, const double* eps_energy,
                                     const double* soft_par, double G, int B, int N, int mode, double dt, int n_steps,
                                     int sample_interval, int n_megno, const int32_t* n_sub, const double* raw_dr,
                                     const double* raw_dv, double k_wall, int barrier_exponent, double* energy_delta,
                                     double* dyn_features, int32_t* status, void* stream);

/* ---- a11/a14 ham_soft construction-time calibration, one thread per system, in place on hs_params / eps_pi:
 *      flags bit0: EpsilonModel.calibrate_from_initial_conditions (hamsoft_eps_model.py:645-729: alpha_run,
 *                  eps_min, eps0) + _calibrate_mu_from_timescales (hamiltonian_softening_integrator.py:251-296);
 *                  on entry ALPHA_RUN holds cfg.alpha, EPS_MIN/EPS_MAX the simulation.py:88-114 defaults
 *      flags bit1: _freeze_production_schedule(dt) (:986-1221) -> n_sub[B] */
int nb_hamsoft_setup_f64(const double* m, const double* q, double G, int B, int N, unsigned flags, double dt,
                         double* hs_params, double* eps_pi, int32_t* n_sub, void* stream);

/* ---- a14 parity tap: eps*(q), H_ext (diagnostics.py:457-549), fallback flag and grad eps* for B systems;
 *      out[B][3+2N] = {eps*, H_ext, used_analytic_fallback, g_0x, g_0y, ...} (hamsoft_eps_model.py:94-234) */
int nb_hamsoft_probe_f64(const double* m, const double* q, const double* v, double G, int B, int N,
                         const double* eps_pi, const double* hs_params, double* out, void* stream);

/* counting sort of systems by n_sub (descending) -> perm[B]; workspace: 128 int32 on the device;
 * on return workspace[64] = number of systems in the "heavy" head of perm (n_sub > workspace[65]) that the run
 * kernels map for latency instead of throughput.  heavy_threshold = -1 (the product setting): automatic, a function of
 * N only (bodies per system, 0 = unknown; max(4, 50 / measured chain speed-up of the latency mapping)) and never of the
 * batch, so a system is integrated by the same arithmetic however the ensemble is sharded; 0..63 fixes it (tests,
 * tuning sweeps).  No process-wide state: every call carries its own setting. */
int nb_sort_by_nsub(const int32_t* n_sub, int B, int N, int32_t* perm, int32_t* workspace, int heavy_threshold,
                    void* stream);

/* ---- the same path end to end with HOST buffers (BatchStabilityAnalyzer.analyze_batch,
 *      batch_stability_analyzer.py:62-80): prepare(flags) -> sort -> run -> features.
 *      v_host is updated with the kicked velocities (the reference mutates the caller's sims). */
int nb_ensemble_analyze_host(const double* m, const double* q, double* v, const double* eps, double G,
                             int B, int N, int mode, unsigned prep_flags, double kick_dt, double sched_dt,
                             double dt, int n_steps, int n_megno, int split_n_max, const double* raw_dr,
                             const double* raw_dv, double* dyn_features, double* static_features,
                             int32_t* n_sub_out, int32_t* status, int device);

/* asynchronous form: enqueues everything on workspace slot `slot` (0..15, each with its own copy-in / compute /
 * copy-out streams and device buffers) and returns; host buffers must stay alive (and should be pinned) until
 * nb_host_sync(slot).  Several buckets -- and consecutive batches -- in flight on different slots overlap their copies
 * and kernels.  The caller's current CUDA device is left unchanged. */
int nb_ensemble_analyze_host_async(const double* m, const double* q, double* v, const double* eps, double G,
                             int B, int N, int mode, unsigned prep_flags, double kick_dt, double sched_dt,
                             double dt, int n_steps, int n_megno, int split_n_max, const double* raw_dr,
                             const double* raw_dv, double* dyn_features, double* static_features,
                             int32_t* n_sub_out, int32_t* status, int device, int slot);

/* extended form: the same call with options.  This is the entry point for the reference's DEFAULT integrator mode
 * (sim_config.py:38 integrator_mode = "ham_soft", what ml_training_pipeline.py:77-84, 211 runs) and for classic
 * adaptive softening:
 *   mode == NB_MODE_HAMSOFT : eps[B] is the constructor softening; the call performs the constructor work of
 *       hamiltonian_softening_integrator.py:47-141 (COM removal if NB_PREP_REMOVE_COM, SimConfig-default parameters
 *       unless opts->hs_params, calibration unless NB_HOST_HS_NO_CALIBRATE, frozen schedule for sched_dt and again for
 *       dt when they differ by > 1 %) and then run_stability_analysis.  opts->eps_pi (in/out) overrides the start state.
 *   NB_HOST_ADAPTIVE (verlet / yoshida4): classic adaptive softening, opts->soft_par[B][3] = {s0, min_softening,
 *       softening_scale}; opts->eps_start / eps_energy default to eps (softening_manager.py:298-336, 423-471, 541-547).
 *   NB_HOST_COMPACT_DYN   : dyn_features is [B][NB_N_DYN_USER] (no E0/E1/L0/L1/t_end taps): 23 % fewer D2H bytes
 *   NB_HOST_KEEP_V        : do not write the kicked velocities back into v (the reference's snapshot() mutation)
 *   NB_HOST_DEVICE_TANGENT: draw the MEGNO tangent vectors on the device (Philox4x32-10 keyed by opts->tangent_seed,
 *       counted by opts->first_index + system index) instead of reading raw_dr / raw_dv: 45 % fewer H2D bytes.  The
 *       host-draw form stays the one that reproduces the reference's np.random.randn stream (evolution_features.py:37-44).
 *   opts->n_chunks: cut the bucket into chunks that pipeline H2D -> kernels -> D2H inside the call (0 / 1 = off; pays
 *       for buckets without sub-step-heavy tails).  Results do not depend on it (bit-identical). */
#define NB_HOST_COMPACT_DYN 1u
#define NB_HOST_KEEP_V 2u
#define NB_HOST_DEVICE_TANGENT 4u
#define NB_HOST_ADAPTIVE 8u
#define NB_HOST_HS_NO_CALIBRATE 16u
typedef struct nb_host_opts {
  uint32_t size;             /* sizeof(nb_host_opts) of the caller (versioning) */
  uint32_t flags;            /* NB_HOST_* */
  int32_t n_chunks;          /* 0 or 1 = one piece */
  int32_t barrier_exponent;  /* adaptive: SimConfig.barrier_exponent (0 = 5) */
  uint64_t tangent_seed;     /* NB_HOST_DEVICE_TANGENT */
  uint64_t first_index;      /* global index of system 0 (sharded ensembles draw the same tangents as unsharded) */
  const double* hs_params;   /* ham_soft: HOST [B][NB_HS_NPARAM], NULL = SimConfig defaults */
  double* eps_pi;            /* ham_soft: HOST [B][2] (epsilon, pi) in/out, NULL = (max(s0, eps_min), 0), not returned */
  const double* soft_par;    /* adaptive: HOST [B][3] */
  const double* eps_start;   /* adaptive: HOST [B] softening the restored copy starts from (NULL = eps) */
  const double* eps_energy;  /* adaptive: HOST [B] constant epsilon of the energy diagnostics (NULL = eps) */
  double* energy_delta;      /* adaptive: HOST [B] out, softening_energy_delta (may be NULL) */
  double k_wall;             /* adaptive: SimConfig.k_wall (0 = 1e9) */
} nb_host_opts;
int nb_ensemble_analyze_host_ex(const double* m, const double* q, double* v, const double* eps, double G,
                             int B, int N, int mode, unsigned prep_flags, double kick_dt, double sched_dt,
                             double dt, int n_steps, int n_megno, int split_n_max, const double* raw_dr,
                             const double* raw_dv, double* dyn_features, double* static_features,
                             int32_t* n_sub_out, int32_t* status, int device, int slot, const nb_host_opts* opts);
int nb_host_sync(int slot);

/* ---- large-N direct sum (new capability, same formula as forces.py:63-75 / 77-112 / potential.py:23-64),
 *      fp32 pair arithmetic, fp64 accumulation across j-tiles.  xym[n_total] = (x, y, m, 0) packed float4.
 *      Rank-local i-range [i0, i0+ni).  acc[ni] float2; sums[2] += {sum_{i in range, j} m_i m_j/rho, sum m_i m_j/rho^3}
 *      (ordered pairs i != j; halve for i<j).  workspace: ni x 2 doubles of DEVICE memory owned by the caller (the fp64
 *      accumulators of this call; concurrent calls pass distinct workspaces).  variant: -1 = default (10: packed f32x2
 *      over j-pairs, 2 i per thread); 0..7 scalar-fp32 kernel (bit0: TMA staging, bits1-2: 4/2/1 i-particles per thread);
 *      8 / 9: packed f32x2 with 4 / 8 i per thread -- A-B tests only, the results agree to fp32 rounding. */
int nb_largeN_accel_f32(const float* xym, int n_total, int i0, int ni, float eps, float G, float* acc,
                        double* sums, double* workspace, int variant, void* stream);
/* kick/drift on the rank-local block and re-pack into the gather buffer (fused pack for the all-gather) */
int nb_largeN_kick_drift_f32(float* xym_local, float* vel, const float* acc, int ni, float kick_h, float drift_h,
                             void* stream);

/* ---- the O(N^2) reductions of the ham_soft epsilon flow for one large-N system (same tile pipeline as the
 *      force kernel; fp32 pair arithmetic, fp64 accumulation; rank-local i-range [i0, i0+ni)):
 *      NB_LN_DENSITY : out[ni][2] = sum_{j!=i} m_j e^{-r^2/h_i^2} {1, r^2}; iparam[ni] = h_i
 *                      (EpsilonModel._solve_hi sweep, hamsoft_eps_model.py:316-400; Sigma and dSigma/dh of
 *                      _production_grad :451-556)
 *      NB_LN_EPSGRAD : out[ni][2] = sum_{j!=i} (q_i-q_j)[A_i m_j e^{-r^2/h_i^2} + A_j m_i e^{-r^2/h_j^2}];
 *                      jaux[round_up(n_total,2)][2] = (-log2(e)/h_j^2, A_j) for ALL particles
 *                      (_production_grad's scatter loop in gather form)
 *      NB_LN_UNITGRAD: out[ni][2] = sum_{j!=i} (q_i-q_j)/r^3 (direction of softening.py:86-131 grad_eps_target)
 *      NB_LN_TAUMIN  : out[ni]    = min_{j!=i} (r^2+eps^2)^{3/2}/(m_i+m_j)
 *                      (tau_grav^2 G, hamiltonian_softening_integrator.py:251-296) */
#define NB_LN_DENSITY 0
#define NB_LN_EPSGRAD 1
#define NB_LN_UNITGRAD 2
#define NB_LN_TAUMIN 3
int nb_largeN_pass_f32(int kind, const float* xym, const float* jaux, int n_total, int i0, int ni,
                       const float* iparam, float eps, double* out, const float* tile_boxes, void* stream);
/* locality culling of the DENSITY / EPSGRAD passes: e^{-r^2/h^2} is evaluated as ex2.approx.ftz and is EXACTLY zero
 * beyond r > 9.35 h, so whole (i-block, j-tile) pairs whose bounding boxes are farther apart than that contribute exact
 * zeros and are skipped when tile_boxes (NB_LN_TILE-particle tiles, NB_LN_BOX_FLOATS floats each: xmin, ymin, xmax, ymax,
 * min |jaux.x|, 3 pad; written by nb_largeN_tile_boxes_f32 for the CURRENT xym / jaux) is passed; NULL = every tile.
 * Bit-identical either way; with a spatially ordered particle array the passes cost O(N x neighbours), not O(N^2). */
#define NB_LN_TILE 512
#define NB_LN_BOX_FLOATS 8
int nb_largeN_tile_boxes_f32(const float* xym, const float* jaux, int n_total, float* tile_boxes, void* stream);

/* ---- on-GPU initial conditions with a counter-based RNG (Philox4x32-10 keyed by seed, counted by the GLOBAL system
 *      index first_index + b): the distributions of InitialConditionGenerator.generate_single
 *      (initial_condition_generator.py:49-104), SpecializedGenerators (specialized_generators.py:23-94) and the cohort
 *      parameters of MLTrainingPipeline.generate_diverse_dataset (ml_training_pipeline.py:44-122).
 *      cohort: 0 random virial, 1 hierarchical triple (N = 3), 2 equal-mass polygon, 3 close encounter,
 *              4 planetary resonant chain, 5 planetary TTV.  Outputs m[B][N], q[B][N][2], v[B][N][2], eps[B]. */
#define NB_GEN_RANDOM 0
#define NB_GEN_HIERARCHICAL 1
#define NB_GEN_POLYGON 2
#define NB_GEN_CLOSE 3
#define NB_GEN_PLANETARY 4
#define NB_GEN_PLANETARY_TTV 5
int nb_generate_ensemble_f64(int cohort, int N, int B, uint64_t seed, uint64_t first_index, double* m, double* q, double* v,
                             double* eps, void* stream);
/* the two randn(N, 2) draws per system of EvolutionFeatures.compute_megno (evolution_features.py:37-44) from the same
 * counter-based generator: dr[B][N][2], dv[B][N][2] ~ N(0, 1), a function of (seed, first_index + system) only */
int nb_generate_tangent_f64(int N, int B, uint64_t seed, uint64_t first_index, double* dr, double* dv, void* stream);

/* ---- stability-classifier inference on the feature tensors (model_zoo.py:18-33 MLP F-128-64-1 with ReLU;
 *      train_mlp.py:141-217 sigmoid + threshold; stability_dataset.py:83-85 nan_to_num; StandardScaler).
 *      feature_index[F]: value c < 63 reads dyn_features[:, c], c >= 64 reads static_features[:, c - 64], and
 *      c == NB_MLP_COL_PATHOLOGICAL (63) is the dataset's derived column pathological_energy = |energy_drift| > 10
 *      (batch_stability_analyzer.py:45-52), which the feature tables a reference-trained model saw contain; F <= 64.
 *      w1[F][128] and w2[128][64] are INPUT-major (the transpose of torch's nn.Linear.weight); prob[B], label[B]. */
#define NB_MLP_COL_PATHOLOGICAL 63
int nb_mlp_classify_f32(const double* dyn_features, const double* static_features, const int32_t* feature_index, int F,
                        const float* mean, const float* inv_scale, const float* w1, const float* b1, const float* w2,
                        const float* b2, const float* w3, float b3, float threshold, int B, float* prob, int32_t* label,
                        void* stream);

/* ---- register-resident FMA micro-benchmarks used for the roofline denominators (TFLOP/s) */
int nb_peak_flops(int which /*0 fp64 DFMA, 1 fp32 FFMA, 2 fp32x2 FFMA2*/, int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
