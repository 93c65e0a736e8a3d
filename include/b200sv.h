/* b200sv.h -- C ABI of libb200sv.so: the B200 (sm_100a) replacement for the two numerical hot paths of
 * ArturSepp/StochVolModels (stochvolmodels 2.2.0).  Plain pointers and sizes; no torch / C++ types.
 *
 * Each entry point names the reference interface it replaces (paths relative to
 * /root/reference/src/stochvolmodels).  The reference is pure Python + Numba, so "the FFI for this path" is a
 * ctypes binding: INTEGRATION.md shows the stub a maintainer adds inside LogSVPricer / HestonPricer.
 *
 * Conventions
 *   - return 0 on success, <0 on error; b200sv_last_error() returns a thread-local message.
 *   - arrays are caller-owned; no pointer is retained after return.
 *   - "host-level" calls (no _dev_ in the name) take HOST pointers, run on the current CUDA device, and are
 *     synchronous: H2D copies of the inputs and D2H copies of the results happen inside the call.
 *   - "device-level" calls (b200sv_dev_*) take DEVICE pointers and a cudaStream_t (as void*; NULL = the legacy
 *     default stream), enqueue work and return without synchronising.  They are what a multi-GPU host composes with
 *     its own collectives (two tiny all-reduces per maturity, see DESIGN.md "multi-GPU").
 *   - option chain layout ("CSR"): strikes[offsets[m] .. offsets[m+1]) belong to maturity m;
 *     types[] uses B200SV_CALL.. codes in the same positions.
 *   - all floating point crossing the ABI is IEEE float64, as in the reference.
 */
#ifndef B200SV_H
#define B200SV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SV_VERSION 100 /* 0.1.0 */

/* option payoff codes: utils/config.py:8-14 OptionType {'C','P','IC','IP'} */
enum { B200SV_CALL = 0, B200SV_PUT = 1, B200SV_INV_CALL = 2, B200SV_INV_PUT = 3 };
/* utils/config.py:17-24 VariableType */
enum { B200SV_LOG_RETURN = 1, B200SV_Q_VAR = 2, B200SV_SIGMA = 3 };
/* pricers/logsv/affine_expansion.py:43-54 ExpansionOrder */
enum { B200SV_ORDER_FIRST = 1, B200SV_ORDER_SECOND = 2 };

/* MC arithmetic flags (bit-or).  Default 0 = fp64 state arithmetic, Gaussians drawn by a float Box-Muller on the SFU
 * and widened (the draws are random inputs, not reference arithmetic).  B200SV_GAUSS_F64 draws them in fp64 as well;
 * B200SV_STATE_F32 keeps x / log-sigma / qvar in float registers (payoff moments stay fp64). */
enum { B200SV_STATE_F64 = 0, B200SV_STATE_F32 = 1, B200SV_GAUSS_F32 = 0, B200SV_GAUSS_F64 = 2,
       /* check mode: the DEFAULT stream's 32-bit uniforms pushed through the fp64 Box-Muller (paired with the default mode path by path,
        * so that price differences isolate the SFU approximation error of the float draws; tests/test_gpu_gauss_bias.py) */
       B200SV_GAUSS_F64_PAIRED = 4 };
/* Heston variance scheme: 0 = the reference's floor-Euler (pricers/heston_pricer.py:369-379, v = max(v, 1e-4)); 1 = opt-in Andersen (2008)
 * quadratic-exponential scheme with central discretisation (BASELINE.json config 2 names it; the reference does not have it) */
enum { B200SV_HESTON_EULER_FLOOR = 0, B200SV_HESTON_QE = 1 };

/* pricers/logsv/logsv_params.py:35-83 LogSvParams (the six model floats; kappa2=None already mapped to kappa1/theta) */
typedef struct {
  double sigma0, theta, kappa1, kappa2, beta, volvol;
} b200sv_logsv_params;

/* pricers/heston_pricer.py:27-41 HestonParams */
typedef struct {
  double v0, theta, kappa, rho, volvol;
} b200sv_heston_params;

/* pricers/hawkes_jd_pricer.py:41-65 HawkesJDParams (the sixteen model floats, in the order of the dataclass with lambda_p / lambda_m moved next
 * to their intensities' parameters as listed here) */
typedef struct {
  double mu, sigma, shift_p, mean_p, shift_m, mean_m;
  double lambda_p, theta_p, kappa_p, beta1_p, beta2_p;
  double lambda_m, theta_m, kappa_m, beta1_m, beta2_m;
} b200sv_hawkes_params;

const char* b200sv_last_error(void);
/* CUDA stream (cudaStream_t as void*) on which the calling thread's HOST-LEVEL calls enqueue their copies and kernels and on which they
 * synchronise before returning; NULL (the default) = the legacy default stream.  Thread-local, like the error string: calls from different
 * threads on different streams run concurrently.  (SURVEY.md 8b asked for a stream on every call; the device-level b200sv_dev_* calls take it
 * as an argument, the host-level ones -- what a ctypes / cgo / JNI binding of the reference's functions binds -- take it from here so that
 * their signatures stay the reference's argument lists.) */
int b200sv_set_stream(void* stream);
void* b200sv_get_stream(void);
int b200sv_version(void);
/* number of this library's kernels launched by the calling thread since the last reset (bench "gpu_launches") */
long long b200sv_launch_count(void);
void b200sv_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Monte Carlo, host-level
 * ------------------------------------------------------------------------------------------------------------------ */

/* replaces logsv_mc_chain_pricer (pricers/logsv_pricer.py:806-867) = chain loop over
 * simulate_logsv_x_vol_terminal (:950-1047) + compute_mc_vars_payoff (utils/mc_payoffs.py:10-88).
 * nb_steps_per_year follows set_time_grid (utils/funcs.py:24-47): S_m = int((ttm_m - ttm_{m-1}) * n) + 1.
 * Single-GPU, synchronous.  etas may be NULL (all 1).  prices_out, stderr_out: offsets[M]-offsets[0] doubles each. */
int b200sv_logsv_mc_chain(const b200sv_logsv_params* params, int M, const double* ttms, const double* forwards,
                          const double* discfactors, const double* etas, const int* offsets, const double* strikes,
                          const int8_t* types, long long nb_path, int nb_steps_per_year, int is_spot_measure,
                          int variable_type, uint64_t seed, int flags, double* prices_out, double* stderr_out);

/* replaces heston_mc_chain_pricer (pricers/heston_pricer.py:285-331) + simulate_heston_x_vol_terminal (:334-381).
 * nb_steps_per_year: the reference hard-wires 360 (:344). */
int b200sv_heston_mc_chain(const b200sv_heston_params* params, int M, const double* ttms, const double* forwards,
                           const double* discfactors, const int* offsets, const double* strikes, const int8_t* types,
                           long long nb_path, int nb_steps_per_year, int variable_type, uint64_t seed, int flags,
                           int scheme, double* prices_out, double* stderr_out);

/* B parameter sets through the fused chain Monte Carlo in ONE call, all on the SAME Philox seed (common random numbers): the objective of
 * CalibrationEngine.MC (pricers/logsv_pricer.py:251-266 -- there: pre-drawn fixed normals W0s/W1s re-used across optimizer iterations;
 * here the counter-based generator re-draws the identical normals for every set, so nothing is stored or streamed).  params[B],
 * etas[B*M] (NULL = 1); prices_out / stderr_out / ivols_out are [B][J] in chain order (ivols_out may be NULL; Black-76 inversion fused into
 * the finalisation kernel).  LOG_RETURN payoffs.  Row b equals b200sv_logsv_mc_chain(params[b], seed) bit for bit. */
int b200sv_logsv_mc_chain_batch(const b200sv_logsv_params* params, int B, int M, const double* ttms, const double* forwards,
                                const double* discfactors, const double* etas, const int* offsets, const double* strikes,
                                const int8_t* types, long long nb_path, int nb_steps_per_year, int is_spot_measure, uint64_t seed,
                                int flags, double* prices_out, double* stderr_out, double* ivols_out);
int b200sv_heston_mc_chain_batch(const b200sv_heston_params* params, int B, int M, const double* ttms, const double* forwards,
                                 const double* discfactors, const int* offsets, const double* strikes, const int8_t* types,
                                 long long nb_path, int nb_steps_per_year, uint64_t seed, int flags, int scheme, double* prices_out,
                                 double* stderr_out, double* ivols_out);

/* replaces LogSVPricer.simulate_terminal_values (pricers/logsv_pricer.py:590-611) -> simulate_logsv_x_vol_terminal.
 * x, sigma, qvar: nb_path doubles each (host, out). */
int b200sv_logsv_terminal(const b200sv_logsv_params* params, double ttm, long long nb_path, int nb_steps_per_year,
                          int is_spot_measure, double eta, uint64_t seed, int flags, double* x, double* sigma,
                          double* qvar);

/* replaces HestonPricer.simulate_terminal_values (pricers/heston_pricer.py:90-108). */
/* the same with PER-PATH initial arrays (pricers/logsv_pricer.py:1007-1020 accepts length-nb_path x0 / sigma0 / qvar0): in/out host arrays;
 * slice_index = Philox sub-stream (0 for a fresh start; m for the m-th consecutive call on one seed, as the chain pricer numbers its slices) */
int b200sv_logsv_terminal_from_state(const b200sv_logsv_params* params, double ttm, long long nb_path, int nb_steps_per_year, int is_spot_measure,
                                     double eta, uint64_t seed, int flags, int slice_index, double* x_inout, double* sigma_inout,
                                     double* qvar_inout);
int b200sv_heston_terminal(const b200sv_heston_params* params, double ttm, long long nb_path, int nb_steps_per_year,
                           uint64_t seed, int flags, int scheme, double* x, double* var, double* qvar);

/* replaces simulate_logsv_x_vol_terminal called with W0, W1, dt (pricers/logsv_pricer.py:1027-1047), the stepping
 * inside logsv_mc_chain_pricer_fixed_randoms (:1100-1162).  W0, W1: unit normals, row-major [S][N] (scaled by
 * sqrt(dt) inside, :1028-1030).  x, sigma, qvar: N doubles, updated in place.  Strict fp64, no FMA contraction,
 * reference evaluation order. */
int b200sv_logsv_step_fixed(double* x, double* sigma, double* qvar, const double* W0, const double* W1, int S,
                            long long N, double dt, const b200sv_logsv_params* params, double eta, int is_spot_measure);

/* Heston twin of the above (loop body pricers/heston_pricer.py:372-379). */
int b200sv_heston_step_fixed(double* x, double* var, double* qvar, const double* W0, const double* W1, int S,
                             long long N, double dt, const b200sv_heston_params* params);

/* replaces simulate_vol_paths (pricers/logsv_pricer.py:870-947; LogSVPricer.simulate_vol_paths :561-587): the whole volatility
 * path matrix sigma_t_out[(S+1)][nb_path] (row 0 = sigma0), S = int(ttm*n)+1.  brownians: optional host array [S][nb_path] of
 * SCALED increments (the reference's `brownians`); NULL draws sqrt(dt)*Z from the device Philox stream.  Strict fp64. */
int b200sv_logsv_vol_paths(const b200sv_logsv_params* params, double ttm, long long nb_path, int nb_steps_per_year,
                           int is_spot_measure, uint64_t seed, const double* brownians, double* sigma_t_out);

/* replaces compute_mc_vars_payoff (utils/mc_payoffs.py:10-88): forward re-centring with the nan-mean over all paths,
 * per-strike discounted nan-mean and nan-std/sqrt(N).  qvar may be NULL for B200SV_LOG_RETURN. */
int b200sv_mc_payoffs(const double* x, const double* qvar, long long N, double ttm, double forward, const double* strikes,
                      const int8_t* types, int J, double discfactor, int variable_type, double* prices_out,
                      double* stderr_out);

/* parity/debug export of the device RNG stream (csrc/philox.cuh): unit normals of paths [path0, path0+n) for chain
 * slice `slice`, steps 0..nsteps-1; z0, z1 row-major [nsteps][n] host doubles. */
int b200sv_device_normals(uint64_t seed, long long path0, long long n, int slice, int nsteps, int flags, double* z0,
                          double* z1);

/* ------------------------------------------------------------------------------------------------------------------
 * Monte Carlo, device-level building blocks (multi-GPU hosts put their all-reduces between them)
 * ------------------------------------------------------------------------------------------------------------------ */

/* one maturity slice of the fused stepper for local paths [0, n_local) whose global ids start at path_offset.
 * state arrays are float64 (B200SV_STATE_F64) or float32 (B200SV_STATE_F32) device arrays of n_local elements;
 * init != 0: start from (0, v_init, 0) instead of loading.  moments_out[2] (device) receives
 * (sum over non-NaN paths of forward*exp(x), count of non-NaN paths) for THIS rank. */
int b200sv_dev_logsv_slice(void* x, void* sigma, void* qvar, long long n_local, long long path_offset, int init,
                           const b200sv_logsv_params* params, double eta, int is_spot_measure, int nsteps, double dt,
                           int slice_index, double forward, uint64_t seed, int flags, double* moments_out, void* p2p_ctx,
                           void* stream);
int b200sv_dev_heston_slice(void* x, void* var, void* qvar, long long n_local, long long path_offset, int init,
                            const b200sv_heston_params* params, int nsteps, double dt, int slice_index, double forward,
                            uint64_t seed, int flags, int scheme, double* moments_out, void* p2p_ctx, void* stream);

/* per-strike payoff sums for local paths given the GLOBAL (all-reduced) re-centring moments[2]:
 * sums_out[3*J] (device) = (sum pay, sum pay^2, count non-NaN) per strike for THIS rank.
 * strikes / types are DEVICE arrays of J entries.  payoff_kinds_hint: what the HOST knows about types[] -- bit 0: some
 * 'C'/'P', bit 1: some 'IC'/'IP'; 1 selects the branch-free vanilla kernel, anything else (0 = unknown) the general one. */
int b200sv_dev_payoff_sums(const void* x, const void* qvar, long long n_local, int flags, double ttm, double forward,
                           const double* strikes, const int8_t* types, int J, int variable_type, int payoff_kinds_hint,
                           const double* moments, double* sums_out, void* p2p_ctx, void* stream);

/* turn GLOBAL sums[3*J] into prices / std errors (device arrays of J): price = df*s1/n, se = df*sqrt(s2/n-(s1/n)^2)/sqrt(N). */
int b200sv_dev_payoff_finalize(const double* sums, int J, double discfactor, long long total_paths, double* prices_out,
                               double* stderr_out, void* p2p_ctx, void* stream);

/* Peer-memory exchange of the two per-maturity messages (csrc/p2p.cuh): with a non-NULL p2p_ctx the calls above fuse the exchange into
 * their kernels instead of leaving it to the host's collective -- *_slice and *_payoff_sums PUBLISH their local values into every
 * peer's mailbox over NVLink from the reduction kernel that produces them, *_payoff_sums and *_payoff_finalize GATHER the global values
 * (summed in rank order) in their prologue; `moments` / `sums` arguments then carry this rank's local values only.  All ranks must
 * issue the same sequence of calls (a publish whose predecessor was never gathered on this rank gets a flag-wait inserted, which keeps
 * the double-buffered mailbox race-free when a maturity has no strikes or a rank has no paths).  create: allocates this rank's mailbox (max_values >= 3 * max strikes per slice) and returns a
 * 64-byte CUDA IPC handle to all-gather among the ranks of the node; connect: maps the peers' mailboxes (handles in rank order). */
int b200sv_p2p_create(int world, int rank, int max_values, void** ctx_out, unsigned char* handle_out);
int b200sv_p2p_connect(void* p2p_ctx, const unsigned char* handles);
int b200sv_p2p_destroy(void* p2p_ctx);
/* A gather that waits longer than the spin limit (default 2^24 polls of ~200 ns, ~3 s) poisons its value with NaN and sets bit r (the
 * peer that never published) of the context's status word.  b200sv_p2p_status synchronises `stream`, returns the word and clears it: a
 * host must call it after the chain's copy back and treat non-zero as an error (multi_gpu.mc_chain_distributed raises P2pTimeout). */
int b200sv_p2p_set_spin_limit(void* p2p_ctx, unsigned int spins);
int b200sv_p2p_status(void* p2p_ctx, unsigned int* status_out, void* stream);
/* test hook: pretend a publish was issued that never arrives, so that the next gather runs into the spin limit */
int b200sv_debug_p2p_lose_publish(void* p2p_ctx);
/* standalone halves of an exchange (ranks without local paths; tests): publish K local values / gather the K global ones */
int b200sv_dev_p2p_publish(void* p2p_ctx, const double* vals, int K, void* stream);
int b200sv_dev_p2p_gather(void* p2p_ctx, int K, double* out, void* stream);

/* device-level twins of b200sv_logsv_step_fixed / b200sv_heston_step_fixed (all arrays on the device; W row-major [S][N]):
 * the calibration inner loop logsv_mc_chain_pricer_fixed_randoms (pricers/logsv_pricer.py:1100-1162) keeps W0s/W1s
 * resident in HBM across optimizer iterations and calls these per maturity.  fast != 0 selects the throughput variant of the
 * LogSV stepper (the fused kernel's per-step update: folded constants, FMA, shared-polynomial exp pair; agrees with the strict
 * kernel to ~1e-14), whose bound is the 16 B/path-step HBM stream of normals. */
int b200sv_dev_logsv_step_fixed(double* x, double* sigma, double* qvar, const double* W0, const double* W1, int S,
                                long long N, double dt, const b200sv_logsv_params* params, double eta,
                                int is_spot_measure, int fast, void* stream);
int b200sv_dev_heston_step_fixed(double* x, double* var, double* qvar, const double* W0, const double* W1, int S,
                                 long long N, double dt, const b200sv_heston_params* params, void* stream);

/* Rough-LogSV multi-factor Monte Carlo chain (Markovian lift of the rough kernel, 1 <= n_factors <= 8):
 * rough_logsv_mc_chain_pricer_fixed_randoms (pricers/logsv_pricer.py:1164-1232) -> log_spot_full_combined
 * (pricers/rough_logsv/split_simulation.py:466-479, f64 branch :340-361).  params[b].volvol carries the ORTHOGONAL vol-of-vol of set b
 * (the reference's `orthog_vol`); weights / nodes are [B][n_factors].  Every maturity m restarts at t = 0 and runs nsteps[m] steps of size
 * hs[m] (the reference's per-maturity `timegrids`), consuming rows [0, nsteps[m]) of the unit normals Z0 / Z1 (host, row-major
 * [z_rows][nb_path]); Z0 == Z1 == NULL draws them in-kernel from the Philox stream of csrc/philox.cuh (seed, flags = B200SV_GAUSS_*).
 * Outputs [B][J_total]: prices, "std errors" exactly as the reference returns them on this route -- discfactor * nanstd(payoff), NOT
 * divided by sqrt(nb_path) (it calls compute_mc_vars_payoff with (1, nb_path) arrays, utils/mc_payoffs.py:88) -- and optionally Black
 * implied vols (ivols_out may be NULL).  states_out (B == 1, may be NULL): [M][n_factors + 2][nb_path] = log-spot, factor values, quadratic
 * variance at each maturity. */
int b200sv_rough_logsv_mc_chain(const b200sv_logsv_params* params, int B, int n_factors, const double* weights, const double* nodes, int M,
                                const double* ttms, const double* forwards, const double* discfactors, const int* offsets, const double* strikes,
                                const int8_t* types, long long nb_path, const int* nsteps, const double* hs, const double* Z0, const double* Z1,
                                long long z_rows, int variable_type, uint64_t seed, int flags, double* prices_out, double* stderr_out,
                                double* ivols_out, double* states_out);

/* Hawkes jump-diffusion Monte Carlo (pricers/hawkes_jd_pricer.py): 1800 steps per year (:752), state (x, lambda_p, lambda_m).
 * b200sv_hawkesjd_mc_chain   = hawkesjd_mc_chain_pricer (:644-715): chained slices, forward-recentred payoffs on x; in-kernel draws (Philox;
 *                              flags = B200SV_GAUSS_* for the normals, the jump clocks and sizes come from 32-bit uniforms through fp64 logs).
 * b200sv_hawkesjd_terminal   = simulate_hawkesjd_terminal (:718-779) with in-kernel draws; use_initial_arrays != 0: the three host arrays hold the
 *                              per-path initial state (in/out), else every path starts from (0, lambda_p, lambda_m) and the arrays are outputs.
 * b200sv_hawkesjd_step_fixed = the same update on caller-supplied inputs in the reference's own form, row-major [S][N] host arrays:
 *                              W0 = sqrt(dt) z, U_P / U_M = -ln(U)/dt, J_P / J_M = shift +- mean Exp(1) (:753-757) -- reference evaluation order,
 *                              no FMA contraction (parity entry).
 * b200sv_hawkesjd_device_draws exports, in that same form, what the in-kernel generator draws for paths [path0, path0 + n) of a slice. */
int b200sv_hawkesjd_mc_chain(const b200sv_hawkes_params* params, int M, const double* ttms, const double* forwards, const double* discfactors,
                             const int* offsets, const double* strikes, const int8_t* types, long long nb_path, int variable_type, uint64_t seed,
                             int flags, double* prices_out, double* stderr_out);
/* Hawkes jump-diffusion, Fourier route.  Replaces hawkesjd_chain_pricer (pricers/hawkes_jd_pricer.py:365-417) and, with a finite
 * risk_premia_gamma, hawkesjd_chain_pricer_with_risk_premia (:420-484) incl. hawkesjd_forwards_under_risk_kernel (:487-515),
 * compute_hawkes_a_mgf_grid / solve_ode_for_a (:518-641: SciPy RK45 per transform point) and the slice pricers of utils/mgf_pricer.py
 * (:174-221, :273-320).  P <= 0: 500 grid points; vol_scaler <= 0: clip(sigma, 0.2, 0.5) sqrt(min(min ttm, 1/12)) (:360-362);
 * risk_premia_gamma = NaN: no risk kernel.  a_out [M][P][3] complex128, log_mgf_out [M][P] complex128, normalizers_out / gamma_forwards_out
 * [M] are optional. */
int b200sv_hawkesjd_price_chain(const b200sv_hawkes_params* params, int M, const double* ttms, const double* forwards, const double* discfactors,
                                const int* offsets, const double* strikes, const int8_t* types, int is_spot_measure, double vol_scaler, int P,
                                double risk_premia_gamma, double* prices_out, double* a_out, double* log_mgf_out, double* normalizers_out,
                                double* gamma_forwards_out);
/* compute_hawkes_a_mgf_grid / solve_a_ode_grid (pricers/hawkes_jd_pricer.py:518-579) on caller-supplied transform grids (complex128 as interleaved
 * doubles; psi may be NULL = zeros): a_inout [P][3] holds A(0) on entry and A(dtau) on return, log_mgf_out [P] = a0 + a_p lambda_p + a_m lambda_m */
int b200sv_hawkesjd_mgf_grid(const double* phi, const double* psi, int P, double dtau, double* a_inout, const b200sv_hawkes_params* params,
                             double* log_mgf_out);
/* slice_pricer_with_mgf_grid_with_gamma (utils/mgf_pricer.py:273-320) on caller-supplied grids (complex128 as interleaved doubles) */
int b200sv_fourier_gamma(const double* log_mgf, const double* phi, int P, double risk_premia_gamma, double forward, double normalizer,
                         double gamma_forward, const double* strikes, const int8_t* types, int J, int is_spot_measure, double* prices_out);
int b200sv_hawkesjd_terminal(const b200sv_hawkes_params* params, double ttm, long long nb_path, uint64_t seed, int flags, int slice_index,
                             int use_initial_arrays, double* x_inout, double* lambda_p_inout, double* lambda_m_inout);
int b200sv_hawkesjd_step_fixed(double* x, double* lambda_p, double* lambda_m, const double* W0, const double* U_P, const double* U_M,
                               const double* J_P, const double* J_M, int S, long long N, double dt, const b200sv_hawkes_params* params);
int b200sv_hawkesjd_device_draws(uint64_t seed, long long path0, long long n, int slice, int S, double dt, const b200sv_hawkes_params* params,
                                 int flags, double* W0, double* U_P, double* U_M, double* J_P, double* J_M);

/* moments_out[2] (device) = (sum over non-NaN paths of forward*exp(x), count) for externally produced float64 states. */
int b200sv_dev_spot_moments(const double* x, long long n, double forward, double* moments_out, void* stream);

/* Sharded slices of the two neighbouring Monte Carlo routes (SURVEY.md 8e applied to 8f #4), same contract as b200sv_dev_logsv_slice:
 * local paths [0, n_local) with global ids from path_offset, float64 device state, moments_out[2] (device) = this rank's re-centring
 * moments, published to the peer-memory mailbox when p2p_ctx is given; b200sv_dev_payoff_sums / _finalize complete the maturity.
 *  - Hawkes jump-diffusion (replaces the loop body of hawkes_jd_pricer.py:687-715): state (x, lambda_p, lambda_m); init != 0 starts from
 *    (0, params->lambda_p, params->lambda_m).
 *  - rough LogSV (logsv_pricer.py:1199-1216 -> rough_logsv/split_simulation.py:466): EVERY maturity restarts at t = 0 on its own grid
 *    (nsteps steps of size h) and draws from slice 0 of the path's stream; vol_factors is [n_factors][n_local] (may be NULL),
 *    weights / nodes are HOST arrays of n_factors entries.  The payoff uses (log_spot, qvar) and, as in the reference on this route,
 *    total_paths = 1 in b200sv_dev_payoff_finalize (its "standard errors" carry no 1/sqrt(N)). */
int b200sv_dev_hawkesjd_slice(double* x, double* lambda_p, double* lambda_m, long long n_local, long long path_offset, int init,
                              const b200sv_hawkes_params* params, int nsteps, double dt, int slice_index, double forward, uint64_t seed,
                              int flags, double* moments_out, void* p2p_ctx, void* stream);
int b200sv_dev_rough_logsv_slice(double* log_spot, double* vol_factors, double* qvar, long long n_local, long long path_offset,
                                 const b200sv_logsv_params* params, int n_factors, const double* weights, const double* nodes, int nsteps,
                                 double h, double forward, uint64_t seed, int flags, double* moments_out, void* p2p_ctx, void* stream);

/* test hook: out[2i] = exp(L[i]), out[2i+1] = exp(-L[i]) through the stepper's shared-polynomial exp pair (host arrays). */
int b200sv_debug_exp_pair(const double* L, long long n, double* out);
/* same for the variant the stepper runs on its table-unit state: out[2i] = exp(Ls[i] ln2/256), out[2i+1] = exp(-Ls[i] ln2/256) */
int b200sv_debug_exp_pair_scaled(const double* Ls, long long n, double* out);

/* ------------------------------------------------------------------------------------------------------------------
 * Fourier / MGF, host-level
 * ------------------------------------------------------------------------------------------------------------------ */

/* replaces logsv_chain_pricer (pricers/logsv_pricer.py:669-739), LOG_RETURN and Q_VAR branches: transform grid
 * (utils/mgf_pricer.py:11-34), per maturity the affine-expansion ODEs integrated with SciPy's RK45 control law
 * (pricers/logsv/affine_expansion.py:229-303, 492-529) carried across maturities, log-MGF contraction (:674-685)
 * and the Simpson/Fourier sums (utils/mgf_pricer.py:174-221).  vol_scaler <= 0 selects set_vol_scaler (:664-666).
 * variable_type = B200SV_Q_VAR prices calls on the annualised quadratic variance on the psi grid -0.5 + 1j*linspace(0, 4000, P)
 * (utils/mgf_pricer.py:37-47, :323-358; only 'C' payoffs, else -5) with phi = 0 (MMA) / 1 (inverse measure) (:79-85).
 * P = grid size, <= 0 selects the reference's (1000 for LOG_RETURN, 40000 for Q_VAR).  Optional outputs (may be NULL):
 * a_out [M][P][n] complex128 interleaved, log_mgf_out [M][P] complex128 interleaved. */
int b200sv_logsv_price_chain(const b200sv_logsv_params* params, int M, const double* ttms, const double* forwards,
                             const double* discfactors, const double* etas, const int* offsets, const double* strikes,
                             const int8_t* types, int is_spot_measure, int variable_type, int expansion_order,
                             double vol_scaler, int P, double* prices_out, double* a_out, double* log_mgf_out);

/* replaces heston_chain_pricer, LOG_RETURN and Q_VAR branches (pricers/heston_pricer.py:217-282) with compute_heston_mgf_grid
 * (:183-214).  vol_scaler <= 0 selects min(0.3, sqrt(v0*ttms[0])) (:234-235). */
int b200sv_heston_price_chain(const b200sv_heston_params* params, int M, const double* ttms, const double* forwards,
                              const double* discfactors, const int* offsets, const double* strikes, const int8_t* types,
                              int variable_type, double vol_scaler, int P, double* prices_out, double* log_mgf_out);

/* B parameter sets on ONE chain in one pass -- the objective of the calibration drivers and its finite-difference gradient:
 * _LogSvCalibrationObjective.__call__ (pricers/logsv_pricer.py:234-294: compute_model_ivols_for_chain = logsv_chain_pricer +
 * option_chain.compute_model_ivols_from_chain_data), which scipy SLSQP evaluates n+1 times per iteration, one parameter set at a time.
 * params[B]; etas[B*M] (NULL = 1); prices_out[B*J], ivols_out[B*J] (NULL = skip the fused Black-76 inversion), J = offsets[M]-offsets[0],
 * row b in chain order.  vol_scaler <= 0: each set gets its own default grid sigma0_b*sqrt(min(min ttm, 1/24)) (:664-666).
 * LOG_RETURN only.  Row b equals b200sv_logsv_price_chain(params[b]) (+ b200sv_bsm_implied_vols) bit for bit. */
int b200sv_logsv_price_chain_batch(const b200sv_logsv_params* params, int B, int M, const double* ttms, const double* forwards,
                                   const double* discfactors, const double* etas, const int* offsets, const double* strikes,
                                   const int8_t* types, int is_spot_measure, int expansion_order, double vol_scaler, int P,
                                   double* prices_out, double* ivols_out);

/* same for Heston: the objective of HestonPricer.calibrate_model_params_to_chain (pricers/heston_pricer.py:111-180). */
int b200sv_heston_price_chain_batch(const b200sv_heston_params* params, int B, int M, const double* ttms, const double* forwards,
                                    const double* discfactors, const int* offsets, const double* strikes, const int8_t* types,
                                    double vol_scaler, int P, double* prices_out, double* ivols_out);

/* replaces compute_logsv_a_mgf_grid, non-analytic branch (pricers/logsv/affine_expansion.py:570-685 -> solve_a_ode_grid
 * :492-529): phi, psi, a (in: A(0), out: A(dtau)), log_mgf_out are complex128 interleaved host arrays of P, P, P*n, P. */
int b200sv_logsv_mgf_grid(const double* phi, const double* psi, int P, double dtau, double* a_inout,
                          const b200sv_logsv_params* params, double eta, int is_spot_measure, int expansion_order,
                          double* log_mgf_out);

/* compute_logsv_a_mgf_grid(is_analytic=True): the semi-analytic branch solve_analytic_ode_for_a (pricers/logsv/affine_expansion.py:306-384) over a
 * transform grid -- business-day steps (year_days = 260 in the reference), linear part exact, quadratic part by 10 fixed-point sweeps.
 * a_inout complex [P][n] in/out, log_mgf_out complex [P]; vol_backbone_eta is ignored on this branch, as in the reference. */
int b200sv_logsv_mgf_grid_analytic(const double* phi, const double* psi, int P, double dtau, double* a_inout, const b200sv_logsv_params* params,
                                   int is_spot_measure, int expansion_order, int year_days, double* log_mgf_out);
/* compute_logsv_a_mgf_grid(is_stiff_solver=True): solve_ivp(method="BDF", jac=func_rhs_jac) per grid point (pricers/logsv/affine_expansion.py:229-303) --
 * a clone of SciPy's BDF control law (scipy 1.18.1 _ivp/bdf.py), default rtol 1e-3 / atol 1e-6.  Same arrays as b200sv_logsv_mgf_grid. */
int b200sv_logsv_mgf_grid_bdf(const double* phi, const double* psi, int P, double dtau, double* a_inout, const b200sv_logsv_params* params,
                              double eta, int is_spot_measure, int expansion_order, double* log_mgf_out);
/* func_a_ode_quadratic_terms (pricers/logsv/affine_expansion.py:67-184): the dense coefficient tensors of A' = A^T M^(k) A + L A + H for
 * P transform points, n = 3 (FIRST) or 5 (SECOND): M_out complex [P][n][n][n] (symmetric in the last two indices), L_out [P][n][n],
 * H_out [P][n], interleaved (re, im).  psi may be NULL (zeros).  Parity entry: the pricing kernels integrate the same rows in sparse form. */
int b200sv_logsv_ode_terms(const double* phi, const double* psi, int P, const b200sv_logsv_params* params, double eta, int is_spot_measure,
                           int expansion_order, double* M_out, double* L_out, double* H_out);
/* func_rhs (affine_expansion.py:187-205) evaluated by the production right-hand side at A [P][n] complex -> rhs_out [P][n] complex. */
int b200sv_logsv_ode_rhs(const double* phi, const double* psi, int P, const double* A, const b200sv_logsv_params* params, double eta,
                         int is_spot_measure, int expansion_order, double* rhs_out);
/* func_rhs with the reference's own signature (caller-supplied dense complex tensors M [n][n][n], L [n][n], H [n]; A [P][n]) */
int b200sv_ode_rhs_dense(const double* A, int P, int n, const double* M, const double* L, const double* H, double* rhs_out);

/* replaces compute_heston_mgf_grid (pricers/heston_pricer.py:183-214); a, b in/out complex128[P]. */
int b200sv_heston_mgf_grid(const double* phi, const double* psi, int P, double dtau, double* a_inout, double* b_inout,
                           const b200sv_heston_params* params, double* log_mgf_out);

/* replaces vanilla_slice_pricer_with_mgf_grid (utils/mgf_pricer.py:174-221). */
int b200sv_fourier_vanilla(const double* log_mgf, const double* phi, int P, double forward, const double* strikes,
                           const int8_t* types, int J, double discfactor, int is_spot_measure, double* prices_out);

/* replaces slice_qvar_pricer_with_a_grid (utils/mgf_pricer.py:323-358): calls on annualised quadratic variance from the log-MGF on
 * the psi grid; price = max(discfactor * S / ttm, 1e-10), S = nansum Re((dp/pi)/psi^2 * exp(K*ttm*psi + log_mgf)). */
int b200sv_fourier_qvar(const double* log_mgf, const double* psi, int P, double ttm, const double* strikes,
                        const int8_t* types, int J, double discfactor, double* prices_out);

/* the sum of pdf_with_mgf_grid (utils/mgf_pricer.py:361-384): out[j] = nansum Re((dp/pi) * exp(z[j]*grid + log_mgf)); the caller
 * applies z = (x - shift)/scale beforehand and the dx / scale factors afterwards (pricers/logsv_pricer.py:778-803). */
int b200sv_fourier_pdf(const double* log_mgf, const double* grid, int P, const double* z, int J, double* out);

/* replaces digital_slice_pricer_with_mgf_grid (utils/mgf_pricer.py:224-269). */
int b200sv_fourier_digital(const double* log_mgf, const double* phi, int P, double forward, const double* strikes,
                           const int8_t* types, int J, double discfactor, double* prices_out);

/* ------------------------------------------------------------------------------------------------------------------
 * Black-76 implied volatilities (the step after every chain pricer in the reference's callers)
 * ------------------------------------------------------------------------------------------------------------------ */

/* replaces option_chain.compute_model_ivols_from_chain_data (data/option_chain.py:327-346 -> third-party
 * vanilla_option_pricers.infer_bsm_ivols_from_model_chain_prices): prices[] and ivols_out[] are flat in chain order
 * (offsets[M]-offsets[0] entries).  'IC'/'IP' quotes are inverted like 'C'/'P'; prices outside the no-arbitrage bounds give NaN. */
int b200sv_bsm_implied_vols(int M, const double* ttms, const double* forwards, const double* discfactors, const int* offsets,
                            const double* strikes, const int8_t* types, const double* prices, double* ivols_out);

#ifdef __cplusplus
}
#endif
#endif /* B200SV_H */
