/* mcig.h — C-ABI of the B200-native Monte Carlo integration engine (libmcig.so).
 *
 * This is the drop-in boundary for the sampling path of DCM-UPB/MCIntegratorPlusPlus. Every entry point names the
 * reference interface it replaces (paths relative to the reference repository). The C++ facade in include/mci/ is
 * implemented on top of these calls, so reference-style programs recompile against it unchanged.
 *
 * Conventions: plain pointers and sizes only; all pointers are caller-owned; a context owns its device memory; calls
 * are blocking; a context is single-threaded like the reference's MCI object. Functions returning int yield 0 on
 * success or an MCIG_ERR_* code, with the message available from mcig_last_error(); the codes map one-to-one onto
 * the std exceptions the reference throws (SURVEY.md §8b "Error conventions").
 *
 * There is NO CPU fallback: every call that samples or estimates fails with MCIG_ERR_CUDA if no sm_100 device or no
 * CUDA driver is available.
 */
#ifndef MCIG_H
#define MCIG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcig_ctx mcig_ctx;

enum {
    MCIG_OK = 0,
    MCIG_ERR_INVALID_ARGUMENT = 1, /* std::invalid_argument */
    MCIG_ERR_DOMAIN = 2,           /* std::domain_error */
    MCIG_ERR_RUNTIME = 3,          /* std::runtime_error */
    MCIG_ERR_CUDA = 4              /* CUDA / NVRTC failure (no reference analogue) */
};

/* include/mci/Factories.hpp:108-145 */
enum { MCIG_MOVE_ALL = 0, MCIG_MOVE_VEC = 1, MCIG_MOVE_MULTISTEP = 2 };
/* SRRDType, include/mci/Factories.hpp:119-133 (default parameters of createSymRRD<>, include/mci/TrialMoveInterface.hpp:113-187) */
enum {
    MCIG_SRRD_UNIFORM = 0, MCIG_SRRD_GAUSSIAN = 1, MCIG_SRRD_STUDENT = 2, MCIG_SRRD_CAUCHY = 3, MCIG_SRRD_EXPONENTIAL = 4,
    MCIG_SRRD_GAMMA = 5, MCIG_SRRD_WEIBULL = 6, MCIG_SRRD_LOGNORMAL = 7, MCIG_SRRD_CHISQ = 8, MCIG_SRRD_FISHER = 9,
    MCIG_SRRD_USER = 10 /* internal: the proposal comes from a user-defined move functor (mcig_set_move_plugin) */
};
/* include/mci/Factories.hpp:52-59 */
enum { MCIG_EST_NOOP = 0, MCIG_EST_UNCORRELATED = 1, MCIG_EST_CORRELATED = 2, MCIG_EST_FCBLOCKER = 3, MCIG_EST_MJBLOCKER = 4 };
/* random number source of the walk kernel */
enum {
    MCIG_RNG_PHILOX32 = 0, /* Philox4x32-10, one 32-bit word per uniform (production default) */
    MCIG_RNG_PHILOX53 = 1, /* Philox4x32-10, 52-bit uniforms */
    MCIG_RNG_REPLAY = 2    /* per-walker std::mt19937_64 + libstdc++ distributions generated on the host and consumed by
                              the kernel in the reference's order: bit-exact reference trajectories (parity mode) */
};
enum { MCIG_PLUGIN_PDF = 0, MCIG_PLUGIN_OBS = 1, MCIG_PLUGIN_CALLBACK = 2, MCIG_PLUGIN_DOMAIN = 3, MCIG_PLUGIN_MOVE = 4 };
/* plugin flags (sampling functions) */
enum {
    MCIG_PLUGIN_HAS_UPDATE = 1,     /* functor overrides updatedAcceptance (selective update for single-vector moves) */
    MCIG_PLUGIN_ELEMENTWISE = 2,    /* proto value k depends on x[k] only; updatedAcceptance touches protonew[changedIdx] only, by plain assignment /
                                     * reads (with HAS_UPDATE the new values of a selective update may live in registers, not in an array) */
    MCIG_PLUGIN_LOG_ACCEPTANCE = 4, /* functor provides logAcceptance(protoold, protonew) = log(acceptanceFunction) */
    MCIG_PLUGIN_PROTO_ELEMENT = 16, /* sampling function (with ELEMENTWISE | HAS_UPDATE) provides protoElement(x_k) = proto value k */
    MCIG_PLUGIN_SUM_ACCEPTANCE = 32,/* sampling function (with PROTO_ELEMENT): acceptanceFunction(po, pn) = exp(sum_k po[k] - sum_k pn[k]), both sums in index
                                     * order from 0 (Gauss, ExpNDPDF: test/common/TestMCIFunctions.hpp:173-177, 240-245). All-moves over >= 64 coordinates then
                                     * spread one walker over several lanes of a warp, each summing its share (state placement 3) */
    MCIG_PLUGIN_DEPENDENT = 8       /* observable: observableFunction(x, out, dep) with dep.proto(i) / dep.obs(k, j) (DependentObservableInterface) */
};

const char * mcig_last_error(void);
int mcig_version(void);
/* number of usable CUDA devices (0 when there is no driver / GPU) */
int mcig_device_count(void);

/* ---- plugin registry: ProtoFunctionInterface / SamplingFunctionInterface / ObservableFunctionInterface subclasses
 *      (include/mci/SamplingFunctionInterface.hpp:36-105, ObservableFunctionInterface.hpp:30-63) re-expressed as
 *      __device__ functors. `source` is CUDA C++ pasted into the JIT translation unit (may be NULL for built-ins),
 *      `type_expr` the functor type with "{ndim}" substituted (e.g. "mcig_builtin::Gauss<{ndim}>").
 *      ndim == 0: any dimension; nvalues == 0: nproto / nobs equals ndim; flags: MCIG_PLUGIN_* bits.
 *      Returns the plugin id (>= 0) or -(error code). */
int mcig_register_plugin(int kind, const char * name, const char * type_expr, const char * source, int ndim, int nvalues,
                         int npar, int flags);
/* id of a registered plugin (the reference's fixtures are pre-registered under their class names), or -1 */
int mcig_lookup_plugin(int kind, const char * name);

/* ---- construction: MCI::MCI(int ndim)  src/MCIntegrator.cpp:624-649 (same defaults) */
mcig_ctx * mcig_create(int ndim);
void mcig_destroy(mcig_ctx * ctx);
int mcig_set_device(mcig_ctx * ctx, int device);

/* ---- MCI::setSeed  src/MCIntegrator.cpp:547-550. Philox modes: the key. Replay mode: walker w is seeded with seed
 *      unless per-walker seeds are given (MPIMCI::setSeed, src/MPIMCI.cpp:38-65: rank r <- seed file entry offset+r). */
int mcig_set_seed(mcig_ctx * ctx, uint64_t seed);
int mcig_set_walker_seeds(mcig_ctx * ctx, const uint64_t * seeds, int64_t n);
/* Philox modes: position of every walker's stream, in draw groups (one group = the uniforms of one proposal + its accept test; the
 * counter-based generator is random-access). integrate advances it; mcig_set_seed resets it to 0. Checkpoint / resume: save the
 * positions (mcig_get_x), the step sizes and this counter. */
int mcig_set_stream_position(mcig_ctx * ctx, uint64_t group);
uint64_t mcig_get_stream_position(mcig_ctx * ctx);
int mcig_set_rng_mode(mcig_ctx * ctx, int mode);

/* ---- walkers: one reference MCI = one chain; many chains exist there only as MPI ranks (src/MPIMCI.cpp:83).
 *      Here one context runs W chains ("virtual ranks"). global_offset / total describe the shard of a multi-GPU job. */
int mcig_set_walkers(mcig_ctx * ctx, int64_t nwalkers, int64_t global_offset, int64_t total_walkers);
int64_t mcig_get_walkers(mcig_ctx * ctx);

/* ---- positions: MCI::setX / getX / centerX / newRandomX / moveX  src/MCIntegrator.cpp:594-619 */
int mcig_set_x(mcig_ctx * ctx, const double * x /*[ndim], broadcast to all walkers*/);
int mcig_set_x_walkers(mcig_ctx * ctx, const double * x /*[nwalkers][ndim]*/);
int mcig_get_x(mcig_ctx * ctx, int64_t walker, double * x /*[ndim]*/);

/* ---- domain: MCI::resetDomain / setIRange  src/MCIntegrator.cpp:390-408 */
int mcig_set_domain_unbound(mcig_ctx * ctx);
int mcig_set_domain_ortho(mcig_ctx * ctx, const double * lbounds, const double * ubounds);
/* MCI::setDomain(const DomainInterface &) with a user-defined domain (the reference's DomainInterface is user-subclassable,
 * include/mci/DomainInterface.hpp:26-55): a device functor of plugin kind MCIG_PLUGIN_DOMAIN,
 *     struct MyDomain { static constexpr int NPAR = ...; const double * par;
 *         __device__ void wrap(int i, double & xi) const;        // applyDomain, coordinate by coordinate
 *         __device__ double scale(int i, double u01) const; };   // scaleToDomain, coordinate by coordinate
 * plus what the host needs from getSizes / getVolume: the lengths of the ndim dimensions (infinite: 2 x float max; they cap the calibrated
 * step sizes, src/MCIntegrator.cpp:151-155) and the volume (0 = infinite; plain sampling without a sampling function needs a finite one).
 * Separable domains only: coordinate i is mapped knowing only coordinate i (reflecting walls, per-coordinate periods, ...). Positions handed
 * to mcig_set_x* are taken as given: the caller applies the domain (the C++ facade calls the host class's applyDomain). */
int mcig_set_domain_plugin(mcig_ctx * ctx, int plugin_id, const double * par, int npar, const double * dim_sizes /*[ndim]*/, double volume);

/* ---- trial move: MCI::setTrialMove(MoveType) / (SRRDType, veclen, ntypes, typeEnds)  src/MCIntegrator.cpp:423-446.
 *      Step sizes are reset to DEFAULT_MRT2STEP = 0.05 (Factories.hpp:145). MultiStep: sub-move = uniform single-vector
 *      move of length veclen, nsteps <= 0 means ndim (MultiStepMove.hpp:46-52); its own sampling functions are added
 *      with mcig_multistep_add_pdf (MultiStepMove::addSamplingFunction). */
int mcig_set_move(mcig_ctx * ctx, int move_type, int srrd, int veclen, int ntypes, const int * type_ends);
/* A move built around a pre-made distribution instead of createSymRRD<>()'s default (the `const SRRD * rdist` argument of the reference's move
 * constructors, include/mci/SRRDAllMove.hpp:45-58, SRRDVecMove.hpp:41-68; test/ut5/main.cpp:110-113 passes student_t(2)). Call after mcig_set_move
 * (which resets to the defaults). par: Gaussian {stddev}; Student {n}; Cauchy {b}; Exponential {lambda}; Gamma {alpha, beta}; Weibull {a, b};
 * Lognormal {m, s}; Chisq {n}; Fisher {m, n}; npar = 0 restores the defaults. Locations stay 0 (a move distribution must be symmetric).
 * Replay mode consumes the libstdc++ outputs of exactly that distribution; the Philox modes sample the same law from a fixed number of uniforms
 * per value: closed forms, for Gamma / Chisq / Fisher shapes that are multiples of 1/2 sums of exponentials, for any other shape Marsaglia & Tsang's
 * test with 6 tries per value (all of them fail with probability < 2e-8; the proposal stays symmetric, so the chain stays exact). */
int mcig_set_srrd_params(mcig_ctx * ctx, int npar, const double * par);
/* MCI::setTrialMove(const TrialMoveInterface &) with a user-defined move (the reference's TrialMoveInterface is user-subclassable,
 * include/mci/TrialMoveInterface.hpp:16-70): a device functor of plugin kind MCIG_PLUGIN_MOVE, registered with nvalues = the number of uniforms in
 * [0,1) it consumes per step (0 = one per coordinate),
 *     struct MyMove { static constexpr int NPAR = ...; const double * par;
 *         template <class XO, class XN, class T, class U>
 *         __device__ double trialMove(const XO & xold, XN & xnew, const double * steps, T typeOf, const U & u) const; };
 * fills xnew[0 .. ndim) from xold, the typed step sizes steps[typeOf.of(i)] and u(0) .. u(nvalues - 1), and returns the move's acceptance factor
 * (1 for symmetric proposals). The engine then applies the domain, evaluates the sampling functions on the whole proposal and accepts when the
 * accept uniform <= pdf acceptance * that factor (src/MCIntegrator.cpp:329-343). Typed step sizes as in mcig_set_move (calibrated by findMRT2Step).
 * Replay mode feeds the outputs of std::uniform_real_distribution<double>(0,1) on the walker's std::mt19937_64, in order. */
int mcig_set_move_plugin(mcig_ctx * ctx, int plugin_id, const double * par, int npar, int ntypes, const int * type_ends);
int mcig_multistep_config(mcig_ctx * ctx, int nsteps);
int mcig_multistep_add_pdf(mcig_ctx * ctx, int plugin_id, const double * par, int npar);
/* MCI::setMRT2Step / getMRT2Step  src/MCIntegrator.cpp:557-585 */
int mcig_get_nsteps_sizes(mcig_ctx * ctx);
int mcig_set_step(mcig_ctx * ctx, int i, double step);
double mcig_get_step(mcig_ctx * ctx, int i);

/* ---- sampling functions / observables: MCI::addSamplingFunction, addObservable(obs, blocksize, nskip, flag_equil,
 *      EstimatorType), pop*, clear*  src/MCIntegrator.cpp:451-490 */
int mcig_add_pdf(mcig_ctx * ctx, int plugin_id, const double * par, int npar);
int mcig_pop_pdf(mcig_ctx * ctx);
int mcig_clear_pdfs(mcig_ctx * ctx);
int mcig_add_obs(mcig_ctx * ctx, int plugin_id, const double * par, int npar, int blocksize, int nskip, int flag_equil, int estim_type);
int mcig_pop_obs(mcig_ctx * ctx);
int mcig_clear_obs(mcig_ctx * ctx);
int mcig_get_nobsdim(mcig_ctx * ctx);

/* ---- MCI::setCallback / clearCallback  include/mci/MCIntegrator.hpp:186-190, called at src/MCIntegrator.cpp:346 and :374 after every
 *      move (in every sampling run: calibration, decorrelation, main), after the accept decision and before the state is committed.
 *      The reference's std::function<void(const MCI &)> runs host code inside the step loop; here the callback is a __device__
 *      functor (plugin kind MCIG_PLUGIN_CALLBACK, nvalues ignored) invoked by every walker:
 *          template <class XO, class XN> __device__ void operator()(const XO & xold, const XN & xnew, bool accepted,
 *                                                                   long long walker (global id), long long step, double * buffer) const;
 *      `buffer` has buffer_doubles doubles of device memory owned by the integrator, zeroed at the start of every integrate call and
 *      readable afterwards with mcig_get_callback_buffer. Walkers run concurrently: index the buffer by walker or use atomicAdd. */
int mcig_set_callback(mcig_ctx * ctx, int plugin_id, const double * par, int npar, int64_t buffer_doubles);
int mcig_clear_callback(mcig_ctx * ctx);
int mcig_get_callback_buffer(mcig_ctx * ctx, double * out, int64_t n);

/* ---- automatic routines: setTargetAcceptanceRate / setNfindMRT2Iterations / setNdecorrelationSteps
 *      include/mci/MCIntegrator.hpp:113-123 (N<0 auto with max |N|, 0 off, N>0 fixed) */
int mcig_set_autotune(mcig_ctx * ctx, int nfind_iterations, int64_t ndecorrelation_steps, double target_acceptance);

/* ---- optional cross-process sum (the MPI_Allreduce(SUM) of src/MCIntegrator.cpp:21-34, :135 and src/MPIMCI.cpp:85-87).
 *      When set, the engine calls it for the acceptance rate during calibration, for the estimates during automatic
 *      decorrelation and for the final [sum avg | sum err^2] so that a job sharded over several processes/GPUs behaves
 *      like one MPI job. buf is a host array of n doubles, summed in place over all processes. */
typedef void (*mcig_allreduce_fn)(double * buf, int n, void * user);
int mcig_set_allreduce(mcig_ctx * ctx, mcig_allreduce_fn fn, void * user);

/* ---- the collective inside the library: MPIMCI::init / myrank / size / finalize (src/MPIMCI.cpp:27-35, 95-104) for one process per GPU.
 *      The communicator is process-global like MPI_COMM_WORLD (NCCL over NVLink / NVSwitch, libnccl.so.2 loaded on first use). Contexts that
 *      are attached to it (mcig_attach_comm) run every MPI_Allreduce of the reference path — the rank-averaged acceptance rate of findMRT2Step
 *      (src/MCIntegrator.cpp:131-138), the equilibration estimates (:21-34) and the final [sum avg | sum err^2] (src/MPIMCI.cpp:85-87) — as
 *      ncclAllReduce on the engine's own stream over device memory, inside the device-resident control loops: no host hop per iteration.
 *      mcig_comm_init_env: rank / size / device from RANK, WORLD_SIZE, LOCAL_RANK (or OMPI_COMM_WORLD_*), the ncclUniqueId travels over TCP
 *      (MASTER_ADDR, MASTER_PORT + 17 or MCIG_COMM_PORT); a process started without WORLD_SIZE is a world of one and never loads NCCL.
 *      mcig_comm_get_unique_id + mcig_comm_init_rank: the caller transports the 128-byte id itself (bench.py: torch.distributed broadcast).
 *      Returns: init_env the rank (>= 0) or -(error code); the others 0 or an error code. */
int mcig_comm_init_env(void);
int mcig_comm_get_unique_id(void * id128);
int mcig_comm_init_rank(const void * id128, int rank, int nranks, int device);
int mcig_comm_rank(void);
int mcig_comm_size(void);
int mcig_comm_finalize(void);
/* on = 1: this context is one rank's shard of a job over mcig_comm_size() processes; its device follows the communicator's. The walker
 * shard itself is set with mcig_set_walkers(ctx, n_local, global_offset, total). */
int mcig_attach_comm(mcig_ctx * ctx, int on);

/* ---- MCI::integrate(Nmc, average, error, doFindMRT2step, doDecorrelation)  src/MCIntegrator.cpp:43-82, combined over
 *      walkers as MPIMCI::integrate does over ranks (src/MPIMCI.cpp:85-92): avg = sum_w avg_w / W, err = sqrt(sum_w err_w^2) / W */
int mcig_integrate(mcig_ctx * ctx, int64_t nmc, double * average, double * error, int do_find_mrt2_step, int do_decorrelation);
/* MCI::getAcceptanceRate  src/MCIntegrator.cpp:587-592 (of the last sampling run, over all walkers) */
double mcig_get_acceptance_rate(mcig_ctx * ctx);

/* ---- results of the last integrate beyond the reference's API */
int mcig_get_walker_results(mcig_ctx * ctx, double * avg /*[nobsdim][nwalkers]*/, double * err /*same*/);
/* local sums [sum_w avg_w | sum_w err_w^2 | sum_w avg_w^2], 3*nobsdim doubles: the all-reduce payload of a sharded job */
int mcig_get_sums(mcig_ctx * ctx, double * sums, int capacity /* doubles the buffer holds: >= 3 * mcig_get_result_nobsdim */);
/* number of observable components of the LAST integrate (the current configuration may have changed since: popObservable) */
int mcig_get_result_nobsdim(mcig_ctx * ctx);
/* standard error of the mean over walkers, sqrt((<a^2>-<a>^2)/(W-1)) (local walkers) */
int mcig_get_cross_walker_error(mcig_ctx * ctx, double * err, int capacity /* >= mcig_get_result_nobsdim */);
/* stored samples of observable iobs of the last integrate (Block/Full accumulators), host order [nstore][nobs] of one walker */
int64_t mcig_get_nstore(mcig_ctx * ctx, int iobs);
int mcig_get_obs_data(mcig_ctx * ctx, int iobs, int64_t walker, double * data);
/* The reference frees every accumulator at the end of integrate (src/MCIntegrator.cpp:79); this engine keeps what it stored until the next
 * call, but a Block / Full accumulator whose estimator is the one-pass uncorrelated one (sum x, sum x^2: src/Estimators.cpp:36-56, 125-155)
 * with at most 8 components does not store at all by default: the walker keeps both sums in registers, nothing is written to or re-read
 * from HBM, and the results have the reference's summation order bit for bit. on = 1 keeps the series as well (for mcig_get_obs_data). */
int mcig_set_keep_samples(mcig_ctx * ctx, int on);
/* device times of the last integrate in ms (CUDA events on the engine's stream): the walk kernel of the main sampling run,
 * the estimation stage, the whole call (calibration + decorrelation + sampling + estimation); and the kernels it launched */
int mcig_get_timings(mcig_ctx * ctx, double * walk_ms, double * estim_ms, double * total_ms, int64_t * kernel_launches);
/* host-clock phases of the last integrate in ms: findMRT2Step, initialDecorrelation, and run-time compilation / module loads
 * (first use of a configuration; contained in whichever phase triggered them) */
int mcig_get_phase_timings(mcig_ctx * ctx, double * find_ms, double * decorr_ms, double * jit_ms);

/* ---- estimators on host data (include/mci/Estimators.hpp:9-45): data x[n][ndim], run on the device */
int mcig_estimate(int estim_type, int64_t n, int ndim, const double * x, double * average, double * error);
/* One/MultiDimBlockEstimator(n, [ndim,] x, nblocks, avg, err)  include/mci/Estimators.hpp:12, :30; src/Estimators.cpp:59-80, 158-185:
 * means of nblocks consecutive blocks of n/nblocks samples (remainder ignored), then the uncorrelated estimator over them */
int mcig_estimate_blocks(int64_t n, int ndim, const double * x, int64_t nblocks, double * average, double * error);

/* ---- engine knobs without reference analogue */
int mcig_set_block_size(mcig_ctx * ctx, int threads_per_block); /* 0 = automatic */
int mcig_set_state_placement(mcig_ctx * ctx, int placement);    /* -1 auto, 0 registers, 1 shared memory, 2 global memory (auto picks 2 when a warp of walkers exceeds 227 KiB),
                                                                  * 3 registers of several lanes per walker (uniform all-moves with a SUM_ACCEPTANCE sampling function and
                                                                  * element-wise observables; auto picks it from 16 coordinates on) */
/* Element-wise observables (plugin flag MCIG_PLUGIN_ELEMENTWISE, e.g. XND, X2) in Simple / Block accumulators under single-vector
 * moves: add value x dwell time when a coordinate changes instead of every component at every step. 1 (default): in the Philox
 * modes, replay mode keeps the reference's summation order; 2: in every mode; 0: never. Sums differ by rounding only. */
int mcig_set_lazy_accumulation(mcig_ctx * ctx, int on);
/* findMRT2Step and initialDecorrelation feedback loops on the device (default 1: sampling launch -> reduction / estimators -> [ncclAllReduce
 * when attached to the communicator] -> controller kernel, enqueued in batches without a host synchronisation per iteration; Philox modes)
 * or on the host (0: one readback per iteration; always used in replay mode and with a host mcig_set_allreduce callback).
 * Same arithmetic, same results. */
int mcig_set_device_calibration(mcig_ctx * ctx, int on);
int mcig_get_calibration_iterations(mcig_ctx * ctx); /* iterations the last findMRT2Step executed */
/* Launches the last main sampling run was split into: 0 = one launch with every stored series resident in HBM; n > 0 = the series did not fit
 * (include/mci/FullAccumulator.hpp:11-13) and was sampled, staged in a double buffer and folded into the estimators' state chunk by chunk
 * (MJBlocker, uncorrelated and Noop estimators; Philox modes, register-resident walkers; anything else fails loudly with the size). */
int64_t mcig_get_staging_chunks(mcig_ctx * ctx);
int mcig_get_decorrelation_chunks(mcig_ctx * ctx);   /* MIN_NMC-step chunks the last automatic initialDecorrelation sampled (src/MCIntegrator.cpp:193-241) */
/* MCI::storeObservablesOnFile (what = 0) / storeWalkerPositionsOnFile (what = 1), src/MCIntegrator.cpp:495-542: text dump of
 * walker 0 every freq-th step of the main sampling run ("ridx v0 v1 ..."), written after the run from device-side shadow
 * accumulators. path NULL or "" switches the dump off (clearObservableFile / clearWalkerFile). */
int mcig_store_on_file(mcig_ctx * ctx, int what, const char * path, int freq);
/* rounds of the Philox4x32 generator: 10 (default, the standard Philox4x32-10) down to 7 (the smallest variant that passes
 * BigCrush according to Salmon et al., SC'11; +13..20 % throughput on RNG-bound integrands) */
int mcig_set_philox_rounds(mcig_ctx * ctx, int rounds);
/* persistent walk kernel that work-steals chunks of steps (balances W that does not fill the SMs evenly): -1 auto, 0 off, 1 on */
int mcig_set_dynamic_scheduling(mcig_ctx * ctx, int mode);
/* compile (JIT) the kernels the current configuration needs without running them; works without a GPU */
int mcig_prebuild(mcig_ctx * ctx);
/* mcig_prebuild plus a rehearsal of the coming integrate(nmc, ., ., do_find, do_decorr) whose every trace is undone (positions, random streams, step
 * sizes, statistics and results restored, no files written): all device buffers of that call exist at their final size and every kernel has run once,
 * so the first real call runs at steady-state speed (tests/test_warm_start.py). Sharded jobs: every rank calls it. */
int mcig_warmup(mcig_ctx * ctx, int64_t nmc, int do_find, int do_decorr);
/* generated CUDA source of the main walk kernel (for inspection); returns bytes needed */
int64_t mcig_get_kernel_source(mcig_ctx * ctx, char * buf, int64_t cap);
/* issue-rate microbenchmarks (roofline denominators): DFMA/s and IMAD/s of the current device */
int mcig_measure_peaks(int device, double * dfma_per_s, double * imad_per_s);
/* Philox4x32-10 blocks per second of the current device when nothing else is issued (the RNG's own issue-rate bound of the walk loop) */
int mcig_measure_philox_peak(int device, double * blocks_per_s);

#ifdef __cplusplus
}
#endif
#endif
