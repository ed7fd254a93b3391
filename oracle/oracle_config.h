/* oracle_config.h — TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * One flat, C-ABI description of a reference `mci::MCI` run, shared by
 *   - oracle/ref_harness.cpp  : drives the UNMODIFIED reference (compiled from /root/reference into oracle/_ref/)
 *   - oracle/mci_oracle.c     : the plain-C restatement of the same algorithm (travels to the GPU box)
 * so that both can be called with the same struct from tests/ (ctypes) and compared bit for bit.
 *
 * Ids name the reference's own fixtures (test/common/TestMCIFunctions.hpp, examples/common/ExampleFunctions.hpp).
 */
#ifndef MCI_ORACLE_CONFIG_H
#define MCI_ORACLE_CONFIG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAXDIM 1024
#define ORC_MAXOBS 8
#define ORC_MAXTYPES 8
#define ORC_MAXOBSDIM 2048

/* sampling functions (reference file:line) */
enum {
    ORC_PDF_NONE = 0,
    ORC_PDF_GAUSS3D = 1, /* ThreeDimGaussianPDF  test/common/TestMCIFunctions.hpp:123-148 */
    ORC_PDF_GAUSS = 2,   /* Gauss(ndim)          :151-187 (has updatedAcceptance) */
    ORC_PDF_EXP1D = 3,   /* Exp1DPDF             :189-215 */
    ORC_PDF_EXPND = 4,   /* ExpNDPDF(ndim)       :217-256 (has updatedAcceptance) */
    ORC_PDF_NORMLINE = 5 /* NormalizedLine       examples/common/ExampleFunctions.hpp:54-82 (1-D, 0.2|x|, ratio acceptance) */
};

/* observables */
enum {
    ORC_OBS_XSQUARED = 1,     /* XSquared      TestMCIFunctions.hpp:260-275 (ndim 3, nobs 1) */
    ORC_OBS_GAUSSXSQUARED = 2,/* GaussXSquared :278-296 */
    ORC_OBS_XYZSQUARED = 3,   /* XYZSquared    :299-317 (nobs 3) */
    ORC_OBS_X1D = 4,          /* X1D           :320-336 */
    ORC_OBS_XND = 5,          /* XND(ndim)     :339-355 */
    ORC_OBS_UPDXND = 6,       /* UpdateableXND :357-379 (updateable) */
    ORC_OBS_CONSTVAL = 7,     /* Constval      :382-398 */
    ORC_OBS_POLYNOM = 8,      /* Polynom       :401-420 */
    ORC_OBS_X2SUM = 9,        /* X2Sum         :423-442 */
    ORC_OBS_X2 = 10,          /* X2(ndim)      :445-473 (updateable) */
    ORC_OBS_PARABOLA = 11,    /* Parabola           examples/common/ExampleFunctions.hpp:11-29 (1-D 4x-x^2) */
    ORC_OBS_NORMPARABOLA = 12,/* NormalizedParabola examples/common/ExampleFunctions.hpp:32-50 */
    ORC_OBS_DEPENDENT = 13    /* reference harness only (oracle/ref_harness.cpp: HarnessDepObs, a DependentObservableInterface): goldens */
};

/* trial moves: include/mci/Factories.hpp:108-145 */
enum { ORC_MOVE_ALL = 0, ORC_MOVE_VEC = 1, ORC_MOVE_MULTISTEP = 2,
       ORC_MOVE_USER_DRIFT = 3 /* reference harness only: a user-defined TrialMoveInterface subclass (oracle/ref_harness.cpp: HarnessDriftMove), goldens for
                                  the device-side move functor; srrd_par[0] = drift */ };
/* SRRD distribution of the move, include/mci/Factories.hpp:119-133. The C restatement covers 0 and 1; 2..9 (Student, Cauchy,
   Exponential, Gamma, Weibull, Lognormal, Chisq, Fisher) are available through the reference harness only (goldens). */
enum { ORC_SRRD_UNIFORM = 0, ORC_SRRD_GAUSSIAN = 1 };
/* estimator types: include/mci/Factories.hpp:52-59 */
enum { ORC_EST_NOOP = 0, ORC_EST_UNCORRELATED = 1, ORC_EST_CORRELATED = 2, ORC_EST_FCBLOCKER = 3, ORC_EST_MJBLOCKER = 4 };
enum { ORC_DOMAIN_UNBOUND = 0, ORC_DOMAIN_ORTHO = 1 };

typedef struct {
    int32_t obs_id;
    int32_t blocksize;  /* 0 Simple, 1 Full, >1 Block   (Factories.hpp:29-43) */
    int32_t nskip;
    int32_t flag_equil;
    int32_t estim_type; /* ORC_EST_* */
} orc_obs_t;

typedef struct {
    int32_t ndim;
    uint64_t seed;
    /* sampling function (0 or 1 pdf; reference containers allow more, fixtures never use >1) */
    int32_t pdf_id;
    /* move */
    int32_t move_type; /* ORC_MOVE_* */
    int32_t srrd;      /* ORC_SRRD_* */
    int32_t veclen;    /* for VEC (and the MultiStep sub-move, which is always a uniform vec move here) */
    int32_t ntypes;
    int32_t type_ends[ORC_MAXTYPES];
    double steps[ORC_MAXTYPES];
    /* MultiStepMove: include/mci/MultiStepMove.hpp:46-52, src/MultiStepMove.cpp:6-47 */
    int32_t ms_nsteps;
    int32_t ms_sub_pdf_id; /* ORC_PDF_* (0 = none: every sub-step accepted, accept draw still consumed) */
    /* domain */
    int32_t domain; /* ORC_DOMAIN_* */
    double lb[ORC_MAXDIM];
    double ub[ORC_MAXDIM];
    /* start */
    double x0[ORC_MAXDIM];
    /* observables */
    int32_t nobs;
    orc_obs_t obs[ORC_MAXOBS];
    /* automatic routines: src/MCIntegrator.cpp:637-639 */
    int32_t nfind;    /* _NfindMRT2Iterations */
    int64_t ndecorr;  /* _NdecorrelationSteps */
    double target_acc;
    /* integrate(Nmc, avg, err, do_find, do_decorr) */
    int64_t nmc;
    int32_t do_find;
    int32_t do_decorr;
    /* emulate R independent MPI ranks (seed r = seeds[r]) combined as src/MPIMCI.cpp:85-92; 0/1 = single chain.
       nranks_for_minstat enters MIN_STAT/MIN_NMC (src/MCIntegrator.cpp:107,193) */
    int32_t nranks_for_minstat;
    /* parameters of the move's distribution (a pre-made distribution passed to the move's constructor, include/mci/SRRDAllMove.hpp:45-58,
       test/ut5/main.cpp:110-113); 0 = createSymRRD<>() defaults. Gaussian: stddev; Student: n; Cauchy: b; Exponential: lambda; Gamma: alpha, beta;
       Weibull: a, b; Lognormal: m, s; Chisq: n; Fisher: m, n. Reference harness only (the C restatement covers the default distributions). */
    int32_t srrd_npar;
    double srrd_par[2];
} orc_config_t;

typedef struct {
    int32_t nobsdim;
    double avg[ORC_MAXOBSDIM];
    double err[ORC_MAXOBSDIM];
    double acc_rate;           /* getAcceptanceRate() after integrate */
    double x_final[ORC_MAXDIM];
    double steps_final[ORC_MAXTYPES];
    int64_t n_acc;             /* _acc of the last sample() run */
    int64_t n_rej;
} orc_result_t;

/* optional per-step trace of the MAIN sampling run (the sample(Nmc, obscont, true) call) */
typedef struct {
    int64_t cap_steps;       /* capacity (in steps) of accepted[]; capacity of draws[] is cap_draws */
    int64_t cap_draws;
    uint8_t* accepted;       /* [nmc] accept decision of every Metropolis step */
    double* draws;           /* distribution OUTPUTS in consumption order (Appendix A of SURVEY.md):
                                uniform_real(-1,1) values, uniform_int values (as double), uniform_real(0,1) accept draws */
    int64_t n_steps;         /* filled */
    int64_t n_draws;         /* filled */
} orc_trace_t;

#ifdef __cplusplus
}
#endif
#endif
