// mcig_functors.cuh — built-in sampling-function / observable functors: the reference's own fixtures re-expressed as
// __device__ functors (test/common/TestMCIFunctions.hpp:123-473, examples/common/ExampleFunctions.hpp:11-82).
//
// Plugin contract (what a user-supplied functor source must look like; it is pasted into the JIT translation unit):
//
//   struct MyPDF {                                  // mirrors mci::SamplingFunctionInterface
//       static constexpr int NPAR = 0;              // doubles of run-time parameters (passed as `par`)
//       static constexpr bool HAS_UPDATE = false;   // overrides updatedAcceptance? (else full recompute is used)
//       static constexpr bool ELEMENTWISE = false;  // proto value k depends on x[k] only and updatedAcceptance touches
//                                                   // protonew[changedIdx] only (lets the smem path commit O(nchanged))
//       const double * par;
//       template <class X, class P> __device__ void protoFunction(const X & in, P & pv) const;
//       template <class P> __device__ double samplingFunction(const P & pv) const;
//       template <class PO, class PN> __device__ double acceptanceFunction(const PO & po, const PN & pn) const;
//       template <class W, class PO, class PN> __device__ double updatedAcceptance(const W & wlk, const PO & po, PN & pn) const;
//       // optional (registration flag MCIG_PLUGIN_LOG_ACCEPTANCE): log of acceptanceFunction. When every sampling function
//       // of an integrator provides it, production modes test u <= exp(sum of logs) with an FP32 pre-filter (accept_log).
//       template <class PO, class PN> __device__ double logAcceptance(const PO & po, const PN & pn) const;
//       // optional (flag MCIG_PLUGIN_PROTO_ELEMENT, with ELEMENTWISE | HAS_UPDATE): proto value k from x[k] alone; single-vector moves on shared- /
//       // global-memory walkers then keep no proto-value array at all (old values are recomputed from the old coordinates)
//       __device__ double protoElement(double xk) const;
//       // with HAS_UPDATE as well: the selective counterpart (updates protonew like updatedAcceptance, returns the log)
//       template <class W, class PO, class PN> __device__ double updatedLogAcceptance(const W & wlk, const PO & po, PN & pn) const;
//   };
//   struct MyObs {                                  // mirrors mci::ObservableFunctionInterface
//       static constexpr int NPAR = 0;
//       const double * par;
//       template <class X, class O> __device__ void observableFunction(const X & in, O & out) const;
//       // optional (registration flag MCIG_PLUGIN_ELEMENTWISE, nobs == ndim): out[j] as a function of in[j] alone. Under single-vector
//       // moves Simple / Block accumulators then add value x dwell time when a coordinate changes instead of every component at
//       // every step (the device counterpart of updatedObservable + flags_xchanged, src/AccumulatorInterface.cpp:55-85)
//       __device__ double observableElement(double xj) const;
//   };
//   struct MyDependentObs {                         // mirrors a class deriving from mci::DependentObservableInterface as well
//       const double * par;                         // (registration flag MCIG_PLUGIN_DEPENDENT)
//       // dep.proto(i): i-th proto value of the sampling functions at the current position (flat over the pdfs in the order added)
//       // dep.obs(k, j): j-th value of observable k (k < own index, its nskip divides this one's) as evaluated in this step
//       template <class X, class DEP> __device__ void observableFunction(const X & in, double * out, const DEP & dep) const;
//   };
//   struct MyCallback {                             // MCI::setCallback (plugin kind MCIG_PLUGIN_CALLBACK)
//       const double * par;
//       // called by every walker after each accept decision, before the commit (xold, xnew as the reference's WalkerState holds
//       // them at src/MCIntegrator.cpp:346), and once per sampling run with step = -1 (initializeSampling, :267);
//       // buffer: the integrator-owned device buffer of mcig_set_callback, shared by all walkers (index by walker or atomicAdd)
//       template <class XO, class XN> __device__ void operator()(const XO & xold, const XN & xnew, bool accepted, long long walker,
//                                                                long long step, double * buffer) const;
//   };
//
// X / P / O are array-like (double*, a strided shared- or global-memory view): index them with [], nothing else.
// ndim / nproto / nobs are declared to the host through the C-ABI registration (include/mcig.h).
//
// Arithmetic note: expressions keep the reference's operation order so that the replay mode (compiled with
// --fmad=false) reproduces the CPU bits; XSquared multiplies by the reciprocal instead of dividing by 3 (an FP64
// division costs ~20 issue slots on the device): <= 1 ulp per sample, inside the 1e-12 sum tolerance.
#pragma once
#include "mcig_device.cuh"

namespace mcig_builtin {

using mcig::exp;

// 1/3 as a constant-bank operand (a 64-bit immediate costs two UMOV per use, re-materialised inside the walk loop)
__constant__ double c_third = 1./3.;

// ---------------------------------------------------------------- sampling functions
struct ThreeDimGaussianPDF { // TestMCIFunctions.hpp:123-148 (ndim 3, nproto 1)
    static constexpr int NPAR = 0;
    static constexpr bool HAS_UPDATE = false;
    static constexpr bool ELEMENTWISE = false;
    const double * par;
    template <class X, class P>
    MCIG_DEV void protoFunction(const X & in, P & pv) const { pv[0] = in[0]*in[0] + in[1]*in[1] + in[2]*in[2]; }
    template <class P>
    MCIG_DEV double samplingFunction(const P & pv) const { return exp(-pv[0]); }
    template <class PO, class PN>
    MCIG_DEV double acceptanceFunction(const PO & po, const PN & pn) const { return exp(-pn[0] + po[0]); }
    template <class PO, class PN>
    MCIG_DEV double logAcceptance(const PO & po, const PN & pn) const { return -pn[0] + po[0]; }
};

template <int NDIM>
struct Gauss { // TestMCIFunctions.hpp:151-187 (nproto = ndim, selective update)
    static constexpr int NPAR = 0;
    static constexpr bool HAS_UPDATE = true;
    static constexpr bool ELEMENTWISE = true;
    const double * par;
    MCIG_DEV double protoElement(double xk) const { return xk*xk; } // proto value k as a function of x[k] alone (MCIG_PLUGIN_PROTO_ELEMENT)
    template <class X, class P>
    MCIG_DEV void protoFunction(const X & in, P & pv) const
    {
#pragma unroll mcig::unroll_n(NDIM)
        for (int i = 0; i < NDIM; ++i) { pv[i] = in[i]*in[i]; }
    }
    template <class P>
    MCIG_DEV double samplingFunction(const P & pv) const
    {
        double s = 0.;
#pragma unroll mcig::unroll_n(NDIM)
        for (int i = 0; i < NDIM; ++i) { s += pv[i]; }
        return exp(-s);
    }
    template <class PO, class PN>
    MCIG_DEV double acceptanceFunction(const PO & po, const PN & pn) const
    {
        double a = 0., b = 0.;
#pragma unroll mcig::unroll_n(NDIM)
        for (int i = 0; i < NDIM; ++i) { a += po[i]; }
#pragma unroll mcig::unroll_n(NDIM)
        for (int i = 0; i < NDIM; ++i) { b += pn[i]; }
        return exp(a - b);
    }
    template <class PO, class PN>
    MCIG_DEV double logAcceptance(const PO & po, const PN & pn) const
    {
        double a = 0., b = 0.;
#pragma unroll mcig::unroll_n(NDIM)
        for (int i = 0; i < NDIM; ++i) { a += po[i]; }
#pragma unroll mcig::unroll_n(NDIM)
        for (int i = 0; i < NDIM; ++i) { b += pn[i]; }
        return a - b;
    }
    template <class W, class PO, class PN>
    MCIG_DEV double updatedAcceptance(const W & wlk, const PO & po, PN & pn) const
    {
        double expf = 0.;
        for (int i = 0; i < wlk.nchanged; ++i) {
            const int k = wlk.changedIdx[i];
            const double xk = wlk.xnew[k];
            const double v = xk*xk;
            pn[k] = v;
            expf += v - po[k];
        }
        return exp(-expf);
    }
    template <class W, class PO, class PN>
    MCIG_DEV double updatedLogAcceptance(const W & wlk, const PO & po, PN & pn) const
    {
        double expf = 0.;
        for (int i = 0; i < wlk.nchanged; ++i) {
            const int k = wlk.changedIdx[i];
            const double xk = wlk.xnew[k];
            const double v = xk*xk;
            pn[k] = v;
            expf += v - po[k];
        }
        return -expf;
    }
};

struct Exp1DPDF { // TestMCIFunctions.hpp:189-215
    static constexpr int NPAR = 0;
    static constexpr bool HAS_UPDATE = false;
    static constexpr bool ELEMENTWISE = false;
    const double * par;
    template <class X, class P>
    MCIG_DEV void protoFunction(const X & in, P & pv) const { pv[0] = fabs(in[0]); }
    template <class P>
    MCIG_DEV double samplingFunction(const P & pv) const { return exp(-pv[0]); }
    template <class PO, class PN>
    MCIG_DEV double acceptanceFunction(const PO & po, const PN & pn) const { return exp(-pn[0] + po[0]); }
    template <class PO, class PN>
    MCIG_DEV double logAcceptance(const PO & po, const PN & pn) const { return -pn[0] + po[0]; }
};

template <int NDIM>
struct ExpNDPDF { // TestMCIFunctions.hpp:217-256
    static constexpr int NPAR = 0;
    static constexpr bool HAS_UPDATE = true;
    static constexpr bool ELEMENTWISE = true;
    const double * par;
    MCIG_DEV double protoElement(double xk) const { return fabs(xk); } // proto value k as a function of x[k] alone (MCIG_PLUGIN_PROTO_ELEMENT)
    template <class X, class P>
    MCIG_DEV void protoFunction(const X & in, P & pv) const
    {
#pragma unroll mcig::unroll_n(NDIM)
        for (int i = 0; i < NDIM; ++i) { pv[i] = fabs(in[i]); }
    }
    template <class P>
    MCIG_DEV double samplingFunction(const P & pv) const
    {
        double s = 0.;
#pragma unroll mcig::unroll_n(NDIM)
        for (int i = 0; i < NDIM; ++i) { s += pv[i]; }
        return exp(-s);
    }
    template <class PO, class PN>
    MCIG_DEV double acceptanceFunction(const PO & po, const PN & pn) const
    {
        double a = 0., b = 0.;
#pragma unroll mcig::unroll_n(NDIM)
        for (int i = 0; i < NDIM; ++i) { a += po[i]; }
#pragma unroll mcig::unroll_n(NDIM)
        for (int i = 0; i < NDIM; ++i) { b += pn[i]; }
        return exp(a - b);
    }
    template <class PO, class PN>
    MCIG_DEV double logAcceptance(const PO & po, const PN & pn) const
    {
        double a = 0., b = 0.;
#pragma unroll mcig::unroll_n(NDIM)
        for (int i = 0; i < NDIM; ++i) { a += po[i]; }
#pragma unroll mcig::unroll_n(NDIM)
        for (int i = 0; i < NDIM; ++i) { b += pn[i]; }
        return a - b;
    }
    template <class W, class PO, class PN>
    MCIG_DEV double updatedAcceptance(const W & wlk, const PO & po, PN & pn) const
    {
        double expf = 0.;
        for (int i = 0; i < wlk.nchanged; ++i) {
            const int k = wlk.changedIdx[i];
            const double v = fabs(wlk.xnew[k]);
            pn[k] = v;
            expf += v - po[k];
        }
        return exp(-expf);
    }
    template <class W, class PO, class PN>
    MCIG_DEV double updatedLogAcceptance(const W & wlk, const PO & po, PN & pn) const
    {
        double expf = 0.;
        for (int i = 0; i < wlk.nchanged; ++i) {
            const int k = wlk.changedIdx[i];
            const double v = fabs(wlk.xnew[k]);
            pn[k] = v;
            expf += v - po[k];
        }
        return -expf;
    }
};

struct NormalizedLine { // ExampleFunctions.hpp:54-82 (not an exponential family: acceptance is a plain ratio)
    static constexpr int NPAR = 0;
    static constexpr bool HAS_UPDATE = false;
    static constexpr bool ELEMENTWISE = false;
    const double * par;
    template <class X, class P>
    MCIG_DEV void protoFunction(const X & in, P & pv) const { pv[0] = 0.2*fabs(in[0]); }
    template <class P>
    MCIG_DEV double samplingFunction(const P & pv) const { return pv[0]; }
    template <class PO, class PN>
    MCIG_DEV double acceptanceFunction(const PO & po, const PN & pn) const
    {
        if (po[0] == 0.) { return (pn[0] != 0.) ? 1. : 0.; }
        return pn[0]/po[0];
    }
};

// ---------------------------------------------------------------- observables
struct XSquared { // TestMCIFunctions.hpp:260-275
    static constexpr int NPAR = 0;
    const double * par;
    template <class X, class O>
    MCIG_DEV void observableFunction(const X & in, O & out) const { out[0] = (in[0]*in[0] + in[1]*in[1] + in[2]*in[2])*c_third; }
};

struct GaussXSquared { // TestMCIFunctions.hpp:278-296
    static constexpr int NPAR = 0;
    const double * par;
    template <class X, class O>
    MCIG_DEV void observableFunction(const X & in, O & out) const
    {
        const double normf = 0.059862374041722184; // 1./sqrt(M_PI*M_PI*M_PI)/3.
        const double x2 = in[0]*in[0] + in[1]*in[1] + in[2]*in[2];
        out[0] = exp(-x2)*x2*normf;
    }
};

struct XYZSquared { // TestMCIFunctions.hpp:299-317
    static constexpr int NPAR = 0;
    const double * par;
    template <class X, class O>
    MCIG_DEV void observableFunction(const X & in, O & out) const
    {
        out[0] = in[0]*in[0];
        out[1] = in[1]*in[1];
        out[2] = in[2]*in[2];
    }
};

struct X1D { // TestMCIFunctions.hpp:320-336
    static constexpr int NPAR = 0;
    const double * par;
    template <class X, class O>
    MCIG_DEV void observableFunction(const X & in, O & out) const { out[0] = in[0]; }
};

template <int NDIM>
struct XND { // TestMCIFunctions.hpp:339-379 (XND and UpdateableXND compute the same values)
    static constexpr int NPAR = 0;
    const double * par;
    template <class X, class O>
    MCIG_DEV void observableFunction(const X & in, O & out) const
    {
#pragma unroll mcig::unroll_n(NDIM)
        for (int i = 0; i < NDIM; ++i) { out[i] = in[i]; }
    }
    MCIG_DEV double observableElement(double xj) const { return xj; } // element-wise: out[j] is a function of in[j] alone
};

struct Constval { // TestMCIFunctions.hpp:382-398
    static constexpr int NPAR = 0;
    const double * par;
    template <class X, class O>
    MCIG_DEV void observableFunction(const X &, O & out) const { out[0] = 1.3; }
};

template <int NDIM>
struct Polynom { // TestMCIFunctions.hpp:401-420
    static constexpr int NPAR = 0;
    const double * par;
    template <class X, class O>
    MCIG_DEV void observableFunction(const X & in, O & out) const
    {
        double s = 0.;
#pragma unroll mcig::unroll_n(NDIM)
        for (int i = 0; i < NDIM; ++i) { s += in[i]; }
        out[0] = s;
    }
};

template <int NDIM>
struct X2Sum { // TestMCIFunctions.hpp:423-442
    static constexpr int NPAR = 0;
    const double * par;
    template <class X, class O>
    MCIG_DEV void observableFunction(const X & in, O & out) const
    {
        double s = 0.;
#pragma unroll mcig::unroll_n(NDIM)
        for (int i = 0; i < NDIM; ++i) { s += in[i]*in[i]; }
        out[0] = s;
    }
};

template <int NDIM>
struct X2 { // TestMCIFunctions.hpp:445-473
    static constexpr int NPAR = 0;
    const double * par;
    template <class X, class O>
    MCIG_DEV void observableFunction(const X & in, O & out) const
    {
#pragma unroll mcig::unroll_n(NDIM)
        for (int i = 0; i < NDIM; ++i) { out[i] = in[i]*in[i]; }
    }
    MCIG_DEV double observableElement(double xj) const { return xj*xj; }
};

struct Parabola { // ExampleFunctions.hpp:11-29
    static constexpr int NPAR = 0;
    const double * par;
    template <class X, class O>
    MCIG_DEV void observableFunction(const X & in, O & out) const { out[0] = 4.*in[0] - in[0]*in[0]; }
};

struct NormalizedParabola { // ExampleFunctions.hpp:32-50
    static constexpr int NPAR = 0;
    const double * par;
    template <class X, class O>
    MCIG_DEV void observableFunction(const X & in, O & out) const
    {
        const double x = in[0];
        double v = (4. - x)*5.;
        if (__double2hiint(x) < 0) { v = -v; } // std::signbit
        out[0] = v;
    }
};

} // namespace mcig_builtin
