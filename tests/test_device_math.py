"""GPU: device building blocks checked through user plugins (which also exercises the plugin registration path):
Philox4x32-10 known-answer vectors (Random123 kat_vectors) and bit-equality of mcig::exp with CUDA's exp()."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KAT = [  # (counter, key) -> output, Random123 examples/kat_vectors: philox4x32 10
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,)*4, (0xffffffff,)*2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers(mcig):
    body = []
    for i, (c, k, _) in enumerate(KAT):
        body.append("{ uint4 r = mcig::philox4x32_10(make_uint4(%du,%du,%du,%du), make_uint2(%du,%du)); out[%d]=r.x; out[%d]=r.y; out[%d]=r.z; out[%d]=r.w; }"
                    % (c + k + (4*i, 4*i + 1, 4*i + 2, 4*i + 3)))
    src = """struct PhiloxKat { static constexpr int NPAR = 0; const double * par;
      template <class X, class O> __device__ void observableFunction(const X &, O & out) const { %s } };""" % "\n".join(body)
    mcig.register_plugin(1, "PhiloxKat", "PhiloxKat", src, ndim=0, nvalues=12)
    mci = mcig.MCI(1)
    mci.setRngMode(0)
    mci.addSamplingFunction(mcig.Exp1DPDF())
    mci.addObservable(mcig.Observable("PhiloxKat"), 1, 1, False, mcig.EstimatorType.Noop)
    mci.integrate(4, False, False)
    data = mci.obsData(0, nobs=12)[0].astype(np.uint64)
    want = np.array([v for _, _, o in KAT for v in o], dtype=np.uint64)
    assert np.array_equal(data, want)


def test_fp64_pipe_mulhi_equals_the_integer_multiplier(mcig):
    """mulhi_f64<M>(a) (one DFMA rounding towards zero on 2^52 + a) is floor(a M / 2^32) for every word: edge values and a pseudo-random sweep
    against __umulhi, both Philox multipliers. (The walk loop keeps the integer multiplier: profiles/r02_f64hi_knobs.log; MCIG_F64HI0/1 switch rounds over.)"""
    src = """struct MulhiCheck { static constexpr int NPAR = 0; const double * par;
      template <class X, class O> __device__ void observableFunction(const X & in, O & out) const {
        const unsigned edge[8] = {0u, 1u, 2u, 0x7fffffffu, 0x80000000u, 0xfffffffeu, 0xffffffffu, 0x00010000u};
        unsigned bad = 0, a = (unsigned)__double2loint(in[0]*1e6) ^ 0x9e3779b9u;
        for (int i = 0; i < 4096; ++i) {
          const unsigned v = (i < 8) ? edge[i] : a;
          bad += (mcig::mulhi_f64<0xD2511F53u>(v) != __umulhi(0xD2511F53u, v)) + (mcig::mulhi_f64<0xCD9E8D57u>(v) != __umulhi(0xCD9E8D57u, v));
          a = a*1664525u + 1013904223u;
        }
        out[0] = (double)bad; out[1] = (double)mcig::mulhi_f64<0xD2511F53u>(0xffffffffu); } };"""
    mcig.register_plugin(1, "MulhiCheck", "MulhiCheck", src, ndim=1, nvalues=2)
    mci = mcig.MCI(1)
    mci.setRngMode(0)
    mci.setNWalkers(64)
    mci.addSamplingFunction(mcig.Exp1DPDF())
    mci.addObservable(mcig.Observable("MulhiCheck"), 1, 1, False, mcig.EstimatorType.Noop)
    mci.integrate(16, False, False)
    for w in (0, 31, 63):
        d = mci.obsData(0, walker=w, nobs=2)
        assert (d[:, 0] == 0).all() and (d[:, 1] == float((0xffffffff*0xD2511F53) >> 32)).all()


def test_exp_bit_equal_to_libdevice(mcig):
    src = """struct ExpCheck { static constexpr int NPAR = 0; const double * par;
      template <class X, class O> __device__ void observableFunction(const X & in, O & out) const {
        const double a = -37.0*in[0]*in[0], b = 11.0*in[0];
        out[0] = a; out[1] = mcig::exp(a); out[2] = ::exp(a); out[3] = b; out[4] = mcig::exp(b); out[5] = ::exp(b);
        out[6] = mcig::exp(-800.0*in[0]*in[0]); out[7] = ::exp(-800.0*in[0]*in[0]); } };"""
    mcig.register_plugin(1, "ExpCheck", "ExpCheck", src, ndim=1, nvalues=8)
    mci = mcig.MCI(1)
    mci.setRngMode(0)
    mci.setSeed(3)
    mci.setNWalkers(64)
    mci.addSamplingFunction(mcig.Exp1DPDF())
    mci.setMRT2Step(3.0)
    mci.addObservable(mcig.Observable("ExpCheck"), 1, 1, False, mcig.EstimatorType.Noop)
    mci.integrate(4096, False, False)
    for w in (0, 17, 63):
        d = mci.obsData(0, walker=w, nobs=8)
        assert np.array_equal(d[:, 1].view(np.uint64), d[:, 2].view(np.uint64))
        assert np.array_equal(d[:, 4].view(np.uint64), d[:, 5].view(np.uint64))
        assert np.array_equal(d[:, 6].view(np.uint64), d[:, 7].view(np.uint64))
        assert np.allclose(d[:, 1], np.exp(d[:, 0]), rtol=4e-16, atol=0)
        assert d[:, 0].min() < -50 and (d[:, 6] == 0).any()  # slow path (underflow) was exercised too


def test_measured_issue_rates_are_sane(mcig):
    dfma, imad = mcig.measure_peaks()
    assert 5e12 < dfma < 4e13 and 5e12 < imad < 8e13  # B200: 64 FP64 lanes/SM/clk x 148 SMs x <= 1.965 GHz = 1.86e13
