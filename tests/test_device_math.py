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
