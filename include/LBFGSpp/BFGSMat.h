// LBFGSpp/BFGSMat.h -- the limited-memory quasi-Newton matrix, resident in HBM.
//
// Front for the S/Y ring kept by liblbfgs_b200 (lbfgs_b200_hist).  It offers the members of the reference's
// BFGSMat<Scalar> that the unconstrained solver uses (reference include/LBFGSpp/BFGSMat.h): reset(n, m) :61-78,
// add_correction(s, y) :81-97 and apply_Hv(v, a, res) :276-302, plus one fused entry the reference spells as four
// Eigen expressions: update(x, xp, g, gp) == { s = x - xp; y = g - gp; if (s'y > eps*y'y) add_correction(s, y) }
// (reference LBFGS.h:159-162).  The n x m matrices never visit the host.
#ifndef LBFGSPP_B200_BFGS_MAT_H
#define LBFGSPP_B200_BFGS_MAT_H

#include <limits>

#include "DeviceVector.h"

namespace LBFGSpp {

template <typename Scalar, bool LBFGSB = false>
class BFGSMat
{
    typedef DeviceVector<Scalar> Vector;

    Device* m_dev;
    lbfgs_b200_hist* m_hist;
    std::ptrdiff_t m_n;
    int m_m;
    int m_algo;  // LBFGS_B200_HV_*

    BFGSMat(const BFGSMat&);
    BFGSMat& operator=(const BFGSMat&);

public:
    BFGSMat() : m_dev(nullptr), m_hist(nullptr), m_n(0), m_m(0), m_algo(LBFGS_B200_HV_AUTO) {}
    ~BFGSMat() { lbfgs_b200_hist_destroy(m_hist); }

    // Which apply_Hv implementation to run (LBFGS_B200_HV_AUTO picks by problem size).
    void set_algorithm(int algo) { m_algo = algo; }
    int algorithm() const { return m_algo; }

    // Forget all pairs.  Unlike the reference, storage is kept when (n, m) did not change, so calling
    // minimize() repeatedly does not reallocate 2*n*m words each time.
    void reset(Device& dev, std::ptrdiff_t n, int m)
    {
        if (m_hist && (m_dev != &dev || m_n != n || m_m != m))
        {
            lbfgs_b200_hist_destroy(m_hist);
            m_hist = nullptr;
        }
        m_dev = &dev;
        m_n = n;
        m_m = m;
        if (!m_hist)
            dev.check(lbfgs_b200_hist_create(dev.ctx(), &m_hist, n, m, int(sizeof(Scalar))));
        else
            dev.check(lbfgs_b200_hist_reset(m_hist));
    }
    void reset(std::ptrdiff_t n, int m) { reset(Device::get_default(), n, m); }

    int num_corrections() const { return lbfgs_b200_hist_ncorr(m_hist); }

    // BFGSMat.h:81-97
    void add_correction(const Vector& s, const Vector& y)
    {
        m_dev->check(detail::Abi<Scalar>::hist_add(m_hist, s.data(), y.data()));
    }

    // LBFGS.h:159-162 in one kernel; returns whether the pair passed the curvature gate
    bool update(const Vector& x, const Vector& xp, const Vector& g, const Vector& gp)
    {
        int accepted = 0;
        m_dev->check(detail::Abi<Scalar>::hist_update(m_hist, x.data(), xp.data(), g.data(), gp.data(),
                                                      std::numeric_limits<Scalar>::epsilon(), &accepted, nullptr));
        return accepted != 0;
    }

    // res = a * H * v  (BFGSMat.h:276-302)
    void apply_Hv(const Vector& v, const Scalar& a, Vector& res)
    {
        res.resize(v.size());
        m_dev->check(detail::Abi<Scalar>::hist_apply_Hv(m_hist, v.data(), a, res.data(), m_algo, nullptr));
    }
    // Same, and also returns v.res from the tail of the last kernel: with v = grad, a = -1 this is the
    // directional derivative `dg = m_grad.dot(m_drt)` of LBFGS.h:123 at no extra memory traffic.
    Scalar apply_Hv_dot(const Vector& v, const Scalar& a, Vector& res)
    {
        res.resize(v.size());
        Scalar vr = Scalar(0);
        m_dev->check(detail::Abi<Scalar>::hist_apply_Hv(m_hist, v.data(), a, res.data(), m_algo, &vr));
        return vr;
    }
};

}  // namespace LBFGSpp

#endif  // LBFGSPP_B200_BFGS_MAT_H
