// LBFGSpp/BFGSMat.h -- the limited-memory quasi-Newton matrix, resident in HBM.
//
// Front for the S/Y ring kept by liblbfgs_b200 (lbfgs_b200_hist).  It offers the members of the reference's
// BFGSMat<Scalar> that the unconstrained solver uses (reference include/LBFGSpp/BFGSMat.h): reset(n, m) :61-78,
// add_correction(s, y) :81-97 and apply_Hv(v, a, res) :276-302, plus one fused entry the reference spells as four
// Eigen expressions: update(x, xp, g, gp) == { s = x - xp; y = g - gp; if (s'y > eps*y'y) add_correction(s, y) }
// (reference LBFGS.h:159-162).  The n x m matrices never visit the host.
#ifndef LBFGSPP_B200_BFGS_MAT_H
#define LBFGSPP_B200_BFGS_MAT_H

#include <algorithm>
#include <limits>
#include <stdexcept>

#include <vector>

#include "DeviceVector.h"
#include "BKLDLT.h"
#include "SmallDense.h"

namespace LBFGSpp {

template <typename Scalar, bool LBFGSB = false>
class BFGSMat
{
    typedef DeviceVector<Scalar> Vector;

    Device* m_dev;
    lbfgs_b200_hist* m_hist;
    bool m_owned;      // false: m_hist belongs to a device-resident solver (borrow())
    std::ptrdiff_t m_n;
    int m_m;
    int m_algo;  // LBFGS_B200_HV_*

    BFGSMat(const BFGSMat&);
    BFGSMat& operator=(const BFGSMat&);

public:
    BFGSMat() : m_dev(nullptr), m_hist(nullptr), m_owned(true), m_n(0), m_m(0), m_algo(LBFGS_B200_HV_AUTO), m_theta_host(1), m_c_host(0), m_box(nullptr) {}
    ~BFGSMat()
    {
        lbfgs_b200_box_destroy(m_box);
        if (m_owned) lbfgs_b200_hist_destroy(m_hist);
    }

    // Which apply_Hv implementation to run (LBFGS_B200_HV_AUTO picks by problem size).
    void set_algorithm(int algo) { m_algo = algo; }
    int algorithm() const { return m_algo; }

    // Forget all pairs.  Unlike the reference, storage is kept when (n, m) did not change, so calling
    // minimize() repeatedly does not reallocate 2*n*m words each time.
    void reset(Device& dev, std::ptrdiff_t n, int m)
    {
        if (m_hist && (!m_owned || m_dev != &dev || m_n != n || m_m != m))
        {
            lbfgs_b200_box_destroy(m_box);
            m_box = nullptr;
            if (m_owned) lbfgs_b200_hist_destroy(m_hist);
            m_hist = nullptr;
            m_owned = true;
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

    // Look at a ring that somebody else owns (the device-resident solver's, after its minimize()): dense() and the Gram accessors
    // then describe that solve's final approximation.  The ring must outlive this object's use of it.
    void borrow(Device& dev, lbfgs_b200_hist* hist, std::ptrdiff_t n, int m)
    {
        lbfgs_b200_box_destroy(m_box);
        m_box = nullptr;
        if (m_owned) lbfgs_b200_hist_destroy(m_hist);
        m_hist = hist;
        m_owned = false;
        m_dev = &dev;
        m_n = n;
        m_m = m;
    }
    void require_history() const
    {
        if (!m_hist || !m_dev) throw std::logic_error("BFGSMat: no history yet (minimize() has not run)");
    }

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

    // update(x, xp, g, gp) followed by apply_Hv_dot(g, a, res) as ONE device pass over x, xp, g, gp (LBFGS.h:159-165): the
    // correction pair is formed inside the Gram-dots kernel.  Returns g.res; `accepted` tells whether the pair passed the gate.
    Scalar update_apply_Hv_dot(const Vector& x, const Vector& xp, const Vector& g, const Vector& gp, const Scalar& a, Vector& res,
                               bool* accepted = nullptr)
    {
        res.resize(g.size());
        int acc = 0;
        Scalar vr = Scalar(0);
        m_dev->check(detail::Abi<Scalar>::hist_update_apply_Hv(m_hist, x.data(), xp.data(), g.data(), gp.data(),
                                                               std::numeric_limits<Scalar>::epsilon(), a, res.data(), m_algo, &acc, &vr));
        if (accepted) *accepted = acc != 0;
        return vr;
    }

    //========== L-BFGS-B part (reference BFGSMat.h:99-146, 307-615) ==========//
    // B = theta*I - W M W' with W = [Y, theta*S] (n x 2c, columns ordered newest pair first) and M = inv(Minv),
    //   Minv = [ -D   L' ]   D = diag(s_a'y_a),  L(a,b) = s_a'y_b when pair a is newer than pair b, else 0.
    //          [  L  theta*S'S ]
    // The c x c Gram blocks are kept on the device (folded pair by pair by the same pass that serves apply_Hv) and
    // downloaded here; the 2c x 2c algebra is host work like in the reference.  W itself is never formed: products with
    // W are masked passes over the S/Y columns (Wt_dot, lincomb, masked_gram).
private:
    Scalar m_theta_host;
    int m_c_host;
    SmallMatrix<Scalar> m_Minv, m_M, m_SS, m_SY;
    BKLDLT<Scalar> m_Msolver;  // Minv is symmetric indefinite (reference BFGSMat.h:47,134-146)
    lbfgs_b200_box* m_box;

public:
    lbfgs_b200_hist* handle() const { return m_hist; }
    Device& device() const { return *m_dev; }
    Scalar theta() const { return m_theta_host; }
    int ncorr() const { return m_c_host; }
    const SmallMatrix<Scalar>& Minv() const { return m_Minv; }
    const SmallMatrix<Scalar>& Mmat() const { return m_M; }

    // Download the Gram blocks and rebuild Minv / M.  Call after reset() and after every accepted update().
    void refresh_middle()
    {
        require_history();
        const int c = num_corrections();
        m_c_host = c;
        m_theta_host = Scalar(1);
        if (c == 0)
        {
            m_dev->check(detail::BoxAbi<Scalar>::hist_gram(m_hist, nullptr, nullptr, nullptr, nullptr, &m_theta_host));
            m_Minv = SmallMatrix<Scalar>();
            m_M = SmallMatrix<Scalar>();
            return;
        }
        m_SY = SmallMatrix<Scalar>(c, c);
        m_SS = SmallMatrix<Scalar>(c, c);
        m_dev->check(detail::BoxAbi<Scalar>::hist_gram(m_hist, m_SY.data(), m_SS.data(), nullptr, nullptr, &m_theta_host));
        m_Minv = SmallMatrix<Scalar>(2 * c, 2 * c);
        for (int a = 0; a < c; a++)
        {
            m_Minv(a, a) = -m_SY(a, a);
            for (int b = 0; b < c; b++)
            {
                if (a < b)  // age a < age b: pair a is the newer one
                {
                    m_Minv(c + a, b) = m_SY(a, b);
                    m_Minv(b, c + a) = m_SY(a, b);
                }
                m_Minv(c + a, c + b) = m_theta_host * m_SS(a, b);
            }
        }
        m_Msolver.compute(m_Minv);
        m_M = SmallMatrix<Scalar>(2 * c, 2 * c);
        std::vector<Scalar> e(size_t(2 * c));
        for (int j = 0; j < 2 * c; j++)
        {
            std::fill(e.begin(), e.end(), Scalar(0));
            e[size_t(j)] = Scalar(1);
            m_Msolver.solve_inplace(e);
            for (int i = 0; i < 2 * c; i++) m_M(i, j) = e[size_t(i)];
        }
    }

    // res = M v   (apply_Mv, BFGSMat.h:361-378)
    std::vector<Scalar> apply_Mv(const std::vector<Scalar>& v) const
    {
        if (m_c_host < 1) return std::vector<Scalar>();
        return m_Msolver.solve(v);
    }

    // W'v for a device vector v (already zero outside the index set of interest): [Y'v ; theta*S'v]
    // (apply_Wtv / apply_WtPv, BFGSMat.h:315-320, 382-433)
    std::vector<Scalar> Wt_dot(const Scalar* v_dev)
    {
        const int c = m_c_host;
        std::vector<Scalar> raw(size_t(2 * c), Scalar(0));
        if (c > 0)
        {
            m_dev->check(detail::BoxAbi<Scalar>::hist_wt_dot(m_hist, v_dev, raw.data()));
            for (int a = 0; a < c; a++) raw[size_t(c + a)] *= m_theta_host;
        }
        return raw;
    }

    // out_i = a0*v0_i + sum_a cy_a*y_a[i] + cs_a*s_a[i] on rows whose class byte intersects `mask`
    void lincomb(Scalar a0, const Scalar* v0_dev, const std::vector<Scalar>& coef, int mask, Scalar* out_dev)
    {
        m_dev->check(detail::BoxAbi<Scalar>::hist_lincomb(m_hist, box(), a0, v0_dev, coef.empty() ? nullptr : coef.data(),
                                                          lbfgs_b200_box_classes(box()), mask, out_dev));
    }

    // inv(P'BP) v on the rows of `mask`; v_dev must be zero outside the mask  (solve_PtBP, BFGSMat.h:529-565)
    void solve_PtBP(int mask, const Scalar* v_dev, Scalar* out_dev)
    {
        const int c = m_c_host;
        const Scalar theta = m_theta_host;
        if (c < 1)
        {
            lincomb(Scalar(1) / theta, v_dev, std::vector<Scalar>(), mask, out_dev);
            return;
        }
        SmallMatrix<Scalar> G(2 * c, 2 * c);
        m_dev->check(detail::BoxAbi<Scalar>::hist_masked_gram(m_hist, box(), lbfgs_b200_box_classes(box()), mask, G.data()));
        SmallMatrix<Scalar> mid(2 * c, 2 * c);
        for (int a = 0; a < c; a++)
            for (int b = 0; b < c; b++)
            {
                mid(a, b) = m_Minv(a, b) - G(a, b) / theta;
                mid(c + a, b) = m_Minv(c + a, b) - G(c + a, b);
                mid(b, c + a) = mid(c + a, b);
                mid(c + a, c + b) = theta * (m_SS(a, b) - G(c + a, c + b));
            }
        BKLDLT<Scalar> midsolver(mid);  // reference BFGSMat.h:551
        std::vector<Scalar> z = Wt_dot(v_dev);      // [Y_P'v ; theta*S_P'v]
        midsolver.solve_inplace(z);
        std::vector<Scalar> coef(size_t(2 * c));
        const Scalar t2 = theta * theta;
        for (int a = 0; a < c; a++)
        {
            coef[size_t(a)] = z[size_t(a)] / t2;
            coef[size_t(c + a)] = theta * z[size_t(c + a)] / t2;
        }
        lincomb(Scalar(1) / theta, v_dev, coef, mask, out_dev);
    }

    // out = -W_rows * (M u) on the rows of `mask`, u a host 2c-vector in W's [Y ; theta*S] convention
    // (apply_PtWMv with scale -1 / apply_PtBQv, BFGSMat.h:435-478, 570-615)
    void minus_W_M(const std::vector<Scalar>& u, int mask, Scalar* out_dev)
    {
        const int c = m_c_host;
        std::vector<Scalar> coef(size_t(2 * c), Scalar(0));
        if (c > 0)
        {
            const std::vector<Scalar> Mu = apply_Mv(u);
            for (int a = 0; a < c; a++)
            {
                coef[size_t(a)] = -Mu[size_t(a)];
                coef[size_t(c + a)] = -m_theta_host * Mu[size_t(c + a)];
            }
        }
        lincomb(Scalar(0), nullptr, coef, mask, out_dev);
    }

    // Explicit n x n approximations for small n (final_approx_hessian / final_approx_inverse_hessian of the reference,
    // BFGSMat.h:150-271), through the compact representation:
    //   B = theta*I - W M W'                with the Minv above,
    //   H = I/theta + Z N Z',  Z = [Y/theta, S],  N = [ 0, -R^-1 ; -R^-T, R^-T (D + Y'Y/theta) R^-1 ],  R = triu(S'Y) in
    //   chronological order.  H is formed as inv(B) here (same matrix; n is small by contract).
    SmallMatrix<Scalar> dense(bool inverse)
    {
        refresh_middle();
        const int n = int(m_n), c = m_c_host;
        if (n > 4096) throw std::invalid_argument("dense Hessian approximations are only formed for n <= 4096");
        SmallMatrix<Scalar> B(n, n);
        for (int i = 0; i < n; i++) B(i, i) = m_theta_host;
        if (c > 0)
        {
            // rows of W = [Y, theta*S] by age, downloaded column by column
            SmallMatrix<Scalar> W(n, 2 * c);
            std::vector<Scalar> col(static_cast<size_t>(n));
            for (int a = 0; a < c; a++)
            {
                m_dev->check(lbfgs_b200_memcpy_d2h(m_dev->ctx(), col.data(), lbfgs_b200_hist_y_col(m_hist, a), sizeof(Scalar) * size_t(n)));
                for (int i = 0; i < n; i++) W(i, a) = col[size_t(i)];
                m_dev->check(lbfgs_b200_memcpy_d2h(m_dev->ctx(), col.data(), lbfgs_b200_hist_s_col(m_hist, a), sizeof(Scalar) * size_t(n)));
                for (int i = 0; i < n; i++) W(i, c + a) = m_theta_host * col[size_t(i)];
            }
            for (int i = 0; i < n; i++)
            {
                std::vector<Scalar> wi(size_t(2 * c));
                for (int q = 0; q < 2 * c; q++) wi[size_t(q)] = W(i, q);
                const std::vector<Scalar> Mw = m_M.times(wi);
                for (int j = 0; j < n; j++)
                {
                    Scalar acc = Scalar(0);
                    for (int q = 0; q < 2 * c; q++) acc += W(j, q) * Mw[size_t(q)];
                    B(j, i) -= acc;
                }
            }
        }
        if (!inverse) return B;
        SmallSolver<Scalar> lu(B);
        SmallMatrix<Scalar> H(n, n);
        std::vector<Scalar> e(static_cast<size_t>(n));
        for (int j = 0; j < n; j++)
        {
            std::fill(e.begin(), e.end(), Scalar(0));
            e[size_t(j)] = Scalar(1);
            lu.solve_inplace(e);
            for (int i = 0; i < n; i++) H(i, j) = e[size_t(i)];
        }
        return H;
    }

    // the n-sized scratch of Cauchy / SubspaceMin, created on first use and tied to the history
    lbfgs_b200_box* box()
    {
        if (!m_box) m_dev->check(lbfgs_b200_box_create(m_hist, &m_box));
        return m_box;
    }
    void drop_box()
    {
        lbfgs_b200_box_destroy(m_box);
        m_box = nullptr;
    }
};

}  // namespace LBFGSpp

#endif  // LBFGSPP_B200_BFGS_MAT_H
