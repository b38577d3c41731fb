// LBFGS.h -- LBFGSSolver<Scalar, LineSearch>: unconstrained L-BFGS with every n-vector in B200 HBM.
//
// Drop-in for the reference's include/LBFGS.h: same class template, same constructor / minimize() /
// final_grad() / final_grad_norm() surface, same convergence rules and return values (reference LBFGS.h:78-173).
// What changes is the Vector type (LBFGSpp::DeviceVector<Scalar>, device memory) and where the arithmetic runs:
// every Eigen expression of the reference's loop body is a call into liblbfgs_b200 (include/lbfgs_b200.h).
//
//   reference (per iteration)                         here
//   m_xp = x; m_gradp = m_grad        LBFGS.h:121-122  O(1) pointer rotation, no copy
//   dg = m_grad.dot(m_drt)            LBFGS.h:123      comes out of the tail of apply_Hv (v.res)
//   line search trials                LBFGS.h:127      one fused kernel per trial (built-in objectives)
//   m_grad.norm(), x.norm()           LBFGS.h:130,137  by-products of the accepted trial's kernel
//   s, y, gate, add_correction        LBFGS.h:159-162  one kernel, written straight into the ring slot
//   apply_Hv(m_grad, -1, m_drt)       LBFGS.h:165      2c+1 fused stage kernels / Gram form / resident kernel
//
// The objective is any callable `Scalar f(const Vector& x, Vector& grad)` working on device vectors (taken by
// non-const reference, called once per trial, exactly like the reference).  Objectives that additionally offer
// fused_trial()/fused_value() (see LBFGSpp/DeviceObjectives.h) get the single-kernel trial.
#ifndef LBFGSPP_B200_LBFGS_H
#define LBFGSPP_B200_LBFGS_H

#include <algorithm>
#include <cmath>
#include <type_traits>
#include <vector>

#include "LBFGSpp/BFGSMat.h"
#include "LBFGSpp/DeviceVector.h"
#include "LBFGSpp/LineSearchBacktracking.h"
#include "LBFGSpp/LineSearchBracketing.h"
#include "LBFGSpp/LineSearchDriver.h"
#include "LBFGSpp/LineSearchMoreThuente.h"
#include "LBFGSpp/LineSearchNocedalWright.h"
#include "LBFGSpp/Param.h"

namespace LBFGSpp {

// A minimal host vector for code that has no Eigen: contiguous storage + the few members user functors rely on.
template <typename Scalar>
class HostVector
{
    std::vector<Scalar> m_v;

public:
    HostVector() {}
    explicit HostVector(std::ptrdiff_t n, Scalar fill = Scalar(0)) : m_v(size_t(n), fill) {}
    std::ptrdiff_t size() const { return std::ptrdiff_t(m_v.size()); }
    void resize(std::ptrdiff_t n) { m_v.resize(size_t(n)); }
    Scalar* data() { return m_v.data(); }
    const Scalar* data() const { return m_v.data(); }
    Scalar& operator[](std::ptrdiff_t i) { return m_v[size_t(i)]; }
    const Scalar& operator[](std::ptrdiff_t i) const { return m_v[size_t(i)]; }
    static HostVector Zero(std::ptrdiff_t n) { return HostVector(n, Scalar(0)); }
    static HostVector Constant(std::ptrdiff_t n, Scalar v) { return HostVector(n, v); }
};

// Wraps a host functor `Scalar f(const HostVec& x, HostVec& grad)` as a device functor (see LBFGSSolver::minimize below).
template <typename Foo, typename HostVec>
class HostFunctorAdapter
{
    Foo& m_f;
    HostVec m_x, m_g;

public:
    HostFunctorAdapter(Foo& f, std::ptrdiff_t n) : m_f(f), m_x(n), m_g(n) {}
    template <typename Scalar>
    Scalar operator()(const DeviceVector<Scalar>& x, DeviceVector<Scalar>& grad)
    {
        x.copy_to_host(m_x.data());
        const Scalar fx = m_f(static_cast<const HostVec&>(m_x), m_g);
        grad.copy_from_host(m_g.data(), x.size());
        return fx;
    }
};

template <typename Scalar, template <class> class LineSearch = LineSearchNocedalWright>
class LBFGSSolver
{
public:
    typedef DeviceVector<Scalar> Vector;

private:
    const LBFGSParam<Scalar>& m_param;  // held by reference, like the reference (LBFGS.h:29)
    BFGSMat<Scalar> m_bfgs;
    std::vector<Scalar> m_fx;           // ring of past objective values (host scalars)
    Vector m_xp, m_grad, m_gradp, m_drt;
    LineSearchWorkspace<Scalar> m_ws;
    Scalar m_gnorm;
    long m_nfev;

    bool small_gradient(Scalar gg, Scalar xx)
    {
        m_gnorm = std::sqrt(gg);
        return m_gnorm <= m_param.epsilon || m_gnorm <= m_param.epsilon_rel * std::sqrt(xx);
    }

public:
    LBFGSSolver(const LBFGSParam<Scalar>& param) : m_param(param), m_gnorm(0), m_nfev(0) { m_param.check_param(); }

    // apply_Hv implementation selector (LBFGS_B200_HV_*); not part of the reference API
    void set_hv_algorithm(int algo) { m_bfgs.set_algorithm(algo); }

    // Minimise f starting from x (device vector, updated in place; its storage may be exchanged with an internal
    // buffer).  Returns the number of iterations; throws what the reference throws.
    template <typename Foo>
    inline int minimize(Foo& f, Vector& x, Scalar& fx)
    {
        using std::abs;
        Device& dev = x.device();
        const std::ptrdiff_t n = x.size();
        const int fpast = m_param.past;

        m_bfgs.reset(dev, n, m_param.m);
        for (Vector* v : {&m_xp, &m_grad, &m_gradp, &m_drt, &m_ws.x_lo, &m_ws.grad_lo})
        {
            if (&v->device() != &dev) *v = Vector(dev);
            v->resize(n);
        }
        if (fpast > 0) m_fx.assign(size_t(fpast), Scalar(0));
        m_nfev = 0;

        // first evaluation and early exit (LBFGS.h:91-103)
        TrialValues<Scalar> at = detail::evaluate_point(f, static_cast<const Vector&>(x), m_grad);
        m_nfev++;
        fx = at.fx;
        if (fpast > 0) m_fx[0] = fx;
        if (small_gradient(at.gg, at.xx)) return 1;

        // steepest-descent start: drt = -grad, first step 1/||drt||  (LBFGS.h:106-108)
        dev.check(detail::Abi<Scalar>::scale_out(dev.ctx(), n, Scalar(-1), m_grad.data(), m_drt.data()));
        Scalar step = Scalar(1) / m_gnorm;
        Scalar dg = -at.gg;  // grad . (-grad): the same products as g.g, negated
        m_ws.gg = at.gg;
        m_ws.xx = at.xx;

        int k = 1;
        for (;;)
        {
            // the current point becomes the "previous" one: rotate buffers instead of copying
            m_xp.swap(x);
            m_gradp.swap(m_grad);

            const Scalar step_max = m_param.max_step;
            run_line_search<typename LineSearch<Scalar>::Machine>(f, m_param, m_xp, m_gradp, m_drt, step_max, step, fx, dg,
                                                                  x, m_grad, m_ws);
            m_nfev += m_ws.evaluations;

            if (small_gradient(m_ws.gg, m_ws.xx)) return k;                 // LBFGS.h:137-140
            if (fpast > 0)                                                   // LBFGS.h:142-149
            {
                const Scalar fxd = m_fx[size_t(k % fpast)];
                if (k >= fpast && abs(fxd - fx) <= m_param.delta * std::max(std::max(abs(fx), abs(fxd)), Scalar(1))) return k;
                m_fx[size_t(k % fpast)] = fx;
            }
            if (m_param.max_iterations != 0 && k >= m_param.max_iterations) return k;  // LBFGS.h:151-154

            m_bfgs.update(x, m_xp, m_grad, m_gradp);                         // LBFGS.h:159-162
            dg = m_bfgs.apply_Hv_dot(m_grad, -Scalar(1), m_drt);             // LBFGS.h:165 (+ :123 of the next pass)
            step = Scalar(1);
            k++;
        }
        return k;
    }

    // ----- host-vector compatibility mode ---------------------------------------------------------------------------
    // Existing LBFGSpp code passes host vectors (Eigen::VectorXd) and a functor over host vectors.  Any `HostVec` with
    // data() / size() / resize() (Eigen::VectorXd, LBFGSpp::HostVector<Scalar>, ...) is accepted here: x is uploaded once, every
    // objective evaluation copies x to the host, calls `f(x_host, grad_host)` and uploads grad (2n words over PCIe per call --
    // meant for small problems and for porting; the vector work of the solver itself still runs on the GPU), and the
    // solution is copied back into x.
    template <typename Foo, typename HostVec>
    typename std::enable_if<!std::is_same<HostVec, Vector>::value, int>::type minimize(Foo& f, HostVec& x, Scalar& fx)
    {
        Device& dev = Device::get_default();
        const std::ptrdiff_t n = std::ptrdiff_t(x.size());
        HostFunctorAdapter<Foo, HostVec> adapter(f, n);
        Vector xd(dev);
        xd.copy_from_host(x.data(), n);
        const int niter = minimize(adapter, xd, fx);
        xd.copy_to_host(x.data());
        return niter;
    }

    const Vector& final_grad() const { return m_grad; }
    Scalar final_grad_norm() const { return m_gnorm; }
    // final_approx_hessian() / final_approx_inverse_hessian() of the reference (LBFGS.h:192-197 -> BFGSMat.h:150-271): explicit
    // n x n matrices, only sensible for small n; returned row-major on the host.
    SmallMatrix<Scalar> final_approx_hessian() { return m_bfgs.dense(false); }
    SmallMatrix<Scalar> final_approx_inverse_hessian() { return m_bfgs.dense(true); }
    // number of objective evaluations of the last minimize() call (not in the reference; used by tests/bench)
    long num_evaluations() const { return m_nfev; }
};

}  // namespace LBFGSpp

#endif  // LBFGSPP_B200_LBFGS_H
