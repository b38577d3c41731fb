// oracle_capi.cpp -- TEST INFRASTRUCTURE ONLY.  C ABI ("orc_" prefix of oracle_api.h) over the
// plain-C++ restatement in lbfgs_oracle.hpp / lbfgsb_oracle.hpp.  Built by oracle/Makefile into
// oracle/liboracle.so; loaded through ctypes by tests/, smoke() and bench.py's CPU legs only.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "lbfgs_oracle.hpp"
#include "lbfgsb_oracle.hpp"
#include "objectives.hpp"
#include "oracle_api.h"

namespace {

int hw_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

template <class T>
struct Functor
{
    int objective;
    const T* d0;
    const T* d1;
    long n;
    int threads;  // > 1: per-thread index ranges, partial sums added in range order
    long nfev;
    double* trace;
    long cap;
    T operator()(const T* x, T* g)
    {
        T fx;
        if (threads <= 1)
            fx = orc::evaluate<T>(objective, d0, d1, n, x, g);
        else
        {
            std::vector<T> part(threads, T(0));
            const long chunk = ((n + threads - 1) / threads + 1) & ~1L;
#pragma omp parallel for num_threads(threads) schedule(static, 1)
            for (int t = 0; t < threads; t++)
            {
                const long lo = std::min(n, t * chunk), hi = std::min(n, lo + chunk);
                part[t] = orc::evaluate_range<T>(objective, d0, d1, lo, hi, n, x, g);
            }
            fx = T(0);
            for (int t = 0; t < threads; t++) fx += part[t];
        }
        if (trace && nfev < cap) trace[nfev] = double(fx);
        nfev++;
        return fx;
    }
};

void set_error(orc_result* out, int code, const char* what)
{
    out->status = code;
    std::strncpy(out->msg, what, sizeof(out->msg) - 1);
    out->msg[sizeof(out->msg) - 1] = 0;
}

template <class Body>
int guarded(orc_result* out, Body body)
{
    std::memset(out, 0, sizeof(*out));
    try { body(); }
    catch (const std::invalid_argument& e) { set_error(out, ORC_INVALID_ARGUMENT, e.what()); }
    catch (const std::logic_error& e) { set_error(out, ORC_LOGIC_ERROR, e.what()); }
    catch (const std::runtime_error& e) { set_error(out, ORC_RUNTIME_ERROR, e.what()); }
    catch (const std::exception& e) { set_error(out, ORC_OTHER_ERROR, e.what()); }
    return out->status;
}

template <class T>
orc::Blas1<T> make_la(int sum_mode)
{
    return orc::Blas1<T>(sum_mode, sum_mode == ORC_SUM_LANES8_OMP ? hw_threads() : 1);
}

template <class T>
int lbfgs_any(int objective, const T* d0, const T* d1, long n, int ls, const orc_param* prm, int sum_mode, T* x,
              T* grad_out, double* trace, long cap, orc_result* out, bool gram)
{
    return guarded(out, [&]() {
        const orc::Blas1<T> la = make_la<T>(sum_mode);
        Functor<T> f{objective, d0, d1, n, la.threads, 0, trace, cap};
        std::vector<T> xv(x, x + n);
        const auto t0 = std::chrono::steady_clock::now();
        orc::LbfgsOutcome<T> res;
        try
        {
            res = gram ? orc::lbfgs_minimize_gram<T>(f, *prm, ls, la, xv) : orc::lbfgs_minimize<T>(f, *prm, ls, la, xv);
        }
        catch (...)
        {
            out->nfev = f.nfev;
            out->trace_len = std::min(f.nfev, cap);
            std::copy(xv.begin(), xv.end(), x);
            throw;
        }
        out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        out->niter = res.niter;
        out->fx = double(res.fx);
        out->gnorm = double(res.gnorm);
        out->nfev = f.nfev;
        out->trace_len = std::min(f.nfev, cap);
        std::copy(xv.begin(), xv.end(), x);
        if (grad_out) std::copy(res.grad.begin(), res.grad.end(), grad_out);
    });
}

}  // namespace

extern "C" {

void orc_default_param(orc_param* p, int lbfgsb)
{
    std::memset(p, 0, sizeof(*p));
    orc::default_param(*p, lbfgsb != 0);
}

int orc_lbfgs_f64(int objective, const double* d0, const double* d1, long n, int ls, const orc_param* prm, int sum_mode,
                  double* x, double* grad_out, double* trace, long cap, orc_result* out)
{
    return lbfgs_any<double>(objective, d0, d1, n, ls, prm, sum_mode, x, grad_out, trace, cap, out, false);
}

int orc_lbfgs_f32(int objective, const float* d0, const float* d1, long n, int ls, const orc_param* prm, int sum_mode,
                  float* x, float* grad_out, double* trace, long cap, orc_result* out)
{
    return lbfgs_any<float>(objective, d0, d1, n, ls, prm, sum_mode, x, grad_out, trace, cap, out, false);
}

int orc_lbfgs_gram_f64(int objective, const double* d0, const double* d1, long n, int ls, const orc_param* prm,
                       int sum_mode, double* x, double* grad_out, double* trace, long cap, orc_result* out)
{
    return lbfgs_any<double>(objective, d0, d1, n, ls, prm, sum_mode, x, grad_out, trace, cap, out, true);
}

int orc_lbfgsb_f64(int objective, const double* d0, const double* d1, long n, const orc_param* prm, int sum_mode,
                   double* x, const double* lb, const double* ub, double* grad_out, double* trace, long cap,
                   orc_result* out)
{
    return guarded(out, [&]() {
        const orc::Blas1<double> la = make_la<double>(sum_mode);
        Functor<double> f{objective, d0, d1, n, la.threads, 0, trace, cap};
        std::vector<double> xv(x, x + n), lbv(lb, lb + n), ubv(ub, ub + n);
        const auto t0 = std::chrono::steady_clock::now();
        orc::LbfgsOutcome<double> res;
        try { res = orc::lbfgsb_minimize<double>(f, *prm, la, xv, lbv, ubv); }
        catch (...)
        {
            out->nfev = f.nfev;
            out->trace_len = std::min(f.nfev, cap);
            throw;
        }
        out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        out->niter = res.niter;
        out->fx = res.fx;
        out->gnorm = res.gnorm;
        out->nfev = f.nfev;
        out->trace_len = std::min(f.nfev, cap);
        std::copy(xv.begin(), xv.end(), x);
        if (grad_out) std::copy(res.grad.begin(), res.grad.end(), grad_out);
    });
}

int orc_bfgs_apply_Hv_f64(long n, int m, int npairs, const double* S, const double* Y, const double* v, double a,
                          int sum_mode, double* res, double* ys_out, double* theta_out)
{
    orc::History<double> h(make_la<double>(sum_mode));
    h.reset(n, m);
    for (int k = 0; k < npairs; k++) h.add(S + size_t(k) * n, Y + size_t(k) * n);
    h.apply_Hv(v, a, res);
    if (ys_out) std::copy(h.ys.begin(), h.ys.end(), ys_out);
    if (theta_out) *theta_out = h.theta;
    return 0;
}

int orc_line_search_f64(int objective, const double* d0, const double* d1, long n, int ls, const orc_param* prm, const double* xp,
                        const double* drt, double step_max, double* step_inout, double* fx_out, double* dg_out, double* x_out,
                        double* grad_out, double* trace, long cap, orc_result* out)
{
    return guarded(out, [&]() {
        const orc::Blas1<double> la(ORC_SUM_SEQUENTIAL, 1);
        Functor<double> f{objective, d0, d1, n, 1, 0, nullptr, 0};
        std::vector<double> vxp(xp, xp + n), vd(drt, drt + n), x(n), grad(n);
        double fx = f(vxp.data(), grad.data());
        double dg = la.dot(grad.data(), vd.data(), n);
        f.nfev = 0;
        f.trace = trace;
        f.cap = cap;
        double step = *step_inout;
        try
        {
            switch (ls)
            {
            case ORC_LS_BACKTRACKING: orc::ls_backtracking(f, *prm, la, vxp, vd, step_max, step, fx, grad, dg, x); break;
            case ORC_LS_BRACKETING: orc::ls_bracketing(f, *prm, la, vxp, vd, step_max, step, fx, grad, dg, x); break;
            case ORC_LS_NOCEDAL_WRIGHT: orc::ls_nocedal_wright(f, *prm, la, vxp, vd, step_max, step, fx, grad, dg, x); break;
            default: orc::ls_more_thuente(f, *prm, la, vxp, vd, step_max, step, fx, grad, dg, x); break;
            }
        }
        catch (...)
        {
            out->nfev = f.nfev;
            out->trace_len = std::min(f.nfev, cap);
            throw;
        }
        *step_inout = step;
        *fx_out = fx;
        *dg_out = dg;
        std::copy(x.begin(), x.end(), x_out);
        std::copy(grad.begin(), grad.end(), grad_out);
        out->nfev = f.nfev;
        out->trace_len = std::min(f.nfev, cap);
        out->fx = fx;
    });
}

double orc_objective_f64(int objective, const double* d0, const double* d1, long n, const double* x, double* grad)
{
    return orc::evaluate<double>(objective, d0, d1, n, x, grad);
}

// Time `reps` apply_Hv calls on a full history (c = m).  Inputs follow SURVEY.md 8d's microbench recipe
// in spirit (s ~ noise, y = s + 0.1 noise so that s'y > 0) from a fixed LCG; returns seconds per call.
double orc_bfgs_apply_Hv_bench_f64(long n, int m, int reps, int sum_mode, int threads)
{
    orc::Blas1<double> la(sum_mode, sum_mode == ORC_SUM_LANES8_OMP ? (threads > 0 ? threads : hw_threads()) : 1);
    orc::History<double> h(la);
    h.reset(n, m);
    std::vector<double> s(n), y(n), v(n), res(n);
    unsigned long long st = 0x9E3779B97F4A7C15ull;
    auto rnd = [&st]() {
        st = st * 6364136223846793005ull + 1442695040888963407ull;
        return double(st >> 11) * (1.0 / 9007199254740992.0) - 0.5;
    };
    for (int k = 0; k < m; k++)
    {
        for (long i = 0; i < n; i++) { s[i] = rnd(); y[i] = s[i] + 0.1 * rnd(); }
        h.add(s.data(), y.data());
    }
    for (long i = 0; i < n; i++) v[i] = rnd();
    h.apply_Hv(v.data(), -1.0, res.data());  // warm-up
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; r++) h.apply_Hv(v.data(), -1.0, res.data());
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return dt / reps;
}

ORC_API int orc_hw_threads() { return hw_threads(); }

}  // extern "C"
