// CPU harness for the line-search cores of the product (include/LBFGSpp/LineSearchCore.h): drives a core with host-side
// trial evaluations written exactly like the CPU checker's (x = xp + step*d element by element, sequential dot), so that
// tests/test_front_cpu.py can demand bit-identical behaviour between the product's decision logic and the restated
// reference line searches -- without a GPU.  Test infrastructure (it includes oracle/objectives.hpp).
#include <cstring>
#include <vector>

#include "../../include/LBFGSpp/LineSearchCore.h"
#include "../../oracle/objectives.hpp"

using namespace LBFGSpp;

template <template <class> class Core>
static int drive(int objective, const double* d0, const double* d1, long n, const LineSearchOptions<double>& opt, const double* xp,
                 const double* drt, double step_max, double* step_io, double* fx_out, double* dg_out, double* x_out, double* grad_out,
                 double* trace, long cap, long* nfev_out)
{
    std::vector<double> x(n), g(n), x_lo(n), g_lo(n), g0(n);
    double fx0 = orc::evaluate<double>(objective, d0, d1, n, xp, g0);
    double dg0 = 0;
    for (long i = 0; i < n; i++) dg0 += g0[i] * drt[i];
    Core<double> core;
    int rc = core.init(opt, fx0, dg0, *step_io, step_max);
    *nfev_out = 0;
    if (rc) return rc;
    bool have_lo = false;
    for (;;)
    {
        for (long i = 0; i < n; i++) x[i] = xp[i] + core.step * drt[i];
        const double fx = orc::evaluate<double>(objective, d0, d1, n, x, g);
        if (*nfev_out < cap) trace[*nfev_out] = fx;
        (*nfev_out)++;
        double dg = 0;
        for (long i = 0; i < n; i++) dg += g[i] * drt[i];
        bool keep = false;
        rc = core.advance(fx, dg, keep);
        if (rc >= LSE_STEP_NOT_POSITIVE) return rc;
        if (keep) { x_lo.swap(x); g_lo.swap(g); have_lo = true; }
        if (rc == LSC_EVALUATE) continue;
        if (rc == LSC_ACCEPT) { *fx_out = fx; *dg_out = dg; }
        else
        {
            if (have_lo) { x.swap(x_lo); g.swap(g_lo); }
            else { std::memcpy(x.data(), xp, sizeof(double) * n); g = g0; }
            *fx_out = core.best_fx;
            *dg_out = core.best_dg;
        }
        *step_io = core.step;
        std::memcpy(x_out, x.data(), sizeof(double) * n);
        std::memcpy(grad_out, g.data(), sizeof(double) * n);
        return 0;
    }
}

extern "C" int core_line_search_f64(int objective, const double* d0, const double* d1, long n, int ls, int linesearch, int max_linesearch,
                                    double min_step, double max_step, double ftol, double wolfe, const double* xp, const double* drt,
                                    double step_max, double* step_io, double* fx_out, double* dg_out, double* x_out, double* grad_out,
                                    double* trace, long cap, long* nfev_out)
{
    LineSearchOptions<double> opt;
    opt.linesearch = (ls == 3) ? 3 : linesearch;
    opt.max_linesearch = max_linesearch;
    opt.min_step = min_step;
    opt.max_step = max_step;
    opt.ftol = ftol;
    opt.wolfe = wolfe;
    switch (ls)
    {
    case 0: return drive<BacktrackingCore>(objective, d0, d1, n, opt, xp, drt, step_max, step_io, fx_out, dg_out, x_out, grad_out, trace, cap, nfev_out);
    case 1: return drive<BracketingCore>(objective, d0, d1, n, opt, xp, drt, step_max, step_io, fx_out, dg_out, x_out, grad_out, trace, cap, nfev_out);
    case 2: return drive<NocedalWrightCore>(objective, d0, d1, n, opt, xp, drt, step_max, step_io, fx_out, dg_out, x_out, grad_out, trace, cap, nfev_out);
    default: return drive<MoreThuenteCore>(objective, d0, d1, n, opt, xp, drt, step_max, step_io, fx_out, dg_out, x_out, grad_out, trace, cap, nfev_out);
    }
}

extern "C" const char* core_error_message(int code) { return ls_error_message(code); }
extern "C" int core_error_kind(int code) { return ls_error_kind(code); }

// ---- host dense algebra of the L-BFGS-B middle matrices (include/LBFGSpp/BKLDLT.h, SmallDense.h) ----------------------------
#include "../../include/LBFGSpp/BKLDLT.h"

// a: n x n row-major (only the `uplo` triangle is read: 0 lower, 1 upper); b in, x out.  Returns info(), or -1 / -2 for the
// invalid_argument / logic_error paths (probe < 0: -1 = non-square input, -2 = solve before compute).
extern "C" int bkldlt_solve_f64(int n, const double* a, int uplo, const double* b, double* x, int probe)
{
    try
    {
        if (probe == -1) { BKLDLT<double> f(SmallMatrix<double>(2, 3)); return 0; }
        if (probe == -2) { BKLDLT<double> f; std::vector<double> v(2); f.solve_inplace(v); return 0; }
        SmallMatrix<double> m(n, n);
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) m(i, j) = a[size_t(i) * n + j];
        BKLDLT<double> f(m, uplo);
        std::vector<double> v(b, b + n);
        f.solve_inplace(v);
        std::copy(v.begin(), v.end(), x);
        return f.info();
    }
    catch (const std::invalid_argument&) { return -1; }
    catch (const std::logic_error&) { return -2; }
}

// ---- option structs (include/LBFGSpp/Param.h): what() of check_param(), "" when the parameters are valid ---------------------
#include "../../include/LBFGSpp/Param.h"

extern "C" const char* param_check_message(int lbfgsb, int m, double epsilon, double epsilon_rel, int past, double delta, int max_iterations,
                                           int linesearch, int max_submin, int max_linesearch, double min_step, double max_step, double ftol,
                                           double wolfe)
{
    static thread_local std::string msg;
    msg.clear();
    try
    {
        if (lbfgsb)
        {
            LBFGSBParam<double> p;
            p.m = m; p.epsilon = epsilon; p.epsilon_rel = epsilon_rel; p.past = past; p.delta = delta; p.max_iterations = max_iterations;
            p.max_submin = max_submin; p.max_linesearch = max_linesearch; p.min_step = min_step; p.max_step = max_step; p.ftol = ftol;
            p.wolfe = wolfe;
            p.check_param();
        }
        else
        {
            LBFGSParam<double> p;
            p.m = m; p.epsilon = epsilon; p.epsilon_rel = epsilon_rel; p.past = past; p.delta = delta; p.max_iterations = max_iterations;
            p.linesearch = linesearch; p.max_linesearch = max_linesearch; p.min_step = min_step; p.max_step = max_step; p.ftol = ftol;
            p.wolfe = wolfe;
            p.check_param();
        }
    }
    catch (const std::invalid_argument& e) { msg = e.what(); }
    return msg.c_str();
}
