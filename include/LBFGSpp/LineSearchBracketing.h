// LBFGSpp/LineSearchBracketing.h -- bisection/doubling bracketing line search as a resumable state machine.
//
// Same decisions as the reference's LineSearchBracketing<Scalar>::LineSearch
// (reference include/LBFGSpp/LineSearchBracketing.h:48-128): keep an interval [lo, hi]; a failed Armijo
// test (or a non-finite f) lowers hi, a too-negative slope raises lo, a too-positive slope lowers hi; the next
// trial is 2*step while hi is infinite and the midpoint afterwards.
#ifndef LBFGSPP_B200_LINE_SEARCH_BRACKETING_H
#define LBFGSPP_B200_LINE_SEARCH_BRACKETING_H

#include <cmath>
#include <limits>
#include <stdexcept>

#include "LineSearchDriver.h"
#include "Param.h"

namespace LBFGSpp {

template <typename Scalar>
class LineSearchBracketing
{
public:
    typedef DeviceVector<Scalar> Vector;

    class Machine
    {
        const LBFGSParam<Scalar>& prm;
        Scalar f0, slope0, armijo_slope, lo, hi;
        int trials;

    public:
        Scalar step;
        Scalar best_fx, best_dg;  // unused

        Machine(const LBFGSParam<Scalar>& param, Scalar fx_init, Scalar dg_init, Scalar step0, Scalar /*step_max*/) :
            prm(param), f0(fx_init), slope0(dg_init), armijo_slope(param.ftol * dg_init), lo(0),
            hi(std::numeric_limits<Scalar>::infinity()), trials(0), step(step0), best_fx(fx_init), best_dg(dg_init)
        {
            if (step0 <= Scalar(0)) throw std::invalid_argument("'step' must be positive");
            if (dg_init > 0) throw std::logic_error("the moving direction increases the objective function value");
        }

        int advance(Scalar fx, Scalar dg, bool& /*keep*/)
        {
            if (fx > f0 + step * armijo_slope || !std::isfinite(fx))
                hi = step;
            else
            {
                if (prm.linesearch == LBFGS_LINESEARCH_BACKTRACKING_ARMIJO) return LS_ACCEPT;
                if (dg < prm.wolfe * slope0)
                    lo = step;
                else
                {
                    if (prm.linesearch == LBFGS_LINESEARCH_BACKTRACKING_WOLFE) return LS_ACCEPT;
                    if (dg > -prm.wolfe * slope0)
                        hi = step;
                    else
                        return LS_ACCEPT;
                }
            }
            if (lo > hi) throw std::runtime_error("the lower bound of the bracketing interval becomes larger than the upper bound");
            if (step < prm.min_step) throw std::runtime_error("the line search step became smaller than the minimum value allowed");
            if (step > prm.max_step) throw std::runtime_error("the line search step became larger than the maximum value allowed");
            step = std::isinf(hi) ? 2 * step : lo / 2 + hi / 2;
            if (++trials >= prm.max_linesearch) throw std::runtime_error("the line search routine reached the maximum number of iterations");
            return LS_EVALUATE;
        }
    };

    // Reference-compatible entry point (`dg` is an output only, LineSearchBracketing.h:60).
    template <typename Foo>
    static void LineSearch(Foo& f, const LBFGSParam<Scalar>& param, const Vector& xp, const Vector& drt, const Scalar& step_max,
                           Scalar& step, Scalar& fx, Vector& grad, Scalar& dg, Vector& x)
    {
        LineSearchWorkspace<Scalar> ws(xp.device());
        const Vector gradp(grad);
        dg = gradp.dot(drt);
        run_line_search<Machine>(f, param, xp, gradp, drt, step_max, step, fx, dg, x, grad, ws);
    }
};

}  // namespace LBFGSpp

#endif  // LBFGSPP_B200_LINE_SEARCH_BRACKETING_H
