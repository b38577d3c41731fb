// LBFGSpp/LineSearchBacktracking.h -- backtracking line search as a resumable state machine.
//
// Same decisions as the reference's LineSearchBacktracking<Scalar>::LineSearch
// (reference include/LBFGSpp/LineSearchBacktracking.h:44-121): shrink the step by 0.5 when the Armijo test
// fails (or f is NaN), grow it by 2.1 when the curvature is still too negative, stop according to
// param.linesearch.  The vector work of each trial lives in LineSearchDriver.h.
#ifndef LBFGSPP_B200_LINE_SEARCH_BACKTRACKING_H
#define LBFGSPP_B200_LINE_SEARCH_BACKTRACKING_H

#include <stdexcept>

#include "LineSearchDriver.h"
#include "Param.h"

namespace LBFGSpp {

template <typename Scalar>
class LineSearchBacktracking
{
public:
    typedef DeviceVector<Scalar> Vector;

    class Machine
    {
        const LBFGSParam<Scalar>& prm;
        Scalar f0, slope0, armijo_slope;
        int trials;

    public:
        Scalar step;
        Scalar best_fx, best_dg;  // unused: this search never falls back to a remembered point

        Machine(const LBFGSParam<Scalar>& param, Scalar fx_init, Scalar dg_init, Scalar step0, Scalar /*step_max*/) :
            prm(param), f0(fx_init), slope0(dg_init), armijo_slope(param.ftol * dg_init), trials(0), step(step0),
            best_fx(fx_init), best_dg(dg_init)
        {
            if (step0 <= Scalar(0)) throw std::invalid_argument("'step' must be positive");
            if (dg_init > 0) throw std::logic_error("the moving direction increases the objective function value");
        }

        int advance(Scalar fx, Scalar dg, bool& /*keep*/)
        {
            const Scalar shrink = Scalar(0.5), grow = Scalar(2.1);
            Scalar factor;
            const bool armijo_fails = (fx > f0 + step * armijo_slope) || (fx != fx);
            if (armijo_fails)
                factor = shrink;
            else
            {
                if (prm.linesearch == LBFGS_LINESEARCH_BACKTRACKING_ARMIJO) return LS_ACCEPT;
                if (dg < prm.wolfe * slope0)
                    factor = grow;
                else
                {
                    if (prm.linesearch == LBFGS_LINESEARCH_BACKTRACKING_WOLFE) return LS_ACCEPT;
                    if (dg > -prm.wolfe * slope0)
                        factor = shrink;
                    else
                        return LS_ACCEPT;
                }
            }
            if (step < prm.min_step) throw std::runtime_error("the line search step became smaller than the minimum value allowed");
            if (step > prm.max_step) throw std::runtime_error("the line search step became larger than the maximum value allowed");
            step *= factor;
            if (++trials >= prm.max_linesearch) throw std::runtime_error("the line search routine reached the maximum number of iterations");
            return LS_EVALUATE;
        }
    };

    // Reference-compatible entry point.  `grad` holds the gradient at xp on entry and at x on return;
    // `dg` is an output only (the reference recomputes grad.dot(drt) itself, LineSearchBacktracking.h:60).
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

#endif  // LBFGSPP_B200_LINE_SEARCH_BACKTRACKING_H
