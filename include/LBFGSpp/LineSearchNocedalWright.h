// LBFGSpp/LineSearchNocedalWright.h -- strong-Wolfe bracket + zoom search as a resumable state machine.
//
// Same decisions as the reference's LineSearchNocedalWright<Scalar>::LineSearch
// (reference include/LBFGSpp/LineSearchNocedalWright.h:84-279, interpolation helper :30-60): a bracketing phase
// that doubles the step until the minimum is bracketed, then a zoom phase that interpolates quadratically
// between the low point (value + slope) and the high point (value).  Ignores step_max / min_step / max_step
// exactly like the reference does.
#ifndef LBFGSPP_B200_LINE_SEARCH_NOCEDAL_WRIGHT_H
#define LBFGSPP_B200_LINE_SEARCH_NOCEDAL_WRIGHT_H

#include <algorithm>
#include <cmath>
#include <stdexcept>

#include "LineSearchDriver.h"
#include "Param.h"

namespace LBFGSpp {

template <typename Scalar>
class LineSearchNocedalWright
{
public:
    typedef DeviceVector<Scalar> Vector;

    class Machine
    {
        enum Phase { BRACKET, ZOOM };
        const LBFGSParam<Scalar>& prm;
        Scalar f0, decrease_slope, curvature_bound;
        Scalar lo, hi, f_lo, f_hi, slope_lo;
        Phase phase;
        int budget_used;

        // minimiser of the parabola through (lo, f_lo) with slope slope_lo and (hi, f_hi); bisect when it is
        // not finite, outside the interval or within 1% of an end point
        Scalar interpolate() const
        {
            using std::abs;
            const Scalar df = f_hi - f_lo, ds = hi - lo, mid = (hi + lo) / Scalar(2);
            Scalar cand = df * lo - mid * ds * slope_lo;
            cand = cand / (df - ds * slope_lo);
            const bool useless = !std::isfinite(cand);
            const Scalar margin = std::min(abs(cand - lo), abs(cand - hi));
            const bool hugging = margin < Scalar(0.01) * abs(ds);
            const bool bisect = useless || cand <= std::min(lo, hi) || cand >= std::max(lo, hi) || hugging;
            return bisect ? mid : cand;
        }

        void remember(Scalar fx, Scalar dg, bool& keep)
        {
            lo = step;
            f_lo = fx;
            slope_lo = dg;
            best_fx = fx;
            best_dg = dg;
            keep = true;
        }

    public:
        Scalar step;
        Scalar best_fx, best_dg;

        Machine(const LBFGSParam<Scalar>& param, Scalar fx_init, Scalar dg_init, Scalar step0, Scalar /*step_max*/) :
            prm(param), f0(fx_init), decrease_slope(param.ftol * dg_init), curvature_bound(-param.wolfe * dg_init),
            lo(0), hi(0), f_lo(fx_init), f_hi(0), slope_lo(dg_init), phase(BRACKET), budget_used(0), step(step0),
            best_fx(fx_init), best_dg(dg_init)
        {
            if (step0 <= Scalar(0)) throw std::invalid_argument("'step' must be positive");
            if (param.linesearch != LBFGS_LINESEARCH_BACKTRACKING_STRONG_WOLFE)
                throw std::invalid_argument("'param.linesearch' must be 'LBFGS_LINESEARCH_BACKTRACKING_STRONG_WOLFE' for LineSearchNocedalWright");
            if (dg_init > Scalar(0)) throw std::logic_error("the moving direction increases the objective function value");
        }

        int advance(Scalar fx, Scalar dg, bool& keep)
        {
            using std::abs;
            const bool too_high = fx - f0 > step * decrease_slope;
            if (phase == BRACKET)
            {
                if (too_high || (Scalar(0) < lo && fx >= f_lo))
                {
                    hi = step;
                    f_hi = fx;
                    phase = ZOOM;
                    step = interpolate();
                    return LS_EVALUATE;
                }
                if (abs(dg) <= curvature_bound) return LS_ACCEPT;
                hi = lo;
                f_hi = f_lo;
                remember(fx, dg, keep);
                if (dg >= Scalar(0))
                {
                    phase = ZOOM;
                    step = interpolate();
                    return LS_EVALUATE;
                }
                if (++budget_used >= prm.max_linesearch) return LS_TAKE_BEST;  // best == the trial just kept
                step *= Scalar(2);
                return LS_EVALUATE;
            }
            // zoom phase
            if (too_high || fx >= f_lo)
            {
                if (step == hi) throw std::runtime_error("the line search routine failed, possibly due to insufficient numeric precision");
                hi = step;
                f_hi = fx;
            }
            else
            {
                if (abs(dg) <= curvature_bound) return LS_ACCEPT;
                if (dg * (hi - lo) >= Scalar(0))
                {
                    hi = lo;
                    f_hi = f_lo;
                }
                if (step == lo) throw std::runtime_error("the line search routine failed, possibly due to insufficient numeric precision");
                remember(fx, dg, keep);
            }
            if (++budget_used >= prm.max_linesearch)
            {
                if (lo <= Scalar(0)) throw std::runtime_error("the line search routine failed, unable to sufficiently decrease the function value");
                step = lo;
                return LS_TAKE_BEST;
            }
            step = interpolate();
            return LS_EVALUATE;
        }
    };

    // Reference-compatible entry point: `grad`/`dg` hold the gradient / slope at xp on entry.
    template <typename Foo>
    static void LineSearch(Foo& f, const LBFGSParam<Scalar>& param, const Vector& xp, const Vector& drt, const Scalar& step_max,
                           Scalar& step, Scalar& fx, Vector& grad, Scalar& dg, Vector& x)
    {
        LineSearchWorkspace<Scalar> ws(xp.device());
        const Vector gradp(grad);
        run_line_search<Machine>(f, param, xp, gradp, drt, step_max, step, fx, dg, x, grad, ws);
    }
};

}  // namespace LBFGSpp

#endif  // LBFGSPP_B200_LINE_SEARCH_NOCEDAL_WRIGHT_H
