// LBFGSpp/LineSearchMoreThuente.h -- More-Thuente strong-Wolfe search as a resumable state machine.
//
// Same decisions as the reference's LineSearchMoreThuente<Scalar>::LineSearch
// (reference include/LBFGSpp/LineSearchMoreThuente.h:213-615) including its simplifications: the search works on
// the auxiliary function psi(a) = f(a) - f(0) - a*ftol*f'(0) throughout (the switch to f itself is commented out
// in the reference, :455-462), trial values are chosen by the safeguarded cubic/quadratic rules of
// step_selection (:116-189), an un-bracketed interval is extrapolated by 1.1, a bracketed one that fails to
// shrink by 0.66 twice in a row is bisected, and while no acceptable-decrease point has been seen the next
// step is capped at 7/12 of the current one ("step_min safeguard").
// All arithmetic below is host scalar work; the vector work of each trial lives in LineSearchDriver.h.
#ifndef LBFGSPP_B200_LINE_SEARCH_MORE_THUENTE_H
#define LBFGSPP_B200_LINE_SEARCH_MORE_THUENTE_H

#include <algorithm>
#include <cmath>
#include <limits>
#include <stdexcept>

#include "LineSearchDriver.h"
#include "Param.h"

namespace LBFGSpp {

template <typename Scalar>
class LineSearchMoreThuente
{
public:
    typedef DeviceVector<Scalar> Vector;

    // One end point (or the trial point) of the search interval: abscissa, psi value, psi slope.
    struct Sample
    {
        Scalar at, f, g;
    };

    // Interpolating polynomials through two samples (reference :34-114).
    struct Interpolation
    {
        // minimiser of the quadratic matching f and f' at p and f at q
        static Scalar quadratic_from_values(const Sample& p, const Sample& q)
        {
            const Scalar span = q.at - p.at;
            const Scalar w = Scalar(0.5) * span * p.g / (p.f - q.f + span * p.g);
            return p.at + w * span;
        }
        // minimiser of the quadratic matching f' at p and at q (secant step)
        static Scalar quadratic_from_slopes(const Sample& p, const Sample& q)
        {
            const Scalar w = p.g / (p.g - q.g);
            return p.at + w * (q.at - p.at);
        }
        // local minimiser of the cubic matching f and f' at p and q; found = false when there is none
        static Scalar cubic(const Sample& p, const Sample& q, bool& found)
        {
            using std::abs;
            using std::sqrt;
            const Scalar a = p.at, b = q.at;
            const Scalar sum = a + b, span = b - a, span2 = span * span;
            const Scalar df = q.f - p.f, dgr = q.g - p.g;
            // derivative of the cubic is (c3*x^2 + c2*x + c1) up to a positive factor
            const Scalar c3 = (p.g + q.g) * span - Scalar(2) * df;
            const Scalar c2 = Scalar(0.5) * (dgr * span2 - Scalar(3) * sum * c3);
            const Scalar c1 = df * span2 - sum * c2 - (a * sum + b * b) * c3;
            const Scalar tiny = std::numeric_limits<Scalar>::epsilon();
            if (abs(c3) < tiny * abs(c2) || abs(c3) < tiny * abs(c1))
            {
                // degenerates to a parabola
                found = (c2 * span > Scalar(0));
                return found ? (-Scalar(0.5) * c1 / c2) : b;
            }
            const Scalar u = c2 / (Scalar(3) * c3), v = c1 / c2;
            const Scalar ratio = v / u;
            found = (ratio <= Scalar(1));
            if (!found) return b;
            Scalar root_a, root_b;
            if (abs(u) >= abs(v))
            {
                const Scalar w = Scalar(1) + sqrt(Scalar(1) - ratio);
                root_a = -u * w;
                root_b = -v / w;
            }
            else
            {
                const Scalar disc = sqrt(abs(u)) * sqrt(abs(v)) * sqrt(1 - u / v);
                root_a = -u - disc;
                root_b = -u + disc;
            }
            return (c3 * span > Scalar(0)) ? (std::max)(root_a, root_b) : (std::min)(root_a, root_b);
        }
    };

    // Safeguarded choice of the next trial inside/outside [lo, hi] given the trial sample t (reference :116-189).
    static Scalar select_step(const Sample& lo, const Sample& hi, const Sample& t)
    {
        using std::abs;
        if (lo.at == hi.at) return lo.at;
        if (!std::isfinite(t.f) || !std::isfinite(t.g)) return (lo.at + t.at) / Scalar(2);
        bool cubic_ok;
        const Scalar ac = Interpolation::cubic(lo, t, cubic_ok);
        const Scalar aq = Interpolation::quadratic_from_values(lo, t);
        if (t.f > lo.f)  // higher value: the minimum is bracketed by lo and t
        {
            if (!cubic_ok) return aq;
            return (abs(ac - lo.at) < abs(aq - lo.at)) ? ac : ((aq + ac) / Scalar(2));
        }
        const Scalar as = Interpolation::quadratic_from_slopes(lo, t);
        if (t.g * lo.g < Scalar(0))  // lower value, slopes of opposite sign
            return (abs(ac - t.at) >= abs(as - t.at)) ? ac : as;
        const Scalar extrapolate = Scalar(1.1), toward_hi = Scalar(0.66);
        if (abs(t.g) < abs(lo.g))  // lower value, same sign, slope magnitude decreases
        {
            const bool use_cubic = cubic_ok && (ac - t.at) * (t.at - lo.at) > Scalar(0) && abs(ac - t.at) < abs(as - t.at);
            const Scalar pick = use_cubic ? ac : as;
            const Scalar limit = t.at + toward_hi * (hi.at - t.at);
            return (t.at > lo.at) ? (std::min)(limit, pick) : (std::max)(limit, pick);
        }
        // lower value, same sign, slope magnitude does not decrease
        if (!std::isfinite(hi.at) || !std::isfinite(hi.f) || !std::isfinite(hi.g)) return t.at + extrapolate * (t.at - lo.at);
        bool unused;
        const Scalar ae = Interpolation::cubic(t, hi, unused);
        const Scalar limit = t.at + toward_hi * (hi.at - t.at);
        return (t.at > lo.at) ? (std::min)(limit, ae) : (std::max)(limit, ae);
    }

    class Machine
    {
        Scalar smin, smax, f0, decrease_slope, curvature_bound;
        Sample lo, hi;       // interval end points in terms of psi
        Scalar psi_lo;       // psi at lo (== lo.f while the search stays on psi)
        bool bracketed, cap_next_step;
        Scalar width, width_before;
        int stalls, trials, budget;

    public:
        Scalar step;
        Scalar best_fx, best_dg;  // f and f' at lo (the point handed back when the budget runs out)

        template <class Param>
        Machine(const Param& param, Scalar fx_init, Scalar dg_init, Scalar step0, Scalar step_max) :
            smin(param.min_step), smax(step_max), f0(fx_init), decrease_slope(param.ftol * dg_init),
            curvature_bound(-param.wolfe * dg_init), psi_lo(0), bracketed(false), cap_next_step(param.min_step > Scalar(0)),
            width(std::numeric_limits<Scalar>::infinity()), width_before(std::numeric_limits<Scalar>::infinity()),
            stalls(0), trials(0), budget(param.max_linesearch), step(step0), best_fx(fx_init), best_dg(dg_init)
        {
            if (step0 <= Scalar(0)) throw std::invalid_argument("'step' must be positive");
            if (step0 < smin) throw std::invalid_argument("'step' is smaller than 'param.min_step'");
            if (step0 > smax) throw std::invalid_argument("'step' exceeds 'step_max'");
            if (dg_init >= Scalar(0)) throw std::logic_error("the moving direction does not decrease the objective function value");
            const Scalar inf = std::numeric_limits<Scalar>::infinity();
            lo.at = Scalar(0);
            lo.f = Scalar(0);
            lo.g = (Scalar(1) - param.ftol) * dg_init;
            hi.at = hi.f = hi.g = inf;
        }

        int advance(Scalar fx, Scalar dg, bool& keep)
        {
            using std::abs;
            const Scalar inf = std::numeric_limits<Scalar>::infinity();
            const Scalar psi = fx - f0 - step * decrease_slope;
            const Scalar dpsi = dg - decrease_slope;

            if (psi <= Scalar(0) && abs(dg) <= curvature_bound) return LS_ACCEPT;          // strong Wolfe
            if (step <= smin && (psi > Scalar(0) || dpsi >= Scalar(0))) return LS_ACCEPT;   // stuck at the lower bound
            if (step >= smax && (psi <= Scalar(0) && dpsi < Scalar(0))) return LS_ACCEPT;   // stuck at the upper bound

            const Sample t = {step, psi, dpsi};
            if (cap_next_step && psi <= Scalar(0) && dpsi < Scalar(0)) cap_next_step = false;

            // lower value and the slope points away from lo: keep marching
            const bool marching = (psi <= psi_lo) && (dpsi * (lo.at - step) > Scalar(0));
            Scalar next;
            if (marching)
                next = (std::min)(smax, step + Scalar(1.1) * (step - lo.at));
            else
            {
                next = select_step(lo, hi, t);
                next = (std::max)(next, smin);
                next = (std::min)(next, smax);
                if (cap_next_step)
                {
                    const Scalar ceiling = (std::max)(smin, (Scalar(7) / Scalar(12)) * step);
                    next = (std::max)(next, smin);
                    next = (std::min)(next, ceiling);
                }
            }

            if (psi > psi_lo)
                hi = t;
            else
            {
                if (!marching) hi = lo;
                lo = t;
                psi_lo = psi;
                best_fx = fx;
                best_dg = dg;
                keep = true;
            }

            if (!bracketed && !marching)
            {
                const Scalar left = (std::min)(lo.at, hi.at), right = (std::max)(lo.at, hi.at);
                bracketed = (left >= smin && right <= smax);
            }
            if (bracketed)
            {
                width_before = width;
                width = abs(hi.at - lo.at);
                if (width_before < inf && width > Scalar(0.66) * width_before)
                    stalls += 1;
                else
                    stalls = 0;
                if (stalls >= 2)
                {
                    next = (lo.at + hi.at) / Scalar(2);
                    stalls = 0;
                }
            }
            step = next;
            if (++trials >= budget)
            {
                step = lo.at;
                return LS_TAKE_BEST;
            }
            return LS_EVALUATE;
        }
    };

    // Reference-compatible entry point (generic over the parameter struct so that LBFGSBSolver can use it):
    // `grad`/`dg` hold the gradient / slope at xp on entry.
    template <typename Foo, typename SolverParam>
    static void LineSearch(Foo& f, const SolverParam& param, const Vector& xp, const Vector& drt, const Scalar& step_max,
                           Scalar& step, Scalar& fx, Vector& grad, Scalar& dg, Vector& x)
    {
        LineSearchWorkspace<Scalar> ws(xp.device());
        const Vector gradp(grad);
        run_line_search<Machine>(f, param, xp, gradp, drt, step_max, step, fx, dg, x, grad, ws);
    }
};

}  // namespace LBFGSpp

#endif  // LBFGSPP_B200_LINE_SEARCH_MORE_THUENTE_H
