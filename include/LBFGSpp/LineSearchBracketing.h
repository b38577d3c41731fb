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

#include "LineSearchCore.h"
#include "LineSearchDriver.h"
#include "Param.h"

namespace LBFGSpp {

template <typename Scalar>
class LineSearchBracketing
{
public:
    typedef DeviceVector<Scalar> Vector;

    // The decisions live in BracketingCore<Scalar> (LineSearchCore.h, shared with the device-resident solve); this adapter gives
    // them the reference's exceptions.
    class Machine : public CoreMachine<Scalar, BracketingCore>
    {
    public:
        Machine(const LBFGSParam<Scalar>& param, Scalar fx_init, Scalar dg_init, Scalar step0, Scalar step_max) :
            CoreMachine<Scalar, BracketingCore>(CoreMachine<Scalar, BracketingCore>::options_of(param, param.linesearch), fx_init, dg_init, step0, step_max) {}
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
