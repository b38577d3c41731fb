// LBFGSpp/Cauchy.h -- generalized Cauchy point for LBFGSBSolver on the GPU.
//
// Same result as the reference's Cauchy<Scalar>::get_cauchy_point (reference include/LBFGSpp/Cauchy.h:86-284): the first
// local minimiser of the quadratic model along the projected steepest-descent path.  Structure here:
//   1. one kernel computes every breakpoint, the moving direction d and the coordinate classes      (reference :111-129)
//   2. the host examines the first segment with 2c-vectors only                                     (reference :146-168)
//   3. if the minimiser is not there, the device sorts the breakpoints, forms the prefix sums that the reference
//      accumulates coordinate by coordinate, and picks the first segment that holds its minimiser   (reference :183-256)
//   4. one kernel builds xcp and the free / newly-active classes                                     (reference :205-216, :258-283)
// Index sets are class bytes on the device (lbfgs_b200_box_classes), not std::vector<int>.
#ifndef LBFGSPP_B200_CAUCHY_H
#define LBFGSPP_B200_CAUCHY_H

#include <algorithm>
#include <limits>
#include <vector>

#include "BFGSMat.h"

#include "PhaseClock.h"

namespace LBFGSpp {

template <typename Scalar>
struct CauchyResult
{
    std::vector<Scalar> vecc;  // W'(xcp - x0), 2c values ([Y ; theta*S] convention)
    long nact;                 // coordinates newly fixed at a bound by the Cauchy path
    long nfree;                // coordinates left free
};

template <typename Scalar>
class Cauchy
{
    typedef DeviceVector<Scalar> Vector;

public:
    // xcp and the classes are left in the box workspace of `bfgs` (lbfgs_b200_box_xcp / _classes).
    static CauchyResult<Scalar> get_cauchy_point(BFGSMat<Scalar, true>& bfgs, const Vector& x0, const Vector& g, const Vector& lb,
                                                 const Vector& ub)
    {
        Device& dev = x0.device();
        lbfgs_b200_box* box = bfgs.box();
        const int c = bfgs.ncorr();
        const Scalar theta = bfgs.theta();
        const Scalar inf = std::numeric_limits<Scalar>::infinity();
        CauchyResult<Scalar> out;
        out.vecc.assign(size_t(2 * c), Scalar(0));

        Scalar info[5];
        PhaseClock::Scope ph_all(dev, "cauchy.total");
        {
        PhaseClock::Scope ph(dev, "cauchy.breakpoints");
        dev.check(detail::BoxAbi<Scalar>::cauchy_breaks(box, x0.data(), g.data(), lb.data(), ub.data(), info));
        }
        const long nfree_inf = long(info[1]), nord = long(info[2]);
        const Scalar dd = info[3], tmin = info[4];
        Scalar counts[2] = {0, 0};
        if (nfree_inf < 1 && nord < 1)
        {
            // every coordinate sits on a bound it cannot leave: xcp = x0 (reference :137-143)
            dev.check(detail::BoxAbi<Scalar>::cauchy_build(box, x0.data(), lb.data(), ub.data(), Scalar(-1), Scalar(0), counts));
            out.nact = 0;
            out.nfree = 0;
            return out;
        }

        // first segment, t in [0, first breakpoint)
        const Scalar* dvec = static_cast<const Scalar*>(lbfgs_b200_box_vector(box, LBFGS_B200_BOXV_DVEC));
        const std::vector<Scalar> p0 = bfgs.Wt_dot(dvec);
        Scalar fp = -dd;
        Scalar fpp = -theta * fp;
        if (c > 0)
        {
            const std::vector<Scalar> Mp = bfgs.apply_Mv(p0);
            for (int q = 0; q < 2 * c; q++) fpp -= p0[size_t(q)] * Mp[size_t(q)];
        }
        Scalar deltatmin = -fp / fpp;
        const Scalar deltat = (nord < 1) ? inf : tmin;

        Scalar t_cross = Scalar(-1), tfinal;
        if (!(deltatmin >= deltat))
        {
            const Scalar eps = std::numeric_limits<Scalar>::epsilon();
            if (fpp < eps) deltatmin = -fp / eps;
            deltatmin = std::max(deltatmin, Scalar(0));
            for (int q = 0; q < 2 * c; q++) out.vecc[size_t(q)] = deltatmin * p0[size_t(q)];
            tfinal = deltatmin;
        }
        else
        {
            std::vector<Scalar> res(size_t(5 + 2 * c));
            PhaseClock::Scope ph(dev, "cauchy.sort+sweep");
            dev.check(detail::BoxAbi<Scalar>::cauchy_sweep(box, g.data(), c > 0 ? bfgs.Mmat().data() : nullptr, c > 0 ? p0.data() : nullptr,
                                                           theta, dd, nord, nfree_inf, res.data()));
            t_cross = res[0];
            tfinal = res[1];
            for (int q = 0; q < 2 * c; q++) out.vecc[size_t(q)] = res[size_t(5 + q)];
        }
        dev.check(detail::BoxAbi<Scalar>::cauchy_build(box, x0.data(), lb.data(), ub.data(), t_cross, tfinal, counts));
        out.nact = long(counts[0]);
        out.nfree = long(counts[1]);
        return out;
    }
};

}  // namespace LBFGSpp

#endif  // LBFGSPP_B200_CAUCHY_H
