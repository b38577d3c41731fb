// LBFGSpp/SubspaceMin.h -- subspace minimisation step of LBFGSBSolver on the GPU.
//
// Same algorithm as the reference's SubspaceMin<Scalar>::subspace_minimize (reference include/LBFGSpp/SubspaceMin.h:122-302):
// minimise the quadratic model over the coordinates left free by the Cauchy point, first without bounds
// (inv(F'BF) via the compact representation), then -- if that point leaves the box -- with up to `maxit` sweeps of the
// BOXCQP active-set iteration, with the reference's three-level fallback.  The free / L / U / P index sets of the reference
// are bits of the per-coordinate class byte; each product with rows of W is one masked pass over the S/Y columns
// (BFGSMat::Wt_dot / lincomb / solve_PtBP); only 2c-vectors and 2c x 2c matrices are touched on the host.
#ifndef LBFGSPP_B200_SUBSPACE_MIN_H
#define LBFGSPP_B200_SUBSPACE_MIN_H

#include <limits>
#include <vector>

#include "BFGSMat.h"
#include "Cauchy.h"

namespace LBFGSpp {

template <typename Scalar>
class SubspaceMin
{
    typedef DeviceVector<Scalar> Vector;

    struct Stepper
    {
        Device& dev;
        lbfgs_b200_box* box;
        const Vector &x0, &g, &lb, &ub;
        Vector& drt;
        Scalar theta;
        // one element-wise pass; returns the three counters of reducing steps
        void run(int op, int flag = 0, Scalar* out3 = nullptr) const
        {
            dev.check(detail::BoxAbi<Scalar>::sub_step(box, op, flag, x0.data(), g.data(), lb.data(), ub.data(), drt.data(), theta, out3));
        }
        Scalar* vec(int which) const { return static_cast<Scalar*>(lbfgs_b200_box_vector(box, which)); }
    };

public:
    // In: the Cauchy point and classes in bfgs.box(), Wd = W'(xcp - x0).  Out: drt (device, n).
    static void subspace_minimize(BFGSMat<Scalar, true>& bfgs, const Vector& x0, const Vector& g, const Vector& lb, const Vector& ub,
                                  const CauchyResult<Scalar>& cp, int maxit, Vector& drt)
    {
        Device& dev = x0.device();
        drt.resize(x0.size());
        const Scalar theta = bfgs.theta();
        const int c = bfgs.ncorr();
        Stepper st{dev, bfgs.box(), x0, g, lb, ub, drt, theta};

        st.run(LBFGS_B200_SUB_INIT);                       // drt = xcp - x0; multipliers and work vectors cleared
        if (cp.nfree < 1) return;

        // vecc = F'B A A'd + g_F   (compute_FtBAb, BFGSMat.h:486-522; SubspaceMin.h:145-158)
        Scalar* vecc = st.vec(LBFGS_B200_BOXV_VECC);
        Scalar* vecy = st.vec(LBFGS_B200_BOXV_VECY);
        Scalar* tmp = st.vec(LBFGS_B200_BOXV_TMP);
        Scalar* tmp2 = st.vec(LBFGS_B200_BOXV_TMP2);
        if (c > 0 && cp.nact > 0)
        {
            st.run(LBFGS_B200_SUB_ACT_DIR);                // tmp = A'd
            const std::vector<Scalar> rhs = bfgs.Wt_dot(tmp);
            bfgs.minus_W_M(rhs, LBFGS_B200_CLS_FREE, vecc);
        }
        st.run(LBFGS_B200_SUB_ADD_G);

        // unconstrained minimiser over the free coordinates (SubspaceMin.h:160-170)
        st.run(LBFGS_B200_SUB_NEG_C_FREE);                 // tmp = -vecc on F
        bfgs.solve_PtBP(LBFGS_B200_CLS_FREE, tmp, vecy);
        Scalar cnt[3];
        st.run(LBFGS_B200_SUB_CHECK_BOUNDS, 0, cnt);
        if (cnt[0] == Scalar(0))
        {
            st.run(LBFGS_B200_SUB_WRITE_DRT, 0, cnt);
            return;
        }

        // BOXCQP sweeps (SubspaceMin.h:172-273)
        dev.check(lbfgs_b200_memcpy_d2d(dev.ctx(), st.vec(LBFGS_B200_BOXV_YFB), vecy, sizeof(Scalar) * size_t(x0.size())));
        int k;
        for (k = 0; k < maxit; k++)
        {
            st.run(LBFGS_B200_SUB_CLASSIFY, 0, cnt);
            const long nL = long(cnt[0]), nU = long(cnt[1]), nP = long(cnt[2]);
            if (nP > 0)
            {
                bool have_terms = false;
                if (c > 0 && (nL > 0 || nU > 0))
                {
                    st.run(LBFGS_B200_SUB_LU_VEC);         // tmp = l on L, u on U
                    const std::vector<Scalar> WQtv = bfgs.Wt_dot(tmp);
                    bfgs.minus_W_M(WQtv, LBFGS_B200_SUB_P, tmp2);   // P'B(L,U) contributions on P
                    have_terms = true;
                }
                st.run(LBFGS_B200_SUB_RHS_P, have_terms ? 1 : 0);   // tmp = -(vecc + terms) on P
                bfgs.solve_PtBP(LBFGS_B200_SUB_P, tmp, vecy);
            }
            if (nL > 0 || nU > 0)
            {
                st.run(LBFGS_B200_SUB_FREE_VEC);           // tmp = y on F
                const std::vector<Scalar> Fy = bfgs.Wt_dot(tmp);
                bfgs.minus_W_M(Fy, LBFGS_B200_SUB_L | LBFGS_B200_SUB_U, tmp2);
                st.run(LBFGS_B200_SUB_MULTIPLIERS);
            }
            st.run(LBFGS_B200_SUB_CONVERGED, 0, cnt);
            if (cnt[0] == Scalar(0) && cnt[1] == Scalar(0) && cnt[2] == Scalar(0)) break;
        }

        if (k >= maxit)  // SubspaceMin.h:276-296
        {
            const Scalar eps = std::numeric_limits<Scalar>::epsilon();
            st.run(LBFGS_B200_SUB_WRITE_DRT, 1, cnt);      // projected last iterate
            if (cnt[0] <= -eps) return;
            st.run(LBFGS_B200_SUB_WRITE_DRT, 3, cnt);      // projected unconstrained solution
            if (cnt[0] <= -eps) return;
            st.run(LBFGS_B200_SUB_WRITE_DRT, 2, cnt);      // unconstrained solution as is
            return;
        }
        st.run(LBFGS_B200_SUB_WRITE_DRT, 0, cnt);
    }
};

}  // namespace LBFGSpp

#endif  // LBFGSPP_B200_SUBSPACE_MIN_H
