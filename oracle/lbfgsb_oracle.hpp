// lbfgsb_oracle.hpp -- TEST INFRASTRUCTURE ONLY.  Restatement of the bound-constrained solver
// (reference include/LBFGSB.h, LBFGSpp/{Cauchy,SubspaceMin,BKLDLT}.h, BFGSMat.h:99-146,307-615).
#ifndef LBFGSB_ORACLE_HPP
#define LBFGSB_ORACLE_HPP
#include "lbfgs_oracle.hpp"
namespace orc {
template <class T, class F>
LbfgsOutcome<T> lbfgsb_minimize(F&, const orc_param&, const Blas1<T>&, Vec<T>&, const Vec<T>&, const Vec<T>&)
{
    throw std::runtime_error("lbfgsb restatement not built yet (use the ref_ build)");
}
}  // namespace orc
#endif
