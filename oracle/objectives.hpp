// objectives.hpp -- TEST INFRASTRUCTURE ONLY (shared by the _ref wrapper and the restatement).
//
// The benchmark objective functions, written once over any vector type that offers
// operator[] so that the reference build (Eigen-shim vectors) and the restatement
// (raw arrays) evaluate bit-identical f and grad.  Each follows a reference example:
//   rosenbrock_paired   examples/example-rosenbrock.cpp:15-27 (pairs (2i,2i+1) independent)
//   quad_shift          examples/example-quadratic.cpp:9-19   (f = |x-d|^2, d_i = i)
//   rosenbrock_chained  examples/example-rosenbrock-box.cpp:18-33
//   quad_tridiag        SURVEY.md 8d config C3 (the reference has no large SPD example):
//                       f = 1/2 x'Ax - b'x,  A = diag(d) + 1/2 tridiag(-1,2,-1)
// Every function works on an index range [lo, hi) and returns that range's contribution
// to f with a sequential "+=" exactly like the examples; evaluate() runs [0, n).
// (The multi-threaded CPU baseline sums per-thread ranges in thread order.)
#ifndef LBFGS_ORACLE_OBJECTIVES_HPP
#define LBFGS_ORACLE_OBJECTIVES_HPP

#include <cmath>
#include "oracle_api.h"

namespace orc {

template <class T, class VX, class VG>
T rosenbrock_paired(long lo, long hi, const VX& x, VG& g)  // lo even
{
    T fx = T(0);
    for (long i = lo; i < hi; i += 2)
    {
        const T t1 = T(1) - x[i];
        const T t2 = T(10) * (x[i + 1] - x[i] * x[i]);
        g[i + 1] = T(20) * t2;
        g[i] = T(-2) * (x[i] * g[i + 1] + t1);
        fx += t1 * t1 + t2 * t2;
    }
    return fx;
}

template <class T, class VX, class VG>
T quad_shift(long lo, long hi, const VX& x, VG& g)
{
    T fx = T(0);
    for (long i = lo; i < hi; i++)
    {
        const T r = x[i] - T(i);
        fx += r * r;
        g[i] = T(2) * r;
    }
    return fx;
}

template <class T, class VX, class VG>
T rosenbrock_chained(long lo, long hi, long n, const VX& x, VG& g)
{
    T fx = T(0);
    for (long i = lo; i < hi; i++)
    {
        if (i == 0)
        {
            fx += (x[0] - T(1)) * (x[0] - T(1));
            g[0] = T(2) * (x[0] - T(1)) + T(16) * (x[0] * x[0] - x[1]) * x[0];
            continue;
        }
        const T u = x[i] - x[i - 1] * x[i - 1];
        fx += T(4) * u * u;
        if (i == n - 1)
            g[i] = T(8) * u;
        else
            g[i] = T(8) * u + T(16) * (x[i] * x[i] - x[i + 1]) * x[i];
    }
    return fx;
}

template <class T, class VX, class VG>
T quad_tridiag(long lo, long hi, long n, const T* d, const T* b, const VX& x, VG& g)
{
    T fx = T(0);
    for (long i = lo; i < hi; i++)
    {
        const T xl = (i > 0) ? T(x[i - 1]) : T(0);
        const T xr = (i + 1 < n) ? T(x[i + 1]) : T(0);
        const T ax = (d[i] + T(1)) * x[i] - T(0.5) * (xl + xr);
        g[i] = ax - b[i];
        fx += x[i] * (T(0.5) * ax - b[i]);
    }
    return fx;
}

template <class T, class VX, class VG>
T evaluate_range(int objective, const T* data0, const T* data1, long lo, long hi, long n, const VX& x, VG& g)
{
    switch (objective)
    {
    case ORC_OBJ_ROSENBROCK_PAIRED: return rosenbrock_paired<T>(lo, hi, x, g);
    case ORC_OBJ_QUAD_SHIFT: return quad_shift<T>(lo, hi, x, g);
    case ORC_OBJ_ROSENBROCK_CHAINED: return rosenbrock_chained<T>(lo, hi, n, x, g);
    case ORC_OBJ_QUAD_TRIDIAG: return quad_tridiag<T>(lo, hi, n, data0, data1, x, g);
    }
    return T(std::nan(""));
}

template <class T, class VX, class VG>
T evaluate(int objective, const T* data0, const T* data1, long n, const VX& x, VG& g)
{
    return evaluate_range<T>(objective, data0, data1, 0, n, n, x, g);
}

}  // namespace orc
#endif
