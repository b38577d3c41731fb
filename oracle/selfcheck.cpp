// selfcheck.cpp -- TEST INFRASTRUCTURE ONLY.  Drives the restatement (lbfgs_oracle.hpp / lbfgsb_oracle.hpp through its C API)
// over a grid of small problems; `make -C oracle sanitize` builds it together with oracle_capi.cpp under
// -fsanitize=address,undefined and runs it, so that an out-of-bounds access or undefined arithmetic in the checker itself shows
// up as a sanitizer report (tests/test_oracle_cpu.py::test_restatement_is_sanitizer_clean).  Prints "selfcheck ok <runs>".
#include <cmath>
#include <cstdio>
#include <vector>

#include "oracle_api.h"

int main()
{
    orc_param p;
    orc_result r;
    int runs = 0, bad = 0;
    for (int n : {2, 10, 25, 100, 1000})
        for (int m : {1, 3, 6, 10})
        {
            std::vector<double> x(n, 3.0), lb(n, 2.0), ub(n, 4.0), g(n), trace(1000);
            for (int i = 0; i < n; i += 3) lb[i] = -INFINITY;
            for (int i = 2; i < n; i += 7) ub[i] = INFINITY;
            for (int i = 1; i < n; i += 5)
                if (std::isfinite(lb[i])) ub[i] = lb[i];
            orc_default_param(&p, 1);
            p.m = m;
            for (int objective : {ORC_OBJ_ROSENBROCK_PAIRED, ORC_OBJ_ROSENBROCK_CHAINED, ORC_OBJ_QUAD_SHIFT})
            {
                if (objective == ORC_OBJ_ROSENBROCK_PAIRED && n % 2) continue;
                for (int submin : {0, 1, 10})
                {
                    p.max_submin = submin;
                    std::vector<double> xx = x;
                    orc_lbfgsb_f64(objective, nullptr, nullptr, n, &p, ORC_SUM_SEQUENTIAL, xx.data(), lb.data(), ub.data(), g.data(),
                                   trace.data(), 1000, &r);
                    runs++;
                    if (r.status != ORC_OK && r.status != ORC_RUNTIME_ERROR && r.status != ORC_LOGIC_ERROR) bad++;
                    if (r.status == ORC_OK && !std::isfinite(r.fx)) bad++;
                }
            }
            if (n % 2) continue;
            orc_default_param(&p, 0);
            p.m = m;
            for (int ls = 0; ls < 4; ls++)
                for (int mode : {ORC_SUM_SEQUENTIAL, ORC_SUM_LANES8, ORC_SUM_LANES8_OMP})
                {
                    std::vector<double> xx(n, 0.0);
                    orc_lbfgs_f64(ORC_OBJ_ROSENBROCK_PAIRED, nullptr, nullptr, n, ls, &p, mode, xx.data(), g.data(), trace.data(), 1000, &r);
                    runs++;
                    if (r.status != ORC_OK) bad++;
                    std::vector<double> xg(n, 0.0);
                    orc_lbfgs_gram_f64(ORC_OBJ_ROSENBROCK_PAIRED, nullptr, nullptr, n, ls, &p, mode, xg.data(), g.data(), trace.data(), 1000, &r);
                    runs++;
                    if (r.status != ORC_OK) bad++;
                }
        }
    std::printf("selfcheck %s %d\n", bad ? "FAILED" : "ok", runs);
    return bad ? 1 : 0;
}
