// front_selfcheck.cpp -- TEST INFRASTRUCTURE.  Drives the product's host code (lbfgspp_b200/csrc/driver.cpp + the header-only front)
// on the host-memory test double (mock_abi.cpp) over a grid of small problems.  tests/test_front_on_mock_cpu.py builds the three
// files together under -fsanitize=address,undefined and runs the result: an out-of-bounds access, use-after-free, leak or undefined
// arithmetic in the front (DeviceVector lifetime and swaps, the line-search drivers, Cauchy / SubspaceMin bookkeeping, the dense
// 2m x 2m algebra) shows up as a sanitizer report.  Prints "front selfcheck ok <runs>".
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

extern "C" {
typedef struct
{
    int m;
    double epsilon, epsilon_rel;
    int past;
    double delta;
    int max_iterations, linesearch, max_submin, max_linesearch;
    double min_step, max_step, ftol, wolfe;
} drv_param;
typedef struct
{
    int status;
    char msg[200];
    int niter;
    long nfev;
    double fx, gnorm;
    long trace_len;
    double seconds, seconds_e2e;
    unsigned long long launches;
    long h2d_bytes, d2h_bytes;
} drv_result;
int lbfgsb200_drv_lbfgs_f64(int, int, const double*, const double*, long, int, const drv_param*, int, int, double*, double*, double*, long, drv_result*);
int lbfgsb200_drv_lbfgs_f32(int, int, const float*, const float*, long, int, const drv_param*, int, int, float*, float*, double*, long, drv_result*);
int lbfgsb200_drv_lbfgsb_f64(int, int, const double*, const double*, long, const drv_param*, double*, const double*, const double*, double*, double*,
                             long, drv_result*);
}

int main()
{
    int runs = 0, bad = 0;
    std::vector<double> trace(4096);
    for (long n : {2L, 10L, 24L, 100L, 1000L})
        for (int m : {1, 3, 6, 10})
        {
            drv_param p = {m, 1e-5, 1e-5, 0, 0.0, 300, 3, 10, 64, 1e-20, 1e20, 1e-4, 0.9};   // every run is bounded: fp32 with m = 1 stalls above the tolerance, in the reference too
            for (int ls = 0; ls < 4; ls++)
                for (int fused = 0; fused < 2; fused++)
                    for (int past : {0, 2})
                    {
                        p.past = past;
                        p.delta = past ? 1e-10 : 0.0;
                        std::vector<double> x(n, 0.0), g(n);
                        drv_result r;
                        if (getenv("FRONT_SELFCHECK_VERBOSE")) std::fprintf(stderr, "lbfgs n=%ld m=%d ls=%d fused=%d past=%d\n", n, m, ls, fused, past);
                        lbfgsb200_drv_lbfgs_f64(0, 0, nullptr, nullptr, n, ls, &p, 0, fused, x.data(), g.data(), trace.data(), 4096, &r);
                        runs++;
                        if (r.status != 0 || std::fabs(x[0] - 1.0) > 1e-2) bad++;   // m = 1 converges slowly but does get there in fp64
                    }
            {
                std::vector<float> x(n, 0.f), g(n);
                drv_result r;
                p.past = 0;
                lbfgsb200_drv_lbfgs_f32(0, 0, nullptr, nullptr, n, 3, &p, 0, 1, x.data(), g.data(), trace.data(), 4096, &r);
                runs++;
                if (r.status != 0) bad++;
            }
            // bound-constrained: mixed finite / infinite / degenerate bounds, 0..10 BOXCQP sweeps
            drv_param q = {m, 1e-5, 1e-5, 1, 1e-10, 300, 3, 10, 20, 1e-20, 1e20, 1e-4, 0.9};
            std::vector<double> lb(n, 2.0), ub(n, 4.0);
            for (long i = 0; i < n; i += 3) lb[i] = -INFINITY;
            for (long i = 2; i < n; i += 7) ub[i] = INFINITY;
            for (long i = 1; i < n; i += 5)
                if (std::isfinite(lb[i])) ub[i] = lb[i];
            for (int objective : {0, 2, 1})
                for (int submin : {0, 1, 10})
                {
                    q.max_submin = submin;
                    std::vector<double> x(n, 3.0), g(n);
                    drv_result r;
                    if (getenv("FRONT_SELFCHECK_VERBOSE")) std::fprintf(stderr, "box n=%ld m=%d obj=%d submin=%d\n", n, m, objective, submin);
                    lbfgsb200_drv_lbfgsb_f64(0, objective, nullptr, nullptr, n, &q, x.data(), lb.data(), ub.data(), g.data(), trace.data(), 4096, &r);
                    runs++;
                    if (r.status == 1 || r.status == 4 || (r.status == 0 && !std::isfinite(r.fx))) bad++;
                }
        }
    // error paths unwind through the front as well
    {
        drv_param p = {0, 1e-5, 1e-5, 0, 0.0, 0, 3, 10, 20, 1e-20, 1e20, 1e-4, 0.9};
        std::vector<double> x(4, 0.0), g(4);
        drv_result r;
        lbfgsb200_drv_lbfgs_f64(0, 0, nullptr, nullptr, 4, 3, &p, 0, 1, x.data(), g.data(), trace.data(), 4096, &r);
        runs++;
        if (r.status != 1) bad++;
        p.m = 6;
        p.max_linesearch = 1;
        lbfgsb200_drv_lbfgs_f64(0, 0, nullptr, nullptr, 4, 0, &p, 0, 1, x.data(), g.data(), trace.data(), 4096, &r);
        runs++;
        if (r.status != 3) bad++;
    }
    std::printf("front selfcheck %s %d\n", bad ? "FAILED" : "ok", runs);
    return bad ? 1 : 0;
}
