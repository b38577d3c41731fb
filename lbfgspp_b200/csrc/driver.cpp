// driver.cpp -- host-buffer entry points over the header-only C++ front (include/LBFGS.h), built into
// lbfgspp_b200/liblbfgs_b200_driver.so.  This is the "reference-facing call with HOST buffers" that tests and
// bench.py drive through ctypes: the caller hands over x0 in host memory, the driver uploads it, runs
// LBFGSpp::LBFGSSolver<Scalar, LineSearch>::minimize() on the GPU and downloads x / grad.  The argument structs
// have the same layout as the CPU checker's (oracle/oracle_api.h) so that a parity test calls both sides with the
// very same objects -- but nothing here includes or links anything from oracle/.
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>

#include "../../include/LBFGS.h"
#include "../../include/LBFGSpp/DeviceObjectives.h"

using namespace LBFGSpp;

extern "C" {

typedef struct
{
    int m;
    double epsilon, epsilon_rel;
    int past;
    double delta;
    int max_iterations;
    int linesearch;
    int max_submin;
    int max_linesearch;
    double min_step, max_step, ftol, wolfe;
} drv_param;

typedef struct
{
    int status;             // 0 ok, 1 invalid_argument, 2 logic_error, 3 runtime_error, 4 other
    char msg[200];
    int niter;
    long nfev;
    double fx;
    double gnorm;
    long trace_len;
    double seconds;         // minimize() alone, operands already in HBM (host wall clock around a synchronised call)
    double seconds_e2e;     // upload of x0 (+data) + minimize() + download of x
    unsigned long long launches;  // kernels launched by minimize()
    long h2d_bytes, d2h_bytes;
} drv_result;

enum { DRV_LS_BACKTRACKING = 0, DRV_LS_BRACKETING = 1, DRV_LS_NOCEDAL_WRIGHT = 2, DRV_LS_MORE_THUENTE = 3 };

}  // extern "C"

namespace {

std::map<int, std::shared_ptr<Device> >& devices()
{
    static std::map<int, std::shared_ptr<Device> > m;
    return m;
}
Device& device(int ordinal)
{
    std::shared_ptr<Device>& d = devices()[ordinal];
    if (!d) d = std::make_shared<Device>(ordinal);
    return *d;
}

// Records f of every evaluation (the reference offers the functor as its only observation point,
// examples/example-rosenbrock-comparison.cpp:14-41) while forwarding the fused hooks when Base has them.
template <class Base, class Scalar, bool Fused>
struct Traced;

template <class Base, class Scalar>
struct Traced<Base, Scalar, true>
{
    typedef DeviceVector<Scalar> Vector;
    Base& base;
    double* trace;
    long cap, count;
    Traced(Base& b, double* t, long c) : base(b), trace(t), cap(c), count(0) {}
    void note(Scalar fx) { if (trace && count < cap) trace[count] = double(fx); count++; }
    Scalar operator()(const Vector& x, Vector& g) { const Scalar fx = base(x, g); note(fx); return fx; }
    void fused_value(const Vector& x, Vector& g, Scalar* o) { base.fused_value(x, g, o); note(o[0]); }
    void fused_trial(const Vector& xp, const Vector& d, Scalar s, Vector& x, Vector& g, Scalar* o) { base.fused_trial(xp, d, s, x, g, o); note(o[0]); }
};
template <class Base, class Scalar>
struct Traced<Base, Scalar, false>
{
    typedef DeviceVector<Scalar> Vector;
    Base& base;
    double* trace;
    long cap, count;
    Traced(Base& b, double* t, long c) : base(b), trace(t), cap(c), count(0) {}
    Scalar operator()(const Vector& x, Vector& g)
    {
        const Scalar fx = base(x, g);
        if (trace && count < cap) trace[count] = double(fx);
        count++;
        return fx;
    }
};

void set_error(drv_result* out, int code, const char* what)
{
    out->status = code;
    std::strncpy(out->msg, what, sizeof(out->msg) - 1);
    out->msg[sizeof(out->msg) - 1] = 0;
}

template <class Body>
int guarded(drv_result* out, Body body)
{
    std::memset(out, 0, sizeof(*out));
    try { body(); }
    catch (const std::invalid_argument& e) { set_error(out, 1, e.what()); }
    catch (const std::logic_error& e) { set_error(out, 2, e.what()); }
    catch (const std::runtime_error& e) { set_error(out, 3, e.what()); }
    catch (const std::exception& e) { set_error(out, 4, e.what()); }
    return out->status;
}

template <class Scalar>
LBFGSParam<Scalar> to_param(const drv_param* q)
{
    LBFGSParam<Scalar> p;
    p.m = q->m;
    p.epsilon = Scalar(q->epsilon);
    p.epsilon_rel = Scalar(q->epsilon_rel);
    p.past = q->past;
    p.delta = Scalar(q->delta);
    p.max_iterations = q->max_iterations;
    p.linesearch = q->linesearch;
    p.max_linesearch = q->max_linesearch;
    p.min_step = Scalar(q->min_step);
    p.max_step = Scalar(q->max_step);
    p.ftol = Scalar(q->ftol);
    p.wolfe = Scalar(q->wolfe);
    return p;
}

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <class Scalar, template <class> class LS, class Objective, bool Fused>
void solve_with(Device& dev, Objective& obj, const drv_param* q, int hv_algo, long n, Scalar* x_host, Scalar* grad_host,
                double* trace, long cap, drv_result* out, double t_begin, long h2d_extra)
{
    typedef DeviceVector<Scalar> Vector;
    const LBFGSParam<Scalar> prm = to_param<Scalar>(q);
    LBFGSSolver<Scalar, LS> solver(prm);
    solver.set_hv_algorithm(hv_algo);
    Traced<Objective, Scalar, Fused> f(obj, trace, cap);
    Vector x(dev);
    x.copy_from_host(x_host, n);
    Scalar fx = Scalar(0);
    const unsigned long long launches0 = lbfgs_b200_launch_count(dev.ctx());
    const double t0 = now();
    int niter = 0;
    try { niter = solver.minimize(f, x, fx); }
    catch (...)
    {
        out->nfev = f.count;
        out->trace_len = f.count < cap ? f.count : cap;
        x.copy_to_host(x_host);
        throw;
    }
    dev.synchronize();
    const double t1 = now();
    x.copy_to_host(x_host);
    if (grad_host) solver.final_grad().copy_to_host(grad_host);
    const double t2 = now();
    out->niter = niter;
    out->fx = double(fx);
    out->gnorm = double(solver.final_grad_norm());
    out->nfev = f.count;
    out->trace_len = f.count < cap ? f.count : cap;
    out->seconds = t1 - t0;
    out->seconds_e2e = t2 - t_begin;
    out->launches = lbfgs_b200_launch_count(dev.ctx()) - launches0;
    out->h2d_bytes = long(sizeof(Scalar)) * n + h2d_extra;
    out->d2h_bytes = long(sizeof(Scalar)) * n * (grad_host ? 2 : 1);
}

template <class Scalar, class Objective, bool Fused>
void solve_ls(int ls, Device& dev, Objective& obj, const drv_param* q, int hv_algo, long n, Scalar* x_host, Scalar* grad_host,
              double* trace, long cap, drv_result* out, double t_begin, long h2d_extra)
{
    switch (ls)
    {
    case DRV_LS_BACKTRACKING: solve_with<Scalar, LineSearchBacktracking, Objective, Fused>(dev, obj, q, hv_algo, n, x_host, grad_host, trace, cap, out, t_begin, h2d_extra); break;
    case DRV_LS_BRACKETING: solve_with<Scalar, LineSearchBracketing, Objective, Fused>(dev, obj, q, hv_algo, n, x_host, grad_host, trace, cap, out, t_begin, h2d_extra); break;
    case DRV_LS_NOCEDAL_WRIGHT: solve_with<Scalar, LineSearchNocedalWright, Objective, Fused>(dev, obj, q, hv_algo, n, x_host, grad_host, trace, cap, out, t_begin, h2d_extra); break;
    case DRV_LS_MORE_THUENTE: solve_with<Scalar, LineSearchMoreThuente, Objective, Fused>(dev, obj, q, hv_algo, n, x_host, grad_host, trace, cap, out, t_begin, h2d_extra); break;
    default: throw std::invalid_argument("unknown line search id");
    }
}

template <class Scalar>
int lbfgs_any(int dev_ordinal, int objective, const Scalar* data0_host, const Scalar* data1_host, long n, int ls,
              const drv_param* q, int hv_algo, int fused, Scalar* x_host, Scalar* grad_host, double* trace, long cap,
              drv_result* out)
{
    return guarded(out, [&]() {
        const double t_begin = now();
        Device& dev = device(dev_ordinal);
        DeviceVector<Scalar> d0(dev), d1(dev);
        long extra = 0;
        if (data0_host) { d0.copy_from_host(data0_host, n); extra += long(sizeof(Scalar)) * n; }
        if (data1_host) { d1.copy_from_host(data1_host, n); extra += long(sizeof(Scalar)) * n; }
        if (fused)
        {
            BuiltinObjective<Scalar> obj(objective, d0.data(), d1.data());
            solve_ls<Scalar, BuiltinObjective<Scalar>, true>(ls, dev, obj, q, hv_algo, n, x_host, grad_host, trace, cap, out, t_begin, extra);
        }
        else
        {
            PlainObjective<Scalar> obj(objective, d0.data(), d1.data());
            solve_ls<Scalar, PlainObjective<Scalar>, false>(ls, dev, obj, q, hv_algo, n, x_host, grad_host, trace, cap, out, t_begin, extra);
        }
    });
}

}  // namespace

extern "C" {

// LBFGSSolver<double, LS>::minimize with host buffers.  fused = 1: BuiltinObjective (single-kernel trials),
// fused = 0: PlainObjective (what a user-written device functor gets).
int lbfgsb200_drv_lbfgs_f64(int device_ordinal, int objective, const double* data0_host, const double* data1_host, long n,
                            int ls, const drv_param* prm, int hv_algo, int fused, double* x_host, double* grad_host,
                            double* fx_trace, long trace_cap, drv_result* out)
{
    return lbfgs_any<double>(device_ordinal, objective, data0_host, data1_host, n, ls, prm, hv_algo, fused, x_host, grad_host, fx_trace, trace_cap, out);
}
int lbfgsb200_drv_lbfgs_f32(int device_ordinal, int objective, const float* data0_host, const float* data1_host, long n,
                            int ls, const drv_param* prm, int hv_algo, int fused, float* x_host, float* grad_host,
                            double* fx_trace, long trace_cap, drv_result* out)
{
    return lbfgs_any<float>(device_ordinal, objective, data0_host, data1_host, n, ls, prm, hv_algo, fused, x_host, grad_host, fx_trace, trace_cap, out);
}

// The context the driver uses for a device (so that Python can call the raw C ABI on the same stream)
void* lbfgsb200_drv_ctx(int device_ordinal)
{
    try { return device(device_ordinal).ctx(); }
    catch (...) { return nullptr; }
}

void lbfgsb200_drv_shutdown(void) { devices().clear(); }

}  // extern "C"
