// resident.cuh -- the device-resident L-BFGS solve: LBFGSSolver::minimize() (reference LBFGS.h:78-173) for a built-in objective
// as ONE CUDA graph launch.  Included at the end of lbfgs_b200.cu.
//
// The host-driven path needs the host after every line-search trial (the scalar decisions of LineSearch*.h) -- T + 2 round
// trips per iteration, which is what limits small shards (n-sharded N = 8: ~230 us per iteration for ~90 us of HBM work).
// Here every scalar decision runs on the device, in the last CTA of the kernel that produced its inputs:
//   * line-search state machine: LBFGSpp::*Core (include/LBFGSpp/LineSearchCore.h, the same code the host front uses)
//   * convergence tests of LBFGS.h:137-154, the curvature gate of :161, the ring bookkeeping of BFGSMat.h:81-97
//   * buffer rotation (x <-> xp, grad <-> gradp, x <-> x_lo ...) = pointer swaps inside the device state
// and control flow is a CUDA graph with conditional nodes (CUDA 12.4+):
//   first_eval -> begin -> WHILE(!finished){ iter_begin -> WHILE(line search){ trial } -> after_ls(+update) -> gram_dots -> gram_combine }
// Every conditional handle is set either by a kernel upstream of its node in the same graph or by a kernel of its own body.
// Kernels take their operands from the device state (pointers rotate), so the instantiated graph is reused for every solve
// of the same shape.  Arithmetic is the host-driven path's, kernel for kernel (same bodies, same grids, same split of the
// Gram pass), so both paths return bit-identical results.  Cross-GPU reductions use the in-kernel NVLink exchange with a
// device-side epoch counter.  No host involvement between launch and completion.
#pragma once

#include "../../include/LBFGSpp/LineSearchCore.h"

namespace lb {

constexpr int kMaxPast = 64;

template <class T> struct DevSolve
{
    // vectors (rotate by pointer swap)
    T *x, *xp, *g, *gp, *drt, *x_lo, *g_lo;
    int64_t n;
    // S/Y ring geometry (device is authoritative during a resident solve)
    int head, ncorr, M, m, gram_cur;
    // options (LBFGSParam)
    T epsilon, epsilon_rel, delta, max_step, eps_gate;
    int past, max_iterations, ls_kind;
    LBFGSpp::LineSearchOptions<T> ls_opt;
    // line-search state
    LBFGSpp::BacktrackingCore<T> bt;
    LBFGSpp::BracketingCore<T> br;
    LBFGSpp::NocedalWrightCore<T> nw;
    LBFGSpp::MoreThuenteCore<T> mt;
    int have_lo, need_restore;
    T lo_gg, lo_xx, start_gg, start_xx;
    // iteration scalars
    T fx, dg, gg, xx, gnorm;
    int k;
    long long nfev;
    int status;     // 0 ok, otherwise a LineSearchError code
    int finished;
    int niter;      // return value of minimize()
    T fx_hist[kMaxPast];
    unsigned long long epoch;   // cross-rank exchange sequence number (continues the context's)
    double* trace;              // optional: f of every evaluation
    long long trace_cap;
    double gate[2];             // {s.y, y.y} of the pair being appended
};

template <class T> __device__ __forceinline__ T& ls_step_ref(DevSolve<T>* st)
{
    switch (st->ls_kind)
    {
    case 0: return st->bt.step;
    case 1: return st->br.step;
    case 2: return st->nw.step;
    default: return st->mt.step;
    }
}
template <class T> __device__ __forceinline__ int ls_init(DevSolve<T>* st, T fx, T dg, T step, T step_max)
{
    switch (st->ls_kind)
    {
    case 0: return st->bt.init(st->ls_opt, fx, dg, step, step_max);
    case 1: return st->br.init(st->ls_opt, fx, dg, step, step_max);
    case 2: return st->nw.init(st->ls_opt, fx, dg, step, step_max);
    default: return st->mt.init(st->ls_opt, fx, dg, step, step_max);
    }
}
template <class T> __device__ __forceinline__ int ls_advance(DevSolve<T>* st, T fx, T dg, bool& keep)
{
    switch (st->ls_kind)
    {
    case 0: return st->bt.advance(fx, dg, keep);
    case 1: return st->br.advance(fx, dg, keep);
    case 2: return st->nw.advance(fx, dg, keep);
    default: return st->mt.advance(fx, dg, keep);
    }
}
template <class T> __device__ __forceinline__ void ls_best(DevSolve<T>* st, T& fx, T& dg)
{
    switch (st->ls_kind)
    {
    case 0: fx = st->bt.best_fx; dg = st->bt.best_dg; break;
    case 1: fx = st->br.best_fx; dg = st->br.best_dg; break;
    case 2: fx = st->nw.best_fx; dg = st->nw.best_dg; break;
    default: fx = st->mt.best_fx; dg = st->mt.best_dg; break;
    }
}
template <class P> __device__ __forceinline__ void dswap(P& a, P& b) { P t = a; a = b; b = t; }

// handles + fixed pointers every resident kernel needs
template <class T> struct ResidentEnv
{
    DevSolve<T>* st;
    ReduceBuf rb;            // partials / ticket / result ; xc ; (epoch filled per launch from st->epoch)
    unsigned* aux_ticket;    // for kernels that do not reduce but need a "last CTA"
    cudaGraphConditionalHandle h_outer, h_inner;
    // history storage (fixed)
    T *S, *Y, *ys, *alpha, *theta;
    T *SY[2], *YY[2], *SS[2];
    int64_t ld;
    double* gram_partials;
    double* gram_raw;
    // objective data
    const T *data0, *data1;
    int64_t index_offset;
};

template <class T> __device__ __forceinline__ ReduceBuf resident_rb(const ResidentEnv<T>& env, double* result)
{
    ReduceBuf rb = env.rb;
    rb.result = result;
    rb.epoch = env.st->epoch + 1ull;
    rb.mail_seq = 0ull;
    return rb;
}
// true in exactly one CTA: the last one to get here (for kernels without a reduction)
__device__ __forceinline__ bool last_cta_arrives(unsigned* ticket)
{
    __shared__ bool s_last_cta;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned t = atomicAdd(ticket, 1u);
        s_last_cta = (t == gridDim.x - 1);
        if (s_last_cta) *ticket = 0u;
    }
    __syncthreads();
    return s_last_cta;
}

template <class T> __device__ __forceinline__ void record_eval(DevSolve<T>* st, T fx)
{
    if (st->trace && st->nfev < st->trace_cap) st->trace[st->nfev] = (double)fx;
    st->nfev++;
}

// ------------------------------------------------------------------------------------------------ first evaluation
template <class T, class OBJ> struct ObjMaker;
template <class T> struct ObjMaker<T, RosenbrockPaired<T> > { static __device__ RosenbrockPaired<T> make(const ResidentEnv<T>&, int64_t n) { return RosenbrockPaired<T>{n}; } };
template <class T> struct ObjMaker<T, QuadShift<T> > { static __device__ QuadShift<T> make(const ResidentEnv<T>& e, int64_t n) { return QuadShift<T>{n, e.index_offset}; } };
template <class T> struct ObjMaker<T, RosenbrockChained<T> > { static __device__ RosenbrockChained<T> make(const ResidentEnv<T>&, int64_t n) { return RosenbrockChained<T>{n, 0, n, nullptr}; } };
template <class T> struct ObjMaker<T, QuadTridiag<T> > { static __device__ QuadTridiag<T> make(const ResidentEnv<T>& e, int64_t n) { return QuadTridiag<T>{n, e.data0, e.data1, 0, n, nullptr}; } };

// fx = f(x, grad), norms, early exit (LBFGS.h:91-103), first step 1/||g|| and dg = -g.g (:106-108,123)
template <class T, class OBJ>
__global__ void __launch_bounds__(kThreads) kg_first_eval(ResidentEnv<T> env)
{
    DevSolve<T>* st = env.st;
    const OBJ obj = ObjMaker<T, OBJ>::make(env, st->n);
    const ReduceBuf rb = resident_rb(env, env.rb.result);
    if (!trial_body<T, OBJ, false, true>(obj, st->n, nullptr, nullptr, T(0), st->x, st->g, rb)) return;
    if (threadIdx.x != 0) return;
    if (rb.xc) st->epoch++;
    const T fx = (T)rb.result[0], gg = (T)rb.result[2], xx = (T)rb.result[3];
    st->nfev = 0;
    record_eval(st, fx);
    st->fx = fx;
    st->gg = gg;
    st->xx = xx;
    st->k = 1;
    st->status = 0;
    st->finished = 0;
    st->need_restore = 0;
    if (st->past > 0) st->fx_hist[0] = fx;
    st->gnorm = sqrt(gg);
    if (st->gnorm <= st->epsilon || st->gnorm <= st->epsilon_rel * sqrt(xx))
    {
        st->finished = 1;
        st->niter = 1;
    }
}

// before the loop: drt = -grad (LBFGS.h:106); the loop runs unless the start point already satisfied the gradient test
template <class T>
__global__ void __launch_bounds__(kThreads) kg_begin(ResidentEnv<T> env)
{
    DevSolve<T>* st = env.st;
    const bool active = !st->finished;
    if (active)
    {
        const T* g = st->g;
        T* d = st->drt;
        const int64_t packs = (st->n + 3) >> 2, stride = (int64_t)gridDim.x * kThreads;
        for (int64_t p = (int64_t)blockIdx.x * kThreads + threadIdx.x; p < packs; p += stride)
        {
            const int64_t i0 = p << 2;
            const int cnt = (st->n - i0 >= 4) ? 4 : int(st->n - i0);
            Pack<T> r = load4<T, Hint::Stream, true>(g, i0, cnt);
#pragma unroll
            for (int k = 0; k < 4; k++) r.v[k] = T(-1) * r.v[k];
            store4<T, Hint::Plain, true>(d, i0, cnt, r);
        }
    }
    if (!last_cta_arrives(env.aux_ticket) || threadIdx.x != 0) return;
    cudaGraphSetConditional(env.h_outer, st->finished ? 0u : 1u);
}

// top of every iteration (LBFGS.h:121-127): the current point becomes the previous one (pointer rotation), the line search is
// armed with step = 1/||g|| on the first iteration and 1 afterwards.  One thread.
template <class T>
__global__ void kg_iter_begin(ResidentEnv<T> env)
{
    DevSolve<T>* st = env.st;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (!st->finished)
    {
        const T step = (st->k == 1) ? T(1) / st->gnorm : T(1);
        if (st->k == 1) st->dg = -st->gg;
        dswap(st->xp, st->x);
        dswap(st->gp, st->g);
        st->have_lo = 0;
        st->start_gg = st->gg;
        st->start_xx = st->xx;
        const int rc = ls_init(st, st->fx, st->dg, step, st->max_step);
        if (rc != 0) { st->status = rc; st->finished = 1; st->niter = st->k; }
    }
    cudaGraphSetConditional(env.h_inner, st->finished ? 0u : 1u);
}

// ------------------------------------------------------------------------------------------------ one line-search trial
template <class T, class OBJ>
__global__ void __launch_bounds__(kThreads) kg_trial(ResidentEnv<T> env)
{
    DevSolve<T>* st = env.st;
    const OBJ obj = ObjMaker<T, OBJ>::make(env, st->n);
    const ReduceBuf rb = resident_rb(env, env.rb.result);
    const T step = ls_step_ref(st);
    if (!trial_body<T, OBJ, true, true>(obj, st->n, st->xp, st->drt, step, st->x, st->g, rb)) return;
    if (threadIdx.x != 0) return;
    if (rb.xc) st->epoch++;
    const T fx = (T)rb.result[0], dg = (T)rb.result[1], gg = (T)rb.result[2], xx = (T)rb.result[3];
    record_eval(st, fx);
    bool keep = false;
    const int rc = ls_advance(st, fx, dg, keep);
    if (keep)
    {
        dswap(st->x, st->x_lo);
        dswap(st->g, st->g_lo);
        st->have_lo = 1;
        st->lo_gg = gg;
        st->lo_xx = xx;
    }
    unsigned more = 0u;
    if (rc == LBFGSpp::LSC_EVALUATE) more = 1u;
    else if (rc == LBFGSpp::LSC_ACCEPT)
    {
        st->fx = fx; st->dg = dg; st->gg = gg; st->xx = xx;
    }
    else if (rc == LBFGSpp::LSC_TAKE_BEST)
    {
        if (st->have_lo)
        {
            dswap(st->x, st->x_lo);
            dswap(st->g, st->g_lo);
            st->gg = st->lo_gg;
            st->xx = st->lo_xx;
        }
        else
        {
            st->need_restore = 1;   // no trial ever improved on the start point: after_ls copies xp/gp back
            st->gg = st->start_gg;
            st->xx = st->start_xx;
        }
        T bf, bd;
        ls_best(st, bf, bd);
        st->fx = bf;
        st->dg = bd;
    }
    else
    {
        st->status = rc;
        st->finished = 1;
        st->niter = st->k;
    }
    cudaGraphSetConditional(env.h_inner, more);
}

// ------------------------------------------------------------------------------------------------ after the line search
// One kernel: (optional restore of the start point) + the pair update s = x - xp, y = g - gp written speculatively into the free
// ring slot with {s.y, y.y} -> st->gate (LBFGS.h:159-160, BFGSMat.h:85-92) + the convergence tests of LBFGS.h:130-154 in the
// last CTA.  When the solve is over the pair is simply never committed.
template <class T>
__global__ void __launch_bounds__(kThreads) kg_after_ls(ResidentEnv<T> env)
{
    DevSolve<T>* st = env.st;
    const bool restore = st->need_restore != 0;
    if (restore)
    {
        const T* xp = st->xp; const T* gp = st->gp;
        T* x = st->x; T* g = st->g;
        for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < st->n; i += (int64_t)gridDim.x * kThreads)
        {
            x[i] = xp[i];
            g[i] = gp[i];
        }
    }
    const ReduceBuf rb = resident_rb(env, st->gate);
    T* s_out = env.S + (int64_t)st->head * env.ld;
    T* y_out = env.Y + (int64_t)st->head * env.ld;
    // after a restore x == xp and g == gp: read the sources so that s = y = 0 without depending on the copy above
    const T* xs = restore ? st->xp : st->x;
    const T* gs = restore ? st->gp : st->g;
    if (!update_body<T, true>(st->n, xs, st->xp, gs, st->gp, s_out, y_out, rb)) return;
    if (threadIdx.x != 0) return;
    if (rb.xc) st->epoch++;
    st->need_restore = 0;
    if (!st->finished)
    {
        const int k = st->k;
        st->gnorm = sqrt(st->gg);
        if (st->gnorm <= st->epsilon || st->gnorm <= st->epsilon_rel * sqrt(st->xx)) { st->finished = 1; st->niter = k; }
        if (!st->finished && st->past > 0)
        {
            const T fxd = st->fx_hist[k % st->past];
            const T fx = st->fx;
            const T afx = fx < T(0) ? -fx : fx, afxd = fxd < T(0) ? -fxd : fxd;
            T big = afx < afxd ? afxd : afx;       // std::max(abs(fx), abs(fxd))
            big = big < T(1) ? T(1) : big;         // std::max(., 1)
            const T diff = (fxd - fx) < T(0) ? -(fxd - fx) : (fxd - fx);
            if (k >= st->past && diff <= st->delta * big) { st->finished = 1; st->niter = k; }
            else st->fx_hist[k % st->past] = fx;
        }
        if (!st->finished && st->max_iterations != 0 && k >= st->max_iterations) { st->finished = 1; st->niter = k; }
    }
    if (st->finished) cudaGraphSetConditional(env.h_outer, 0u);
}

// ------------------------------------------------------------------------------------------------ apply_Hv
// ring geometry after the curvature gate (LBFGS.h:161): every thread derives it from the state and {s.y, y.y}
template <class T> struct Geometry
{
    int accepted, c, new_slot, head_after;
    T sy, yy;
};
template <class T> __device__ __forceinline__ Geometry<T> gate_geometry(const DevSolve<T>* st)
{
    Geometry<T> q;
    q.sy = (T)st->gate[0];
    q.yy = (T)st->gate[1];
    q.accepted = (q.sy > st->eps_gate * q.yy) ? 1 : 0;
    q.c = q.accepted ? (st->ncorr < st->m ? st->ncorr + 1 : st->m) : st->ncorr;
    q.new_slot = q.accepted ? st->head : -1;
    q.head_after = q.accepted ? (st->head + 1) % st->M : st->head;
    return q;
}
__device__ __forceinline__ int slot_by_age(int head, int M, int age) { return ((head - 1 - age) % M + M) % M; }

template <class T, int ROUNDS>
__global__ void __launch_bounds__(kGramMaxThreads, 1) kg_gram_dots(ResidentEnv<T> env)
{
    DevSolve<T>* st = env.st;
    if (st->finished) return;   // the solve ended in kg_after_ls: nothing to prepare
    const Geometry<T> q = gate_geometry(st);
    GramDotsArgs<T> a;
    a.n = st->n; a.ld = env.ld; a.v = st->g; a.S = env.S; a.Y = env.Y;
    a.c = q.c; a.new_slot = q.new_slot;
    int split = 8;
    while (split > 1 && q.c * split > kGramMaxWarps) split >>= 1;
    a.split = split;
    a.cols_per_round = q.c < kGramMaxWarps / split ? q.c : kGramMaxWarps / split;
    a.use_tma = 1;
    for (int age = 0; age < q.c; age++) a.slots[age] = (unsigned char)slot_by_age(q.head_after, st->M, age);
    const unsigned long long epoch = st->epoch + 1ull;
    gram_dots_body<T, ROUNDS>(a, env.gram_partials, env.rb.ticket, env.gram_raw, env.rb.xc, epoch);
}
// the epoch bump of kg_gram_dots is folded into kg_gram_combine's controller (gram_dots_body has no "last CTA" return)

template <class T>
__global__ void __launch_bounds__(kThreads) kg_gram_combine(ResidentEnv<T> env)
{
    DevSolve<T>* st = env.st;
    if (st->finished) return;
    const Geometry<T> q = gate_geometry(st);
    GramCombineArgs<T> a;
    a.n = st->n; a.ld = env.ld; a.v = st->g; a.S = env.S; a.Y = env.Y; a.res = st->drt; a.want_dot = 1;
    GramSolveArgs<T>& g = a.solve;
    g.c = q.c; g.M = st->M; g.new_slot = q.new_slot; g.with_v = 1; g.a = T(-1);
    g.raw = env.gram_raw;
    const int in = st->gram_cur, out = q.accepted ? 1 - st->gram_cur : st->gram_cur;
    g.SY_in = env.SY[in]; g.YY_in = env.YY[in]; g.SS_in = env.SS[in];
    g.SY_out = env.SY[out]; g.YY_out = env.YY[out]; g.SS_out = env.SS[out];
    g.ys = env.ys; g.alpha = env.alpha; g.theta = env.theta;
    g.ov_slot = q.new_slot; g.ov_ys = q.sy; g.ov_theta_on = q.accepted; g.ov_theta = q.accepted ? q.yy / q.sy : T(1);
    for (int age = 0; age < q.c; age++) g.slots[age] = (unsigned char)slot_by_age(q.head_after, st->M, age);
    ReduceBuf rb = resident_rb(env, env.rb.result);
    rb.epoch = st->epoch + 2ull;   // kg_gram_dots took st->epoch + 1 without bumping the counter
    if (!gram_combine_body<T, true>(a, rb)) return;
    if (threadIdx.x != 0) return;
    if (rb.xc) st->epoch = rb.epoch;
    // commit the pair (BFGSMat.h:89-97)
    if (q.accepted)
    {
        env.ys[st->head] = q.sy;
        *env.theta = q.yy / q.sy;
        st->head = q.head_after;
        st->ncorr = q.c;
        st->gram_cur = 1 - st->gram_cur;
    }
    // LBFGS.h:165-169: the direction is in drt, dg = grad.drt for the next line search, k++
    st->dg = (T)rb.result[0];
    st->k += 1;
}

}  // namespace lb

// =====================================================================================================================
// host side: solver handle, graph construction, C ABI
// =====================================================================================================================
struct lbfgs_b200_solver
{
    lbfgs_b200_ctx* ctx = nullptr;
    lbfgs_b200_hist* hist = nullptr;
    int64_t n = 0;
    int m = 0, elem = 8;
    void* vec[7] = {};          // x, xp, g, gp, drt, x_lo, g_lo
    void* d_state = nullptr;    // DevSolve<T>
    unsigned* aux_ticket = nullptr;
    double* d_trace = nullptr;
    long long trace_cap = 0;
    // cached graph
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int g_objective = -1;
    const void *g_data0 = nullptr, *g_data1 = nullptr;
    const void* g_xc = nullptr;
    void* final_g = nullptr;    // device pointer of the final gradient (one of vec[])
    void* final_x = nullptr;
};

template <class T, class OBJ>
static lbfgs_b200_status build_graph(lbfgs_b200_solver* s, const lb::ResidentEnv<T>& env_in)
{
    using namespace lb;
    lbfgs_b200_ctx* ctx = s->ctx;
    if (s->exec) { cudaGraphExecDestroy(s->exec); s->exec = nullptr; }
    if (s->graph) { cudaGraphDestroy(s->graph); s->graph = nullptr; }
    CU(ctx, cudaGraphCreate(&s->graph, 0));
    ResidentEnv<T> env = env_in;
    CU(ctx, cudaGraphConditionalHandleCreate(&env.h_outer, s->graph, 0, cudaGraphCondAssignDefault));
    CU(ctx, cudaGraphConditionalHandleCreate(&env.h_inner, s->graph, 0, cudaGraphCondAssignDefault));

    const int64_t n = s->n;
    const int g_stream2 = grid_for(ctx, n, 2), g_stream1 = grid_for(ctx, n, 1);
    void* args[] = {&env};
    auto kernel_node = [&](cudaGraph_t g, cudaGraphNode_t* node, const cudaGraphNode_t* deps, size_t ndeps, void* fn, int grid, int block,
                           size_t smem) -> cudaError_t {
        cudaKernelNodeParams kp = {};
        kp.func = fn; kp.gridDim = dim3((unsigned)grid); kp.blockDim = dim3((unsigned)block); kp.sharedMemBytes = (unsigned)smem;
        kp.kernelParams = args;
        return cudaGraphAddKernelNode(node, g, deps, ndeps, &kp);
    };
    auto cond_node = [&](cudaGraph_t g, cudaGraphNode_t* node, const cudaGraphNode_t* deps, size_t ndeps, cudaGraphConditionalHandle h,
                         cudaGraphConditionalNodeType type, cudaGraph_t* body) -> cudaError_t {
        cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
        p.conditional.handle = h; p.conditional.type = type; p.conditional.size = 1;
        cudaError_t e = cudaGraphAddNode(node, g, deps, ndeps, &p);
        if (e == cudaSuccess) *body = p.conditional.phGraph_out[0];
        return e;
    };

    cudaGraphNode_t n_first, n_begin, n_outer, n_iter, n_inner, n_trial, n_after, n_dots, n_combine;
    cudaGraph_t g_outer, g_inner;
    CU(ctx, kernel_node(s->graph, &n_first, nullptr, 0, (void*)kg_first_eval<T, OBJ>, g_stream2, kThreads, 0));
    CU(ctx, kernel_node(s->graph, &n_begin, &n_first, 1, (void*)kg_begin<T>, g_stream2, kThreads, 0));
    CU(ctx, cond_node(s->graph, &n_outer, &n_begin, 1, env.h_outer, cudaGraphCondTypeWhile, &g_outer));
    CU(ctx, kernel_node(g_outer, &n_iter, nullptr, 0, (void*)kg_iter_begin<T>, 1, 32, 0));
    CU(ctx, cond_node(g_outer, &n_inner, &n_iter, 1, env.h_inner, cudaGraphCondTypeWhile, &g_inner));
    CU(ctx, kernel_node(g_inner, &n_trial, nullptr, 0, (void*)kg_trial<T, OBJ>, g_stream2, kThreads, 0));
    CU(ctx, kernel_node(g_outer, &n_after, &n_inner, 1, (void*)kg_after_ls<T>, g_stream2, kThreads, 0));
    // Gram pass: launch geometry for the largest history (the kernel derives the actual one from the state)
    int split = 8;
    while (split > 1 && s->m * split > kGramMaxWarps) split >>= 1;
    const int per_round = s->m < kGramMaxWarps / split ? s->m : kGramMaxWarps / split;
    const int rounds = (s->m + per_round - 1) / per_round;
    const int64_t ntiles = (n + kGramTE - 1) / kGramTE;
    const int g_dots = (int)(ntiles < ctx->sm_count ? ntiles : ctx->sm_count);
    const size_t smem_dots = (size_t)kGramStages * 3 * kGramTE * sizeof(T);
    void* dots_fn = rounds <= 1 ? (void*)kg_gram_dots<T, 1> : rounds == 2 ? (void*)kg_gram_dots<T, 2> : (void*)kg_gram_dots<T, 3>;
    CU(ctx, cudaFuncSetAttribute(dots_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dots));
    CU(ctx, kernel_node(g_outer, &n_dots, &n_after, 1, dots_fn, g_dots, kGramMaxThreads, smem_dots));
    const size_t smem_comb = gram_solve_smem_elems(s->m) * sizeof(T);
    if (smem_comb > 40 * 1024)
        CU(ctx, cudaFuncSetAttribute((void*)kg_gram_combine<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_comb));
    CU(ctx, kernel_node(g_outer, &n_combine, &n_dots, 1, (void*)kg_gram_combine<T>, g_stream1, kThreads, smem_comb));
    CU(ctx, cudaGraphInstantiate(&s->exec, s->graph, 0));
    return LBFGS_B200_OK;
}

template <class T>
static lbfgs_b200_status solver_minimize(lbfgs_b200_solver* s, int objective, const T* data0, const T* data1, const lbfgs_b200_param* prm,
                                         int ls_kind, T* x_inout, double* trace_host, long long trace_cap, lbfgs_b200_outcome* out)
{
    using namespace lb;
    lbfgs_b200_ctx* ctx = s->ctx;
    REQUIRE(ctx, s->elem == (int)sizeof(T), "solver element size mismatch");
    REQUIRE(ctx, prm && out && x_inout, "solver_minimize: NULL argument");
    REQUIRE(ctx, prm->m == s->m, "solver was created for m = %d, called with m = %d", s->m, prm->m);
    REQUIRE(ctx, ls_kind >= 0 && ls_kind <= 3, "unknown line search %d", ls_kind);
    REQUIRE(ctx, prm->past <= kMaxPast, "past > %d is not supported by the device-resident solve", kMaxPast);
    REQUIRE(ctx, ctx->nranks == 1 || ctx->x_active, "the device-resident solve needs the in-kernel exchange (comm_p2p) when sharded");
    REQUIRE(ctx, ctx->nranks == 1 || objective == LBFGS_B200_OBJ_ROSENBROCK_PAIRED || objective == LBFGS_B200_OBJ_QUAD_SHIFT,
            "objective %d couples neighbouring coordinates: under n-sharding use the host-driven loop (the device-resident graph has no halo exchange)", objective);
    if (objective == LBFGS_B200_OBJ_ROSENBROCK_PAIRED) REQUIRE(ctx, s->n % 2 == 0, "paired Rosenbrock needs an even n");
    if (objective == LBFGS_B200_OBJ_QUAD_TRIDIAG) REQUIRE(ctx, data0 && data1, "quad_tridiag needs data0 = diag, data1 = rhs");
    lbfgs_b200_hist* h = s->hist;
    if (auto st = lbfgs_b200_hist_reset(h)) return st;

    // (re)build the graph when the objective or its data changed
    const void* xc_now = ctx->x_active ? (const void*)ctx->x_comm : nullptr;
    if (!s->exec || s->g_objective != objective || s->g_data0 != data0 || s->g_data1 != data1 || s->g_xc != xc_now)
    {
        ResidentEnv<T> env{};
        env.st = static_cast<DevSolve<T>*>(s->d_state);
        env.rb = ctx->rb;
        env.rb.xc = ctx->x_active ? ctx->x_comm : nullptr;
        env.aux_ticket = s->aux_ticket;
        env.S = static_cast<T*>(h->S); env.Y = static_cast<T*>(h->Y); env.ys = static_cast<T*>(h->ys);
        env.alpha = static_cast<T*>(h->alpha); env.theta = static_cast<T*>(h->theta);
        for (int b = 0; b < 2; b++) { env.SY[b] = static_cast<T*>(h->SY[b]); env.YY[b] = static_cast<T*>(h->YY[b]); env.SS[b] = static_cast<T*>(h->SS[b]); }
        env.ld = h->ld; env.gram_partials = ctx->gram_partials; env.gram_raw = ctx->gram_raw;
        env.data0 = data0; env.data1 = data1; env.index_offset = ctx->index_offset;
        lbfgs_b200_status st = LBFGS_B200_OK;
        switch (objective)
        {
        case LBFGS_B200_OBJ_ROSENBROCK_PAIRED: st = build_graph<T, RosenbrockPaired<T> >(s, env); break;
        case LBFGS_B200_OBJ_QUAD_SHIFT: st = build_graph<T, QuadShift<T> >(s, env); break;
        case LBFGS_B200_OBJ_ROSENBROCK_CHAINED: st = build_graph<T, RosenbrockChained<T> >(s, env); break;
        case LBFGS_B200_OBJ_QUAD_TRIDIAG: st = build_graph<T, QuadTridiag<T> >(s, env); break;
        default: return fail(ctx, LBFGS_B200_ERR_INVALID, "unknown objective id %d", objective);
        }
        if (st) return st;
        s->g_objective = objective; s->g_data0 = data0; s->g_data1 = data1; s->g_xc = xc_now;
    }

    if (trace_host && trace_cap > s->trace_cap)
    {
        cudaFree(s->d_trace);
        s->d_trace = nullptr;
        CU(ctx, cudaMalloc(&s->d_trace, sizeof(double) * (size_t)trace_cap));
        s->trace_cap = trace_cap;
    }

    DevSolve<T> hs{};
    T** v = reinterpret_cast<T**>(s->vec);
    hs.x = v[0]; hs.xp = v[1]; hs.g = v[2]; hs.gp = v[3]; hs.drt = v[4]; hs.x_lo = v[5]; hs.g_lo = v[6];
    hs.n = s->n;
    hs.head = 0; hs.ncorr = 0; hs.M = h->M; hs.m = h->m; hs.gram_cur = h->gram_cur;
    hs.epsilon = (T)prm->epsilon; hs.epsilon_rel = (T)prm->epsilon_rel; hs.delta = (T)prm->delta; hs.max_step = (T)prm->max_step;
    hs.eps_gate = std::numeric_limits<T>::epsilon();
    hs.past = prm->past; hs.max_iterations = prm->max_iterations; hs.ls_kind = ls_kind;
    hs.ls_opt.linesearch = (ls_kind == 3) ? 3 : prm->linesearch;
    hs.ls_opt.max_linesearch = prm->max_linesearch;
    hs.ls_opt.min_step = (T)prm->min_step; hs.ls_opt.max_step = (T)prm->max_step; hs.ls_opt.ftol = (T)prm->ftol; hs.ls_opt.wolfe = (T)prm->wolfe;
    hs.epoch = ctx->x_epoch;
    hs.trace = trace_host ? s->d_trace : nullptr;
    hs.trace_cap = trace_host ? trace_cap : 0;
    const size_t vb = sizeof(T) * (size_t)s->n;
    CU(ctx, cudaMemcpyAsync(hs.x, x_inout, vb, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(s->d_state, &hs, sizeof(hs), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaGraphLaunch(s->exec, ctx->stream));
    ctx->launches++;
    CU(ctx, cudaMemcpyAsync(&hs, s->d_state, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    // the state tells where the result lives after all the pointer rotations
    CU(ctx, cudaMemcpyAsync(x_inout, hs.x, vb, cudaMemcpyDeviceToDevice, ctx->stream));
    s->final_g = hs.g;
    s->final_x = hs.x;
    if (trace_host)
    {
        const long long cnt = hs.nfev < trace_cap ? hs.nfev : trace_cap;
        CU(ctx, cudaMemcpyAsync(trace_host, s->d_trace, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    // mirror the ring state so that the history object stays usable through the ordinary entry points
    h->head = hs.head; h->ncorr = hs.ncorr; h->gram_cur = hs.gram_cur; h->pending = -1;
    ctx->x_epoch = hs.epoch;
    out->status = hs.status;
    out->niter = hs.niter;
    out->nfev = hs.nfev;
    out->fx = (double)hs.fx;
    out->gnorm = (double)hs.gnorm;
    return LBFGS_B200_OK;
}

extern "C" {

lbfgs_b200_status lbfgs_b200_solver_create(lbfgs_b200_ctx* ctx, int64_t n, int m, int elem_bytes, lbfgs_b200_solver** out)
{
    REQUIRE(ctx, ctx && out, "solver_create: NULL argument");
    *out = nullptr;
    lbfgs_b200_solver* s = new (std::nothrow) lbfgs_b200_solver();
    if (!s) return fail(ctx, LBFGS_B200_ERR_ALLOC, "out of host memory");
    s->ctx = ctx; s->n = n; s->m = m; s->elem = elem_bytes;
    lbfgs_b200_status st = lbfgs_b200_hist_create(ctx, &s->hist, n, m, elem_bytes);
    if (st) { delete s; return st; }
    cudaError_t e = cudaSuccess;
    const size_t vb = ((size_t)n * elem_bytes + 255) & ~size_t(255);
    for (int k = 0; k < 7 && e == cudaSuccess; k++) e = cudaMalloc(&s->vec[k], vb);
    const size_t state_bytes = elem_bytes == 8 ? sizeof(lb::DevSolve<double>) : sizeof(lb::DevSolve<float>);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_state, state_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&s->aux_ticket, sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(s->aux_ticket, 0, sizeof(unsigned));
    if (e != cudaSuccess)
    {
        lbfgs_b200_solver_destroy(s);
        return fail(ctx, e == cudaErrorMemoryAllocation ? LBFGS_B200_ERR_ALLOC : LBFGS_B200_ERR_CUDA, "solver_create: %s", cudaGetErrorString(e));
    }
    *out = s;
    return LBFGS_B200_OK;
}

void lbfgs_b200_solver_destroy(lbfgs_b200_solver* s)
{
    if (!s) return;
    if (s->ctx && s->ctx->stream) cudaStreamSynchronize(s->ctx->stream);
    if (s->exec) cudaGraphExecDestroy(s->exec);
    if (s->graph) cudaGraphDestroy(s->graph);
    for (void* p : s->vec) cudaFree(p);
    cudaFree(s->d_state);
    cudaFree(s->aux_ticket);
    cudaFree(s->d_trace);
    lbfgs_b200_hist_destroy(s->hist);
    delete s;
}

const void* lbfgs_b200_solver_final_grad(const lbfgs_b200_solver* s) { return s ? s->final_g : nullptr; }
lbfgs_b200_hist* lbfgs_b200_solver_history(lbfgs_b200_solver* s) { return s ? s->hist : nullptr; }

lbfgs_b200_status lbfgs_b200_solver_minimize_f64(lbfgs_b200_solver* s, int objective, const double* data0, const double* data1,
                                                 const lbfgs_b200_param* prm, int line_search, double* x_inout, double* trace_host,
                                                 long long trace_cap, lbfgs_b200_outcome* out)
{
    if (!s || !s->ctx) return LBFGS_B200_ERR_INVALID;
    return solver_minimize<double>(s, objective, data0, data1, prm, line_search, x_inout, trace_host, trace_cap, out);
}
lbfgs_b200_status lbfgs_b200_solver_minimize_f32(lbfgs_b200_solver* s, int objective, const float* data0, const float* data1,
                                                 const lbfgs_b200_param* prm, int line_search, float* x_inout, double* trace_host,
                                                 long long trace_cap, lbfgs_b200_outcome* out)
{
    if (!s || !s->ctx) return LBFGS_B200_ERR_INVALID;
    return solver_minimize<float>(s, objective, data0, data1, prm, line_search, x_inout, trace_host, trace_cap, out);
}

}  // extern "C"
