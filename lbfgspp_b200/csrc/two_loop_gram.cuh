// two_loop_gram.cuh -- "vector-free" form of BFGSMat::apply_Hv (reference BFGSMat.h:276-302).
//
// The two-loop recursion only ever combines the vectors {v, s_j, y_j}; every inner product it takes
// (s_j'q in the backward sweep, y_j'r in the forward sweep) is a linear combination of the entries of
//   b_s[j] = s_j'v,  b_y[j] = y_j'v,  SY[i][j] = s_i'y_j,  YY[i][j] = y_i'y_j .
// So the recursion can be carried out on 2c coefficients by one thread (k_gram_solve) once those inner
// products are known, and the result is a single linear combination  res = cv*v + sum_j cy_j*y_j + cs_j*s_j.
// HBM traffic per call: one pass over S,Y,v for the dots + one pass over S,Y,v for the combination
//   = (4c+3) n words   vs  (8c+4) n words for the stage-by-stage recursion (and 2c collectives -> 1).
// SY / YY are kept incrementally: the pair appended last contributes one new row/column, whose 3c dots
// (S'y_new, Y'y_new, Y's_new) are taken in the SAME pass that computes b_s, b_y -- no extra traffic.
// The arithmetic differs from the literal recursion only by rounding (same operations on the same exact
// quantities, re-associated); tests/test_gpu_* bound the difference and oracle/lbfgs_oracle.hpp carries
// a CPU twin (History::apply_Hv_gram) used to study it.
//
// k_gram_dots: tall-skinny [S Y]'[v s_new y_new].  One warp (or `split` warps) per history column pair;
// the three right-hand vectors are staged tile by tile into shared memory with TMA bulk copies
// (cp.async.bulk + mbarrier, 3-stage ring) and shared by all warps of the CTA; the S/Y columns stream from
// HBM straight into registers with 256-bit evict-first loads.  Deterministic: fixed-slot block partials,
// fixed-order final sum, integer ticket only.
#pragma once

#include "device_utils.cuh"

namespace lb {

constexpr int kGramTE = 2048;      // tile length in elements (16 KB of fp64 per staged vector)
constexpr int kGramStages = 3;
constexpr int kGramWarps = 16;
constexpr int kGramThreads = kGramWarps * 32;
constexpr int kMaxM = 64;
constexpr int kGramVals = 5;       // per column pair: s.v, y.v, s.ynew, y.ynew, y.snew

template <class T> struct GramDotsArgs
{
    int64_t n, ld;
    const T* v;        // may be nullptr (refresh only)
    const T* S;
    const T* Y;
    int c;             // number of valid pairs
    int new_slot;      // physical slot of the pair whose Gram row/column is still missing, or -1
    int split;         // warps cooperating on one column pair (power of two)
    int use_tma;       // v / columns 16-byte aligned
    unsigned char slots[kMaxM];  // physical slot by age (0 = newest)
};

// ---- mbarrier / TMA helpers (sm_90+ PTX, assembled for sm_100a) -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// pack of 4 from shared memory (two 128-bit LDS)
template <class T> __device__ __forceinline__ Pack<T> lds_pack(const T* p)
{
    Pack<T> r;
#pragma unroll
    for (int k = 0; k < 4; k++) r.v[k] = p[k];
    return r;
}

template <class T, int ROUNDS>
__global__ void __launch_bounds__(kGramThreads, 1) k_gram_dots(GramDotsArgs<T> a, double* partials, unsigned* ticket, double* result)
{
    extern __shared__ __align__(128) unsigned char gram_smem[];
    T* tiles = reinterpret_cast<T*>(gram_smem);                  // [stage][3][TE]
    __shared__ __align__(8) uint64_t full_bar[kGramStages];
    __shared__ double s_red[kGramWarps][ROUNDS * kGramVals];
    __shared__ bool s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool with_v = a.v != nullptr, with_new = a.new_slot >= 0;
    const T* snew = with_new ? a.S + (int64_t)a.new_slot * a.ld : nullptr;
    const T* ynew = with_new ? a.Y + (int64_t)a.new_slot * a.ld : nullptr;
    const int64_t ntiles = (a.n + kGramTE - 1) / kGramTE;
    const int cols_per_round = kGramWarps / a.split;
    const int my_col = warp / a.split, my_part = warp % a.split;
    const int part_len = kGramTE / a.split;                        // elements of a tile handled by this warp

    if (tid == 0)
    {
        for (int s = 0; s < kGramStages; s++) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // stage a tile of the right-hand vectors: TMA for full aligned tiles, guarded element loads for the tail
    auto stage_tile = [&](int64_t tile, int stage) {
        T* dst = tiles + (size_t)stage * 3 * kGramTE;
        const int64_t e0 = tile * kGramTE;
        const int64_t len = (a.n - e0 < kGramTE) ? (a.n - e0) : kGramTE;
        if (a.use_tma && len == kGramTE)
        {
            if (tid == 0)
            {
                const unsigned bytes = kGramTE * sizeof(T);
                mbar_expect_tx(&full_bar[stage], bytes * ((with_v ? 1 : 0) + (with_new ? 2 : 0)));
                if (with_v) tma_load_1d(dst, a.v + e0, bytes, &full_bar[stage]);
                if (with_new)
                {
                    tma_load_1d(dst + kGramTE, snew + e0, bytes, &full_bar[stage]);
                    tma_load_1d(dst + 2 * kGramTE, ynew + e0, bytes, &full_bar[stage]);
                }
            }
        }
        else
        {
            for (int i = tid; i < kGramTE; i += kGramThreads)
            {
                const bool ok = i < len;
                dst[i] = (with_v && ok) ? a.v[e0 + i] : T(0);
                dst[kGramTE + i] = (with_new && ok) ? snew[e0 + i] : T(0);
                dst[2 * kGramTE + i] = (with_new && ok) ? ynew[e0 + i] : T(0);
            }
        }
    };
    auto tile_is_tma = [&](int64_t tile) { return a.use_tma && (a.n - tile * kGramTE >= kGramTE); };

    T acc[ROUNDS][kGramVals];
#pragma unroll
    for (int r = 0; r < ROUNDS; r++)
#pragma unroll
        for (int k = 0; k < kGramVals; k++) acc[r][k] = T(0);

    // prologue: fill the ring
    int64_t next_tile = blockIdx.x;
    for (int s = 0; s < kGramStages; s++, next_tile += gridDim.x)
        if (next_tile < ntiles) stage_tile(next_tile, s);

    unsigned phase_bits = 0;  // one parity bit per stage
    int stage = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    {
        if (tile_is_tma(tile))
        {
            mbar_wait(&full_bar[stage], (phase_bits >> stage) & 1u);
            phase_bits ^= (1u << stage);
        }
        else
            __syncthreads();  // element-wise staged tile: make the stores visible
        const T* vt = tiles + (size_t)stage * 3 * kGramTE;
        const T* snt = vt + kGramTE;
        const T* ynt = vt + 2 * kGramTE;
        const int64_t e0 = tile * kGramTE;
        const int64_t len = (a.n - e0 < kGramTE) ? (a.n - e0) : kGramTE;
        const bool full_tile = (len == kGramTE) && a.use_tma;

#pragma unroll
        for (int r = 0; r < ROUNDS; r++)
        {
            const int j = r * cols_per_round + my_col;
            if (j < a.c)
            {
                const T* scol = a.S + (int64_t)a.slots[j] * a.ld + e0;
                const T* ycol = a.Y + (int64_t)a.slots[j] * a.ld + e0;
                // this warp's part of the tile, 2 packs (8 elements) per lane per step
                for (int base = my_part * part_len + lane * 4; base < (my_part + 1) * part_len; base += 256)
                {
                    Pack<T> ps[2], py[2];
#pragma unroll
                    for (int u = 0; u < 2; u++)
                    {
                        const int off = base + u * 128;
                        if (full_tile)
                        {
                            ps[u] = ld_pack<Hint::Stream>(scol + off);
                            py[u] = ld_pack<Hint::Stream>(ycol + off);
                        }
                        else
                        {
#pragma unroll
                            for (int k = 0; k < 4; k++)
                            {
                                const bool ok = off + k < len;
                                ps[u].v[k] = ok ? scol[off + k] : T(0);
                                py[u].v[k] = ok ? ycol[off + k] : T(0);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; u++)
                    {
                        const int off = base + u * 128;
                        if (with_v)
                        {
                            const Pack<T> pv = lds_pack(vt + off);
#pragma unroll
                            for (int k = 0; k < 4; k++)
                            {
                                acc[r][0] += ps[u].v[k] * pv.v[k];
                                acc[r][1] += py[u].v[k] * pv.v[k];
                            }
                        }
                        if (with_new)
                        {
                            const Pack<T> pyn = lds_pack(ynt + off), psn = lds_pack(snt + off);
#pragma unroll
                            for (int k = 0; k < 4; k++)
                            {
                                acc[r][2] += ps[u].v[k] * pyn.v[k];
                                acc[r][3] += py[u].v[k] * pyn.v[k];
                                acc[r][4] += py[u].v[k] * psn.v[k];
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with this stage's tile
        if (next_tile < ntiles) stage_tile(next_tile, stage);
        next_tile += gridDim.x;
        stage = (stage + 1 == kGramStages) ? 0 : stage + 1;
    }

    // ---- block reduction: lanes -> warp, then the `split` warps of a column ------------------------------
#pragma unroll
    for (int r = 0; r < ROUNDS; r++)
#pragma unroll
        for (int k = 0; k < kGramVals; k++)
        {
            const double w = warp_sum((double)acc[r][k]);
            if (lane == 0) s_red[warp][r * kGramVals + k] = w;
        }
    __syncthreads();
    const int nvals = a.c * kGramVals;
    for (int idx = tid; idx < nvals; idx += kGramThreads)
    {
        const int j = idx / kGramVals, k = idx % kGramVals;
        const int r = j / cols_per_round, col = j % cols_per_round;
        double t = 0.0;
        for (int p = 0; p < a.split; p++) t += s_red[col * a.split + p][r * kGramVals + k];
        partials[(size_t)blockIdx.x * (kMaxM * kGramVals) + idx] = t;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int idx = tid; idx < nvals; idx += kGramThreads)
    {
        double t = 0.0;
        for (unsigned b = 0; b < gridDim.x; b++) t += __ldcg(&partials[(size_t)b * (kMaxM * kGramVals) + idx]);
        result[idx] = t;
    }
    if (tid == 0) *ticket = 0u;
}

// ---- the O(c^2) recursion on coefficients: one thread ----------------------------------------------------
template <class T> struct GramSolveArgs
{
    int c, M, new_slot, with_v;
    T a;                     // scale of v
    const double* raw;       // [c][5] reduced dots (after the all-reduce)
    T* SY;                   // [M][M] by physical slot: SY[i*M+j] = s_i'y_j
    T* YY;                   // [M][M]
    const T* ys;             // [M]
    T* alpha;                // [M]
    const T* theta;
    T* coef;                 // out: [0] = cv, [1 + age] = cy_age, [1 + c + age] = cs_age   (age order, newest first)
    unsigned char slots[kMaxM];
};

template <class T> __global__ void k_gram_solve(GramSolveArgs<T> g)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int c = g.c, M = g.M;
    // 1. fold the new pair's row/column into the Gram matrices
    if (g.new_slot >= 0)
    {
        const int nw = g.new_slot;
        for (int i = 0; i < c; i++)
        {
            const int j = g.slots[i];
            g.SY[j * M + nw] = (T)g.raw[i * kGramVals + 2];   // s_j'y_new
            g.YY[j * M + nw] = (T)g.raw[i * kGramVals + 3];   // y_j'y_new
            g.YY[nw * M + j] = (T)g.raw[i * kGramVals + 3];
            g.SY[nw * M + j] = (T)g.raw[i * kGramVals + 4];   // s_new'y_j
        }
    }
    if (!g.with_v) return;
    const T theta = *g.theta;
    T cs[kMaxM];
    // 2. backward sweep (BFGSMat.h:285-290): alpha_j = s_j'q / ys_j with q = a*v - sum_{newer t} alpha_t y_t
    for (int i = 0; i < c; i++)
    {
        const int j = g.slots[i];
        T sq = g.a * (T)g.raw[i * kGramVals + 0];
        for (int t = 0; t < i; t++) sq -= g.alpha[g.slots[t]] * g.SY[j * M + g.slots[t]];
        g.alpha[j] = sq / g.ys[j];
    }
    // 3. forward sweep (BFGSMat.h:293-301): r = q/theta + sum_{older t} (alpha_t - beta_t) s_t ; beta_j = y_j'r / ys_j
    for (int i = c - 1; i >= 0; i--)
    {
        const int j = g.slots[i];
        T yq = g.a * (T)g.raw[i * kGramVals + 1];
        for (int t = 0; t < c; t++) yq -= g.alpha[g.slots[t]] * g.YY[j * M + g.slots[t]];
        T yr = yq / theta;
        for (int t = c - 1; t > i; t--) yr += cs[t] * g.SY[g.slots[t] * M + j];
        const T beta = yr / g.ys[j];
        cs[i] = g.alpha[j] - beta;
    }
    g.coef[0] = g.a / theta;
    for (int i = 0; i < c; i++)
    {
        g.coef[1 + i] = -(g.alpha[g.slots[i]] / theta);
        g.coef[1 + c + i] = cs[i];
    }
}

// ---- res = cv*v + sum_j cy_j*y_j + cs_j*s_j  (+ v.res) ----------------------------------------------------
template <class T> struct GramCombineArgs
{
    int64_t n, ld;
    const T* v;
    const T* S;
    const T* Y;
    T* res;
    const T* coef;
    int c;
    int want_dot;
    unsigned char slots[kMaxM];
};

template <class T, bool VEC>
__global__ void __launch_bounds__(kThreads) k_gram_combine(GramCombineArgs<T> a, ReduceBuf rb)
{
    __shared__ T s_coef[2 * kMaxM + 1];
    for (int i = threadIdx.x; i < 2 * a.c + 1; i += kThreads) s_coef[i] = a.coef[i];
    __syncthreads();
    const T cv = s_coef[0];
    T dot = T(0);
    const int64_t packs = (a.n + 3) >> 2, stride = (int64_t)gridDim.x * kThreads;
    for (int64_t p = (int64_t)blockIdx.x * kThreads + threadIdx.x; p < packs; p += stride)
    {
        const int64_t i0 = p << 2;
        const int cnt = (a.n - i0 >= 4) ? 4 : int(a.n - i0);
        const Pack<T> pv = load4<T, Hint::Stream, VEC>(a.v, i0, cnt);
        T r[4];
#pragma unroll
        for (int k = 0; k < 4; k++) r[k] = cv * pv.v[k];
        // y terms newest -> oldest, then s terms oldest -> newest (the order the recursion would add them)
#pragma unroll 4
        for (int j = 0; j < a.c; j++)
        {
            const Pack<T> py = load4<T, Hint::Stream, VEC>(a.Y + (int64_t)a.slots[j] * a.ld, i0, cnt);
            const T cy = s_coef[1 + j];
#pragma unroll
            for (int k = 0; k < 4; k++) r[k] += cy * py.v[k];
        }
#pragma unroll 4
        for (int j = a.c - 1; j >= 0; j--)
        {
            const Pack<T> ps = load4<T, Hint::Stream, VEC>(a.S + (int64_t)a.slots[j] * a.ld, i0, cnt);
            const T cs = s_coef[1 + a.c + j];
#pragma unroll
            for (int k = 0; k < 4; k++) r[k] += cs * ps.v[k];
        }
        Pack<T> out;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            out.v[k] = r[k];
            dot += (k < cnt) ? pv.v[k] * r[k] : T(0);
        }
        store4<T, Hint::Plain, VEC>(a.res, i0, cnt, out);
    }
    if (a.want_dot)
    {
        double dacc[1] = {(double)dot};
        grid_reduce<1>(dacc, rb);
    }
}

}  // namespace lb
