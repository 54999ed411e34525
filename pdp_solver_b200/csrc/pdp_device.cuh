// pdp_device.cuh -- grid-wide phases of the p-d-p loop, shared by the persistent cooperative kernel
// (pdp_loop.cu) and by the step-wise entry points.  Every phase is a __device__ function executed
// by ALL threads of a cooperative grid; the caller separates phases with grid.sync().
//
// The SP sweep exists twice:
//  * blocked passes (blk_clause_pass / blk_var_pass): the product path.  One CTA per block of whole
//    nodes, messages staged through shared memory, contiguous global reads and piece-wise contiguous
//    global writes (layout: pdp_common.cuh);
//  * generic passes (gen_*): thread per node with per-edge indexed global accesses into the same
//    layout.  They serve what the blocked passes leave out: the full [E,3] state, pi != 0, graphs whose
//    node degrees do not fit a block, and the problems on the sticky-NaN path.
//
// Work distribution of everything else: "warp-strided" loops -- warp w of the grid handles nodes [32w, 32w+32), then
// jumps by 32 * (warps in the grid).  Loads are coalesced and every lane sees a monotone sequence of
// problem ids (batches are laid out problem after problem), so per-problem reductions run as lane-local
// running accumulators that are flushed on a key change and merged warp-wide (__match_any_sync +
// redux) and block-wide (shared memory) before touching the per-problem atomics.
#pragma once
#include <cooperative_groups.h>

#include "pdp_common.cuh"

namespace cg = cooperative_groups;

struct KArgs {
    pdp_graph g;
    pdp_state s;
    int32_t* trace;      // optional decimation trace: triples (iteration, variable, sign)
    int32_t trace_cap;
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int64_t gwarp() { return ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; }
__device__ __forceinline__ int64_t gwarps() { return ((int64_t)gridDim.x * blockDim.x) >> 5; }
__device__ __forceinline__ int64_t gtid() { return (int64_t)blockIdx.x * blockDim.x + threadIdx.x; }
__device__ __forceinline__ int64_t gthreads() { return (int64_t)gridDim.x * blockDim.x; }

#define WARP_STRIDED(i, N) \
    for (int64_t i##_base = gwarp() * 32, i = i##_base + lane_id(); i##_base < (N); i##_base += gwarps() * 32, i = i##_base + lane_id())

// ------------------------------------------------------------------------------------------------
// keyed reduction helper.  ACC needs: void reset(); void merge_shfl(unsigned mask) [warp-reduce over
// lanes in `mask`, all of which hold the same key]; void commit(const pdp_state&, int key).
// ------------------------------------------------------------------------------------------------
template <typename ACC>
struct KeyedReducer {
    ACC acc;
    int key;
    __device__ __forceinline__ KeyedReducer() : key(-1) { acc.reset(); }
    // call for every item; commits the running accumulator when the key changes
    __device__ __forceinline__ void touch(const pdp_state& s, int k) {
        if (k != key) {
            if (key >= 0) acc.commit(s, key);
            key = k;
            acc.reset();
        }
    }
    // call once, by ALL threads of the block, outside of divergent code.  Warps whose lanes all hold
    // the same problem merge in registers and hand their partial to a block-level merge in shared
    // memory (one commit per block and problem run: a 1M-variable problem costs ~600 atomics per
    // pass instead of ~10^5); mixed warps commit per lane group.
    __device__ __forceinline__ void finish(const pdp_state& s) {
        __shared__ ACC sm_acc[32];
        __shared__ int sm_key[32];
        const unsigned m = __match_any_sync(0xffffffffu, key);
        const bool uniform = (m == 0xffffffffu);
        const int warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
        if (uniform) {
            acc.merge_full();
            if (lane_id() == 0) { sm_acc[warp] = acc; sm_key[warp] = key; }
        } else {
            acc.merge_group(m);
            if (lane_id() == (__ffs(m) - 1) && key >= 0) acc.commit(s, key);
            if (lane_id() == 0) sm_key[warp] = -1;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int w = 0;
            while (w < nwarp) {
                const int k = sm_key[w];
                if (k < 0) { ++w; continue; }
                ACC a = sm_acc[w];
                int v = w + 1;
                while (v < nwarp && sm_key[v] == k) { a.merge(sm_acc[v]); ++v; }
                a.commit(s, k);
                w = v;
            }
        }
        __syncthreads();
        key = -1;       // everything has been committed: a later touch() starts from scratch
        acc.reset();
    }
};

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int o) {
    unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, o), hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), o);
    return ((unsigned long long)hi << 32) | lo;
}

// ------------------------------------------------------------------------------------------------
// accumulators
// ------------------------------------------------------------------------------------------------
struct StatAcc {   // SequentialDecimator statistics: per problem max/min of two smooth-max vectors
    uint32_t mx0, mn0, mx1, mn1, nan, nav;
    __device__ __forceinline__ void reset() { mx0 = 0u; mn0 = 0x7f800000u; mx1 = 0u; mn1 = 0x7f800000u; nan = 0u; nav = 0u; }
    __device__ __forceinline__ void add(float v0, float v1, bool has1, uint32_t act) {
        if (v0 != v0) nan |= 1u; else { uint32_t u = f2u(v0); mx0 = max(mx0, u); mn0 = min(mn0, u); }
        if (has1) { if (v1 != v1) nan |= 2u; else { uint32_t u = f2u(v1); mx1 = max(mx1, u); mn1 = min(mn1, u); } }
        nav += act;
    }
    __device__ __forceinline__ void merge_shfl(unsigned m) {
        mx0 = __reduce_max_sync(m, mx0); mn0 = __reduce_min_sync(m, mn0);
        mx1 = __reduce_max_sync(m, mx1); mn1 = __reduce_min_sync(m, mn1);
        nan = __reduce_or_sync(m, nan); nav = __reduce_add_sync(m, nav);
    }
    __device__ __forceinline__ void merge_full() { merge_shfl(0xffffffffu); }
    __device__ __forceinline__ void merge_group(unsigned m) { merge_shfl(m); }
    __device__ __forceinline__ void merge(const StatAcc& o) {
        mx0 = max(mx0, o.mx0); mn0 = min(mn0, o.mn0); mx1 = max(mx1, o.mx1); mn1 = min(mn1, o.mn1); nan |= o.nan; nav += o.nav;
    }
    __device__ __forceinline__ void commit(const pdp_state& s, int b) const {
        atomicMax(&s.st_max[2 * b], mx0); atomicMin(&s.st_min[2 * b], mn0);
        atomicMax(&s.st_max[2 * b + 1], mx1); atomicMin(&s.st_min[2 * b + 1], mn1);
        if (nan) atomicOr(&s.st_nan[b], nan);
        if (nav) atomicAdd(&s.nav[b], (int)nav);
    }
};

struct CoefAcc {   // decimation coefficients |score| * active: per problem max / min / NaN
    uint32_t mx, mn, nan;
    __device__ __forceinline__ void reset() { mx = 0u; mn = 0x7f800000u; nan = 0u; }
    __device__ __forceinline__ void add(float c) {
        if (c != c) nan = 1u; else { uint32_t u = f2u(c); mx = max(mx, u); mn = min(mn, u); }
    }
    __device__ __forceinline__ void merge_shfl(unsigned m) {
        mx = __reduce_max_sync(m, mx); mn = __reduce_min_sync(m, mn); nan = __reduce_or_sync(m, nan);
    }
    __device__ __forceinline__ void merge_full() { merge_shfl(0xffffffffu); }
    __device__ __forceinline__ void merge_group(unsigned m) { merge_shfl(m); }
    __device__ __forceinline__ void merge(const CoefAcc& o) { mx = max(mx, o.mx); mn = min(mn, o.mn); nan |= o.nan; }
    __device__ __forceinline__ void commit(const pdp_state& s, int b) const {
        atomicMax(&s.c_max[b], mx); atomicMin(&s.c_min[b], mn);
        if (nan) atomicOr(&s.c_nan[b], nan);
    }
};

struct CountAcc {  // integer count into s.n_unsat
    int n;
    __device__ __forceinline__ void reset() { n = 0; }
    __device__ __forceinline__ void merge_shfl(unsigned m) { n = __reduce_add_sync(m, n); }
    __device__ __forceinline__ void merge_full() { merge_shfl(0xffffffffu); }
    __device__ __forceinline__ void merge_group(unsigned m) { merge_shfl(m); }
    __device__ __forceinline__ void merge(const CountAcc& o) { n += o.n; }
    __device__ __forceinline__ void commit(const pdp_state& s, int b) const { if (n) atomicAdd(&s.n_unsat[b], n); }
};

struct EnergyAcc { // integer count into s.energy
    int n;
    __device__ __forceinline__ void reset() { n = 0; }
    __device__ __forceinline__ void merge_shfl(unsigned m) { n = __reduce_add_sync(m, n); }
    __device__ __forceinline__ void merge_full() { merge_shfl(0xffffffffu); }
    __device__ __forceinline__ void merge_group(unsigned m) { merge_shfl(m); }
    __device__ __forceinline__ void merge(const EnergyAcc& o) { n += o.n; }
    __device__ __forceinline__ void commit(const pdp_state& s, int b) const { if (n) atomicAdd(&s.energy[b], n); }
};

// WalkSAT candidate selection (solver.py:453-458): per problem
//   kg = min over variables of (delta, index)            -> first index of the minimum energy delta
//   kr = max over variables of (fl(x+1) bits, ~index)    -> first index of the maximum random key
//   mn / mx = min / max of x (float bits) for the exact handling of min x > 0
struct PickAcc {
    unsigned long long kg, kr;
    uint32_t mn, mx;
    __device__ __forceinline__ void reset() { kg = ~0ull; kr = 0ull; mn = 0x7f800000u; mx = 0u; }
    __device__ __forceinline__ void add(unsigned long long g, unsigned long long r, uint32_t xb) {
        kg = g < kg ? g : kg; kr = r > kr ? r : kr; mn = min(mn, xb); mx = max(mx, xb);
    }
    __device__ __forceinline__ void merge(const PickAcc& o) {
        kg = o.kg < kg ? o.kg : kg; kr = o.kr > kr ? o.kr : kr; mn = min(mn, o.mn); mx = max(mx, o.mx);
    }
    __device__ __forceinline__ void merge_full() {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long g = shfl_xor_u64(kg, o), r = shfl_xor_u64(kr, o);
            kg = g < kg ? g : kg; kr = r > kr ? r : kr;
        }
        mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
    }
    // mixed warp: serialise over the (few) distinct keys; every lane of a group ends with the group's result
    __device__ __forceinline__ void merge_group(unsigned m) {
        mn = __reduce_min_sync(m, mn); mx = __reduce_max_sync(m, mx);
        unsigned todo = m;
        unsigned long long g = kg, r = kr;
        while (todo) {
            const int src = __ffs(todo) - 1;
            const unsigned long long og = __shfl_sync(m, g, src), orr = __shfl_sync(m, r, src);
            kg = og < kg ? og : kg; kr = orr > kr ? orr : kr;
            todo &= todo - 1;
        }
    }
    __device__ __forceinline__ void commit(const pdp_state& s, int b) const {
        atomicMin(&s.ws_key[2 * b], kg); atomicMax(&s.ws_key[2 * b + 1], kr);
        atomicMin(&s.ws_best[2 * b], mn); atomicMax(&s.ws_best[2 * b + 1], mx);
    }
};


// ------------------------------------------------------------------------------------------------
// activity masks.  The edge mask em(e) = active_variable(i(e)) * active_function(a(e)) of the reference
// (solver.py:370-371) is kept as one bit per edge, indexed by the edge's position in each of the two
// message layouts (g.vmask / g.qmask) and set where a node is de-activated: a pass reads the mask bits
// of its region with the same contiguous access as the messages and never gathers the node masks.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mask_edge(const pdp_graph& g, int vpos, int qpos) {
    atomicOr(&g.vmask[vpos >> 5], 1u << (vpos & 31));
    atomicOr(&g.qmask[qpos >> 5], 1u << (qpos & 31));
}
__device__ __forceinline__ void deactivate_variable(const pdp_graph& g, const pdp_state& s, int i) {
    s.av[i] = 0;
    for (int p = g.var_ptr[i]; p < g.var_ptr[i + 1]; ++p) mask_edge(g, g.p_vpos[p], g.p_qpos[p]);
}
__device__ __forceinline__ void deactivate_clause(const pdp_graph& g, const pdp_state& s, int a) {
    s.af[a] = 0;
    for (int c = g.cl_ptr[a]; c < g.cl_ptr[a + 1]; ++c) mask_edge(g, g.c_vpos[c], g.c_qpos[c]);
}
// the words change between passes: read around L1
__device__ __forceinline__ bool mbit(const uint32_t* words, int pos) { return (__ldcg(words + (pos >> 5)) >> (pos & 31)) & 1u; }

// variable side of the SP update with the per-variable part hoisted (pi == 0): for one sign s the terms
// 0.5(1+s)P + 0.5(1-s)N, opp and exp(opp) do not depend on the edge.  Same operations, same order as
// sp_var_update_qu.
__device__ __forceinline__ void sp_var_prepare(float P, float N, float s, float& same_base, float& opp, float& O) {
    same_base = 0.5f * (1.f + s) * P + 0.5f * (1.f - s) * N;
    opp = 0.5f * (1.f - s) * P + 0.5f * (1.f + s) * N;
    opp += 0.f;
    O = X30(opp);
}
__device__ __forceinline__ float sp_var_finish(float same_base, float opp, float O, float y) {
    float same = same_base - y;
    same += 0.f;
    const float dc = X30(same + opp);
    const float S = X30(same);
    const float u = S * (1.f - O), v = O * (1.f - S);
    const float total = u + v + dc;
    return pdp_divf(u, total);
}

// ------------------------------------------------------------------------------------------------
// generic passes: thread per node, indexed global accesses.  MODE selects the problems.
// ------------------------------------------------------------------------------------------------
enum { GEN_ALL = 0, GEN_NAN = 1 };
template <int MODE>
__device__ __forceinline__ bool gen_take(const pdp_state& s, int b) {
    if (!s.active[b]) return false;
    return (MODE == GEN_ALL) ? true : (s.nanflag[b] != 0);
}

// clause side: eta'(e) = exp(min(sum_{e' in a(e)} x_e' - x_e, 30)), x = log(max(q_u,1e-40)) * em
// (pdp_propagate.py:166-175).  r = buffer of the previous surveys.
template <int MODE>
__device__ __forceinline__ void gen_clause_side(const KArgs& A, int r, bool use_mask) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    const float* __restrict__ qin = s.qu;
    const float* __restrict__ eold = s.eta[r];
    float* __restrict__ eout = s.eta[r ^ 1];
    WARP_STRIDED(a, g.F) {
        if (a >= g.F) continue;
        const int b = g.bfm[a];
        if (!gen_take<MODE>(s, b)) continue;
        const bool um = use_mask && s.masked[b];
        // the reference blends `mask*new + (1-mask)*old` arithmetically, so a NaN message is sticky
        // (0*NaN); problems that have produced a NaN re-read the old value
        const bool sticky = s.nanflag[b] != 0;
        bool made_nan = false;
        const int beg = g.cl_ptr[a], end = g.cl_ptr[a + 1];
        float tot = 0.f;
        for (int c = beg; c < end; ++c) {
            const int qp = g.c_qpos[c];
            float v = L40(qin[qp]);
            if (um && mbit(g.qmask, qp)) v = v * 0.f;
            tot += v;
        }
        for (int c = beg; c < end; ++c) {
            const int qp = g.c_qpos[c];
            float v = L40(qin[qp]);
            if (um && mbit(g.qmask, qp)) v = v * 0.f;
            const int pos = g.c_vpos[c];
            float nv = X30(tot - v);
            if (sticky) { const float ov = eold[pos]; if (ov != ov) nv = ov; }
            made_nan |= (nv != nv);
            eout[pos] = nv;
        }
        if (made_nan && !sticky) s.nanpend[b] = 1;
    }
}

// variable side (pdp_propagate.py:184-218): q(t) from eta[r], written in place
template <int MODE, bool FULL>
__device__ __forceinline__ void gen_var_side(const KArgs& A, int r, bool use_mask, float pi) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    const float* __restrict__ ein = s.eta[r];
    WARP_STRIDED(i, g.V) {
        if (i >= g.V) continue;
        const int b = g.bvm[i];
        if (!gen_take<MODE>(s, b)) continue;
        const bool um = use_mask && s.masked[b];
        const int beg = g.var_ptr[i], end = g.var_ptr[i + 1];
        const bool sticky = s.nanflag[b] != 0;
        bool made_nan = false;
        float P = 0.f, N = 0.f;
        for (int p = beg; p < end; ++p) {
            const int vp = g.p_vpos[p];
            const bool neg = (g.v_cedge[p] & PDP_SIGN_BIT) != 0u;
            float y = L40(1.f - ein[vp]);
            if (um && mbit(g.vmask, vp)) y = y * 0.f;
            // the reference's pos/neg incidence matrices hold explicit zeros: 0*y keeps NaN alive
            P += (neg ? 0.f : 1.f) * y;
            N += (neg ? 1.f : 0.f) * y;
        }
        for (int p = beg; p < end; ++p) {
            const int vp = g.p_vpos[p];
            float y = L40(1.f - ein[vp]);
            if (um && mbit(g.vmask, vp)) y = y * 0.f;
            const float sg = (g.v_cedge[p] & PDP_SIGN_BIT) ? -1.f : 1.f;
            const int qp = g.p_qpos[p];
            if (FULL || pi != 0.f) {
                float u, v, d;
                sp_var_update(P, N, y, sg, s.ext[p], pi, u, v, d);
                if (sticky) {
                    const float ou = s.qu[qp]; if (ou != ou) u = ou;
                    if (FULL) { const float ov = s.qs[qp], od = s.qd[qp]; if (ov != ov) v = ov; if (od != od) d = od; }
                }
                made_nan |= (u != u);
                s.qu[qp] = u;
                if (FULL) { s.qs[qp] = v; s.qd[qp] = d; }
            } else {
                float u = sp_var_update_qu(P, N, y, sg);
                if (sticky) { const float ou = s.qu[qp]; if (ou != ou) u = ou; }
                made_nan |= (u != u);
                s.qu[qp] = u;
            }
        }
        if (made_nan && !sticky) s.nanpend[b] = 1;
    }
}

// decimator statistics (pdp_decimate.py:127-143, util.py:282-286): per variable smooth-max of the new
// surveys and of |eta_prev - eta_new| * edge_mask, times active_variables; per problem max / min.
// w = buffer holding the new surveys.
template <int MODE>
__device__ __forceinline__ void gen_stats(const KArgs& A, int w, bool has_prev, bool em_set) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    const float* __restrict__ en = s.eta[w];
    const float* __restrict__ eo = s.eta[w ^ 1];
    KeyedReducer<StatAcc> red;
    WARP_STRIDED(i, g.V) {
        if (i >= g.V) continue;
        const int b = g.bvm[i];
        if (!gen_take<MODE>(s, b)) continue;
        red.touch(s, b);
        const int beg = g.var_ptr[i], end = g.var_ptr[i + 1];
        const uint32_t act = s.av[i];
        const bool um = em_set && s.masked[b];
        float n0 = 0.f, d0 = 0.f, n1 = 0.f, d1 = 0.f;
        for (int p = beg; p < end; ++p) {
            const int pos = g.p_vpos[p];
            const float v = en[pos];
            const float c = X30(30.f * v);
            n0 += v * c; d0 += c;
            if (has_prev) {
                float d = fabsf(eo[pos] - v);
                if (um && mbit(g.vmask, pos)) d = d * 0.f;
                const float cd = X30(30.f * d);
                n1 += d * cd; d1 += cd;
            }
        }
        const float sm0 = pdp_divs(n0, tmaxf(d0, 1.0f)) * (float)act;
        const float sm1 = pdp_divs(n1, tmaxf(d1, 1.0f)) * (float)act;
        red.acc.add(sm0, sm1, has_prev, act);
    }
    red.finish(s);
}

// ------------------------------------------------------------------------------------------------
// blocked passes.  Dynamic shared memory (PDP_SWEEP_SMEM bytes):
//   [0, 4*PDP_BLK_C)           clause pass: one plane X;  variable pass: planes PA | PB (PDP_BLK_V each)
//   [4*PDP_BLK_C, +PDP_BLK_C/8) skip bits: slots of nodes the pass leaves alone (frozen / sticky-NaN
//                               problems inside a block that has work)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool blk_problem_runs(const pdp_state& s, int b) { return s.active[b] && !s.nanflag[b]; }

// true when no problem in [b0, b1] is to be processed by the blocked passes (uniform over the CTA)
__device__ __forceinline__ bool blk_idle(const pdp_state& s, int b0, int b1) {
    if (b0 == b1) return !blk_problem_runs(s, b0);
    int any = 0;
    for (int b = b0 + (int)threadIdx.x; b <= b1; b += (int)blockDim.x) any |= blk_problem_runs(s, b) ? 1 : 0;
    return __syncthreads_or(any) == 0;
}

__device__ __forceinline__ void blk_mark_skip(uint32_t* skip, int lo, int hi) {
    for (int l = lo; l < hi; ++l) atomicOr(&skip[l >> 5], 1u << (l & 31));
}

#define NT PDP_SWEEP_THREADS   // compile-time stride of the block-wide loops (immediate address offsets)
// PDP_PHASE_TIMING (profiling builds only): thread 0 of every CTA adds the clock cycles (>> 10) it spent in
// each phase of the blocked passes to the trace buffer: [0..2] clause load / node / write-out, [3..5] variable
#ifdef PDP_PHASE_TIMING
#define PHASE_T0() long long _pt = clock64()
#define PHASE_ADD(slot) do { if (threadIdx.x == 0 && A.trace) { const long long _n = clock64(); atomicAdd(&A.trace[slot], (int)((_n - _pt) >> 10)); _pt = _n; } } while (0)
#else
#define PHASE_T0() do {} while (0)
#define PHASE_ADD(slot) do {} while (0)
#endif

// pulls [base, base + bytes) into L2 (one request per 128-byte line).  The per-node phases are issue bound
// and leave the memory system idle: they start by prefetching what the NEXT block of this CTA will load and
// the write-out tables of the current one, so that the latency-bound load / write-out phases hit L2.
__device__ __forceinline__ void blk_prefetch_l2(const void* base, size_t bytes) {
    const char* p = reinterpret_cast<const char*>(base);
    const char* end = p + bytes;
    p = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)127);
    for (p += (size_t)threadIdx.x * 128; p < end; p += (size_t)NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// write-out: slots [0, ne) of the block in ascending destination order; consecutive slots mostly hit
// consecutive destinations (runs), so a warp's stores coalesce into a few sectors
template <bool SKIP>
__device__ __forceinline__ void blk_write_out_t(const uint16_t* __restrict__ src, const int32_t* __restrict__ dst, int ne,
                                                const float* plane, const uint32_t* skip, float* __restrict__ out) {
    int w = threadIdx.x;
    constexpr int U = 8;
    for (; w + (U - 1) * NT < ne; w += U * NT) {
        int l[U], d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { l[u] = src[w + u * NT]; d[u] = dst[w + u * NT]; }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (!SKIP || !((skip[l[u] >> 5] >> (l[u] & 31)) & 1u)) out[d[u]] = plane[l[u]];
    }
    for (; w < ne; w += NT) {
        const int l = src[w];
        if (!SKIP || !((skip[l >> 5] >> (l & 31)) & 1u)) out[dst[w]] = plane[l];
    }
}
__device__ __forceinline__ void blk_write_out(const uint16_t* __restrict__ src, const int32_t* __restrict__ dst, int ne,
                                              const float* plane, const uint32_t* skip, bool any_skip, float* __restrict__ out) {
    if (any_skip) blk_write_out_t<true>(src, dst, ne, plane, skip, out);
    else blk_write_out_t<false>(src, dst, ne, plane, skip, out);
}

// clause pass, load phase: x = log(max(q_u, 1e-40)) * em, scattered into clause-major order.
// e0 = first C-layout position of the block (the mask bits are indexed by position).
template <bool MASKED>
__device__ __forceinline__ void blk_clause_load(const float* __restrict__ qsrc, const uint16_t* __restrict__ inv,
                                                const uint32_t* __restrict__ qmask, int e0, int ne, float* X) {
    int x = threadIdx.x;
    constexpr int U = 8;
    for (; x + (U - 1) * NT < ne; x += U * NT) {
        float q[U]; int l[U]; bool m[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { q[u] = qsrc[x + u * NT]; l[u] = inv[x + u * NT]; m[u] = MASKED ? mbit(qmask, e0 + x + u * NT) : false; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float v = L40(q[u]);
            if (MASKED && m[u]) v = v * 0.f;
            X[l[u]] = v;
        }
    }
    for (; x < ne; x += NT) {
        float v = L40(qsrc[x]);
        if (MASKED && mbit(qmask, e0 + x)) v = v * 0.f;
        X[inv[x]] = v;
    }
}

// one clause of K literals held in X[lo .. lo+K): surveys in place.  Returns whether a NaN was produced.
template <int K>
__device__ __forceinline__ bool blk_clause_body(float* X, int lo) {
    float x[K];
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) { x[j] = X[lo + j]; tot += x[j]; }
    bool made_nan = false;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const float nv = X30(tot - x[j]);
        made_nan |= (nv != nv);
        X[lo + j] = nv;
    }
    return made_nan;
}
__device__ __forceinline__ bool blk_clause_body_any(float* X, int lo, int k) {
    float tot = 0.f;
    for (int j = 0; j < k; ++j) tot += X[lo + j];
    bool made_nan = false;
    for (int j = 0; j < k; ++j) {
        const float nv = X30(tot - X[lo + j]);
        made_nan |= (nv != nv);
        X[lo + j] = nv;
    }
    return made_nan;
}

// clause pass of iteration t: eta(t) [buffer r^1, V-layout] from q(t-1) [C-layout]
__device__ __forceinline__ void blk_clause_pass(const KArgs& A, int r, bool use_mask, unsigned char* smem) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    float* X = reinterpret_cast<float*>(smem);
    uint32_t* skip = reinterpret_cast<uint32_t*>(smem + 4 * PDP_BLK_C);
    __shared__ int sm_any_skip;
    const float* __restrict__ qin = s.qu;
    float* __restrict__ eout = s.eta[r ^ 1];
    const int tid = threadIdx.x;
    for (int blk = blockIdx.x; blk < g.ncb; blk += gridDim.x) {
        const int a0 = g.cb_ptr[blk], a1 = g.cb_ptr[blk + 1];
        if (a1 <= a0) continue;
        const int b0 = g.bfm[a0], b1 = g.bfm[a1 - 1];
        if (blk_idle(s, b0, b1)) continue;
        const bool multi = (b0 != b1);
        const int e0 = g.cl_ptr[a0], ne = g.cl_ptr[a1] - e0;
        const int ku = g.cb_k[blk];
        for (int i = tid; i < (ne + 31) / 32; i += NT) skip[i] = 0u;
        if (tid == 0) sm_any_skip = 0;
        // ---- load (contiguous), log, edge mask, scatter into clause-major order
        PHASE_T0();
        if (use_mask && (multi || s.masked[b0])) blk_clause_load<true>(qin + e0, g.cinv + e0, g.qmask, e0, ne, X);
        else blk_clause_load<false>(qin + e0, g.cinv + e0, g.qmask, e0, ne, X);
        __syncthreads();
        PHASE_ADD(0);
        {   // L2 prefetch: this block's write-out tables, the next block's inputs
            blk_prefetch_l2(g.csrc + e0, (size_t)ne * 2);
            blk_prefetch_l2(g.cdst + e0, (size_t)ne * 4);
            const int nb = blk + gridDim.x;
            if (nb < g.ncb) {
                const int n0 = g.cl_ptr[g.cb_ptr[nb]], n1 = g.cl_ptr[g.cb_ptr[nb + 1]];
                blk_prefetch_l2(qin + n0, (size_t)(n1 - n0) * 4);
                blk_prefetch_l2(g.cinv + n0, (size_t)(n1 - n0) * 2);
            }
        }
        // ---- thread per clause
        for (int a = a0 + tid; a < a1; a += NT) {
            int lo, k;
            if (ku) { k = ku; lo = (a - a0) * ku; }
            else { lo = g.cl_ptr[a] - e0; k = g.cl_ptr[a + 1] - e0 - lo; }
            int b = b0;
            if (multi) {
                b = g.bfm[a];
                if (!blk_problem_runs(s, b)) { blk_mark_skip(skip, lo, lo + k); sm_any_skip = 1; continue; }
            }
            bool made_nan;
            switch (k) {
                case 3: made_nan = blk_clause_body<3>(X, lo); break;
                case 4: made_nan = blk_clause_body<4>(X, lo); break;
                case 5: made_nan = blk_clause_body<5>(X, lo); break;
                case 2: made_nan = blk_clause_body<2>(X, lo); break;
                default: made_nan = blk_clause_body_any(X, lo, k); break;
            }
            if (made_nan) s.nanpend[b] = 1;
        }
        __syncthreads();
        PHASE_ADD(1);
        blk_write_out(g.csrc + e0, g.cdst + e0, ne, X, skip, sm_any_skip != 0, eout);
        __syncthreads();
        PHASE_ADD(2);
    }
}

// statistics of multi-problem blocks: per block-local problem in shared memory
#define PDP_STAT_SLOTS 64
struct BlkStats {
    uint32_t mx0[PDP_STAT_SLOTS], mn0[PDP_STAT_SLOTS], mx1[PDP_STAT_SLOTS], mn1[PDP_STAT_SLOTS], nan[PDP_STAT_SLOTS], nav[PDP_STAT_SLOTS];
};

// branch-free select (the compiler would otherwise split the update loop into one path per literal sign)
__device__ __forceinline__ float fsel(uint32_t mask, float a, float b) {
    return __uint_as_float((__float_as_uint(a) & mask) | (__float_as_uint(b) & ~mask));
}

// variable pass, load phase.  The surveys are non-negative, so their sign bits carry the two per-edge
// flags the variable loops need: PA (new survey) sign = edge masked, PB (old survey) sign = negative literal.
template <bool MASKED>
__device__ __forceinline__ void blk_var_load(const float* __restrict__ sn, const float* __restrict__ so, const uint16_t* __restrict__ inv,
                                             const uint32_t* __restrict__ vmask, int e0, int ne, float* PA, float* PB) {
    int x = threadIdx.x;
    constexpr int U = 6;
    for (; x + (U - 1) * NT < ne; x += U * NT) {
        uint32_t n[U], o[U], iv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            n[u] = __float_as_uint(sn[x + u * NT]); o[u] = __float_as_uint(so[x + u * NT]); iv[u] = inv[x + u * NT];
            if (MASKED) n[u] |= mbit(vmask, e0 + x + u * NT) ? 0x80000000u : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int l = iv[u] & 0x7fff;
            PA[l] = __uint_as_float(n[u]); PB[l] = __uint_as_float(o[u] ^ ((iv[u] & PDP_VINV_NEG) << 16));
        }
    }
    for (; x < ne; x += NT) {
        uint32_t n0 = __float_as_uint(sn[x]);
        const uint32_t o0 = __float_as_uint(so[x]), i0 = inv[x];
        if (MASKED) n0 |= mbit(vmask, e0 + x) ? 0x80000000u : 0u;
        const int l0 = i0 & 0x7fff;
        PA[l0] = __uint_as_float(n0); PB[l0] = __uint_as_float(o0 ^ ((i0 & PDP_VINV_NEG) << 16));
    }
}

// variable pass of iteration t: the decimator statistics of eta(t) [buffer r^1] against eta(t-1)
// [buffer r], and q(t) [C-layout, in place] from eta(t-1).  Requires eta(t-1) >= +0 or NaN without sign.
__device__ __forceinline__ void blk_var_pass(const KArgs& A, int r, bool use_mask, bool has_prev, bool em_set, unsigned char* smem) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    float* PA = reinterpret_cast<float*>(smem);   // eta(t), then q(t)
    float* PB = PA + PDP_BLK_V;                   // eta(t-1), then y
    uint32_t* skip = reinterpret_cast<uint32_t*>(smem + 4 * PDP_BLK_C);
    __shared__ int sm_any_skip;
    __shared__ BlkStats sm_st;
    const float* __restrict__ en = s.eta[r ^ 1];
    const float* __restrict__ eo = s.eta[r];
    const int tid = threadIdx.x;
    KeyedReducer<StatAcc> red;
    for (int blk = blockIdx.x; blk < g.nvb; blk += gridDim.x) {
        const int v0 = g.vb_ptr[blk], v1 = g.vb_ptr[blk + 1];
        if (v1 <= v0) continue;
        const int b0 = g.bvm[v0], b1 = g.bvm[v1 - 1];
        if (blk_idle(s, b0, b1)) continue;
        const bool multi = (b0 != b1);
        const bool local_stats = multi && (b1 - b0 < PDP_STAT_SLOTS);   // else: registers (one problem) or global atomics
        const int e0 = g.var_ptr[v0], ne = g.var_ptr[v1] - e0;
        for (int i = tid; i < (ne + 31) / 32; i += NT) skip[i] = 0u;
        if (tid == 0) sm_any_skip = 0;
        if (local_stats && tid <= b1 - b0) {
            sm_st.mx0[tid] = 0u; sm_st.mn0[tid] = 0x7f800000u; sm_st.mx1[tid] = 0u; sm_st.mn1[tid] = 0x7f800000u;
            sm_st.nan[tid] = 0u; sm_st.nav[tid] = 0u;
        }
        // ---- load both survey regions (contiguous), scatter into variable-major order
        PHASE_T0();
        if ((use_mask || em_set) && (multi || s.masked[b0])) blk_var_load<true>(en + e0, eo + e0, g.vinv + e0, g.vmask, e0, ne, PA, PB);
        else blk_var_load<false>(en + e0, eo + e0, g.vinv + e0, g.vmask, e0, ne, PA, PB);
        __syncthreads();
        PHASE_ADD(3);
        {   // L2 prefetch: this block's write-out tables, the next block's inputs
            blk_prefetch_l2(g.vsrc + e0, (size_t)ne * 2);
            blk_prefetch_l2(g.vdst + e0, (size_t)ne * 4);
            const int nb = blk + gridDim.x;
            if (nb < g.nvb) {
                const int n0 = g.var_ptr[g.vb_ptr[nb]], n1 = g.var_ptr[g.vb_ptr[nb + 1]];
                blk_prefetch_l2(en + n0, (size_t)(n1 - n0) * 4);
                blk_prefetch_l2(eo + n0, (size_t)(n1 - n0) * 4);
                blk_prefetch_l2(g.vinv + n0, (size_t)(n1 - n0) * 2);
            }
        }
        // ---- thread per variable (descending degree): ordered sums, statistics, update.
        // rounds alternate direction over the degree-sorted list: every thread gets high and low degrees
        for (int base = v0, round = 0; base < v1; base += NT, ++round) {
            const int t = (round & 1) ? (base + NT - 1 - tid) : (base + tid);
            if (t >= v1) continue;
            const int2 ve = __ldg(&g.vsort[t]);
            const int i = ve.x, lo = ve.y & 0xffff, deg = ve.y >> 16;
            int b = b0;
            if (multi) {
                b = g.bvm[i];
                if (!blk_problem_runs(s, b)) { blk_mark_skip(skip, lo, lo + deg); sm_any_skip = 1; continue; }
            }
            const uint32_t act = s.av[i];
            float P = 0.f, N = 0.f, n0 = 0.f, d0 = 0.f, n1 = 0.f, d1 = 0.f;
            for (int j = 0; j < deg; ++j) {
                const uint32_t nb = __float_as_uint(PA[lo + j]), ob = __float_as_uint(PB[lo + j]);
                const bool m = (nb >> 31) != 0u;                 // edge masked
                const uint32_t negm = (uint32_t)((int32_t)ob >> 31);   // all ones: negative literal
                const float xn = __uint_as_float(nb & 0x7fffffffu), xo = __uint_as_float(ob & 0x7fffffffu);
                float y = L40(1.f - xo);
                if (use_mask && m) y = y * 0.f;
                // y <= +0 (or NaN): stored as |y| under the literal's sign bit
                PB[lo + j] = __uint_as_float((__float_as_uint(y) & 0x7fffffffu) | (ob & 0x80000000u));
                // the reference's pos/neg incidence matrices hold explicit zeros: 0*y keeps NaN alive
                const float zy = 0.f * y;
                P += fsel(negm, zy, y);
                N += fsel(negm, y, zy);
                const float c = X30(30.f * xn);
                n0 += xn * c; d0 += c;
                if (has_prev) {
                    float d = fabsf(xo - xn);
                    if (em_set && m) d = d * 0.f;
                    const float cd = X30(30.f * d);
                    n1 += d * cd; d1 += cd;
                }
            }
            const float sm0 = pdp_divs(n0, tmaxf(d0, 1.0f)) * (float)act;
            const float sm1 = pdp_divs(n1, tmaxf(d1, 1.0f)) * (float)act;
            if (!multi) {
                red.touch(s, b);
                red.acc.add(sm0, sm1, has_prev, act);
            } else {
                StatAcc one;
                one.reset();
                one.add(sm0, sm1, has_prev, act);
                if (local_stats) {
                    const int lb = b - b0;
                    atomicMax(&sm_st.mx0[lb], one.mx0); atomicMin(&sm_st.mn0[lb], one.mn0);
                    if (has_prev) { atomicMax(&sm_st.mx1[lb], one.mx1); atomicMin(&sm_st.mn1[lb], one.mn1); }
                    if (one.nan) atomicOr(&sm_st.nan[lb], one.nan);
                    if (act) atomicAdd(&sm_st.nav[lb], act);
                } else {
                    one.commit(s, b);
                }
            }
            float sb_pos, opp_pos, O_pos, sb_neg, opp_neg, O_neg;
            sp_var_prepare(P, N, 1.f, sb_pos, opp_pos, O_pos);
            sp_var_prepare(P, N, -1.f, sb_neg, opp_neg, O_neg);
            bool made_nan = false;
            for (int j = 0; j < deg; ++j) {
                const uint32_t yb = __float_as_uint(PB[lo + j]);
                const uint32_t negm = (uint32_t)((int32_t)yb >> 31);
                const float y = __uint_as_float(yb | 0x80000000u);   // -|y|; -0 for +0 is erased by `same += 0`
                const float u = sp_var_finish(fsel(negm, sb_neg, sb_pos), fsel(negm, opp_neg, opp_pos), fsel(negm, O_neg, O_pos), y);
                made_nan |= (u != u);
                PA[lo + j] = u;
            }
            if (made_nan) s.nanpend[b] = 1;
        }
        __syncthreads();
        if (local_stats && tid <= b1 - b0) {
            StatAcc a;
            a.mx0 = sm_st.mx0[tid]; a.mn0 = sm_st.mn0[tid]; a.mx1 = sm_st.mx1[tid]; a.mn1 = sm_st.mn1[tid];
            a.nan = sm_st.nan[tid]; a.nav = sm_st.nav[tid];
            if (a.mn0 != 0x7f800000u || a.mx0 != 0u || a.nan || a.nav || a.mn1 != 0x7f800000u) a.commit(s, b0 + tid);
        }
        PHASE_ADD(4);
        blk_write_out(g.vsrc + e0, g.vdst + e0, ne, PA, skip, sm_any_skip != 0, s.qu);
        red.finish(s);   // block-level merge of the statistics; its barriers also fence the planes
        PHASE_ADD(5);
    }
}
#undef NT

// ------------------------------------------------------------------------------------------------
// SurveyScorer over the converged problems (pdp_predict.py:155-192) + coefficient statistics
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float score_variable(const pdp_graph& g, const pdp_state& s, const float* __restrict__ eta,
                                                int i, float pi) {
    const int beg = g.var_ptr[i], end = g.var_ptr[i + 1];
    float extsum = 0.f, ps = 0.f, ns = 0.f, as = 0.f;
    for (int p = beg; p < end; ++p) {
        extsum += s.ext[p];
        const float f = L10(1.f - eta[g.p_vpos[p]]) * (float)s.af[g.v_cls[p]];
        const bool neg = (g.v_cedge[p] & PDP_SIGN_BIT) != 0u;
        ps += (neg ? 0.f : 1.f) * f;
        ns += (neg ? 1.f : 0.f) * f;
        as += f;
    }
    return sp_score_tail(ps, ns, as, sgnf(extsum), pi);
}

__device__ __forceinline__ void score_phase(const KArgs& A, int w, float pi) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    KeyedReducer<CoefAcc> red;
    WARP_STRIDED(i, g.V) {
        if (i >= g.V) continue;
        const int b = g.bvm[i];
        if (!s.conv[b]) continue;
        red.touch(s, b);
        const float sc = score_variable(g, s, s.eta[w], (int)i, pi);
        s.score[i] = sc;
        red.acc.add(fabsf(sc) * (float)s.av[i]);
    }
    red.finish(s);
}

// first index attaining max of fl(fl(c - min) + 1)  (util.py:257-265)
__device__ __forceinline__ void argmax_phase(const KArgs& A) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    WARP_STRIDED(i, g.V) {
        if (i >= g.V) continue;
        const int b = g.bvm[i];
        if (!s.conv[b] || s.c_nan[b]) continue;
        const float m = u2f(s.c_min[b]);
        const float kmax = argmax_key(u2f(s.c_max[b]), m);
        const float c = fabsf(s.score[i]) * (float)s.av[i];
        if (argmax_key(c, m) == kmax) atomicMin(&s.arg_idx[b], (int)i);
    }
}

// fixing one variable (the per-variable form of _set_variable_core, solver.py:205-226): clauses
// holding a now-true occurrence are de-activated, the variable is de-activated, solution updated
__device__ __forceinline__ void fix_variable(const pdp_graph& g, const pdp_state& s, int i, float sg) {
    const int beg = g.var_ptr[i], end = g.var_ptr[i + 1];
    for (int p = beg; p < end; ++p) {
        const float lit = (g.v_cedge[p] & PDP_SIGN_BIT) ? -1.f : 1.f;
        const int a = g.v_cls[p];
        if (lit * sg > 0.f && s.af[a]) deactivate_clause(g, s, a);
    }
    deactivate_variable(g, s, i);
    s.sol[i] = (sg + 1.f) / 2.0f;
}


// ------------------------------------------------------------------------------------------------
// unit propagation round (solver.py:228-273), split at its data dependencies
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void up_find_units(const KArgs& A, int flag_slot) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    WARP_STRIDED(a, g.F) {
        if (a >= g.F) continue;
        if (!s.af[a]) continue;
        const int b = g.bfm[a];
        if (!s.dirty[b]) continue;
        const int beg = g.cl_ptr[a], end = g.cl_ptr[a + 1];
        int deg = 0; uint32_t hit = 0;
        for (int c = beg; c < end; ++c) {
            const uint32_t w = g.c_var[c];
            if (s.av[w & PDP_IDX_MASK]) { ++deg; hit = w; }
        }
        if (deg == 1) {
            s.single[a] = 1;
            const int j = (int)(hit & PDP_IDX_MASK);
            atomicAdd(&s.up_cnt[j], 1);
            atomicAdd(&s.up_ev[j], (hit & PDP_SIGN_BIT) ? -1 : 1);
            s.ctrl[flag_slot] = 1;
        }
    }
}

__device__ __forceinline__ void up_find_conflicts(const KArgs& A) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    WARP_STRIDED(i, g.V) {
        if (i >= g.V) continue;
        const int cnt = s.up_cnt[i];
        if (cnt > 0 && abs(s.up_ev[i]) != cnt) atomicAdd(&s.conflicts[g.bvm[i]], 1);
    }
}

// quirk kept from the reference (solver.py:257,261): the problem is wiped only when its conflict
// COUNT equals exactly one
__device__ __forceinline__ void up_apply_conflicts(const KArgs& A) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    const int64_t n = g.V > g.F ? g.V : g.F;
    WARP_STRIDED(i, n) {
        if (i < g.F) {
            if (s.single[i]) { deactivate_clause(g, s, (int)i); s.single[i] = 0; s.masked[g.bfm[i]] = 1; }
            else if (s.af[i] && s.conflicts[g.bfm[i]] == 1) deactivate_clause(g, s, (int)i);
        }
        if (i < g.V) {
            if (s.av[i] && s.conflicts[g.bvm[i]] == 1) deactivate_variable(g, s, (int)i);
        }
        if (i < g.B) {
            if (s.conflicts[i] >= 1) { s.is_sat[i] = 0.f; s.flags[i] |= PDP_FLAG_UP_CONFLICT; s.masked[i] = 1; }
        }
    }
}

__device__ __forceinline__ void up_assign(const KArgs& A) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    const int64_t n = g.V > g.B ? g.V : g.B;
    WARP_STRIDED(i, n) {
        if (i < g.V) {
            const int cnt = s.up_cnt[i];
            if (cnt > 0) {
                const int ev = s.up_ev[i];
                s.up_cnt[i] = 0; s.up_ev[i] = 0;
                if (s.av[i] && abs(ev) == cnt) fix_variable(g, s, (int)i, ev > 0 ? 1.f : -1.f);
            }
        }
    }
}

__device__ __forceinline__ void up_clear_conflicts(const KArgs& A) {
    const pdp_state& s = A.s;
    WARP_STRIDED(i, A.g.B) { if (i < A.g.B) s.conflicts[i] = 0; }
}

// ------------------------------------------------------------------------------------------------
// pure-literal peeling round (solver.py:180-203)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void peel_find(const KArgs& A, int flag_slot) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    WARP_STRIDED(i, g.V) {
        if (i >= g.V) continue;
        if (!s.av[i]) continue;
        const int b = g.bvm[i];
        if (!s.dirty[b]) continue;
        const int beg = g.var_ptr[i], end = g.var_ptr[i + 1];
        int deg = 0, sdeg = 0;
        for (int p = beg; p < end; ++p) {
            if (s.af[g.v_cls[p]]) { ++deg; sdeg += (g.v_cedge[p] & PDP_SIGN_BIT) ? -1 : 1; }
        }
        if (deg == abs(sdeg)) {
            s.pure[i] = 1;
            s.sol[i] = ((sdeg > 0 ? 1.f : (sdeg < 0 ? -1.f : 0.f)) + 1.f) / 2.0f;
            s.ctrl[flag_slot] = 1;
        }
    }
}

__device__ __forceinline__ void peel_apply(const KArgs& A) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    WARP_STRIDED(i, g.V) {
        if (i >= g.V) continue;
        if (!s.pure[i]) continue;
        s.pure[i] = 0;
        const int beg = g.var_ptr[i], end = g.var_ptr[i + 1];
        for (int p = beg; p < end; ++p) {
            const int a = g.v_cls[p];
            if (s.af[a]) deactivate_clause(g, s, a);
        }
        deactivate_variable(g, s, (int)i);
        s.masked[g.bvm[i]] = 1;
    }
}

// UP closure then peel closure for the dirty problems.  Called by ALL threads of the cooperative grid.
// Flag slots alternate by round parity so a flag is never reset while it can still be read.
__device__ __forceinline__ void closure(const KArgs& A, cg::grid_group& grid) {
    const pdp_state& s = A.s;
    int round = 0;
    for (;;) {   // solver.py:234-273
        const int slot = CTRL_FLAG_A + (round & 1);
        const int other = CTRL_FLAG_A + ((round + 1) & 1);
        up_find_units(A, slot);
        if (gtid() == 0) s.ctrl[other] = 0;
        grid.sync();
        if (!s.ctrl[slot]) break;
        up_find_conflicts(A);
        grid.sync();
        up_apply_conflicts(A);
        grid.sync();
        up_assign(A);
        up_clear_conflicts(A);
        grid.sync();
        ++round;
    }
    round = 0;
    for (;;) {   // solver.py:188-203
        const int slot = CTRL_FLAG_C + (round & 1);
        const int other = CTRL_FLAG_C + ((round + 1) & 1);
        peel_find(A, slot);
        if (gtid() == 0) s.ctrl[other] = 0;
        grid.sync();
        if (!s.ctrl[slot]) break;
        peel_apply(A);
        grid.sync();
        ++round;
    }
    // every flag slot is zero again here; UP and peel use disjoint slots so a thread that has left one
    // loop can never disturb a flag another thread is still reading
}

// ------------------------------------------------------------------------------------------------
// full-formula satisfaction count (SatCNFEvaluator on _solution, util.py:210-236) for dirty problems
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool literal_true(float sgn, float p) {
    float ev = sgn * p;
    ev = ev + (1.f - sgn) / 2.f;
    return ev > 0.5f;
}

__device__ __forceinline__ void cnf_count_dirty(const KArgs& A) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    KeyedReducer<CountAcc> red;
    WARP_STRIDED(a, g.F) {
        if (a >= g.F) continue;
        const int b = g.bfm[a];
        if (!s.dirty[b]) continue;
        red.touch(s, b);
        const int beg = g.cl_ptr[a], end = g.cl_ptr[a + 1];
        bool sat = false;
        for (int c = beg; c < end; ++c) {
            const uint32_t w = g.c_var[c];
            if (literal_true((w & PDP_SIGN_BIT) ? -1.f : 1.f, s.sol[w & PDP_IDX_MASK])) { sat = true; break; }
        }
        red.acc.n += sat ? 0 : 1;
    }
    red.finish(s);
}
