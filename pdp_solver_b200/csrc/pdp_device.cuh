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
    int32_t stagger_c, stagger_v;   // cycles the second CTA of an SM waits at the start of a clause / variable pass
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int64_t gwarp() { return ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; }
__device__ __forceinline__ int64_t gwarps() { return ((int64_t)gridDim.x * blockDim.x) >> 5; }
__device__ __forceinline__ int64_t gtid() { return (int64_t)blockIdx.x * blockDim.x + threadIdx.x; }
__device__ __forceinline__ int64_t gthreads() { return (int64_t)gridDim.x * blockDim.x; }

#define WARP_STRIDED(i, N) \
    for (int64_t i##_base = gwarp() * 32, i = i##_base + lane_id(); i##_base < (N); i##_base += gwarps() * 32, i = i##_base + lane_id())

// named barriers (ids 1..15; 0 is __syncthreads): producer / consumer hand-over between warp groups
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

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
        const int gthr = blockDim.x, t = threadIdx.x;
        const unsigned m = __match_any_sync(0xffffffffu, key);
        const bool uniform = (m == 0xffffffffu);
        const int warp = t >> 5, nwarp = (gthr + 31) >> 5;
        if (uniform) {
            acc.merge_full();
            if (lane_id() == 0) { sm_acc[warp] = acc; sm_key[warp] = key; }
        } else {
            acc.merge_group(m);
            if (lane_id() == (__ffs(m) - 1) && key >= 0) acc.commit(s, key);
            if (lane_id() == 0) sm_key[warp] = -1;
        }
        __syncthreads();
        if (t == 0) {
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
    for (int c = g.cl_ptr[a]; c < g.cl_ptr[a + 1]; ++c) mask_edge(g, cvpos(g, c), cqpos(g, c));
}
// the words change between passes: read around L1
__device__ __forceinline__ bool mbit(const uint32_t* words, int pos) { return (__ldcg(words + (pos >> 5)) >> (pos & 31)) & 1u; }

// variable side of the SP update with the per-variable part hoisted (pi == 0): for one sign s the terms
// 0.5(1+s)P + 0.5(1-s)N, opp and exp(opp) do not depend on the edge.  Same operations, same order as
// sp_var_update_qu.
__device__ __forceinline__ void sp_var_prepare(float P, float N, float s, float& same_base, float& opp, float& O) {
    same_base = 0.5f * (1.f + s) * P + 0.5f * (1.f - s) * N;
    opp = 0.5f * (1.f - s) * P + 0.5f * (1.f + s) * N;   // (the reference's `+ log(1 - 0)` = +0 only turns a -0 into +0: exp is blind to it)
    O = X30(opp);
}
__device__ __forceinline__ float sp_var_finish(float same_base, float opp, float O, float y) {
    const float same = same_base - y;
    const float S = X30(same);
    const float dc = X30(same + opp);
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
            const int qp = cqpos(g, c);
            float v = L40(qin[qp]);
            if (um && mbit(g.qmask, qp)) v = v * 0.f;
            tot += v;
        }
        for (int c = beg; c < end; ++c) {
            const int qp = cqpos(g, c);
            float v = L40(qin[qp]);
            if (um && mbit(g.qmask, qp)) v = v * 0.f;
            const int pos = cvpos(g, c);
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
            float y = L40_1m(ein[vp]);
            if (um && mbit(g.vmask, vp)) y = y * 0.f;
            // the reference's pos/neg incidence matrices hold explicit zeros: 0*y keeps NaN alive
            P += (neg ? 0.f : 1.f) * y;
            N += (neg ? 1.f : 0.f) * y;
        }
        for (int p = beg; p < end; ++p) {
            const int vp = g.p_vpos[p];
            float y = L40_1m(ein[vp]);
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
            const float c = X30S(v);
            n0 += v * c; d0 += c;
            if (has_prev) {
                float d = fabsf(eo[pos] - v);
                if (um && mbit(g.vmask, pos)) d = d * 0.f;
                const float cd = X30S(d);
                n1 += d * cd; d1 += cd;
            }
        }
        const float sm0 = pdp_divs(n0, tmaxf(d0, 1.0f)) * (float)act;
        const float sm1 = pdp_divs(n1, tmaxf(d1, 1.0f)) * (float)act;
        red.acc.add(sm0, sm1, has_prev, act);
    }
    red.finish(s);
}

#include "pdp_sweep.cuh"

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

// A batch of a few large problems: the grid walks the node range of each flagged problem (no per-node problem
// look-up, no dependent loads, one block-level merge per problem) instead of scanning every node of the batch.
#define PDP_RANGE_SCAN_MAX_B 64
__device__ __forceinline__ bool range_scans(const pdp_graph& g) { return g.contiguous_problems && g.B <= PDP_RANGE_SCAN_MAX_B; }

__device__ __forceinline__ void score_phase(const KArgs& A, int w, float pi) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    KeyedReducer<CoefAcc> red;
    if (range_scans(g)) {
        for (int b = 0; b < (int)g.B; ++b) {
            if (!s.conv[b]) continue;
            const int v1 = g.prob_vptr[b + 1];
            const bool have = s.have_score[b] != 0;
            red.touch(s, b);
            if (have) {
#pragma unroll 4
                for (int i = g.prob_vptr[b] + (int)gtid(); i < v1; i += (int)gthreads())
                    red.acc.add(fabsf(s.score[i]) * (float)s.av[i]);
            } else {
                for (int i = g.prob_vptr[b] + (int)gtid(); i < v1; i += (int)gthreads()) {
                    const float sc = score_variable(g, s, s.eta[w], i, pi);
                    s.score[i] = sc;
                    red.acc.add(fabsf(sc) * (float)s.av[i]);
                }
            }
            red.finish(s);
        }
        return;
    }
    WARP_STRIDED(i, g.V) {
        if (i >= g.V) continue;
        const int b = g.bvm[i];
        if (!s.conv[b]) continue;
        red.touch(s, b);
        float sc;
        if (s.have_score[b]) sc = s.score[i];     // written by this iteration's variable pass (ph_var_score)
        else { sc = score_variable(g, s, s.eta[w], (int)i, pi); s.score[i] = sc; }
        red.acc.add(fabsf(sc) * (float)s.av[i]);
    }
    red.finish(s);
}

// first index attaining max of fl(fl(c - min) + 1)  (util.py:257-265)
__device__ __forceinline__ void argmax_phase(const KArgs& A) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    if (range_scans(g)) {
        for (int b = 0; b < (int)g.B; ++b) {
            if (!s.conv[b] || s.c_nan[b]) continue;
            const float m = u2f(s.c_min[b]);
            const float kmax = argmax_key(u2f(s.c_max[b]), m);
            const int v1 = g.prob_vptr[b + 1];
#pragma unroll 4
            for (int i = g.prob_vptr[b] + (int)gtid(); i < v1; i += (int)gthreads()) {
                const float c = fabsf(s.score[i]) * (float)s.av[i];
                if (argmax_key(c, m) == kmax) atomicMin(&s.arg_idx[b], i);
            }
        }
        return;
    }
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
__device__ PDP_COLD void closure(const KArgs& A, cg::grid_group& grid) {
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
// Frontier closure (large problems, grid-wide).  Once every problem is closed under unit propagation and peeling
// (after simplify()), fixing a variable can only create unit clauses among the clauses that hold a variable
// de-activated in the previous round, and pure literals among the variables of clauses de-activated since: the
// rounds below are the rounds of closure() -- same phases, same synchronous semantics, same helper arithmetic --
// run over those lists instead of over all F clauses and V variables (8.9 -> ms per decimation iteration at
// 8 x n = 1 M, where a round's full scan gathers 100 M node flags to find a dozen nodes).
//   clause list (ping-pong by UP round)   : clauses to test for unit-ness
//   variable list (ping-pong by peel round): variables to test for purity; filled during the whole UP stage
//   unit list (by UP round parity)         : variables some unit clause of the round points at
// List entries are unique per epoch (stamps).  A list that overflows, or a problem wiped by the single-conflict
// quirk, sends the rest of the closure through the full scans.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fr_push(const pdp_state& s, int list, int32_t* stamp, int id, int epoch) {
    if (atomicExch(&stamp[id], epoch) == epoch) return;
    const int k = atomicAdd(&s.ctrl[CTRL_FR_N + list], 1);
    if (k < s.fr_cap) s.fr_list[list][k] = id; else s.ctrl[CTRL_FR_OVER] = 1;
}
__device__ __forceinline__ int fr_len(const pdp_state& s, int slot) {
    const int n = s.ctrl[slot];
    return n < s.fr_cap ? n : s.fr_cap;
}
// de-activate clause a and queue its still-active variables for the purity test
__device__ __forceinline__ void fr_deactivate_clause(const pdp_graph& g, const pdp_state& s, int a, int vlist, int epv) {
    deactivate_clause(g, s, a);
    for (int c = g.cl_ptr[a]; c < g.cl_ptr[a + 1]; ++c) {
        const int j = (int)(g.c_var[c] & PDP_IDX_MASK);
        if (s.av[j]) fr_push(s, vlist, s.stamp_v, j, epv);
    }
}
// fix_variable that also queues: clauses made true -> their variables (purity); the other clauses of i -> unit test
__device__ __forceinline__ void fr_fix_variable(const pdp_graph& g, const pdp_state& s, int i, float sg, int clist, int epc,
                                                int vlist, int epv) {
    const int beg = g.var_ptr[i], end = g.var_ptr[i + 1];
    for (int p = beg; p < end; ++p) {
        const float lit = (g.v_cedge[p] & PDP_SIGN_BIT) ? -1.f : 1.f;
        const int a = g.v_cls[p];
        if (!s.af[a]) continue;
        if (lit * sg > 0.f) fr_deactivate_clause(g, s, a, vlist, epv);
        else fr_push(s, clist, s.stamp_c, a, epc);
    }
    deactivate_variable(g, s, i);
    s.sol[i] = (sg + 1.f) / 2.0f;
}

// fix_variable / fr_fix_variable by the 32 lanes of a warp (lanes over the variable's edges): the decimation step of a
// large problem fixes ONE variable, and a single thread walking its ~40 dependent global updates was 0.4 ms
__device__ __forceinline__ void warp_fix_variable(const pdp_graph& g, const pdp_state& s, int i, float sg, bool frontier, int epc, int epv) {
    const int beg = g.var_ptr[i], end = g.var_ptr[i + 1];
    for (int p = beg + lane_id(); p < end; p += 32) {
        const float lit = (g.v_cedge[p] & PDP_SIGN_BIT) ? -1.f : 1.f;
        const int a = g.v_cls[p];
        if (s.af[a]) {
            if (lit * sg > 0.f) { if (frontier) fr_deactivate_clause(g, s, a, 2, epv); else deactivate_clause(g, s, a); }
            else if (frontier) fr_push(s, 0, s.stamp_c, a, epc);
        }
        mask_edge(g, g.p_vpos[p], g.p_qpos[p]);
    }
    if (lane_id() == 0) { s.av[i] = 0; s.sol[i] = (sg + 1.f) / 2.0f; }
}

__device__ __forceinline__ void fr_find_units(const KArgs& A, int clist, int ulist, int flag_slot) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    const int n = fr_len(s, CTRL_FR_N + clist);
    WARP_STRIDED(x, n) {
        if (x >= n) continue;
        const int a = s.fr_list[clist][x];
        if (!s.af[a]) continue;
        const int beg = g.cl_ptr[a], end = g.cl_ptr[a + 1];
        int deg = 0; uint32_t hit = 0;
        for (int c = beg; c < end; ++c) {
            const uint32_t w = g.c_var[c];
            if (s.av[w & PDP_IDX_MASK]) { ++deg; hit = w; }
        }
        if (deg == 1) {
            s.single[a] = 1;
            const int j = (int)(hit & PDP_IDX_MASK);
            if (atomicAdd(&s.up_cnt[j], 1) == 0) {
                const int k = atomicAdd(&s.ctrl[CTRL_FR_NU + ulist], 1);
                if (k < s.fr_cap) s.fr_unit[ulist][k] = j; else s.ctrl[CTRL_FR_OVER] = 1;
            }
            atomicAdd(&s.up_ev[j], (hit & PDP_SIGN_BIT) ? -1 : 1);
            s.ctrl[flag_slot] = 1;
        }
    }
}
__device__ __forceinline__ void fr_find_conflicts(const KArgs& A, int ulist) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    const int n = fr_len(s, CTRL_FR_NU + ulist);
    WARP_STRIDED(x, n) {
        if (x >= n) continue;
        const int i = s.fr_unit[ulist][x];
        const int cnt = s.up_cnt[i];
        if (cnt > 0 && abs(s.up_ev[i]) != cnt) {
            if (atomicAdd(&s.conflicts[g.bvm[i]], 1) == 0) s.ctrl[CTRL_FR_WIPE] = 1;   // one conflict may mean a wipe
        }
    }
}
// up_apply_conflicts over the lists; the wipe of single-conflict problems scans their nodes (rare)
__device__ __forceinline__ void fr_apply_conflicts(const KArgs& A, int clist, int vlist, int epv) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    const int n = fr_len(s, CTRL_FR_N + clist);
    WARP_STRIDED(x, n) {
        if (x >= n) continue;
        const int a = s.fr_list[clist][x];
        if (s.single[a]) { fr_deactivate_clause(g, s, a, vlist, epv); s.single[a] = 0; s.masked[g.bfm[a]] = 1; }
    }
    if (s.ctrl[CTRL_FR_WIPE]) {
        const int64_t m = g.V > g.F ? g.V : g.F;
        WARP_STRIDED(i, m) {
            if (i < g.F && s.af[i] && !s.single[i] && s.conflicts[g.bfm[i]] == 1) deactivate_clause(g, s, (int)i);
            if (i < g.V && s.av[i] && s.conflicts[g.bvm[i]] == 1) deactivate_variable(g, s, (int)i);
        }
    }
    WARP_STRIDED(b, g.B) {
        if (b < g.B && s.conflicts[b] >= 1) { s.is_sat[b] = 0.f; s.flags[b] |= PDP_FLAG_UP_CONFLICT; s.masked[b] = 1; }
    }
}
__device__ __forceinline__ void fr_assign(const KArgs& A, int ulist, int clist_next, int epc, int vlist, int epv) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    const int n = fr_len(s, CTRL_FR_NU + ulist);
    WARP_STRIDED(x, n) {
        if (x >= n) continue;
        const int i = s.fr_unit[ulist][x];
        const int cnt = s.up_cnt[i];
        if (cnt > 0) {
            const int ev = s.up_ev[i];
            s.up_cnt[i] = 0; s.up_ev[i] = 0;
            if (s.av[i] && abs(ev) == cnt) fr_fix_variable(g, s, i, ev > 0 ? 1.f : -1.f, clist_next, epc, vlist, epv);
        }
    }
}
__device__ __forceinline__ void fr_peel_find(const KArgs& A, int vlist, int flag_slot) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    const int n = fr_len(s, CTRL_FR_N + vlist);
    WARP_STRIDED(x, n) {
        if (x >= n) continue;
        const int i = s.fr_list[vlist][x];
        if (!s.av[i]) continue;
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
__device__ __forceinline__ void fr_peel_apply(const KArgs& A, int vlist, int vlist_next, int epv) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    const int n = fr_len(s, CTRL_FR_N + vlist);
    WARP_STRIDED(x, n) {
        if (x >= n) continue;
        const int i = s.fr_list[vlist][x];
        if (!s.pure[i]) continue;
        s.pure[i] = 0;
        const int beg = g.var_ptr[i], end = g.var_ptr[i + 1];
        for (int p = beg; p < end; ++p) {
            const int a = g.v_cls[p];
            if (s.af[a]) fr_deactivate_clause(g, s, a, vlist_next, epv);
        }
        deactivate_variable(g, s, i);
        s.masked[g.bvm[i]] = 1;
    }
}

// Called by ALL threads of the cooperative grid after select_and_fix_phase<frontier> and a grid barrier: clause list 0
// and variable list 2 hold what the fixes touched (epochs epc0 / epv0, both already stored in the control block).
__device__ PDP_COLD void closure_frontier(const KArgs& A, cg::grid_group& grid) {
    const pdp_state& s = A.s;
    int epc = s.ctrl[CTRL_FR_EPC], epv = s.ctrl[CTRL_FR_EPV];
    int round = 0, cur = 0;
    bool fallback = false;
    for (;;) {   // unit propagation (solver.py:234-273)
        const int slot = CTRL_FLAG_A + (round & 1), other = CTRL_FLAG_A + ((round + 1) & 1);
        const int ul = round & 1;
        if (s.ctrl[CTRL_FR_OVER]) { fallback = true; break; }        // uniform: written before the last barrier
        fr_find_units(A, cur, ul, slot);
        if (gtid() == 0) { s.ctrl[other] = 0; s.ctrl[CTRL_FR_N + (cur ^ 1)] = 0; s.ctrl[CTRL_FR_NU + (ul ^ 1)] = 0; }
        grid.sync();
        if (!s.ctrl[slot]) break;
        fr_find_conflicts(A, ul);
        grid.sync();
        fr_apply_conflicts(A, cur, 2, epv);
        grid.sync();
        ++epc;
        fr_assign(A, ul, cur ^ 1, epc, 2, epv);
        up_clear_conflicts(A);
        if (gtid() == 0) s.ctrl[CTRL_FR_WIPE] = 0;
        grid.sync();
        cur ^= 1;
        ++round;
    }
    if (!fallback && s.ctrl[CTRL_FR_OVER]) fallback = true;
    if (!fallback) {
        int vc = 2;
        round = 0;
        for (;;) {   // peeling (solver.py:188-203)
            const int slot = CTRL_FLAG_C + (round & 1), other = CTRL_FLAG_C + ((round + 1) & 1);
            if (s.ctrl[CTRL_FR_OVER]) { fallback = true; break; }
            fr_peel_find(A, vc, slot);
            if (gtid() == 0) { s.ctrl[other] = 0; s.ctrl[CTRL_FR_N + (vc ^ 1)] = 0; }
            grid.sync();
            if (!s.ctrl[slot]) break;
            ++epv;
            fr_peel_apply(A, vc, vc ^ 1, epv);
            grid.sync();
            vc ^= 1;
            ++round;
        }
    }
    grid.sync();
    if (gtid() == 0) {
        s.ctrl[CTRL_FR_EPC] = epc + 1; s.ctrl[CTRL_FR_EPV] = epv + 1;
        for (int i = 0; i < 4; ++i) s.ctrl[CTRL_FR_N + i] = 0;
        s.ctrl[CTRL_FR_NU] = 0; s.ctrl[CTRL_FR_NU + 1] = 0; s.ctrl[CTRL_FR_OVER] = 0; s.ctrl[CTRL_FR_WIPE] = 0;
    }
    if (fallback) {
        // per-round scratch is clean at a round boundary; the full scans finish the closure from the current masks
        if (gtid() == 0) { s.ctrl[CTRL_FLAG_A] = 0; s.ctrl[CTRL_FLAG_A + 1] = 0; s.ctrl[CTRL_FLAG_C] = 0; s.ctrl[CTRL_FLAG_C + 1] = 0; }
        grid.sync();
        closure(A, grid);
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-local decimation.  Everything the decimator does to a converged problem -- scoring, arg-max, fixing
// the variable, unit propagation and pure-literal peeling to closure, the CNF check and the termination
// decision -- touches that problem only, so for problems that one CTA can walk (the common case: thousands of
// n ~ 100 problems per batch) a CTA does all of it with block barriers, where the grid-wide phases above need a
// few dozen grid barriers per iteration.  Same arithmetic and the same synchronous rounds as the grid-wide
// phases (which stay for large problems and batch replication).
// ------------------------------------------------------------------------------------------------
struct LocSmem {
    uint32_t cmax, cmin, cnan;
    int arg, flag, conflicts, nunsat, fixed;
};

__device__ __forceinline__ bool literal_true(float sgn, float p);

#define LSYNC() bar_sync(bar_id, nthr)
__device__ __forceinline__ void loc_closure(const KArgs& A, int b, int v0, int v1, int f0, int f1, LocSmem& ls, int tid, int nthr, int bar_id) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    for (;;) {   // unit propagation rounds, solver.py:234-273
        if (tid == 0) { ls.flag = 0; ls.conflicts = 0; }
        LSYNC();
        for (int a = f0 + tid; a < f1; a += nthr) {
            if (!s.af[a]) continue;
            int deg = 0; uint32_t hit = 0;
            for (int c = g.cl_ptr[a]; c < g.cl_ptr[a + 1]; ++c) {
                const uint32_t w = g.c_var[c];
                if (s.av[w & PDP_IDX_MASK]) { ++deg; hit = w; }
            }
            if (deg == 1) {
                s.single[a] = 1;
                const int j = (int)(hit & PDP_IDX_MASK);
                atomicAdd(&s.up_cnt[j], 1);
                atomicAdd(&s.up_ev[j], (hit & PDP_SIGN_BIT) ? -1 : 1);
                ls.flag = 1;
            }
        }
        LSYNC();
        if (!ls.flag) break;
        for (int i = v0 + tid; i < v1; i += nthr) {
            const int cnt = s.up_cnt[i];
            if (cnt > 0 && abs(s.up_ev[i]) != cnt) atomicAdd(&ls.conflicts, 1);
        }
        LSYNC();
        const int nc = ls.conflicts;   // the `== 1` quirk of solver.py:257,261
        for (int a = f0 + tid; a < f1; a += nthr) {
            if (s.single[a]) { deactivate_clause(g, s, a); s.single[a] = 0; s.masked[b] = 1; }
            else if (s.af[a] && nc == 1) deactivate_clause(g, s, a);
        }
        for (int i = v0 + tid; i < v1; i += nthr)
            if (s.av[i] && nc == 1) deactivate_variable(g, s, i);
        if (tid == 0 && nc >= 1) { s.is_sat[b] = 0.f; s.flags[b] |= PDP_FLAG_UP_CONFLICT; s.masked[b] = 1; }
        LSYNC();
        for (int i = v0 + tid; i < v1; i += nthr) {
            const int cnt = s.up_cnt[i];
            if (cnt > 0) {
                const int ev = s.up_ev[i];
                s.up_cnt[i] = 0; s.up_ev[i] = 0;
                if (s.av[i] && abs(ev) == cnt) fix_variable(g, s, i, ev > 0 ? 1.f : -1.f);
            }
        }
        LSYNC();
    }
    for (;;) {   // pure-literal peeling rounds, solver.py:188-203
        LSYNC();
        if (tid == 0) ls.flag = 0;
        LSYNC();
        for (int i = v0 + tid; i < v1; i += nthr) {
            if (!s.av[i]) continue;
            int deg = 0, sdeg = 0;
            for (int p = g.var_ptr[i]; p < g.var_ptr[i + 1]; ++p)
                if (s.af[g.v_cls[p]]) { ++deg; sdeg += (g.v_cedge[p] & PDP_SIGN_BIT) ? -1 : 1; }
            if (deg == abs(sdeg)) {
                s.pure[i] = 1;
                s.sol[i] = ((sdeg > 0 ? 1.f : (sdeg < 0 ? -1.f : 0.f)) + 1.f) / 2.0f;
                ls.flag = 1;
            }
        }
        LSYNC();
        if (!ls.flag) break;
        for (int i = v0 + tid; i < v1; i += nthr) {
            if (!s.pure[i]) continue;
            s.pure[i] = 0;
            for (int p = g.var_ptr[i]; p < g.var_ptr[i + 1]; ++p) {
                const int a = g.v_cls[p];
                if (s.af[a]) deactivate_clause(g, s, a);
            }
            deactivate_variable(g, s, i);
            s.masked[b] = 1;
        }
    }
}

// one converged problem: pdp_decimate.py:152-171 + solver.py:275-285 + trainer.py:150-162
__device__ __forceinline__ void loc_decimate_problem(const KArgs& A, int b, int iter, int w, float pi, bool check_termination, LocSmem& ls,
                                                     int tid, int nthr, int bar_id) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    const int v0 = g.prob_vptr[b], v1 = g.prob_vptr[b + 1], f0 = g.prob_fptr[b], f1 = g.prob_fptr[b + 1];
    if (tid == 0) { ls.cmax = 0u; ls.cmin = 0x7f800000u; ls.cnan = 0u; ls.arg = 0x7fffffff; ls.fixed = 0; ls.nunsat = 0; }
    LSYNC();
    for (int i = v0 + tid; i < v1; i += nthr) {
        const float sc = score_variable(g, s, s.eta[w], i, pi);
        s.score[i] = sc;
        const float c = fabsf(sc) * (float)s.av[i];
        if (c != c) ls.cnan = 1u; else { atomicMax(&ls.cmax, f2u(c)); atomicMin(&ls.cmin, f2u(c)); }
    }
    LSYNC();
    if (!ls.cnan) {   // first index attaining max of fl(fl(c - min) + 1), util.py:257-265
        const float m = u2f(ls.cmin);
        const float kmax = argmax_key(u2f(ls.cmax), m);
        for (int i = v0 + tid; i < v1; i += nthr) {
            const float c = fabsf(s.score[i]) * (float)s.av[i];
            if (argmax_key(c, m) == kmax) atomicMin(&ls.arg, i);
        }
    }
    LSYNC();
    if (tid == 0 && !ls.cnan && u2f(ls.cmax) > 0.f && ls.arg != 0x7fffffff) {
        const int i = ls.arg;
        const float sg = sgnf(s.score[i]);
        if (sg != 0.f && s.av[i]) {
            fix_variable(g, s, i, sg);
            s.masked[b] = 1; s.dirty[b] = 1;
            ls.fixed = 1;
            if (A.trace) {
                const int k = atomicAdd(&s.ctrl[CTRL_TRACE_LEN], 1);
                if (k < A.trace_cap) { A.trace[3 * k] = iter; A.trace[3 * k + 1] = i; A.trace[3 * k + 2] = (int)sg; }
            }
        }
    }
    LSYNC();
    if (ls.fixed) loc_closure(A, b, v0, v1, f0, f1, ls, tid, nthr, bar_id);
    LSYNC();
    if (s.dirty[b]) {
        if (check_termination) {   // SatCNFEvaluator on _solution over the full formula, then trainer.py:150-162
            int n = 0;
            for (int a = f0 + tid; a < f1; a += nthr) {
                bool sat = false;
                for (int c = g.cl_ptr[a]; c < g.cl_ptr[a + 1]; ++c) {
                    const uint32_t wv = g.c_var[c];
                    if (literal_true((wv & PDP_SIGN_BIT) ? -1.f : 1.f, s.sol[wv & PDP_IDX_MASK])) { sat = true; break; }
                }
                n += sat ? 0 : 1;
            }
            n = __reduce_add_sync(0xffffffffu, n);
            if ((tid & 31) == 0 && n) atomicAdd(&ls.nunsat, n);
            LSYNC();
            if (tid == 0) {
                if (ls.nunsat == 0) {
                    if (s.active[b]) { s.active[b] = 0; s.freeze_iter[b] = iter; atomicSub(&s.ctrl[CTRL_NUM_ACTIVE], 1); }
                    s.flags[b] |= PDP_FLAG_SOLVED;
                }
                s.dirty[b] = 0; s.n_unsat[b] = 0;
            }
        } else if (tid == 0) {
            s.ctrl[CTRL_ANY_DIRTY] = 1;
        }
    }
    if (tid == 0) s.conv[b] = 0;
    LSYNC();
}

// ------------------------------------------------------------------------------------------------
// shared-memory tier of the CTA-local decimation.  A problem whose masks, adjacency (16-bit local indices) and
// scratch fit the group's share of the (idle) sweep buffer is decimated entirely in shared memory: the closure
// is a chain of dozens of dependent rounds, each a couple of microseconds through global memory and tens of
// nanoseconds here.  Same rounds, same arithmetic as loc_closure; what changed is written back at the end.
// ------------------------------------------------------------------------------------------------
struct TinyView {
    uint8_t *av, *af, *pure, *single;
    float *sol, *score;
    int *cnt, *ev;
    uint16_t *clp, *cvar, *vp, *vcls;   // local CSR / CSC: pointers, (local index | sign << 15)
};
__device__ __forceinline__ size_t tiny_bytes(int n, int m, int E) {
    return (size_t)n * (1 + 1 + 4 + 4 + 4 + 4) + (size_t)m * 2 + (size_t)(m + 1 + n + 1 + 2 * E) * 2 + 64;
}
__device__ __forceinline__ TinyView tiny_carve(unsigned char* base, int n, int m, int E) {
    TinyView t;
    float* f = reinterpret_cast<float*>(base);
    t.sol = f; t.score = f + n;
    t.cnt = reinterpret_cast<int*>(f + 2 * n); t.ev = t.cnt + n;
    uint16_t* h = reinterpret_cast<uint16_t*>(t.ev + n);
    t.clp = h; t.vp = t.clp + (m + 1); t.cvar = t.vp + (n + 1); t.vcls = t.cvar + E;
    uint8_t* q = reinterpret_cast<uint8_t*>(t.vcls + E);
    t.av = q; t.pure = q + n; t.af = q + 2 * n; t.single = t.af + m;
    return t;
}

__device__ __forceinline__ void tiny_fix(const TinyView& t, int i, float sg) {
    for (int p = t.vp[i]; p < t.vp[i + 1]; ++p) {
        const uint32_t w = t.vcls[p];
        const float lit = (w & 0x8000u) ? -1.f : 1.f;
        if (lit * sg > 0.f) t.af[w & 0x7fffu] = 0;
    }
    t.av[i] = 0;
    t.sol[i] = (sg + 1.f) / 2.0f;
}

#define LSYNC() bar_sync(bar_id, nthr)
__device__ __forceinline__ void loc_decimate_tiny(const KArgs& A, int b, int iter, int w, float pi, bool check_termination, LocSmem& ls,
                                                  int tid, int nthr, int bar_id, unsigned char* area, int share) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    const int v0 = g.prob_vptr[b], v1 = g.prob_vptr[b + 1], f0 = g.prob_fptr[b], f1 = g.prob_fptr[b + 1];
    const int n = v1 - v0, m = f1 - f0;
    const int ec0 = g.cl_ptr[f0], ep0 = g.var_ptr[v0], E = g.cl_ptr[f1] - ec0;
    const TinyView t = tiny_carve(area, n, m, E);
    // the surveys of the problem in variable-major order, when they fit too (pi == 0: the external force column drops out)
    const size_t eta_off = (tiny_bytes(n, m, E) + 15) & ~(size_t)15;
    const bool stage_eta = (pi == 0.f) && (eta_off + (size_t)E * 4 <= (size_t)share);
    float* eta_s = reinterpret_cast<float*>(area + eta_off);
    if (stage_eta) {
        const float* __restrict__ eta = s.eta[w];
        for (int e = tid; e < E; e += nthr) eta_s[e] = eta[g.p_vpos[ep0 + e]];
    }
    // ---- stage
    for (int i = tid; i < n; i += nthr) {
        t.av[i] = s.av[v0 + i]; t.pure[i] = 0; t.sol[i] = s.sol[v0 + i]; t.cnt[i] = 0; t.ev[i] = 0;
    }
    for (int i = tid; i <= n; i += nthr) t.vp[i] = (uint16_t)(g.var_ptr[v0 + i] - ep0);
    for (int a = tid; a < m; a += nthr) { t.af[a] = s.af[f0 + a]; t.single[a] = 0; }
    for (int a = tid; a <= m; a += nthr) t.clp[a] = (uint16_t)(g.cl_ptr[f0 + a] - ec0);
    for (int e = tid; e < E; e += nthr) {
        const uint32_t cw = g.c_var[ec0 + e];
        t.cvar[e] = (uint16_t)(((cw & PDP_IDX_MASK) - v0) | ((cw & PDP_SIGN_BIT) ? 0x8000u : 0u));
        t.vcls[e] = (uint16_t)((g.v_cls[ep0 + e] - f0) | ((g.v_cedge[ep0 + e] & PDP_SIGN_BIT) ? 0x8000u : 0u));
    }
    if (tid == 0) { ls.cmax = 0u; ls.cmin = 0x7f800000u; ls.cnan = 0u; ls.arg = 0x7fffffff; ls.fixed = 0; ls.nunsat = 0; }
    LSYNC();
    // ---- score, arg-max, fix (pdp_decimate.py:152-171)
    for (int i = tid; i < n; i += nthr) {
        float sc;
        if (stage_eta) {   // score_variable on the staged copies (same order, same operations)
            float ps = 0.f, ns = 0.f, as = 0.f;
            for (int p = t.vp[i]; p < t.vp[i + 1]; ++p) {
                const uint32_t cw = t.vcls[p];
                const float f = L10(1.f - eta_s[p]) * (float)t.af[cw & 0x7fffu];
                const bool neg = (cw & 0x8000u) != 0u;
                ps += (neg ? 0.f : 1.f) * f;
                ns += (neg ? 1.f : 0.f) * f;
                as += f;
            }
            sc = sp_score_tail(ps, ns, as, 0.f, 0.f);
        } else {
            sc = score_variable(g, s, s.eta[w], v0 + i, pi);
        }
        t.score[i] = sc;
        const float c = fabsf(sc) * (float)t.av[i];
        if (c != c) ls.cnan = 1u; else { atomicMax(&ls.cmax, f2u(c)); atomicMin(&ls.cmin, f2u(c)); }
    }
    LSYNC();
    if (!ls.cnan) {
        const float mn = u2f(ls.cmin);
        const float kmax = argmax_key(u2f(ls.cmax), mn);
        for (int i = tid; i < n; i += nthr) {
            const float c = fabsf(t.score[i]) * (float)t.av[i];
            if (argmax_key(c, mn) == kmax) atomicMin(&ls.arg, i);
        }
    }
    LSYNC();
    if (tid == 0 && !ls.cnan && u2f(ls.cmax) > 0.f && ls.arg != 0x7fffffff) {
        const int i = ls.arg;
        const float sg = sgnf(t.score[i]);
        if (sg != 0.f && t.av[i]) {
            tiny_fix(t, i, sg);
            s.masked[b] = 1; s.dirty[b] = 1;
            ls.fixed = 1;
            if (A.trace) {
                const int k = atomicAdd(&s.ctrl[CTRL_TRACE_LEN], 1);
                if (k < A.trace_cap) { A.trace[3 * k] = iter; A.trace[3 * k + 1] = v0 + i; A.trace[3 * k + 2] = (int)sg; }
            }
        }
    }
    LSYNC();
    if (ls.fixed) {
        for (;;) {   // unit propagation rounds, solver.py:234-273
            if (tid == 0) { ls.flag = 0; ls.conflicts = 0; }
            LSYNC();
            for (int a = tid; a < m; a += nthr) {
                if (!t.af[a]) continue;
                int deg = 0; uint32_t hit = 0;
                for (int c = t.clp[a]; c < t.clp[a + 1]; ++c) {
                    const uint32_t cw = t.cvar[c];
                    if (t.av[cw & 0x7fffu]) { ++deg; hit = cw; }
                }
                if (deg == 1) {
                    t.single[a] = 1;
                    const int j = (int)(hit & 0x7fffu);
                    atomicAdd(&t.cnt[j], 1);
                    atomicAdd(&t.ev[j], (hit & 0x8000u) ? -1 : 1);
                    ls.flag = 1;
                }
            }
            LSYNC();
            if (!ls.flag) break;
            for (int i = tid; i < n; i += nthr) {
                const int cnt = t.cnt[i];
                if (cnt > 0 && abs(t.ev[i]) != cnt) atomicAdd(&ls.conflicts, 1);
            }
            LSYNC();
            const int nc = ls.conflicts;   // the `== 1` quirk of solver.py:257,261
            for (int a = tid; a < m; a += nthr) {
                if (t.single[a]) { t.af[a] = 0; t.single[a] = 0; }
                else if (t.af[a] && nc == 1) t.af[a] = 0;
            }
            for (int i = tid; i < n; i += nthr)
                if (t.av[i] && nc == 1) t.av[i] = 0;
            if (tid == 0) { s.masked[b] = 1; if (nc >= 1) { s.is_sat[b] = 0.f; s.flags[b] |= PDP_FLAG_UP_CONFLICT; } }
            LSYNC();
            for (int i = tid; i < n; i += nthr) {
                const int cnt = t.cnt[i];
                if (cnt > 0) {
                    const int ev = t.ev[i];
                    t.cnt[i] = 0; t.ev[i] = 0;
                    if (t.av[i] && abs(ev) == cnt) tiny_fix(t, i, ev > 0 ? 1.f : -1.f);
                }
            }
            LSYNC();
        }
        for (;;) {   // pure-literal peeling rounds, solver.py:188-203
            LSYNC();
            if (tid == 0) ls.flag = 0;
            LSYNC();
            for (int i = tid; i < n; i += nthr) {
                if (!t.av[i]) continue;
                int deg = 0, sdeg = 0;
                for (int p = t.vp[i]; p < t.vp[i + 1]; ++p) {
                    const uint32_t cw = t.vcls[p];
                    if (t.af[cw & 0x7fffu]) { ++deg; sdeg += (cw & 0x8000u) ? -1 : 1; }
                }
                if (deg == abs(sdeg)) {
                    t.pure[i] = 1;
                    t.sol[i] = ((sdeg > 0 ? 1.f : (sdeg < 0 ? -1.f : 0.f)) + 1.f) / 2.0f;
                    ls.flag = 1;
                }
            }
            LSYNC();
            if (!ls.flag) break;
            for (int i = tid; i < n; i += nthr) {
                if (!t.pure[i]) continue;
                t.pure[i] = 0;
                for (int p = t.vp[i]; p < t.vp[i + 1]; ++p) t.af[t.vcls[p] & 0x7fffu] = 0;
                t.av[i] = 0;
            }
            if (tid == 0) s.masked[b] = 1;
        }
        // ---- write back what changed: masks (+ edge-mask bits), solutions
        for (int i = tid; i < n; i += nthr) {
            if (!t.av[i] && s.av[v0 + i]) deactivate_variable(g, s, v0 + i);
            s.sol[v0 + i] = t.sol[i];
        }
        for (int a = tid; a < m; a += nthr)
            if (!t.af[a] && s.af[f0 + a]) deactivate_clause(g, s, f0 + a);
    }
    LSYNC();
    if (s.dirty[b]) {
        if (check_termination) {   // SatCNFEvaluator on _solution over the full formula, then trainer.py:150-162
            int nun = 0;
            for (int a = tid; a < m; a += nthr) {
                bool sat = false;
                for (int c = t.clp[a]; c < t.clp[a + 1]; ++c) {
                    const uint32_t cw = t.cvar[c];
                    if (literal_true((cw & 0x8000u) ? -1.f : 1.f, t.sol[cw & 0x7fffu])) { sat = true; break; }
                }
                nun += sat ? 0 : 1;
            }
            nun = __reduce_add_sync(0xffffffffu, nun);
            if ((tid & 31) == 0 && nun) atomicAdd(&ls.nunsat, nun);
            LSYNC();
            if (tid == 0) {
                if (ls.nunsat == 0) {
                    if (s.active[b]) { s.active[b] = 0; s.freeze_iter[b] = iter; atomicSub(&s.ctrl[CTRL_NUM_ACTIVE], 1); }
                    s.flags[b] |= PDP_FLAG_SOLVED;
                }
                s.dirty[b] = 0; s.n_unsat[b] = 0;
            }
        } else if (tid == 0) {
            s.ctrl[CTRL_ANY_DIRTY] = 1;
        }
    }
    if (tid == 0) s.conv[b] = 0;
    LSYNC();
}
#undef LSYNC

__device__ __forceinline__ bool loc_problem_is_small(const pdp_graph& g, int b) {
    return (g.prob_vptr[b + 1] - g.prob_vptr[b] <= PDP_LOCAL_MAX_V) && (g.prob_fptr[b + 1] - g.prob_fptr[b] <= PDP_LOCAL_MAX_F);
}

#undef LSYNC
// groups of PDP_LOCAL_GROUP threads (named barriers 8..15) take one problem each
#define PDP_LOCAL_GROUP 128
// `area` / `area_bytes`: the CTA's dynamic shared memory (the sweep buffer, idle during the decimation), split evenly
// between the groups; null when the kernel has none
__device__ __forceinline__ void loc_decimate_all(const KArgs& A, int iter, int w, float pi, bool check_termination,
                                                 unsigned char* area, int area_bytes) {
    __shared__ LocSmem ls[8];
    const pdp_state& s = A.s;
    int ngroups = blockDim.x / PDP_LOCAL_GROUP;
    if (ngroups > 8) ngroups = 8;
    const int gthr = blockDim.x / ngroups;           // threads per group (a multiple of 32)
    const int grp = threadIdx.x / gthr, gt = threadIdx.x % gthr;
    // the groups drain the queue decide_phase filled (the order is irrelevant: problems do not interact)
    const int count = s.ctrl[CTRL_LOC_COUNT];
    for (;;) {
        if (gt == 0) ls[grp].flag = atomicAdd(&s.ctrl[CTRL_LOC_NEXT], 1);
        bar_sync(8 + grp, gthr);
        const int k = ls[grp].flag;
        bar_sync(8 + grp, gthr);
        if (k >= count) break;
        const int b = s.loc_list[k];
        const int share = area ? ((area_bytes / ngroups) & ~15) : 0;
        const int n = A.g.prob_vptr[b + 1] - A.g.prob_vptr[b], m = A.g.prob_fptr[b + 1] - A.g.prob_fptr[b];
        const int E = A.g.cl_ptr[A.g.prob_fptr[b + 1]] - A.g.cl_ptr[A.g.prob_fptr[b]];
        if (n < 32768 && m < 32768 && E < 65536 && tiny_bytes(n, m, E) <= (size_t)share)
            loc_decimate_tiny(A, b, iter, w, pi, check_termination, ls[grp], gt, gthr, 8 + grp, area + (size_t)grp * share, share);
        else
            loc_decimate_problem(A, b, iter, w, pi, check_termination, ls[grp], gt, gthr, 8 + grp);
    }
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
    const bool native = s.ctrl[CTRL_NATIVE] != 0;
    if (range_scans(g)) {
        for (int b = 0; b < (int)g.B; ++b) {
            if (!s.dirty[b]) continue;
            const int f1 = g.prob_fptr[b + 1];
            red.touch(s, b);
            for (int a = g.prob_fptr[b] + (int)gtid(); a < f1; a += (int)gthreads()) {
                if (native && s.af[a]) { red.acc.n += 1; continue; }
                bool sat = false;
                for (int c = g.cl_ptr[a]; c < g.cl_ptr[a + 1]; ++c) {
                    const uint32_t w = g.c_var[c];
                    if (literal_true((w & PDP_SIGN_BIT) ? -1.f : 1.f, s.sol[w & PDP_IDX_MASK])) { sat = true; break; }
                }
                red.acc.n += sat ? 0 : 1;
            }
            red.finish(s);
        }
        return;
    }
    WARP_STRIDED(a, g.F) {
        if (a >= g.F) continue;
        const int b = g.bfm[a];
        if (!s.dirty[b]) continue;
        red.touch(s, b);
        if (native && s.af[a]) { red.acc.n += 1; continue; }   // an active clause holds no true literal (CTRL_NATIVE)
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
