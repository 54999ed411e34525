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

// named barriers (ids 1..15; 0 is __syncthreads): producer / consumer hand-over between warp groups
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void group_sync(int bar_id, int n) { if (bar_id == 0) __syncthreads(); else bar_sync(bar_id, n); }

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
    // bar_id == 0: the whole CTA (__syncthreads); otherwise a named barrier of `gthr` threads whose local
    // thread index is `t` (a warp group of the pipelined passes)
    __device__ __forceinline__ void finish(const pdp_state& s, int bar_id = 0, int gthr = 0, int t = -1) {
        __shared__ ACC sm_acc[32];
        __shared__ int sm_key[32];
        if (bar_id == 0) { gthr = blockDim.x; t = threadIdx.x; }
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
        group_sync(bar_id, gthr);
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
        group_sync(bar_id, gthr);
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
// The serial passes apply the sticky-NaN rule themselves at write-out; the pipelined / TMA variants leave the
// problems on the sticky-NaN path to the generic passes.
#define PDP_STICKY_INLINE (!(PDP_PIPELINE || PDP_TMA))
__device__ __forceinline__ bool blk_problem_runs(const pdp_state& s, int b) { return s.active[b] && (PDP_STICKY_INLINE || !s.nanflag[b]); }

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

#ifndef PDP_VN_UNROLL
#define PDP_VN_UNROLL 1    // edge loops of the variable node phase
#endif
#ifndef PDP_UNROLL_WO
#define PDP_UNROLL_WO 6    // elements per thread in flight: write-out (2 loads each)
#endif
#ifndef PDP_UNROLL_CL
#define PDP_UNROLL_CL 8    // clause load (2-3 loads each)
#endif
#ifndef PDP_UNROLL_VL
#define PDP_UNROLL_VL 6    // variable load (3-4 loads each)
#endif
// four-slot groups per thread in flight (PDP_VEC4)
#ifndef PDP_UNROLL_WO4
#define PDP_UNROLL_WO4 4
#endif
#ifndef PDP_UNROLL_CL4
#define PDP_UNROLL_CL4 4
#endif
#ifndef PDP_UNROLL_VL4
#define PDP_UNROLL_VL4 2    // 3 costs a spilled register under the 64-register cap, same speed
#endif

// ================================================================================================
// phase bodies of the blocked passes, written for a GROUP of G threads with local index t: the whole CTA
// in the serial passes, one of the two warp groups in the pipelined passes
// ================================================================================================

// write-out: slots [0, ne) of the block in ascending destination order; consecutive slots mostly hit
// consecutive destinations (runs), so a warp's stores coalesce into a few sectors
// STICKY: 0 = off, 1 = slots flagged in `sticky` bits, 2 = every slot.  A sticky slot keeps a NaN that is already
// stored at its destination in `old` (the reference blends mask*new + (1-mask)*old arithmetically: 0*NaN = NaN,
// pdp_propagate.py:175,218).
// PDP_VEC4 = 1: the memory phases read their contiguous streams four slots per thread and instruction (128-bit loads
// of the fp32 streams, 64-bit loads of the 16-bit tables).  They are latency-bound (measured, clock64 phase timers):
// what counts is bytes in flight per thread.  A block's region starts at an arbitrary element of 256-byte aligned
// arrays, so up to three head and three tail slots go through the scalar path.
#ifndef PDP_VEC4
#define PDP_VEC4 1
#endif
// the write-out keeps one slot per thread and instruction: with four consecutive slots per thread a warp's stores are
// strided by four elements and every destination sector is written four times (measured: +30 % on the phase)
#ifndef PDP_VEC4_WO
#define PDP_VEC4_WO 0
#endif
#ifndef PDP_INPASS_SCORE
#define PDP_INPASS_SCORE 1
#endif
#ifndef PDP_COLD
#define PDP_COLD __forceinline__
#endif
#ifndef PDP_WO_PIPE
#define PDP_WO_PIPE 0
#endif
struct Vec4Range { int head, nvec, tail0; };
template <typename T>
__device__ __forceinline__ Vec4Range vec4_range(const T* p32, int ne) {   // p32: the region's start in a 4-byte array
    Vec4Range R;
    R.head = (int)((16u - ((unsigned)(uintptr_t)p32 & 15u)) & 15u) >> 2;
    if (R.head > ne) R.head = ne;
    R.nvec = (ne - R.head) >> 2;
    R.tail0 = R.head + 4 * R.nvec;
    return R;
}
__device__ __forceinline__ uint32_t mnib(const uint32_t* words, int pos) { return (__ldcg(words + (pos >> 5)) >> (pos & 31)) & 15u; }

template <int G, bool SKIP, int STICKY>
__device__ __forceinline__ void ph_write_out_t(int t, const uint16_t* __restrict__ src, const int32_t* __restrict__ dst, int ne,
                                               const float* plane, const uint32_t* skip, const uint32_t* sticky,
                                               const float* old, float* out) {   // old may alias out (q is updated in place)
    auto one = [&](int l, int d) {
        if (SKIP && ((skip[l >> 5] >> (l & 31)) & 1u)) return;
        float v = plane[l];
        if (STICKY == 2 || (STICKY == 1 && ((sticky[l >> 5] >> (l & 31)) & 1u))) { const float ov = old[d]; if (ov != ov) v = ov; }
        out[d] = v;
    };
#if PDP_VEC4_WO
    const Vec4Range R = vec4_range(dst, ne);
    if (t < R.head) one(src[t], dst[t]);
    if (t < ne - R.tail0) one(src[R.tail0 + t], dst[R.tail0 + t]);
    const uint2* __restrict__ s4 = reinterpret_cast<const uint2*>(src + R.head);
    const int4* __restrict__ d4 = reinterpret_cast<const int4*>(dst + R.head);
    constexpr int U = PDP_UNROLL_WO4;
    int w = t;
    for (; w + (U - 1) * G < R.nvec; w += U * G) {
        uint2 l[U]; int4 d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { l[u] = s4[w + u * G]; d[u] = d4[w + u * G]; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            one((int)(l[u].x & 0xffffu), d[u].x); one((int)(l[u].x >> 16), d[u].y);
            one((int)(l[u].y & 0xffffu), d[u].z); one((int)(l[u].y >> 16), d[u].w);
        }
    }
    for (; w < R.nvec; w += G) {
        const uint2 l = s4[w]; const int4 d = d4[w];
        one((int)(l.x & 0xffffu), d.x); one((int)(l.x >> 16), d.y); one((int)(l.y & 0xffffu), d.z); one((int)(l.y >> 16), d.w);
    }
#else
    int w = t;
    constexpr int U = PDP_UNROLL_WO;
#if PDP_WO_PIPE
    if (ne <= 0) return;
    // register double buffer: the index loads of the next batch are in flight while this batch gathers and stores
    // (the phase waits on exactly these loads: long-scoreboard stalls on the gather, profiles/r1_sp_run_ncu.md)
    int l[U], d[U];
    bool have = w + (U - 1) * G < ne;
    {
        const int w0 = have ? w : 0;      // (no full batch: the loads below are dummies of valid addresses, ne >= 1 here)
#pragma unroll
        for (int u = 0; u < U; ++u) { l[u] = src[have ? w0 + u * G : 0]; d[u] = dst[have ? w0 + u * G : 0]; }
    }
    while (have) {
        const int wn = w + U * G;
        const bool more = wn + (U - 1) * G < ne;
        const int wl = more ? wn : w;      // last batch: re-read the current indices (in range, discarded)
        int ln[U], dn[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { ln[u] = src[wl + u * G]; dn[u] = dst[wl + u * G]; }
#pragma unroll
        for (int u = 0; u < U; ++u) one(l[u], d[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) { l[u] = ln[u]; d[u] = dn[u]; }
        w = wn;
        have = more;
    }
#else
    for (; w + (U - 1) * G < ne; w += U * G) {
        int l[U], d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { l[u] = src[w + u * G]; d[u] = dst[w + u * G]; }
#pragma unroll
        for (int u = 0; u < U; ++u) one(l[u], d[u]);
    }
#endif
    for (; w < ne; w += G) one(src[w], dst[w]);
#endif
}
// flags: bit 0 = some slots are skipped, bit 1 = some slots are sticky, bit 2 = every slot is sticky
template <int G>
__device__ __forceinline__ void ph_write_out(int t, const uint16_t* __restrict__ src, const int32_t* __restrict__ dst, int ne,
                                             const float* plane, const uint32_t* skip, int flags, float* out,
                                             const uint32_t* sticky = nullptr, const float* old = nullptr) {
    if (flags & 4) {
        if (flags & 1) ph_write_out_t<G, true, 2>(t, src, dst, ne, plane, skip, sticky, old, out);
        else ph_write_out_t<G, false, 2>(t, src, dst, ne, plane, skip, sticky, old, out);
    } else if (flags & 2) {
        if (flags & 1) ph_write_out_t<G, true, 1>(t, src, dst, ne, plane, skip, sticky, old, out);
        else ph_write_out_t<G, false, 1>(t, src, dst, ne, plane, skip, sticky, old, out);
    } else {
        if (flags & 1) ph_write_out_t<G, true, 0>(t, src, dst, ne, plane, skip, sticky, old, out);
        else ph_write_out_t<G, false, 0>(t, src, dst, ne, plane, skip, sticky, old, out);
    }
}

// clause pass, load phase: x = log(max(q_u, 1e-40)) * em, scattered into clause-major order.
// e0 = first C-layout position of the block (the mask bits are indexed by position).
template <int G, bool MASKED>
__device__ __forceinline__ void ph_clause_load(int t, const float* __restrict__ qsrc, const uint16_t* __restrict__ inv,
                                               const uint32_t* __restrict__ qmask, int e0, int ne, float* X) {
    auto put = [&](float q, int l, bool m) {
        float v = L40(q);
        if (MASKED && m) v = v * 0.f;
        X[l] = v;
    };
#if PDP_VEC4
    const Vec4Range R = vec4_range(qsrc, ne);
    if (t < R.head) put(qsrc[t], inv[t], MASKED ? mbit(qmask, e0 + t) : false);
    if (t < ne - R.tail0) put(qsrc[R.tail0 + t], inv[R.tail0 + t], MASKED ? mbit(qmask, e0 + R.tail0 + t) : false);
    const float4* __restrict__ q4 = reinterpret_cast<const float4*>(qsrc + R.head);
    const uint2* __restrict__ i4 = reinterpret_cast<const uint2*>(inv + R.head);
    const int pos0 = e0 + R.head;     // a multiple of 4: the four mask bits of a group sit in one word
    constexpr int U = PDP_UNROLL_CL4;
    int x = t;
    for (; x + (U - 1) * G < R.nvec; x += U * G) {
        float4 q[U]; uint2 l[U]; uint32_t m[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { q[u] = q4[x + u * G]; l[u] = i4[x + u * G]; m[u] = MASKED ? mnib(qmask, pos0 + 4 * (x + u * G)) : 0u; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            put(q[u].x, (int)(l[u].x & 0xffffu), m[u] & 1u); put(q[u].y, (int)(l[u].x >> 16), m[u] & 2u);
            put(q[u].z, (int)(l[u].y & 0xffffu), m[u] & 4u); put(q[u].w, (int)(l[u].y >> 16), m[u] & 8u);
        }
    }
    for (; x < R.nvec; x += G) {
        const float4 q = q4[x]; const uint2 l = i4[x]; const uint32_t m = MASKED ? mnib(qmask, pos0 + 4 * x) : 0u;
        put(q.x, (int)(l.x & 0xffffu), m & 1u); put(q.y, (int)(l.x >> 16), m & 2u);
        put(q.z, (int)(l.y & 0xffffu), m & 4u); put(q.w, (int)(l.y >> 16), m & 8u);
    }
#else
    int x = t;
    constexpr int U = PDP_UNROLL_CL;
    for (; x + (U - 1) * G < ne; x += U * G) {
        float q[U]; int l[U]; bool m[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { q[u] = qsrc[x + u * G]; l[u] = inv[x + u * G]; m[u] = MASKED ? mbit(qmask, e0 + x + u * G) : false; }
#pragma unroll
        for (int u = 0; u < U; ++u) put(q[u], l[u], m[u]);
    }
    for (; x < ne; x += G) put(qsrc[x], inv[x], MASKED ? mbit(qmask, e0 + x) : false);
#endif
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

// geometry of one block of a pass
struct BlkGeo {
    int n0, n1;      // node range
    int e0, ne;      // first slot / slots
    int b0, b1;      // problem range
    __device__ __forceinline__ bool multi() const { return b0 != b1; }
};
__device__ __forceinline__ BlkGeo clause_block(const pdp_graph& g, int blk) {
    BlkGeo B;
    B.n0 = g.cb_ptr[blk]; B.n1 = g.cb_ptr[blk + 1];
    if (B.n1 <= B.n0) { B.e0 = 0; B.ne = 0; B.b0 = B.b1 = 0; return B; }
    B.e0 = g.cl_ptr[B.n0]; B.ne = g.cl_ptr[B.n1] - B.e0;
    B.b0 = g.bfm[B.n0]; B.b1 = g.bfm[B.n1 - 1];
    return B;
}
__device__ __forceinline__ BlkGeo var_block(const pdp_graph& g, int blk) {
    BlkGeo B;
    B.n0 = g.vb_ptr[blk]; B.n1 = g.vb_ptr[blk + 1];
    if (B.n1 <= B.n0) { B.e0 = 0; B.ne = 0; B.b0 = B.b1 = 0; return B; }
    B.e0 = g.var_ptr[B.n0]; B.ne = g.var_ptr[B.n1] - B.e0;
    B.b0 = g.bvm[B.n0]; B.b1 = g.bvm[B.n1 - 1];
    return B;
}

// clause pass, node phase: thread per clause
template <int G>
__device__ __forceinline__ void ph_clause_node(int t, const pdp_graph& g, const pdp_state& s, const BlkGeo& B, int ku,
                                               float* X, uint32_t* skip, int* any_skip, uint32_t* sticky = nullptr) {
    const bool multi = B.multi();
    for (int a = B.n0 + t; a < B.n1; a += G) {
        int lo, k;
        if (ku) { k = ku; lo = (a - B.n0) * ku; }
        else { lo = g.cl_ptr[a] - B.e0; k = g.cl_ptr[a + 1] - B.e0 - lo; }
        int b = B.b0;
        if (multi) {
            b = g.bfm[a];
            if (!blk_problem_runs(s, b)) { blk_mark_skip(skip, lo, lo + k); atomicOr(any_skip, 1); continue; }
            if (PDP_STICKY_INLINE && sticky && s.nanflag[b]) { blk_mark_skip(sticky, lo, lo + k); atomicOr(any_skip, 2); }
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
template <int G, bool MASKED>
__device__ __forceinline__ void ph_var_load(int t, const float* __restrict__ sn, const float* __restrict__ so, const uint16_t* __restrict__ inv,
                                            const uint32_t* __restrict__ vmask, int e0, int ne, float* PA, float* PB) {
    auto put = [&](float n, float o, uint32_t iv, bool m) {
        const int l = iv & 0x7fff;
        PA[l] = __uint_as_float(__float_as_uint(n) | ((MASKED && m) ? 0x80000000u : 0u));
        PB[l] = __uint_as_float(__float_as_uint(o) ^ ((iv & PDP_VINV_NEG) << 16));
    };
#if PDP_VEC4
    const Vec4Range R = vec4_range(sn, ne);
    if (t < R.head) put(sn[t], so[t], inv[t], MASKED ? mbit(vmask, e0 + t) : false);
    if (t < ne - R.tail0) put(sn[R.tail0 + t], so[R.tail0 + t], inv[R.tail0 + t], MASKED ? mbit(vmask, e0 + R.tail0 + t) : false);
    const float4* __restrict__ n4 = reinterpret_cast<const float4*>(sn + R.head);
    const float4* __restrict__ o4 = reinterpret_cast<const float4*>(so + R.head);
    const uint2* __restrict__ i4 = reinterpret_cast<const uint2*>(inv + R.head);
    const int pos0 = e0 + R.head;
    constexpr int U = PDP_UNROLL_VL4;
    int x = t;
    for (; x + (U - 1) * G < R.nvec; x += U * G) {
        float4 n[U], o[U]; uint2 l[U]; uint32_t m[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            n[u] = n4[x + u * G]; o[u] = o4[x + u * G]; l[u] = i4[x + u * G];
            m[u] = MASKED ? mnib(vmask, pos0 + 4 * (x + u * G)) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            put(n[u].x, o[u].x, l[u].x & 0xffffu, m[u] & 1u); put(n[u].y, o[u].y, l[u].x >> 16, m[u] & 2u);
            put(n[u].z, o[u].z, l[u].y & 0xffffu, m[u] & 4u); put(n[u].w, o[u].w, l[u].y >> 16, m[u] & 8u);
        }
    }
    for (; x < R.nvec; x += G) {
        const float4 n = n4[x], o = o4[x]; const uint2 l = i4[x]; const uint32_t m = MASKED ? mnib(vmask, pos0 + 4 * x) : 0u;
        put(n.x, o.x, l.x & 0xffffu, m & 1u); put(n.y, o.y, l.x >> 16, m & 2u);
        put(n.z, o.z, l.y & 0xffffu, m & 4u); put(n.w, o.w, l.y >> 16, m & 8u);
    }
#else
    int x = t;
    constexpr int U = PDP_UNROLL_VL;
    for (; x + (U - 1) * G < ne; x += U * G) {
        float n[U], o[U]; uint32_t iv[U]; bool m[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            n[u] = sn[x + u * G]; o[u] = so[x + u * G]; iv[u] = inv[x + u * G];
            m[u] = MASKED ? mbit(vmask, e0 + x + u * G) : false;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) put(n[u], o[u], iv[u], m[u]);
    }
    for (; x < ne; x += G) put(sn[x], so[x], inv[x], MASKED ? mbit(vmask, e0 + x) : false);
#endif
}

// variable pass, SurveyScorer (pdp_predict.py:155-192) of the variables whose problem asked for it (want_score): same
// operations and order as score_variable() on the new surveys held in PA (sign bit = edge masked; for an active variable
// the edge mask is the clause mask the scorer multiplies with, an inactive variable's score is never looked at) and the
// literal signs held in PB.  pi == 0 on the blocked path: the external force does not enter.
template <int G>
__device__ __forceinline__ void ph_var_score(int t, const pdp_graph& g, const pdp_state& s, const BlkGeo& B,
                                             const float* __restrict__ PA, const float* __restrict__ PB) {
    for (int base = B.n0, round = 0; base < B.n1; base += G, ++round) {
        const int ti = (round & 1) ? (base + G - 1 - t) : (base + t);
        if (ti >= B.n1) continue;
        const int2 ve = __ldg(&g.vsort[ti]);
        const int i = ve.x, lo = ve.y & 0xffff, deg = ve.y >> 16;
        if (!s.want_score[B.multi() ? g.bvm[i] : B.b0]) continue;
        float ps = 0.f, ns = 0.f, as = 0.f;
        for (int j = 0; j < deg; ++j) {
            const uint32_t nb = __float_as_uint(PA[lo + j]), ob = __float_as_uint(PB[lo + j]);
            const uint32_t negm = (uint32_t)((int32_t)ob >> 31);
            const float f = L10(1.f - __uint_as_float(nb & 0x7fffffffu)) * ((nb >> 31) ? 0.f : 1.f);
            const float zf = 0.f * f;
            ps += fsel(negm, zf, f);
            ns += fsel(negm, f, zf);
            as += f;
        }
        s.score[i] = sp_score_tail(ps, ns, as, 0.f, 0.f);
    }
}

// variable pass, node phase: thread per variable (descending degree, rounds alternate direction so that
// every thread gets high and low degrees): ordered sums, decimator statistics, update.
// Requires eta(t-1) >= +0 or NaN without sign (the sign bits are borrowed, see ph_var_load).
template <int G>
__device__ __forceinline__ void ph_var_node(int t, const pdp_graph& g, const pdp_state& s, const BlkGeo& B, bool use_mask, bool has_prev,
                                            bool em_set, float* __restrict__ PA, float* __restrict__ PB, uint32_t* skip, int* any_skip,
                                            KeyedReducer<StatAcc>& red, BlkStats& sm_st, bool local_stats, uint32_t* sticky = nullptr) {
    const bool multi = B.multi();
    constexpr int VNU = PDP_VN_UNROLL;
    for (int base = B.n0, round = 0; base < B.n1; base += G, ++round) {
        const int ti = (round & 1) ? (base + G - 1 - t) : (base + t);
        if (ti >= B.n1) continue;
        const int2 ve = __ldg(&g.vsort[ti]);
        const int i = ve.x, lo = ve.y & 0xffff, deg = ve.y >> 16;
        int b = B.b0;
        if (multi) {
            b = g.bvm[i];
            if (!blk_problem_runs(s, b)) { blk_mark_skip(skip, lo, lo + deg); atomicOr(any_skip, 1); continue; }
            if (PDP_STICKY_INLINE && sticky && s.nanflag[b]) { blk_mark_skip(sticky, lo, lo + deg); atomicOr(any_skip, 2); }
        }
        const uint32_t act = s.av[i];
        float P = 0.f, N = 0.f, n0 = 0.f, d0 = 0.f, n1 = 0.f, d1 = 0.f;
#pragma unroll VNU
        for (int j = 0; j < deg; ++j) {
            const uint32_t nb = __float_as_uint(PA[lo + j]), ob = __float_as_uint(PB[lo + j]);
            const bool m = (nb >> 31) != 0u;                        // edge masked
            const uint32_t negm = (uint32_t)((int32_t)ob >> 31);    // all ones: negative literal
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
                const int lb = b - B.b0;
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
#pragma unroll VNU
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
}

__device__ __forceinline__ void stats_slots_reset(BlkStats& sm_st, int t, int nprob) {
    if (t < nprob) {
        sm_st.mx0[t] = 0u; sm_st.mn0[t] = 0x7f800000u; sm_st.mx1[t] = 0u; sm_st.mn1[t] = 0x7f800000u;
        sm_st.nan[t] = 0u; sm_st.nav[t] = 0u;
    }
}
__device__ __forceinline__ void stats_slots_commit(const pdp_state& s, BlkStats& sm_st, int t, int nprob, int b0) {
    if (t < nprob) {
        StatAcc a;
        a.mx0 = sm_st.mx0[t]; a.mn0 = sm_st.mn0[t]; a.mx1 = sm_st.mx1[t]; a.mn1 = sm_st.mn1[t];
        a.nan = sm_st.nan[t]; a.nav = sm_st.nav[t];
        if (a.mn0 != 0x7f800000u || a.mx0 != 0u || a.nan || a.nav || a.mn1 != 0x7f800000u) a.commit(s, b0 + t);
    }
}

// ================================================================================================
// pipelined passes.  The memory phases (load, write-out) are bound by DRAM bandwidth / latency and issue
// few instructions; the node phases are bound by instruction issue and touch no global memory.  Run back
// to back by the same threads they add up, so the CTA is split into two warp groups working on two
// shared-memory slots: the MEMORY group loads block i+1 and writes block i-1 out while the COMPUTE group
// runs the node phase of block i.  Hand-over through named barriers (bar.arrive / bar.sync):
//   FULL[slot]  memory -> compute: the slot holds a loaded block
//   DONE[slot]  compute -> memory: the node phase of the slot's block is finished
// ================================================================================================
// PDP_PHASE_TIMING (profiling builds only): thread 0 of every CTA adds the clock cycles (>> 10) it spent in each
// phase of the blocked passes to the trace buffer: [0..2] clause wait-for-load / node / write-out, [3..5] variable
#ifdef PDP_PHASE_TIMING
#define PHASE_T0() long long _pt = clock64()
#define PHASE_ADD(slot_) do { if (threadIdx.x == 0 && A.trace) { const long long _n = clock64(); atomicAdd(&A.trace[slot_], (int)((_n - _pt) >> 10)); _pt = _n; } } while (0)
#else
#define PHASE_T0() do {} while (0)
#define PHASE_ADD(slot_) do {} while (0)
#endif
#if PDP_PIPELINE || PDP_TMA
#define PIPE_MEM_THREADS (32 * PDP_PIPE_MEM_WARPS)
#define PIPE_CMP_THREADS (PDP_SWEEP_THREADS - PIPE_MEM_THREADS)
#define PIPE_BAR_FULL 1    // +slot
#define PIPE_BAR_DONE 3    // +slot
#define PIPE_BAR_CMP 5     // compute group internal
#define PIPE_SLOT_BYTES (4 * PDP_BLK_C)
#define PIPE_MAX_BLOCKS 32 // blocks of one CTA per pass handled per pipeline run

#ifdef PDP_PHASE_TIMING
#define PT_DECL() long long _pt = clock64()
#define PT_ADD(slot_) do { if (t == 0 && A.trace) { const long long _n = clock64(); atomicAdd(&A.trace[slot_], (int)((_n - _pt) >> 10)); _pt = _n; } } while (0)
#else
#define PT_DECL() do {} while (0)
#define PT_ADD(slot_) do {} while (0)
#endif

struct PipeSmem {
    int any_skip[2];
    int blk[PIPE_MAX_BLOCKS];
    int nblk;
};

// the CTA's non-idle blocks of this pass, PIPE_MAX_BLOCKS at a time (uniform over the CTA)
template <bool VAR>
__device__ __forceinline__ int pipe_collect(const pdp_graph& g, const pdp_state& s, PipeSmem& ps, int& next_blk) {
    const int total = VAR ? g.nvb : g.ncb;
    int n = 0;
    while (next_blk < total && n < PIPE_MAX_BLOCKS) {
        const int blk = next_blk;
        next_blk += gridDim.x;
        const BlkGeo B = VAR ? var_block(g, blk) : clause_block(g, blk);
        if (B.n1 <= B.n0) continue;
        if (blk_idle(s, B.b0, B.b1)) continue;
        if (threadIdx.x == 0) ps.blk[n] = blk;
        ++n;
    }
    __syncthreads();
    return n;
}

__device__ __forceinline__ void pipe_clause_pass(const KArgs& A, int r, bool use_mask, unsigned char* smem) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    __shared__ PipeSmem ps;
    const float* __restrict__ qin = s.qu;
    float* __restrict__ eout = s.eta[r ^ 1];
    const bool is_mem = threadIdx.x < PIPE_MEM_THREADS;
    const int t = is_mem ? threadIdx.x : (threadIdx.x - PIPE_MEM_THREADS);
    int next_blk = blockIdx.x;
    for (;;) {
        const int n = pipe_collect<false>(g, s, ps, next_blk);
        if (n == 0) break;
        PT_DECL();
        if (is_mem) {
            for (int i = 0; i < n + 2; ++i) {
                const int slot = i & 1;
                float* X = reinterpret_cast<float*>(smem + slot * PIPE_SLOT_BYTES);
                uint32_t* skip = reinterpret_cast<uint32_t*>(smem + 2 * PIPE_SLOT_BYTES + slot * (PDP_BLK_C / 8));
                if (i >= 2) {   // the slot's previous block: wait for its node phase, write it out
                    PT_ADD(8);
                    bar_sync(PIPE_BAR_DONE + slot, PDP_SWEEP_THREADS);
                    PT_ADD(10);
                    const BlkGeo B = clause_block(g, ps.blk[i - 2]);
                    ph_write_out<PIPE_MEM_THREADS>(t, g.csrc + B.e0, g.cdst + B.e0, B.ne, X, skip, ps.any_skip[slot] != 0, eout);
                    bar_sync(PIPE_BAR_CMP + 1, PIPE_MEM_THREADS);
                    PT_ADD(9);   // memory group: the slot is free
                }
                if (i < n) {
                    const BlkGeo B = clause_block(g, ps.blk[i]);
                    for (int w = t; w < (B.ne + 31) / 32; w += PIPE_MEM_THREADS) skip[w] = 0u;
                    if (t == 0) ps.any_skip[slot] = 0;
                    if (use_mask && (B.multi() || s.masked[B.b0])) ph_clause_load<PIPE_MEM_THREADS, true>(t, qin + B.e0, g.cinv + B.e0, g.qmask, B.e0, B.ne, X);
                    else ph_clause_load<PIPE_MEM_THREADS, false>(t, qin + B.e0, g.cinv + B.e0, g.qmask, B.e0, B.ne, X);
                    __threadfence_block();
                    bar_arrive(PIPE_BAR_FULL + slot, PDP_SWEEP_THREADS);
                }
            }
        } else {
            for (int i = 0; i < n; ++i) {
                const int slot = i & 1;
                float* X = reinterpret_cast<float*>(smem + slot * PIPE_SLOT_BYTES);
                uint32_t* skip = reinterpret_cast<uint32_t*>(smem + 2 * PIPE_SLOT_BYTES + slot * (PDP_BLK_C / 8));
                PT_ADD(11);
                bar_sync(PIPE_BAR_FULL + slot, PDP_SWEEP_THREADS);
                PT_ADD(12);
                const int blk = ps.blk[i];
                const BlkGeo B = clause_block(g, blk);
                ph_clause_node<PIPE_CMP_THREADS>(t, g, s, B, g.cb_k[blk], X, skip, &ps.any_skip[slot]);
                __threadfence_block();
                bar_arrive(PIPE_BAR_DONE + slot, PDP_SWEEP_THREADS);
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void pipe_var_pass(const KArgs& A, int r, bool use_mask, bool has_prev, bool em_set, unsigned char* smem) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    __shared__ PipeSmem ps;
    __shared__ BlkStats sm_st;
    const float* __restrict__ en = s.eta[r ^ 1];
    const float* __restrict__ eo = s.eta[r];
    const bool is_mem = threadIdx.x < PIPE_MEM_THREADS;
    const int t = is_mem ? threadIdx.x : (threadIdx.x - PIPE_MEM_THREADS);
    KeyedReducer<StatAcc> red;
    int next_blk = blockIdx.x;
    for (;;) {
        const int n = pipe_collect<true>(g, s, ps, next_blk);
        if (n == 0) break;
        PT_DECL();
        if (is_mem) {
            for (int i = 0; i < n + 2; ++i) {
                const int slot = i & 1;
                float* PA = reinterpret_cast<float*>(smem + slot * PIPE_SLOT_BYTES);
                float* PB = PA + PDP_BLK_V;
                uint32_t* skip = reinterpret_cast<uint32_t*>(smem + 2 * PIPE_SLOT_BYTES + slot * (PDP_BLK_C / 8));
                if (i >= 2) {
                    PT_ADD(16);
                    bar_sync(PIPE_BAR_DONE + slot, PDP_SWEEP_THREADS);
                    PT_ADD(18);
                    const BlkGeo B = var_block(g, ps.blk[i - 2]);
                    ph_write_out<PIPE_MEM_THREADS>(t, g.vsrc + B.e0, g.vdst + B.e0, B.ne, PA, skip, ps.any_skip[slot] != 0, s.qu);
                    bar_sync(PIPE_BAR_CMP + 1, PIPE_MEM_THREADS);
                    PT_ADD(17);
                }
                if (i < n) {
                    const BlkGeo B = var_block(g, ps.blk[i]);
                    for (int w = t; w < (B.ne + 31) / 32; w += PIPE_MEM_THREADS) skip[w] = 0u;
                    if (t == 0) ps.any_skip[slot] = 0;
                    if ((use_mask || em_set) && (B.multi() || s.masked[B.b0])) ph_var_load<PIPE_MEM_THREADS, true>(t, en + B.e0, eo + B.e0, g.vinv + B.e0, g.vmask, B.e0, B.ne, PA, PB);
                    else ph_var_load<PIPE_MEM_THREADS, false>(t, en + B.e0, eo + B.e0, g.vinv + B.e0, g.vmask, B.e0, B.ne, PA, PB);
                    __threadfence_block();
                    bar_arrive(PIPE_BAR_FULL + slot, PDP_SWEEP_THREADS);
                }
            }
        } else {
            for (int i = 0; i < n; ++i) {
                const int slot = i & 1;
                float* PA = reinterpret_cast<float*>(smem + slot * PIPE_SLOT_BYTES);
                float* PB = PA + PDP_BLK_V;
                uint32_t* skip = reinterpret_cast<uint32_t*>(smem + 2 * PIPE_SLOT_BYTES + slot * (PDP_BLK_C / 8));
                const BlkGeo B = var_block(g, ps.blk[i]);
                const bool local_stats = B.multi() && (B.b1 - B.b0 < PDP_STAT_SLOTS);
                if (local_stats) { stats_slots_reset(sm_st, t, B.b1 - B.b0 + 1); bar_sync(PIPE_BAR_CMP, PIPE_CMP_THREADS); }
                PT_ADD(19);
                bar_sync(PIPE_BAR_FULL + slot, PDP_SWEEP_THREADS);
                PT_ADD(20);
                ph_var_node<PIPE_CMP_THREADS>(t, g, s, B, use_mask, has_prev, em_set, PA, PB, skip, &ps.any_skip[slot], red, sm_st, local_stats);
                __threadfence_block();
                bar_arrive(PIPE_BAR_DONE + slot, PDP_SWEEP_THREADS);
                if (local_stats) { bar_sync(PIPE_BAR_CMP, PIPE_CMP_THREADS); stats_slots_commit(s, sm_st, t, B.b1 - B.b0 + 1, B.b0); }
                red.finish(s, PIPE_BAR_CMP, PIPE_CMP_THREADS, t);
            }
        }
        __syncthreads();
    }
}

#endif  // PDP_PIPELINE || PDP_TMA

// ================================================================================================
// TMA-staged passes.  Everything a block's node phase reads -- the message regions, the 16-bit permutation
// of the block's slots and the edge-mask words -- is contiguous in global memory, so ONE thread brings it
// into a shared-memory slot with bulk asynchronous copies (cp.async.bulk, completion on an mbarrier) while
// the CTA still works on the previous block in the other slot.  The node phases gather through the
// permutation (slot of the node's j-th edge -> staged position) instead of reading a scattered copy, and
// leave their results at the same staged positions; the write-out reads them through vsrc2 / csrc2.
// ================================================================================================
#if PDP_TMA
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
// 16-byte aligned source / destination, size a multiple of 16
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// byte offsets inside a slot
#define TS_VAR_SA 0
#define TS_VAR_SB ((PDP_BLK_V + 4) * 4)
#define TS_VAR_PERM (2 * (PDP_BLK_V + 4) * 4)
#define TS_VAR_MASK (TS_VAR_PERM + (PDP_BLK_V + 8) * 2)
#define TS_VAR_SKIP (TS_VAR_MASK + (PDP_BLK_V / 32 + 8) * 4)
#define TS_CL_X 0
#define TS_CL_PERM ((PDP_BLK_C + 4) * 4)
#define TS_CL_MASK (TS_CL_PERM + (PDP_BLK_C + 8) * 2)
#define TS_CL_SKIP (TS_CL_MASK + (PDP_BLK_C / 32 + 8) * 4)
static_assert(TS_VAR_SKIP + PDP_BLK_V / 8 + 16 <= PDP_TMA_SLOT_BYTES, "variable slot too large");
static_assert(TS_CL_SKIP + PDP_BLK_C / 8 + 16 <= PDP_TMA_SLOT_BYTES, "clause slot too large");
static_assert(TS_VAR_SB % 16 == 0 && TS_VAR_PERM % 16 == 0 && TS_VAR_MASK % 16 == 0 && TS_CL_PERM % 16 == 0 && TS_CL_MASK % 16 == 0, "TMA alignment");

struct TmaState {          // per-CTA pipeline state living in registers across the whole kernel
    uint32_t parity[2];
};
struct TmaSmem {
    uint64_t bar[2];
    int any_skip[2];
    int blk[PIPE_MAX_BLOCKS];
};

// staged geometry of a block: region [base, base + n_st) with base = e0 & ~3
struct Staged {
    int base, off, n_st;       // off = e0 - base; n_st = staged floats (multiple of 4)
    int pbase, poff, n_perm;   // permutation table: 8-element aligned start, staged 16-bit entries (multiple of 8)
    int wbase, woff, n_w;      // mask words: 4-word aligned start
};
__device__ __forceinline__ Staged staged_of(int e0, int ne) {
    Staged S;
    S.base = e0 & ~3; S.off = e0 - S.base; S.n_st = (S.off + ne + 3) & ~3;
    S.pbase = e0 & ~7; S.poff = e0 - S.pbase; S.n_perm = (S.poff + ne + 7) & ~7;
    const int w0 = S.base >> 5, w1 = (S.base + S.n_st - 1) >> 5;
    S.wbase = w0 & ~3; S.woff = w0 - S.wbase; S.n_w = (w1 - S.wbase + 1 + 3) & ~3;
    return S;
}
// is the edge at staged position x masked?  (mask words staged from word S.wbase)
__device__ __forceinline__ bool staged_mbit(const uint32_t* mw, const Staged& S, int x) {
    const int bitpos = (S.base & 31) + x;
    return (mw[S.woff + (bitpos >> 5)] >> (bitpos & 31)) & 1u;
}

__device__ __forceinline__ void tma_issue_clause(const pdp_graph& g, const pdp_state& s, const BlkGeo& B, bool masked,
                                                 unsigned char* slot, uint64_t* bar) {
    const Staged S = staged_of(B.e0, B.ne);
    const bool any = B.ne > 0;
    const uint32_t bytes = any ? ((uint32_t)S.n_st * 4u + (uint32_t)S.n_perm * 2u + (masked ? (uint32_t)S.n_w * 4u : 0u)) : 0u;
    mbar_expect_tx(bar, bytes);
    if (!any) return;
    tma_load_1d(slot + TS_CL_X, s.qu + S.base, (uint32_t)S.n_st * 4u, bar);
    tma_load_1d(slot + TS_CL_PERM, g.cperm + S.pbase, (uint32_t)S.n_perm * 2u, bar);
    if (masked) tma_load_1d(slot + TS_CL_MASK, g.qmask + S.wbase, (uint32_t)S.n_w * 4u, bar);
}
__device__ __forceinline__ void tma_issue_var(const pdp_graph& g, const pdp_state& s, const BlkGeo& B, bool masked, const float* en,
                                              const float* eo, unsigned char* slot, uint64_t* bar) {
    const Staged S = staged_of(B.e0, B.ne);
    const bool any = B.ne > 0;
    const uint32_t bytes = any ? (2u * (uint32_t)S.n_st * 4u + (uint32_t)S.n_perm * 2u + (masked ? (uint32_t)S.n_w * 4u : 0u)) : 0u;
    mbar_expect_tx(bar, bytes);
    if (!any) return;
    tma_load_1d(slot + TS_VAR_SA, en + S.base, (uint32_t)S.n_st * 4u, bar);
    tma_load_1d(slot + TS_VAR_SB, eo + S.base, (uint32_t)S.n_st * 4u, bar);
    tma_load_1d(slot + TS_VAR_PERM, g.vperm + S.pbase, (uint32_t)S.n_perm * 2u, bar);
    if (masked) tma_load_1d(slot + TS_VAR_MASK, g.vmask + S.wbase, (uint32_t)S.n_w * 4u, bar);
}

// the CTA's non-idle blocks of this pass into sm.blk (uniform over the CTA)
template <bool VAR>
__device__ __forceinline__ int tma_collect(const pdp_graph& g, const pdp_state& s, TmaSmem& sm, int& next_blk) {
    const int total = VAR ? g.nvb : g.ncb;
    int n = 0;
    while (next_blk < total && n < PIPE_MAX_BLOCKS) {
        const int blk = next_blk;
        next_blk += gridDim.x;
        const BlkGeo B = VAR ? var_block(g, blk) : clause_block(g, blk);
        if (B.n1 <= B.n0) continue;
        if (blk_idle(s, B.b0, B.b1)) continue;
        if (threadIdx.x == 0) sm.blk[n] = blk;
        ++n;
    }
    __syncthreads();
    return n;
}

#define NT PDP_SWEEP_THREADS
__device__ __forceinline__ void tma_clause_pass(const KArgs& A, int r, bool use_mask, unsigned char* smem, TmaSmem& sm, TmaState& st) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    float* __restrict__ eout = s.eta[r ^ 1];
    const int tid = threadIdx.x;
    int next_blk = blockIdx.x;
    for (;;) {
        const int n = tma_collect<false>(g, s, sm, next_blk);
        if (n == 0) break;
        if (tid == 0) {
            const BlkGeo B0 = clause_block(g, sm.blk[0]);
            tma_issue_clause(g, s, B0, use_mask && (B0.multi() || s.masked[B0.b0]), smem, &sm.bar[0]);
        }
        for (int i = 0; i < n; ++i) {
            const int slot = i & 1;
            unsigned char* base = smem + slot * PDP_TMA_SLOT_BYTES;
            float* X = reinterpret_cast<float*>(base + TS_CL_X);
            const uint16_t* perm = reinterpret_cast<const uint16_t*>(base + TS_CL_PERM);
            const uint32_t* mw = reinterpret_cast<const uint32_t*>(base + TS_CL_MASK);
            uint32_t* skip = reinterpret_cast<uint32_t*>(base + TS_CL_SKIP);
            const int blk = sm.blk[i];
            const BlkGeo B = clause_block(g, blk);
            const Staged S = staged_of(B.e0, B.ne);
            const bool masked = use_mask && (B.multi() || s.masked[B.b0]);
            if (i + 1 < n && tid == 0) {   // the other slot is free (its block was written out before the last barrier)
                const BlkGeo Bn = clause_block(g, sm.blk[i + 1]);
                tma_issue_clause(g, s, Bn, use_mask && (Bn.multi() || s.masked[Bn.b0]), smem + (slot ^ 1) * PDP_TMA_SLOT_BYTES, &sm.bar[slot ^ 1]);
            }
            for (int w = tid; w < (S.n_st + 31) / 32; w += NT) skip[w] = 0u;
            if (tid == 0) sm.any_skip[slot] = 0;
            PHASE_T0();
            mbar_wait(&sm.bar[slot], st.parity[slot]);
            st.parity[slot] ^= 1u;
            __syncthreads();
            PHASE_ADD(0);
            // ---- thread per clause: x = log(max(q,1e-40)) * em gathered through the permutation
            const bool multi = B.multi();
            const int ku = g.cb_k[blk];
            for (int a = B.n0 + tid; a < B.n1; a += NT) {
                int lo, k;
                if (ku) { k = ku; lo = (a - B.n0) * ku; }
                else { lo = g.cl_ptr[a] - B.e0; k = g.cl_ptr[a + 1] - B.e0 - lo; }
                int b = B.b0;
                if (multi) {
                    b = g.bfm[a];
                    if (!blk_problem_runs(s, b)) {
                        for (int j = 0; j < k; ++j) { const int x = perm[S.poff + lo + j]; atomicOr(&skip[x >> 5], 1u << (x & 31)); }
                        sm.any_skip[slot] = 1;
                        continue;
                    }
                }
                bool made_nan = false;
                if (k <= 8) {
                    float xv[8]; int xs[8];
                    float tot = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < k) {
                            const int x = perm[S.poff + lo + j];
                            float v = L40(X[x]);
                            if (masked && staged_mbit(mw, S, x)) v = v * 0.f;
                            xs[j] = x; xv[j] = v; tot += v;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < k) {
                            const float nv = X30(tot - xv[j]);
                            made_nan |= (nv != nv);
                            X[xs[j]] = nv;
                        }
                    }
                } else {
                    float tot = 0.f;
                    for (int j = 0; j < k; ++j) {
                        const int x = perm[S.poff + lo + j];
                        float v = L40(X[x]);
                        if (masked && staged_mbit(mw, S, x)) v = v * 0.f;
                        X[x] = v; tot += v;
                    }
                    for (int j = 0; j < k; ++j) {
                        const int x = perm[S.poff + lo + j];
                        const float nv = X30(tot - X[x]);
                        made_nan |= (nv != nv);
                        X[x] = nv;
                    }
                }
                if (made_nan) s.nanpend[b] = 1;
            }
            __syncthreads();
            PHASE_ADD(1);
            ph_write_out<NT>(tid, g.csrc2 + B.e0, g.cdst + B.e0, B.ne, X, skip, sm.any_skip[slot] != 0, eout);
            fence_proxy_async();   // generic-proxy accesses of this slot before the next bulk copy into it
            __syncthreads();
            PHASE_ADD(2);
        }
    }
}

__device__ __forceinline__ void tma_var_pass(const KArgs& A, int r, bool use_mask, bool has_prev, bool em_set, unsigned char* smem,
                                             TmaSmem& sm, TmaState& st) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    __shared__ BlkStats sm_st;
    const float* __restrict__ en = s.eta[r ^ 1];
    const float* __restrict__ eo = s.eta[r];
    const int tid = threadIdx.x;
    KeyedReducer<StatAcc> red;
    int next_blk = blockIdx.x;
    for (;;) {
        const int n = tma_collect<true>(g, s, sm, next_blk);
        if (n == 0) break;
        if (tid == 0) {
            const BlkGeo B0 = var_block(g, sm.blk[0]);
            tma_issue_var(g, s, B0, (use_mask || em_set) && (B0.multi() || s.masked[B0.b0]), en, eo, smem, &sm.bar[0]);
        }
        for (int i = 0; i < n; ++i) {
            const int slot = i & 1;
            unsigned char* base = smem + slot * PDP_TMA_SLOT_BYTES;
            float* SA = reinterpret_cast<float*>(base + TS_VAR_SA);   // eta(t), then q(t)
            float* SB = reinterpret_cast<float*>(base + TS_VAR_SB);   // eta(t-1), then y
            const uint16_t* perm = reinterpret_cast<const uint16_t*>(base + TS_VAR_PERM);
            const uint32_t* mw = reinterpret_cast<const uint32_t*>(base + TS_VAR_MASK);
            uint32_t* skip = reinterpret_cast<uint32_t*>(base + TS_VAR_SKIP);
            const BlkGeo B = var_block(g, sm.blk[i]);
            const Staged S = staged_of(B.e0, B.ne);
            const bool masked = (use_mask || em_set) && (B.multi() || s.masked[B.b0]);
            if (i + 1 < n && tid == 0) {
                const BlkGeo Bn = var_block(g, sm.blk[i + 1]);
                tma_issue_var(g, s, Bn, (use_mask || em_set) && (Bn.multi() || s.masked[Bn.b0]), en, eo,
                              smem + (slot ^ 1) * PDP_TMA_SLOT_BYTES, &sm.bar[slot ^ 1]);
            }
            const bool multi = B.multi();
            const bool local_stats = multi && (B.b1 - B.b0 < PDP_STAT_SLOTS);
            for (int w = tid; w < (S.n_st + 31) / 32; w += NT) skip[w] = 0u;
            if (tid == 0) sm.any_skip[slot] = 0;
            if (local_stats) stats_slots_reset(sm_st, tid, B.b1 - B.b0 + 1);
            PHASE_T0();
            mbar_wait(&sm.bar[slot], st.parity[slot]);
            st.parity[slot] ^= 1u;
            __syncthreads();
            PHASE_ADD(3);
            // ---- thread per variable (descending degree, alternating round direction)
            for (int vbase = B.n0, round = 0; vbase < B.n1; vbase += NT, ++round) {
                const int ti = (round & 1) ? (vbase + NT - 1 - tid) : (vbase + tid);
                if (ti >= B.n1) continue;
                const int2 ve = __ldg(&g.vsort[ti]);
                const int i_var = ve.x, lo = ve.y & 0xffff, deg = ve.y >> 16;
                const uint16_t* pj = perm + S.poff + lo;
                int b = B.b0;
                if (multi) {
                    b = g.bvm[i_var];
                    if (!blk_problem_runs(s, b)) {
                        for (int j = 0; j < deg; ++j) { const int x = pj[j] & 0x7fff; atomicOr(&skip[x >> 5], 1u << (x & 31)); }
                        sm.any_skip[slot] = 1;
                        continue;
                    }
                }
                const uint32_t act = s.av[i_var];
                float P = 0.f, N = 0.f, n0 = 0.f, d0 = 0.f, n1 = 0.f, d1 = 0.f;
                for (int j = 0; j < deg; ++j) {
                    const uint32_t pw = pj[j];
                    const int x = pw & 0x7fff;
                    const uint32_t negm = 0u - (pw >> 15);     // all ones: negative literal
                    const bool m = masked && staged_mbit(mw, S, x);
                    const float xn = SA[x], xo = SB[x];
                    float y = L40(1.f - xo);
                    if (use_mask && m) y = y * 0.f;
                    SB[x] = y;
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
                        const int lb = b - B.b0;
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
                    const uint32_t pw = pj[j];
                    const int x = pw & 0x7fff;
                    const uint32_t negm = 0u - (pw >> 15);
                    const float u = sp_var_finish(fsel(negm, sb_neg, sb_pos), fsel(negm, opp_neg, opp_pos), fsel(negm, O_neg, O_pos), SB[x]);
                    made_nan |= (u != u);
                    SA[x] = u;
                }
                if (made_nan) s.nanpend[b] = 1;
            }
            __syncthreads();
            PHASE_ADD(4);
            if (local_stats) stats_slots_commit(s, sm_st, tid, B.b1 - B.b0 + 1, B.b0);
            ph_write_out<NT>(tid, g.vsrc2 + B.e0, g.vdst + B.e0, B.ne, SA, skip, sm.any_skip[slot] != 0, s.qu);
            fence_proxy_async();
            red.finish(s);   // block-level merge of the statistics; its barriers also close the slot
            PHASE_ADD(5);
        }
    }
}
#undef NT
#endif  // PDP_TMA

// L2 prefetch of [ptr, ptr + bytes) by one bulk instruction (16-byte granularity).  The node phases are issue bound and
// leave the memory system idle: thread 0 starts them by pulling in what the next block of this CTA will load and the
// write-out tables of the current block.
__device__ __forceinline__ void l2_prefetch(const void* ptr, size_t bytes) {
#if PDP_L2_PREFETCH
    if (bytes == 0) return;
    const uintptr_t a = reinterpret_cast<uintptr_t>(ptr) & ~(uintptr_t)15;
    const uint32_t n = (uint32_t)((reinterpret_cast<uintptr_t>(ptr) + bytes - a + 15) & ~(size_t)15);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(n) : "memory");
#endif
}

// (Dynamic block scheduling through a global counter was measured and dropped: with the 2-7 equal-size blocks a CTA
// gets per pass it evens out nothing, and smaller blocks cost more per edge than they balance.)

// ================================================================================================
// serial passes: the whole CTA runs load, node phase and write-out of a block back to back
// ================================================================================================

// Block hand-out of one pass.  Static (block b to CTA b mod grid) when one CTA owns the SM: the CTAs then finish within
// 1-2 % of each other.  With two CTAs per SM the pair drifts apart, the early one waits at the grid barrier and its
// partner finishes alone at half the SM's warps (10 % of the iteration, measured), so the blocks after a CTA's first one
// come from a counter; the next index is fetched while the current block is processed.
#ifndef PDP_DYN_BLOCKS
#define PDP_DYN_BLOCKS 1
#endif
// thread 0, at the top of a block: the index of the block after `blk`, left in slot[par] for feed_advance
template <bool DYN>
__device__ __forceinline__ int feed_fetch(int* ctr, int* slot, int par, int blk) {
    const int nx = DYN ? (int)gridDim.x + atomicAdd(ctr, 1) : blk + (int)gridDim.x;
    slot[par] = nx;
    return nx;
}
// all threads, after the block's last use of shared memory
__device__ __forceinline__ int feed_advance(const int* slot, int& par) {
    __syncthreads();
    const int nx = slot[par];
    par ^= 1;
    return nx;
}

// clause pass of iteration t: eta(t) [buffer r^1, V-layout] from q(t-1) [C-layout]
template <int CTAS>
__device__ __forceinline__ void blk_clause_pass(const KArgs& A, int r, bool use_mask, unsigned char* smem) {
    constexpr int NT = SweepCfg<CTAS>::kThreads, BLK_C = SweepCfg<CTAS>::kBlkC;
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    float* X = reinterpret_cast<float*>(smem);
    uint32_t* skip = reinterpret_cast<uint32_t*>(smem + 4 * BLK_C);
    uint32_t* sticky = skip + BLK_C / 32;
    __shared__ int sm_any_skip;
    const float* __restrict__ qin = s.qu;
    float* __restrict__ eout = s.eta[r ^ 1];
    const int tid = threadIdx.x;
    __shared__ int sm_feed[2];
    __shared__ int sm_nb;
    constexpr bool DYN = PDP_DYN_BLOCKS && CTAS == 2;
    int par = 0;
    for (int blk = blockIdx.x; blk < g.ncb; blk = feed_advance(sm_feed, par)) {
        if (tid == 0) sm_nb = feed_fetch<DYN>(&s.ctrl[CTRL_NEXT_CBLK], sm_feed, par, blk);
        const BlkGeo B = clause_block(g, blk);
        if (B.n1 <= B.n0) continue;
        if (blk_idle(s, B.b0, B.b1)) continue;
        for (int i = tid; i < (B.ne + 31) / 32; i += NT) { skip[i] = 0u; sticky[i] = 0u; }
        if (tid == 0) sm_any_skip = (!B.multi() && s.nanflag[B.b0]) ? 4 : 0;
        PHASE_T0();
        if (use_mask && (B.multi() || s.masked[B.b0])) ph_clause_load<NT, true>(tid, qin + B.e0, g.cinv + B.e0, g.qmask, B.e0, B.ne, X);
        else ph_clause_load<NT, false>(tid, qin + B.e0, g.cinv + B.e0, g.qmask, B.e0, B.ne, X);
        __syncthreads();
        PHASE_ADD(0);
        if (tid == 0) {
            l2_prefetch(g.csrc + B.e0, (size_t)B.ne * 2);
            l2_prefetch(g.cdst + B.e0, (size_t)B.ne * 4);
            const int nb = sm_nb;
            if (nb < g.ncb) {
                const BlkGeo Bn = clause_block(g, nb);
                l2_prefetch(qin + Bn.e0, (size_t)Bn.ne * 4);
                l2_prefetch(g.cinv + Bn.e0, (size_t)Bn.ne * 2);
            }
        }
        ph_clause_node<NT>(tid, g, s, B, g.cb_k[blk], X, skip, &sm_any_skip, sticky);
        __syncthreads();
        PHASE_ADD(1);
        ph_write_out<NT>(tid, g.csrc + B.e0, g.cdst + B.e0, B.ne, X, skip, sm_any_skip, eout, sticky, s.eta[r]);
        __syncthreads();
        PHASE_ADD(2);
    }
}

// variable pass of iteration t: the decimator statistics of eta(t) [buffer r^1] against eta(t-1)
// [buffer r], and q(t) [C-layout, in place] from eta(t-1)
template <int CTAS>
__device__ __forceinline__ void blk_var_pass(const KArgs& A, int r, bool use_mask, bool has_prev, bool em_set, unsigned char* smem) {
    constexpr int NT = SweepCfg<CTAS>::kThreads, BLK_V = SweepCfg<CTAS>::kBlkV, BLK_C = SweepCfg<CTAS>::kBlkC;
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    float* PA = reinterpret_cast<float*>(smem);   // eta(t), then q(t)
    float* PB = PA + BLK_V;                       // eta(t-1), then y
    uint32_t* skip = reinterpret_cast<uint32_t*>(smem + 4 * BLK_C);
    uint32_t* sticky = skip + BLK_C / 32;
    __shared__ int sm_any_skip;
    __shared__ BlkStats sm_st;
    const float* __restrict__ en = s.eta[r ^ 1];
    const float* __restrict__ eo = s.eta[r];
    const int tid = threadIdx.x;
    KeyedReducer<StatAcc> red;
    __shared__ int sm_feed[2];
    __shared__ int sm_nb;
    constexpr bool DYN = PDP_DYN_BLOCKS && CTAS == 2;
    int par = 0;
    for (int blk = blockIdx.x; blk < g.nvb; blk = feed_advance(sm_feed, par)) {
        if (tid == 0) sm_nb = feed_fetch<DYN>(&s.ctrl[CTRL_NEXT_VBLK], sm_feed, par, blk);
        const BlkGeo B = var_block(g, blk);
        if (B.n1 <= B.n0) continue;
        if (blk_idle(s, B.b0, B.b1)) continue;
        const bool local_stats = B.multi() && (B.b1 - B.b0 < PDP_STAT_SLOTS);   // else: registers (one problem) or global atomics
        for (int i = tid; i < (B.ne + 31) / 32; i += NT) { skip[i] = 0u; sticky[i] = 0u; }
        if (tid == 0) sm_any_skip = (!B.multi() && s.nanflag[B.b0]) ? 4 : 0;
        if (local_stats) stats_slots_reset(sm_st, tid, B.b1 - B.b0 + 1);
        PHASE_T0();
        if ((use_mask || em_set) && (B.multi() || s.masked[B.b0])) ph_var_load<NT, true>(tid, en + B.e0, eo + B.e0, g.vinv + B.e0, g.vmask, B.e0, B.ne, PA, PB);
        else ph_var_load<NT, false>(tid, en + B.e0, eo + B.e0, g.vinv + B.e0, g.vmask, B.e0, B.ne, PA, PB);
        __syncthreads();
        PHASE_ADD(3);
        if (tid == 0) {
            l2_prefetch(g.vsrc + B.e0, (size_t)B.ne * 2);
            l2_prefetch(g.vdst + B.e0, (size_t)B.ne * 4);
            const int nb = sm_nb;
            if (nb < g.nvb) {
                const BlkGeo Bn = var_block(g, nb);
                l2_prefetch(en + Bn.e0, (size_t)Bn.ne * 4);
                l2_prefetch(eo + Bn.e0, (size_t)Bn.ne * 4);
                l2_prefetch(g.vinv + Bn.e0, (size_t)Bn.ne * 2);
            }
        }
        // SurveyScorer of problems about to converge, while the new surveys are still in the planes (the node phase
        // overwrites them); its own loop, so that the node phase's code is the same with and without it
        if (PDP_INPASS_SCORE && (s.want_score[B.b0] || s.want_score[B.b1])) ph_var_score<NT>(tid, g, s, B, PA, PB);
        ph_var_node<NT>(tid, g, s, B, use_mask, has_prev, em_set, PA, PB, skip, &sm_any_skip, red, sm_st, local_stats, sticky);
        __syncthreads();
        PHASE_ADD(4);
        if (local_stats) stats_slots_commit(s, sm_st, tid, B.b1 - B.b0 + 1, B.b0);
        ph_write_out<NT>(tid, g.vsrc + B.e0, g.vdst + B.e0, B.ne, PA, skip, sm_any_skip, s.qu, sticky, s.qu);
        red.finish(s);   // block-level merge of the statistics; its barriers also fence the planes
        PHASE_ADD(5);
    }
}

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
