// pdp_device.cuh -- grid-wide phases of the p-d-p loop, shared by the persistent cooperative kernel
// (pdp_loop.cu) and by the step-wise entry points.  Every phase is a __device__ function executed
// by ALL threads of a cooperative grid; the caller separates phases with grid.sync().
//
// Work distribution: "warp-strided" loops -- warp w of the grid handles nodes [32w, 32w+32), then
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
// SP sweep, clause side: eta'(e) = exp(min(sum_{e' in a(e)} x_e' - x_e, 30)), x = log(max(q_u,1e-40)) * em
// (pdp_propagate.py:166-175).  Thread per clause; k <= 8 keeps x in registers.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sweep_clause_side(const KArgs& A, int r, bool use_mask_global) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    const float* __restrict__ qin = s.qu[r];
    float* __restrict__ eout = s.eta[r ^ 1];
    WARP_STRIDED(a, g.F) {
        if (a >= g.F) continue;
        const int b = g.bfm[a];
        if (!s.active[b]) continue;
        const bool um = use_mask_global && s.masked[b];
        // the reference blends `mask*new + (1-mask)*old` arithmetically, so a NaN message is sticky
        // (0*NaN); problems that have produced a NaN take the path that re-reads the old value
        const bool sticky = s.nanflag[b] != 0;
        const float* __restrict__ eold = s.eta[r];
        bool made_nan = false;
        const int beg = g.cl_ptr[a], end = g.cl_ptr[a + 1];
        const int k = end - beg;
        const float afa = um ? (float)s.af[a] : 1.f;
        if (k <= 8) {
            float x[8]; int pos[8];
            float tot = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < k) {
                    const int c = beg + j;
                    const int p = g.c_pos[c];
                    float v = L40(qin[p]);
                    if (um) v = v * ((float)s.av[g.c_var[c] & PDP_IDX_MASK] * afa);
                    x[j] = v; pos[j] = p;
                    tot += v;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < k) {
                    float nv = X30(tot - x[j]);
                    if (sticky) { const float ov = eold[pos[j]]; if (ov != ov) nv = ov; }
                    made_nan |= (nv != nv);
                    eout[pos[j]] = nv;
                }
        } else {
            float tot = 0.f;
            for (int c = beg; c < end; ++c) {
                float v = L40(qin[g.c_pos[c]]);
                if (um) v = v * ((float)s.av[g.c_var[c] & PDP_IDX_MASK] * afa);
                tot += v;
            }
            for (int c = beg; c < end; ++c) {
                const int p = g.c_pos[c];
                float v = L40(qin[p]);
                if (um) v = v * ((float)s.av[g.c_var[c] & PDP_IDX_MASK] * afa);
                float nv = X30(tot - v);
                if (sticky) { const float ov = eold[p]; if (ov != ov) nv = ov; }
                made_nan |= (nv != nv);
                eout[p] = nv;
            }
        }
        if (made_nan && !sticky) s.nanflag[b] = 1;
    }
}

// ------------------------------------------------------------------------------------------------
// SP sweep, variable side (pdp_propagate.py:184-218).  Thread per variable.
// ------------------------------------------------------------------------------------------------
template <bool FULL>
__device__ __forceinline__ void sweep_var_side(const KArgs& A, int r, bool use_mask_global, float pi) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    const float* __restrict__ ein = s.eta[r];
    float* __restrict__ qout = s.qu[r ^ 1];
    WARP_STRIDED(i, g.V) {
        if (i >= g.V) continue;
        const int b = g.bvm[i];
        if (!s.active[b]) continue;
        const bool um = use_mask_global && s.masked[b];
        const int beg = g.var_ptr[i], end = g.var_ptr[i + 1];
        const float avi = um ? (float)s.av[i] : 1.f;
        const bool sticky = s.nanflag[b] != 0;
        bool made_nan = false;
        float P = 0.f, N = 0.f;
        for (int p = beg; p < end; ++p) {
            float y = L40(1.f - ein[p]);
            if (um) y = y * (avi * (float)s.af[g.v_cls[p]]);
            const bool neg = (g.v_cedge[p] & PDP_SIGN_BIT) != 0u;
            // the reference's pos/neg incidence matrices hold explicit zeros: 0*y keeps NaN alive
            P += (neg ? 0.f : 1.f) * y;
            N += (neg ? 1.f : 0.f) * y;
        }
        for (int p = beg; p < end; ++p) {
            float y = L40(1.f - ein[p]);
            if (um) y = y * (avi * (float)s.af[g.v_cls[p]]);
            const float sg = (g.v_cedge[p] & PDP_SIGN_BIT) ? -1.f : 1.f;
            if (FULL || pi != 0.f) {
                float u, v, d;
                sp_var_update(P, N, y, sg, s.ext[p], pi, u, v, d);
                if (sticky) {
                    const float ou = s.qu[r][p]; if (ou != ou) u = ou;
                    if (FULL) { const float ov = s.qs[r][p], od = s.qd[r][p]; if (ov != ov) v = ov; if (od != od) d = od; }
                }
                made_nan |= (u != u);
                qout[p] = u;
                if (FULL) { s.qs[r ^ 1][p] = v; s.qd[r ^ 1][p] = d; }
            } else {
                float u = sp_var_update_qu(P, N, y, sg);
                if (sticky) { const float ou = s.qu[r][p]; if (ou != ou) u = ou; }
                made_nan |= (u != u);
                qout[p] = u;
            }
        }
        if (made_nan && !sticky) s.nanflag[b] = 1;
    }
}

// ------------------------------------------------------------------------------------------------
// decimator statistics (pdp_decimate.py:127-143, util.py:282-286): per variable smooth-max of the new
// surveys and of |eta_prev - eta_new| * edge_mask, times active_variables; per problem max / min.
// w = buffer holding the new surveys.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stats_phase(const KArgs& A, int w, bool has_prev, bool em_set) {
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    const float* __restrict__ en = s.eta[w];
    const float* __restrict__ eo = s.eta[w ^ 1];
    KeyedReducer<StatAcc> red;
    WARP_STRIDED(i, g.V) {
        if (i >= g.V) continue;
        const int b = g.bvm[i];
        if (!s.active[b]) continue;
        red.touch(s, b);
        const int beg = g.var_ptr[i], end = g.var_ptr[i + 1];
        const uint32_t act = s.av[i];
        const bool um = em_set && s.masked[b];
        float n0 = 0.f, d0 = 0.f, n1 = 0.f, d1 = 0.f;
        for (int p = beg; p < end; ++p) {
            const float v = en[p];
            const float c = X30(30.f * v);
            n0 += v * c; d0 += c;
            if (has_prev) {
                float d = fabsf(eo[p] - v);
                if (um) d = d * ((float)act * (float)s.af[g.v_cls[p]]);
                const float cd = X30(30.f * d);
                n1 += d * cd; d1 += cd;
            }
        }
        const float sm0 = (n0 / tmaxf(d0, 1.0f)) * (float)act;
        const float sm1 = (n1 / tmaxf(d1, 1.0f)) * (float)act;
        red.acc.add(sm0, sm1, has_prev, act);
    }
    red.finish(s);
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
        const float f = L10(1.f - eta[p]) * (float)s.af[g.v_cls[p]];
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
        if (lit * sg > 0.f) s.af[g.v_cls[p]] = 0;
    }
    s.av[i] = 0;
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
            if (s.single[i]) { s.af[i] = 0; s.single[i] = 0; s.masked[g.bfm[i]] = 1; }
            else if (s.af[i] && s.conflicts[g.bfm[i]] == 1) s.af[i] = 0;
        }
        if (i < g.V) {
            if (s.av[i] && s.conflicts[g.bvm[i]] == 1) s.av[i] = 0;
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
        for (int p = beg; p < end; ++p) s.af[g.v_cls[p]] = 0;
        s.av[i] = 0;
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
