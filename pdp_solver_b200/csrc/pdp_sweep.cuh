// pdp_sweep.cuh -- the blocked shared-memory passes of the SP sweep (product path of pdp_sp_run, included by
// pdp_device.cuh).  Layout: pdp_common.cuh / DESIGN.md.  One CTA walks blocks of whole nodes; per block
//     load      contiguous 128-bit reads of the block's message region, scattered into NODE order in shared memory
//               through a 16-bit local index (g.vinv / g.cinv);
//     node      the per-node arithmetic, in place, without shared-memory bank conflicts:
//                 clauses of a uniform-degree block: thread a reads X[k * a + j] (k odd: conflict-free);
//                 variables: a WARP owns 32 variables of (nearly) equal degree and the block stores their edges
//                 transposed -- row j of the group holds the j-th edge of every member that has one, so lane l reads
//                 word `row start + l`: consecutive banks, and the ascending-edge accumulation order of the reference
//                 (torch.mm(sparse, dense) on CPU) is the row order;
//     write-out the results in ascending destination order.  Consecutive slots hit consecutive destinations (a RUN:
//               the edges between this block and one block of the other side), so destinations are not stored per edge:
//               one bit per slot marks the start of a run, one 32-bit count per 32 slots gives the run index, one
//               32-bit offset per run gives `destination - slot`.  The run offsets of a block are staged in shared memory.
// Dynamic shared memory of a CTA (SweepCfg<CTAS>::kSmem bytes):
//   [0, 4 * kBlkC)      clause pass: plane X;  variable pass: planes PA | PB (kBlkV words each)
//   then 4 * kBitWords  skip bits: slots of nodes the pass leaves alone (frozen problems inside a block that has work)
//   then 4 * kBitWords  sticky bits: slots of problems on the sticky-NaN path
//   then 4 * kAdjCap B  run offsets of the block being written out
#pragma once

__device__ __forceinline__ bool blk_problem_runs(const pdp_state& s, int b) { return s.active[b] != 0; }

// true when no problem in [b0, b1] is to be processed by the blocked passes (uniform over the CTA)
__device__ __forceinline__ bool blk_idle(const pdp_state& s, int b0, int b1) {
    if (b0 == b1) return !blk_problem_runs(s, b0);
    int any = 0;
    for (int b = b0 + (int)threadIdx.x; b <= b1; b += (int)blockDim.x) any |= blk_problem_runs(s, b) ? 1 : 0;
    return __syncthreads_or(any) == 0;
}

__device__ __forceinline__ void blk_mark_skip(uint32_t* bits, int lo, int hi) {
    for (int l = lo; l < hi; ++l) atomicOr(&bits[l >> 5], 1u << (l & 31));
}
// the same for one row of a transposed variable group: the slots `off + lane` of the lanes in `lanes`
__device__ __forceinline__ void row_mark(uint32_t* bits, int off, unsigned lanes) {
    if (!lanes) return;   // warp-uniform
    const int w = off >> 5, sh = off & 31;
    if (lane_id() == 0) atomicOr(&bits[w], lanes << sh);
    if (lane_id() == 1 && sh) { const unsigned hi = lanes >> (32 - sh); if (hi) atomicOr(&bits[w + 1], hi); }
}

#ifndef PDP_UNROLL_WO
#define PDP_UNROLL_WO 6    // write-out: slots per thread in flight (two loads each)
#endif
#ifndef PDP_UNROLL_CL4
#define PDP_UNROLL_CL4 4   // clause load: four-slot groups per thread in flight
#endif
#ifndef PDP_UNROLL_VL4
#define PDP_UNROLL_VL4 3   // variable load (measured: 2 -> 3 is +1 %)
#endif
#ifndef PDP_INPASS_SCORE
#define PDP_INPASS_SCORE 1
#endif
#ifndef PDP_COLD
#define PDP_COLD __forceinline__
#endif

// The memory phases read their contiguous streams four slots per thread and instruction (128-bit loads of the fp32
// streams, 64-bit loads of the 16-bit tables).  A block's region starts at an arbitrary element of 256-byte aligned
// arrays, so up to three head and three tail slots go through the scalar path.
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

// geometry of one block of a pass: one 32-byte descriptor (two 16-byte loads)
struct BlkGeo {
    int n0, n1;      // node range
    int e0, ne;      // first slot / slots
    int b0, b1;      // problem range
    int run0, nruns; // write-out runs of the block: [run0, run0 + nruns)
    __device__ __forceinline__ bool multi() const { return b0 != b1; }
};
__device__ __forceinline__ BlkGeo load_block(const pdp_blk* __restrict__ desc, int blk) {
    const int4* p = reinterpret_cast<const int4*>(desc + blk);
    const int4 a = __ldg(p), b = __ldg(p + 1);
    BlkGeo B;
    B.n0 = a.x; B.n1 = a.y; B.e0 = a.z; B.ne = a.w; B.b0 = b.x; B.b1 = b.y; B.run0 = b.z; B.nruns = b.w;
    return B;
}
__device__ __forceinline__ BlkGeo clause_block(const pdp_graph& g, int blk) { return load_block(g.cb_desc, blk); }
__device__ __forceinline__ BlkGeo var_block(const pdp_graph& g, int blk) { return load_block(g.vb_desc, blk); }

// the block's run offsets -> shared memory (when they fit; the write-out reads them from global memory otherwise)
template <int G, int CAP>
__device__ __forceinline__ void ph_stage_runs(int t, const int32_t* __restrict__ wadj, const BlkGeo& B, int32_t* adj_s) {
    if (B.nruns > CAP) return;
    for (int i = t; i < B.nruns; i += G) adj_s[i] = __ldg(&wadj[B.run0 + i]);
}

// write-out: slots [0, ne) of the block = the positions of its own region, in load order (`src` is the load table:
// position -> local node slot, SMASK strips its flag bit); destinations ascend.
// STICKY: 0 = off, 1 = slots flagged in `sticky` bits, 2 = every slot.  A sticky slot keeps a NaN that is already
// stored at its destination in `old` (the reference blends mask*new + (1-mask)*old arithmetically: 0*NaN = NaN,
// pdp_propagate.py:175,218).  One slot per thread and instruction: with four consecutive slots per thread a warp's
// stores would be strided by four elements and every destination sector written four times.
template <int G, bool SKIP, int STICKY, bool ADJ_S, unsigned SMASK>
__device__ __forceinline__ void ph_write_out_t(int t, const uint16_t* __restrict__ src, const uint2* __restrict__ wrun,
                                               const int32_t* __restrict__ adj, const BlkGeo& B, const float* plane,
                                               const uint32_t* skip, const uint32_t* sticky,
                                               const float* old, float* out) {   // old may alias out (q is updated in place)
    // A warp takes 32 slots that share one word of the run table: slot w = w0 + 32 k + lane with w0 a multiple of 32, so
    // the lane's bit mask is fixed and the table word is one broadcast load.  Slots before e0 / from e0 + ne on are idle.
    const int lane = t & 31;
    const int w0 = (B.e0 & ~31) + 32 * (t >> 5);
    const int wend = B.e0 + B.ne;
    const uint32_t lmask = 0xffffffffu >> (31 - lane);
    adj -= ADJ_S ? B.run0 : 0;
    auto one = [&](int w, int l, uint2 rb) {
        if (SKIP && ((skip[l >> 5] >> (l & 31)) & 1u)) return;
        const int d = adj[(int)rb.y + __popc(rb.x & lmask)] + w;
        float v = plane[l];
        if (STICKY == 2 || (STICKY == 1 && ((sticky[l >> 5] >> (l & 31)) & 1u))) { const float ov = old[d]; if (ov != ov) v = ov; }
        out[d] = v;
    };
    constexpr int U = PDP_UNROLL_WO;
    constexpr int STEP = G;      // slots between two iterations of a warp (G / 32 warps x 32 slots)
    int w = w0 + lane;
    if (w < B.e0 && w < wend) w += STEP;      // (only the first 32 slots can lie before the region)
    else if (w < B.e0) return;
    // the head iteration above is folded in: a lane whose first slot precedes e0 starts one step later
    for (; w + (U - 1) * STEP < wend; w += U * STEP) {
        int l[U]; uint2 rb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { l[u] = src[w + u * STEP] & SMASK; rb[u] = __ldg(&wrun[(w + u * STEP) >> 5]); }
#pragma unroll
        for (int u = 0; u < U; ++u) one(w + u * STEP, l[u], rb[u]);
    }
    for (; w < wend; w += STEP) one(w, src[w] & SMASK, __ldg(&wrun[w >> 5]));
}
// flags: bit 0 = some slots are skipped, bit 1 = some slots are sticky, bit 2 = every slot is sticky
template <int G, int CAP, unsigned SMASK>
__device__ __forceinline__ void ph_write_out(int t, const uint16_t* __restrict__ src, const uint2* __restrict__ wrun,
                                             const int32_t* __restrict__ wadj, const int32_t* adj_s, const BlkGeo& B,
                                             const float* plane, const uint32_t* skip, int flags, float* out,
                                             const uint32_t* sticky, const float* old) {
    if (B.nruns <= CAP) {
        if (flags == 0) ph_write_out_t<G, false, 0, true, SMASK>(t, src, wrun, adj_s, B, plane, skip, sticky, old, out);
        else if (flags & 4) ph_write_out_t<G, true, 2, true, SMASK>(t, src, wrun, adj_s, B, plane, skip, sticky, old, out);
        else ph_write_out_t<G, true, 1, true, SMASK>(t, src, wrun, adj_s, B, plane, skip, sticky, old, out);
    } else {
        if (flags & 4) ph_write_out_t<G, true, 2, false, SMASK>(t, src, wrun, wadj, B, plane, skip, sticky, old, out);
        else ph_write_out_t<G, true, 1, false, SMASK>(t, src, wrun, wadj, B, plane, skip, sticky, old, out);
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

// clause pass, node phase: thread per clause
template <int G>
__device__ __forceinline__ void ph_clause_node(int t, const pdp_graph& g, const pdp_state& s, const BlkGeo& B, int ku,
                                               float* X, uint32_t* skip, int* any_skip, uint32_t* sticky) {
    const bool multi = B.multi();
    for (int a = B.n0 + t; a < B.n1; a += G) {
        int lo, k;
        if (ku) { k = ku; lo = (a - B.n0) * ku; }
        else { lo = g.cl_ptr[a] - B.e0; k = g.cl_ptr[a + 1] - B.e0 - lo; }
        int b = B.b0;
        if (multi) {
            b = g.bfm[a];
            if (!blk_problem_runs(s, b)) { blk_mark_skip(skip, lo, lo + k); atomicOr(any_skip, 1); continue; }
            if (s.nanflag[b]) { blk_mark_skip(sticky, lo, lo + k); atomicOr(any_skip, 2); }
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

// branch-free select on a sign mask (all ones / all zeros): one LOP3
__device__ __forceinline__ float fsel(uint32_t mask, float a, float b) {
    return __uint_as_float((__float_as_uint(a) & mask) | (__float_as_uint(b) & ~mask));
}
__device__ __forceinline__ float fand(uint32_t mask, float a) { return __uint_as_float(__float_as_uint(a) & mask); }

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
}

// One warp-group of a variable block: 32 variables of consecutive rank in the block's descending-degree order
// (g.vsort), lane l <-> rank 32 * group + l.  The group owns 32 * (degree of its first member) slots from its base
// (a multiple of 32): row j = slots [base + 32 j, base + 32 j + 32), lane l's j-th edge sits at base + 32 j + l.
// Slots of members with fewer edges than the first one are padding (never written, never read): 2-5 % of a block.
struct VarGroup {
    int i, deg, base, maxdeg;
    bool have;
};
__device__ __forceinline__ VarGroup var_group(const pdp_graph& g, const BlkGeo& B, int grp) {
    VarGroup V;
    const int ti = B.n0 + grp * 32 + lane_id();
    V.have = ti < B.n1;
    V.i = 0; V.deg = 0; V.base = 0;
    if (V.have) { const int2 ve = __ldg(&g.vsort[ti]); V.i = ve.x; V.base = ve.y & 0xffff; V.deg = (int)((unsigned)ve.y >> 16); }
    V.base = __shfl_sync(0xffffffffu, V.base, 0);     // every entry of a group carries the group's base; lane 0 exists
    V.maxdeg = __shfl_sync(0xffffffffu, V.deg, 0);
    return V;
}
// groups of a block are handed to the warps round by round, odd rounds in reverse (every warp gets high and low degrees)
#define VAR_GROUP_LOOP(grp, B, G, t)                                                                        \
    for (int _gb = 0, _rd = 0, _ng = ((B).n1 - (B).n0 + 31) >> 5, grp; _gb < _ng; _gb += (G) / 32, ++_rd)   \
        if ((grp = (_rd & 1) ? _gb + ((G) / 32 - 1 - ((t) >> 5)) : _gb + ((t) >> 5)) < _ng)

// variable pass, SurveyScorer (pdp_predict.py:155-192) of the variables whose problem asked for it (want_score): same
// operations and order as score_variable() on the new surveys held in PA (sign bit = edge masked; for an active variable
// the edge mask is the clause mask the scorer multiplies with, an inactive variable's score is never looked at) and the
// literal signs held in PB.  pi == 0 on the blocked path: the external force does not enter.
// (The reference's pos / neg incidence products hold explicit zeros, 0 * f: they only matter when f is NaN, and then
// the sum over all edges is NaN as well and with it bias and the score: the per-sign sums skip them.)
template <int G>
__device__ __forceinline__ void ph_var_score(int t, const pdp_graph& g, const pdp_state& s, const BlkGeo& B,
                                             const float* __restrict__ PA, const float* __restrict__ PB) {
    VAR_GROUP_LOOP(grp, B, G, t) {
        const VarGroup V = var_group(g, B, grp);
        if (!(V.have && s.want_score[B.multi() ? g.bvm[V.i] : B.b0])) continue;
        float ps = 0.f, ns = 0.f, as = 0.f;
        const float* __restrict__ pa = PA + V.base + lane_id();
        const float* __restrict__ pb = PB + V.base + lane_id();
        for (int j = 0; j < V.deg; ++j) {
            const uint32_t nb = __float_as_uint(pa[32 * j]), ob = __float_as_uint(pb[32 * j]);
            const uint32_t negm = (uint32_t)((int32_t)ob >> 31);
            const float f = L10(1.f - __uint_as_float(nb & 0x7fffffffu)) * ((nb >> 31) ? 0.f : 1.f);
            ps += fand(~negm, f);
            ns += fand(negm, f);
            as += f;
        }
        s.score[V.i] = sp_score_tail(ps, ns, as, 0.f, 0.f);
    }
}

// variable pass, node phase: ordered sums, decimator statistics, update.
// Requires eta(t-1) >= +0 or NaN without sign (the sign bits are borrowed, see ph_var_load).
// (As in the scorer, the explicit zeros 0 * y of the reference's per-sign sums are skipped: a NaN y makes the sum of its
// own sign NaN, and every message of the variable reads both sums -- `same` the one of its sign, `opp` the other.)
// MULTI: the block holds several problems: frozen ones are left alone (skip bits), those on the sticky-NaN path are
// flagged (sticky bits); a row is one 32-bit word of the bit arrays, written whole by the warp that owns the group.
template <int G, bool MULTI, bool MASKED, bool PREV>
__device__ __forceinline__ void ph_var_node(int t, const pdp_graph& g, const pdp_state& s, const BlkGeo& B, bool use_mask,
                                            bool em_set, float* __restrict__ PA, float* __restrict__ PB, uint32_t* skip, int* any_skip,
                                            KeyedReducer<StatAcc>& red, BlkStats& sm_st, bool local_stats, uint32_t* sticky) {
    const int lane = t & 31;
    VAR_GROUP_LOOP(grp, B, G, t) {
        const VarGroup V = var_group(g, B, grp);
        int b = B.b0;
        bool runs = V.have;
        if (MULTI) {
            bool stk = false;
            if (V.have) { b = g.bvm[V.i]; runs = blk_problem_runs(s, b); stk = runs && s.nanflag[b]; }
            const unsigned skip_l = __ballot_sync(0xffffffffu, V.have && !runs);
            const unsigned stk_l = __ballot_sync(0xffffffffu, stk);
            if (lane == 0 && (skip_l | stk_l)) atomicOr(any_skip, (skip_l ? 1 : 0) | (stk_l ? 2 : 0));
            for (int j = 0; j < V.maxdeg; ++j) {
                const unsigned on = __ballot_sync(0xffffffffu, j < V.deg);
                if (lane == 0) { skip[(V.base >> 5) + j] = on & skip_l; sticky[(V.base >> 5) + j] = on & stk_l; }
            }
            if (!__any_sync(0xffffffffu, runs)) continue;
        }
        if (!runs) continue;
        const uint32_t act = (uint32_t)s.av[V.i];
        float* __restrict__ pa = PA + V.base + lane;
        float* __restrict__ pb = PB + V.base + lane;
        float P = 0.f, N = 0.f, n0 = 0.f, d0 = 0.f, n1 = 0.f, d1 = 0.f;
#pragma unroll 4
        for (int j = 0; j < V.deg; ++j) {
            const uint32_t nb = __float_as_uint(pa[32 * j]), ob = __float_as_uint(pb[32 * j]);
            const uint32_t negm = (uint32_t)((int32_t)ob >> 31);    // all ones: negative literal
            const float xn = __uint_as_float(nb & 0x7fffffffu), xo = __uint_as_float(ob & 0x7fffffffu);
            float y = L40_1m(xo);
            if (MASKED && use_mask && (int32_t)nb < 0) y = y * 0.f;      // edge masked
            // y <= +0 (or NaN): stored as |y| under the literal's sign bit
            pb[32 * j] = __uint_as_float((__float_as_uint(y) & 0x7fffffffu) | (ob & 0x80000000u));
            P += fand(~negm, y);
            N += fand(negm, y);
            const float c = X30S(xn);
            n0 += xn * c; d0 += c;
            if (PREV) {
                float d = fabsf(xo - xn);
                if (MASKED && em_set && (int32_t)nb < 0) d = d * 0.f;
                const float cd = X30S(d);
                n1 += d * cd; d1 += cd;
            }
        }
        {
            const float sm0 = pdp_divs(n0, tmaxf(d0, 1.0f)) * (float)act;
            const float sm1 = pdp_divs(n1, tmaxf(d1, 1.0f)) * (float)act;
            if (!MULTI) {
                red.touch(s, b);
                red.acc.add(sm0, sm1, PREV, act);
            } else {
                StatAcc one;
                one.reset();
                one.add(sm0, sm1, PREV, act);
                if (local_stats) {
                    const int lb = b - B.b0;
                    atomicMax(&sm_st.mx0[lb], one.mx0); atomicMin(&sm_st.mn0[lb], one.mn0);
                    if (PREV) { atomicMax(&sm_st.mx1[lb], one.mx1); atomicMin(&sm_st.mn1[lb], one.mn1); }
                    if (one.nan) atomicOr(&sm_st.nan[lb], one.nan);
                    if (act) atomicAdd(&sm_st.nav[lb], act);
                } else {
                    one.commit(s, b);
                }
            }
        }
        float sb_pos, opp_pos, O_pos, sb_neg, opp_neg, O_neg;
        sp_var_prepare(P, N, 1.f, sb_pos, opp_pos, O_pos);
        sp_var_prepare(P, N, -1.f, sb_neg, opp_neg, O_neg);
        bool made_nan = false;
#pragma unroll 4
        for (int j = 0; j < V.deg; ++j) {
            const uint32_t yb = __float_as_uint(pb[32 * j]);
            const uint32_t negm = (uint32_t)((int32_t)yb >> 31);
            const float y = __uint_as_float(yb | 0x80000000u);   // -|y|
            const float u = sp_var_finish(fsel(negm, sb_neg, sb_pos), fsel(negm, opp_neg, opp_pos), fsel(negm, O_neg, O_pos), y);
            made_nan |= (u != u);
            pa[32 * j] = u;
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

// PDP_PHASE_TIMING (profiling builds only): thread 0 of every CTA adds the clock cycles (>> 10) it spent in each
// phase of the blocked passes to the trace buffer: [0..2] clause load / node / write-out, [3..5] variable
#ifdef PDP_PHASE_TIMING
#define PHASE_T0() long long _pt = clock64()
#define PHASE_ADD(slot_) do { if (threadIdx.x == 0 && A.trace) { const long long _n = clock64(); atomicAdd(&A.trace[slot_], (int)((_n - _pt) >> 10)); _pt = _n; } } while (0)
#else
#define PHASE_T0() do {} while (0)
#define PHASE_ADD(slot_) do {} while (0)
#endif

// Block hand-out of one pass.  Static (block b to CTA b mod grid) when one CTA owns the SM: the CTAs then finish within
// 1-2 % of each other.  With two CTAs per SM the pair drifts apart, the early one waits at the grid barrier and its
// partner finishes alone at half the SM's warps (10 % of the iteration, measured), so the blocks after a CTA's first one
// come from a counter; the next index is fetched while the current block is processed.
// thread 0, at the top of a block: the index of the block after `blk`, left in slot[par] for feed_advance
template <bool DYN>
__device__ __forceinline__ void feed_fetch(int* ctr, int* slot, int par, int blk) {
    slot[par] = DYN ? (int)gridDim.x + atomicAdd(ctr, 1) : blk + (int)gridDim.x;
}
// all threads, after the block's last use of shared memory
__device__ __forceinline__ int feed_advance(const int* slot, int& par) {
    __syncthreads();
    const int nx = slot[par];
    par ^= 1;
    return nx;
}

// Two CTAs share an SM so that one's memory phases (load, write-out: DRAM latency, idle issue slots) overlap the other's
// node phase (issue bound, no DRAM traffic).  Both leave a grid barrier at the same time, and with equal blocks they
// would then run the same phase at the same time for the whole pass: the second CTA of an SM starts a pass about half a
// block period late.
__device__ __forceinline__ void blk_stagger(int sm_rank, int cycles) {
    if (sm_rank == 0 || cycles <= 0) return;
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) __nanosleep(256);
}

// clause pass of iteration t: eta(t) [buffer r^1, V-layout] from q(t-1) [C-layout]
template <int CTAS>
__device__ __forceinline__ void blk_clause_pass(const KArgs& A, int r, bool use_mask, unsigned char* smem, int sm_rank) {
    using Cfg = SweepCfg<CTAS>;
    constexpr int NT = Cfg::kThreads, BLK_C = Cfg::kBlkC, CAP = Cfg::kAdjCap;
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    float* X = reinterpret_cast<float*>(smem);
    uint32_t* skip = reinterpret_cast<uint32_t*>(smem + 4 * BLK_C);
    uint32_t* sticky = skip + Cfg::kBitWords;
    int32_t* adj_s = reinterpret_cast<int32_t*>(sticky + Cfg::kBitWords);
    __shared__ int sm_any_skip;
    const float* __restrict__ qin = s.qu;
    float* __restrict__ eout = s.eta[r ^ 1];
    const int tid = threadIdx.x;
    __shared__ int sm_feed[2];
    constexpr bool DYN = CTAS == 2;
    int par = 0;
    if (CTAS == 2) blk_stagger(sm_rank, A.stagger_c);
    for (int blk = blockIdx.x; blk < g.ncb; blk = feed_advance(sm_feed, par)) {
        if (tid == 0) feed_fetch<DYN>(&s.ctrl[CTRL_NEXT_CBLK], sm_feed, par, blk);
        const BlkGeo B = clause_block(g, blk);
        if (B.n1 <= B.n0) continue;
        if (blk_idle(s, B.b0, B.b1)) continue;
        for (int i = tid; i < (B.ne + 31) / 32; i += NT) { skip[i] = 0u; sticky[i] = 0u; }
        if (tid == 0) sm_any_skip = (!B.multi() && s.nanflag[B.b0]) ? 4 : 0;
        PHASE_T0();
        ph_stage_runs<NT, CAP>(tid, g.c_wadj, B, adj_s);
        if (use_mask && (B.multi() || s.masked[B.b0])) ph_clause_load<NT, true>(tid, qin + B.e0, g.cinv + B.e0, g.qmask, B.e0, B.ne, X);
        else ph_clause_load<NT, false>(tid, qin + B.e0, g.cinv + B.e0, g.qmask, B.e0, B.ne, X);
        __syncthreads();
        PHASE_ADD(0);
        ph_clause_node<NT>(tid, g, s, B, g.cb_k[blk], X, skip, &sm_any_skip, sticky);
        __syncthreads();
        PHASE_ADD(1);
        ph_write_out<NT, CAP, 0xffffu>(tid, g.cinv, g.c_wrun, g.c_wadj, adj_s, B, X, skip, sm_any_skip, eout, sticky, s.eta[r]);
        PHASE_ADD(2);
    }
}

// variable pass of iteration t: the decimator statistics of eta(t) [buffer r^1] against eta(t-1)
// [buffer r], and q(t) [C-layout, in place] from eta(t-1)
template <int CTAS>
__device__ __forceinline__ void blk_var_pass(const KArgs& A, int r, bool use_mask, bool has_prev, bool em_set, unsigned char* smem, int sm_rank) {
    using Cfg = SweepCfg<CTAS>;
    constexpr int NT = Cfg::kThreads, BLK_V = Cfg::kBlkV, BLK_C = Cfg::kBlkC, CAP = Cfg::kAdjCap;
    const pdp_graph& g = A.g; const pdp_state& s = A.s;
    float* PA = reinterpret_cast<float*>(smem);   // eta(t), then q(t)
    float* PB = PA + BLK_V;                       // eta(t-1), then y
    uint32_t* skip = reinterpret_cast<uint32_t*>(smem + 4 * BLK_C);
    uint32_t* sticky = skip + Cfg::kBitWords;
    int32_t* adj_s = reinterpret_cast<int32_t*>(sticky + Cfg::kBitWords);
    __shared__ int sm_any_skip;
    __shared__ BlkStats sm_st;
    const float* __restrict__ en = s.eta[r ^ 1];
    const float* __restrict__ eo = s.eta[r];
    const int tid = threadIdx.x;
    KeyedReducer<StatAcc> red;
    int acc_key = -1;
    __shared__ int sm_feed[2];
    constexpr bool DYN = CTAS == 2;
    int par = 0;
    if (CTAS == 2) blk_stagger(sm_rank, A.stagger_v);
    for (int blk = blockIdx.x; blk < g.nvb; blk = feed_advance(sm_feed, par)) {
        if (tid == 0) feed_fetch<DYN>(&s.ctrl[CTRL_NEXT_VBLK], sm_feed, par, blk);
        const BlkGeo B = var_block(g, blk);
        if (B.n1 <= B.n0) continue;
        if (blk_idle(s, B.b0, B.b1)) continue;
        const bool local_stats = B.multi() && (B.b1 - B.b0 < PDP_STAT_SLOTS);   // else: registers (one problem) or global atomics
        // statistics of single-problem blocks: the per-thread accumulators run across the blocks of one problem and are
        // merged block-wide when the CTA moves on to another problem (and at the end of the pass)
        if (!B.multi() && acc_key != B.b0) { if (acc_key >= 0) red.finish(s); acc_key = B.b0; }
        for (int i = tid; i < BLK_V / 32; i += NT) { skip[i] = 0u; sticky[i] = 0u; }   // (padded slots: the whole plane)
        if (tid == 0) sm_any_skip = (!B.multi() && s.nanflag[B.b0]) ? 4 : 0;
        if (local_stats) stats_slots_reset(sm_st, tid, B.b1 - B.b0 + 1);
        PHASE_T0();
        ph_stage_runs<NT, CAP>(tid, g.v_wadj, B, adj_s);
        if ((use_mask || em_set) && (B.multi() || s.masked[B.b0])) ph_var_load<NT, true>(tid, en + B.e0, eo + B.e0, g.vinv + B.e0, g.vmask, B.e0, B.ne, PA, PB);
        else ph_var_load<NT, false>(tid, en + B.e0, eo + B.e0, g.vinv + B.e0, g.vmask, B.e0, B.ne, PA, PB);
        __syncthreads();
        PHASE_ADD(3);
        // SurveyScorer of problems about to converge, while the new surveys are still in the planes (the node phase
        // overwrites them); its own loop, so that the node phase's code is the same with and without it
        if (PDP_INPASS_SCORE && (s.want_score[B.b0] || s.want_score[B.b1])) ph_var_score<NT>(tid, g, s, B, PA, PB);
        {
            const bool masked = (use_mask || em_set) && (B.multi() || s.masked[B.b0]);
#define VN_CALL(MU, MA, PR) ph_var_node<NT, MU, MA, PR>(tid, g, s, B, use_mask, em_set, PA, PB, skip, &sm_any_skip, red, sm_st, local_stats, sticky)
            if (B.multi()) { if (has_prev) VN_CALL(true, true, true); else VN_CALL(true, true, false); }
            else if (masked) { if (has_prev) VN_CALL(false, true, true); else VN_CALL(false, true, false); }
            else { if (has_prev) VN_CALL(false, false, true); else VN_CALL(false, false, false); }
#undef VN_CALL
        }
        __syncthreads();
        PHASE_ADD(4);
        if (local_stats) stats_slots_commit(s, sm_st, tid, B.b1 - B.b0 + 1, B.b0);
        ph_write_out<NT, CAP, 0x7fffu>(tid, g.vinv, g.v_wrun, g.v_wadj, adj_s, B, PA, skip, sm_any_skip, s.qu, sticky, s.qu);
        PHASE_ADD(5);
    }
    red.finish(s);
}
