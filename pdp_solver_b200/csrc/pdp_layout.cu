// pdp_layout.cu -- builds the blocked message layout of the SP sweep (pdp_common.cuh):
// block partitions of both edge orders, the V-layout / C-layout positions of every edge, the 16-bit
// local scatter / gather tables of the two passes and their write-out pieces.  Runs once per batch
// inside pdp_create, entirely on the device (four stable radix sorts on block ids, two scans, two
// stream compactions); the message arrays are not live yet and serve as scratch.
#include <cub/cub.cuh>
#include <stdlib.h>

#include "pdp_common.cuh"

namespace {

#define GS(i, n) for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

// first node of block t: the first node whose first slot is >= t * stride (lower bound over ptr[0..n))
__global__ void k_block_ptr(const int32_t* __restrict__ ptr, int64_t n, int32_t stride, int32_t nblk, int32_t* blk_ptr) {
    GS(t, (int64_t)nblk + 1) {
        if (t == nblk) { blk_ptr[t] = (int32_t)n; continue; }
        const int64_t target = t * (int64_t)stride;
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (ptr[mid] < target) lo = mid + 1; else hi = mid;
        }
        blk_ptr[t] = (int32_t)lo;
    }
}

// keys of the V-layout sort, in clause-major order: variable block of the edge
__global__ void k_key_vblock_of_c(pdp_graph g, int32_t* key, int32_t* val) {
    GS(c, g.E) {
        const int var = (int)(g.c_var[c] & PDP_IDX_MASK);
        key[c] = g.var_ptr[var] / g.sv;
        val[c] = (int32_t)c;
    }
}
// keys of the C-layout sort, in variable-major order: clause block of the edge
__global__ void k_key_cblock_of_p(pdp_graph g, int32_t* key, int32_t* val) {
    GS(p, g.E) {
        key[p] = g.cl_ptr[g.v_cls[p]] / g.sc;
        val[p] = (int32_t)p;
    }
}

// x = V-layout position, lv[x] = clause-major slot stored there, ki[x] = its variable block
__global__ void k_fill_vlayout(pdp_graph g, const int32_t* __restrict__ ki, const int32_t* __restrict__ lv) {
    GS(x, g.E) {
        const int c = lv[x];
        const int p = g.c_pos[c];
        g.p_vpos[p] = (int32_t)x;
        g.vinv[x] = (uint16_t)((p - g.var_ptr[g.vb_ptr[ki[x]]]) | ((g.v_cedge[p] & PDP_SIGN_BIT) ? PDP_VINV_NEG : 0u));
    }
}
// x = C-layout position, lq[x] = variable-major slot stored there, kj[x] = its clause block
__global__ void k_fill_clayout(pdp_graph g, const int32_t* __restrict__ kj, const int32_t* __restrict__ lq) {
    GS(x, g.E) {
        const int p = lq[x];
        const int c = (int)(g.v_cedge[p] & PDP_IDX_MASK);
        g.p_qpos[p] = (int32_t)x;
        g.cinv[x] = (uint16_t)(c - g.cl_ptr[g.cb_ptr[kj[x]]]);
    }
}

// write-out order of the variable blocks: the edges of a block sorted by their C-layout position.
// Input in C-layout order x (lq[x] = variable-major slot): key = variable block, value = x.
// block of a slot of the node-major order: the stride region the slot lies in, or an earlier one when the slot's node
// started before that region (block t begins at the first node whose first slot is >= t * stride, k_block_ptr).  Two
// loads from the small block tables instead of three dependent random gathers through the adjacency.
__device__ __forceinline__ int block_of_slot(const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ blk_ptr, int stride, int slot) {
    int b = slot / stride;
    while (b > 0 && slot < node_ptr[blk_ptr[b]]) --b;
    return b;
}
__global__ void k_key_vblock_of_q(pdp_graph g, const int32_t* __restrict__ lq, int32_t* key, int32_t* val) {
    GS(x, g.E) {
        key[x] = block_of_slot(g.var_ptr, g.vb_ptr, g.sv, lq[x]);
        val[x] = (int32_t)x;
    }
}
__global__ void k_key_cblock_of_v(pdp_graph g, const int32_t* __restrict__ lv, int32_t* key, int32_t* val) {
    GS(x, g.E) {
        key[x] = block_of_slot(g.cl_ptr, g.cb_ptr, g.sc, lv[x]);
        val[x] = (int32_t)x;
    }
}

// w = write-out slot; kb[w] = block, dst[w] = destination position, slot_of_dst[dst] = producer-order slot.
// src16[w] = local producer-order index
__global__ void k_fill_writeout(int64_t E, const int32_t* __restrict__ kb, const int32_t* __restrict__ dst,
                                const int32_t* __restrict__ slot_of_dst, const int32_t* __restrict__ node_ptr,
                                const int32_t* __restrict__ blk_ptr, uint16_t* src16, int32_t* dst_out) {
    GS(w, E) {
        const int d = dst[w];
        src16[w] = (uint16_t)(slot_of_dst[d] - node_ptr[blk_ptr[kb[w]]]);
        dst_out[w] = d;
    }
}

#if PDP_TMA
// tables of the TMA-staged passes.  A block's region is staged from its 16-byte aligned start
// (first position & ~3), so a staged position is (layout position) - (first position & ~3).
__global__ void k_fill_staged_tables(pdp_graph g) {
    GS(p, g.E) {   // variable side
        const int var = (int)(g.c_var[g.v_cedge[p] & PDP_IDX_MASK] & PDP_IDX_MASK);
        const int e0 = g.var_ptr[g.vb_ptr[g.var_ptr[var] / g.sv]];
        g.vperm[p] = (uint16_t)((g.p_vpos[p] - (e0 & ~3)) | ((g.v_cedge[p] & PDP_SIGN_BIT) ? PDP_VINV_NEG : 0u));
        // write-out slot w = p ranges over the same block: its source is the edge at local index vsrc[w]
        g.vsrc2[p] = (uint16_t)(g.p_vpos[e0 + g.vsrc[p]] - (e0 & ~3));
    }
    GS(c, g.E) {   // clause side
        const int cl = g.v_cls[g.c_pos[c]];
        const int e0 = g.cl_ptr[g.cb_ptr[g.cl_ptr[cl] / g.sc]];
        g.cperm[c] = (uint16_t)(cqpos(g, c) - (e0 & ~3));
        g.csrc2[c] = (uint16_t)(cqpos(g, e0 + g.csrc[c]) - (e0 & ~3));
    }
}
#endif

// degree-sorted variable order: key = block << 14 | (16383 - degree)
__global__ void k_key_vsort(pdp_graph g, int32_t* key, int32_t* val) {
    GS(v, g.V) {
        const int deg = g.var_ptr[v + 1] - g.var_ptr[v];
        key[v] = ((g.var_ptr[v] / g.sv) << 14) | (16383 - (deg > 16383 ? 16383 : deg));
        val[v] = (int32_t)v;
    }
}
// order == nullptr: identity
__global__ void k_fill_vsort(pdp_graph g, const int32_t* __restrict__ order) {
    GS(t, g.V) {
        const int v = order ? order[t] : (int)t;
        const int blk = g.var_ptr[v] / g.sv;
        const int lo = g.var_ptr[v] - g.var_ptr[g.vb_ptr[blk]];
        const int deg = g.var_ptr[v + 1] - g.var_ptr[v];
        g.vsort[t] = make_int2(v, lo | (deg << 16));
    }
}

// cb_k[blk] = common degree of the block's clauses (1..8) or 0: seeded with the first clause's degree,
// cleared by any clause that disagrees
__global__ void k_clause_block_degree_init(pdp_graph g) {
    GS(blk, g.ncb) {
        const int a0 = g.cb_ptr[blk], a1 = g.cb_ptr[blk + 1];
        const int k = (a1 > a0) ? g.cl_ptr[a0 + 1] - g.cl_ptr[a0] : 0;
        g.cb_k[blk] = (k >= 1 && k <= 8) ? k : 0;
    }
}
__global__ void k_clause_block_degree_check(pdp_graph g) {
    GS(a, g.F) {
        const int blk = g.cl_ptr[a] / g.sc;
        const int k = g.cb_k[blk];
        if (k != 0 && g.cl_ptr[a + 1] - g.cl_ptr[a] != k) g.cb_k[blk] = 0;
    }
}

__global__ void k_identity_layout(pdp_graph g) {
    GS(c, g.E) {
        const int p = g.c_pos[c];
        g.p_vpos[p] = p; g.p_qpos[p] = p;
    }
}

int bits_for(int64_t n) {
    int bits = 1;
    while (((int64_t)1 << bits) < n && bits < 31) ++bits;
    return bits;
}

}  // namespace

#define LCK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { pdp_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); return PDP_ERR_CUDA; } } while (0)
#define LLK() do { c->launches++; cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) { pdp_set_error("%s:%d: launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); return PDP_ERR_CUDA; } } while (0)
#define G1(n) pdp_grid((n), 256, nsm), 256, 0, stream

// block stride: as large as a block allows, shrunk so that the block count is a multiple of the SM count
static int32_t pick_stride(int64_t E, int32_t blk, int32_t max_degree, int nsm) {
    const int64_t smax = (int64_t)blk - max_degree + 1;
    const int64_t rounds = (E + smax * nsm - 1) / (smax * nsm);
    int64_t s = (E + rounds * nsm - 1) / (rounds * nsm);
    if (s < 64) s = 64;
    if (s > smax) s = smax;
    return (int32_t)s;
}

int pdp_build_layout(pdp_ctx* c, cudaStream_t stream, bool monotone_maps) {
    pdp_graph& g = c->g;
    const int nsm = c->num_sms;
    const int64_t E = g.E;
    g.blocked_ok = 0; g.nvb = 0; g.ncb = 0; g.sv = 1; g.sc = 1;
    if (E == 0) return PDP_OK;
    g.ctas = 1;
    if (!(PDP_PIPELINE || PDP_TMA)) {
        const char* forced = getenv("PDP_B200_CTAS");
        if (forced) g.ctas = (atoi(forced) == 2) ? 2 : 1;
        else if (g.B > 0 && E / g.B >= PDP_CTAS_EDGES_PER_PROBLEM) g.ctas = 2;
    }
    const int blk_v = PDP_BLK_V / g.ctas, blk_c = PDP_BLK_C / g.ctas;
    const bool ok = monotone_maps && g.max_var_degree <= blk_v / 2 && g.max_clause_degree <= blk_c / 2 &&
                    g.V > 0 && g.F > 0 && getenv("PDP_B200_NO_BLOCKED") == nullptr;
    if (!ok) {
        k_identity_layout<<<G1(E)>>>(g);
        LLK();
        return PDP_OK;
    }
    g.sv = pick_stride(E, blk_v, g.max_var_degree, nsm * g.ctas);
    g.sc = pick_stride(E, blk_c, g.max_clause_degree, nsm * g.ctas);
    g.nvb = (int32_t)(E / g.sv + 1);
    g.ncb = (int32_t)(E / g.sc + 1);
    if (nsm > PDP_MAX_SMS || g.nvb > E / (PDP_BLK_V / 4) + 2 * PDP_MAX_SMS + 2 || g.ncb > E / (PDP_BLK_C / 4) + 2 * PDP_MAX_SMS + 2) {
        pdp_set_error("pdp_create: block tables too small (nvb=%d ncb=%d)", g.nvb, g.ncb);
        return PDP_ERR_WORKSPACE;
    }
    k_block_ptr<<<G1(g.nvb + 1)>>>(g.var_ptr, g.V, g.sv, g.nvb, g.vb_ptr);
    LLK();
    k_block_ptr<<<G1(g.ncb + 1)>>>(g.cl_ptr, g.F, g.sc, g.ncb, g.cb_ptr);
    LLK();

    // scratch: the message arrays
    int32_t* S0 = reinterpret_cast<int32_t*>(c->s.eta[0]);
    int32_t* S1 = reinterpret_cast<int32_t*>(c->s.eta[1]);
    int32_t* S2 = reinterpret_cast<int32_t*>(c->s.qu);
    int32_t* S3 = reinterpret_cast<int32_t*>(c->s.qs);
    int32_t* LV = reinterpret_cast<int32_t*>(c->s.qd);    // V-layout position -> clause-major slot
    int32_t* LQ = reinterpret_cast<int32_t*>(c->s.ext);   // C-layout position -> variable-major slot

    // stable sort of (key, value) pairs held in S0 / S2; returns the buffers holding the result
    auto sort_pairs = [&](int64_t count, int64_t nkeys, int32_t** keys_out, int32_t** vals_out, int32_t** keys_free, int32_t** vals_free) -> int {
        cub::DoubleBuffer<int32_t> dk(S0, S1);
        cub::DoubleBuffer<int32_t> dv(S2, S3);
        size_t q = 0;
        const int bits = bits_for(nkeys);
        if (cub::DeviceRadixSort::SortPairs(nullptr, q, dk, dv, (int)count, 0, bits, stream) != cudaSuccess) return PDP_ERR_CUDA;
        if (q > c->cub_tmp_bytes) { pdp_set_error("pdp_create: sort scratch %zu > reserve %zu", q, c->cub_tmp_bytes); return PDP_ERR_WORKSPACE; }
        size_t tb = c->cub_tmp_bytes;
        if (cub::DeviceRadixSort::SortPairs(c->cub_tmp, tb, dk, dv, (int)count, 0, bits, stream) != cudaSuccess) return PDP_ERR_CUDA;
        c->launches++;
        *keys_out = dk.Current(); *vals_out = dv.Current(); *keys_free = dk.Alternate(); *vals_free = dv.Alternate();
        return PDP_OK;
    };
    int32_t *K, *L, *KF, *LF;
    int rc;

    // ---- V-layout: edges sorted by (variable block, clause-major slot)
    k_key_vblock_of_c<<<G1(E)>>>(g, S0, S2);
    LLK();
    if ((rc = sort_pairs(E, g.nvb, &K, &L, &KF, &LF)) != PDP_OK) return rc;
    k_fill_vlayout<<<G1(E)>>>(g, K, L);
    LLK();
    LCK(cudaMemcpyAsync(LV, L, sizeof(int32_t) * (size_t)E, cudaMemcpyDeviceToDevice, stream));
    // ---- C-layout: edges sorted by (clause block, variable-major slot)
    k_key_cblock_of_p<<<G1(E)>>>(g, S0, S2);
    LLK();
    if ((rc = sort_pairs(E, g.ncb, &K, &L, &KF, &LF)) != PDP_OK) return rc;
    k_fill_clayout<<<G1(E)>>>(g, K, L);
    LLK();
    LCK(cudaMemcpyAsync(LQ, L, sizeof(int32_t) * (size_t)E, cudaMemcpyDeviceToDevice, stream));

    // ---- write-out orders: the edges of a block sorted by their destination position
    for (int side = 0; side < 2; ++side) {
        const bool var_side = (side == 0);
        if (var_side) k_key_vblock_of_q<<<G1(E)>>>(g, LQ, S0, S2);
        else k_key_cblock_of_v<<<G1(E)>>>(g, LV, S0, S2);
        LLK();
        if ((rc = sort_pairs(E, var_side ? g.nvb : g.ncb, &K, &L, &KF, &LF)) != PDP_OK) return rc;
        // K[w] = block, L[w] = destination position
        k_fill_writeout<<<G1(E)>>>(E, K, L, var_side ? LQ : LV, var_side ? g.var_ptr : g.cl_ptr,
                                    var_side ? g.vb_ptr : g.cb_ptr, var_side ? g.vsrc : g.csrc, var_side ? g.vdst : g.cdst);
        LLK();
    }
#if PDP_TMA
    k_fill_staged_tables<<<G1(E)>>>(g);
    LLK();
#endif
    // ---- per block: variables by descending degree (uniform trip counts inside a warp), clause degree
    if (g.V <= E && g.max_var_degree < 16384 && g.nvb < (1 << 17)) {
        k_key_vsort<<<G1(g.V)>>>(g, S0, S2);
        LLK();
        if ((rc = sort_pairs(g.V, (int64_t)g.nvb << 14, &K, &L, &KF, &LF)) != PDP_OK) return rc;
        k_fill_vsort<<<G1(g.V)>>>(g, L);
    } else {
        k_fill_vsort<<<G1(g.V)>>>(g, nullptr);
    }
    LLK();
    k_clause_block_degree_init<<<G1(g.ncb)>>>(g);
    LLK();
    k_clause_block_degree_check<<<G1(g.F)>>>(g);
    LLK();
    g.blocked_ok = 1;
    return PDP_OK;
}

// ------------------------------------------------------------------------------------------------
// self-check of the layout tables (tests): counts violated invariants into errs[0..7]
//   0 position maps inconsistent between the two edge orders     1 V-layout position outside its block / vinv wrong
//   2 C-layout position outside its block / cinv wrong           3 variable write-out does not land on p_qpos
//   4 clause write-out does not land on c_vpos                   5 vsort / cb_k wrong
//   6 distinct write-out sources of the variable blocks (= E)    7 ... of the clause blocks (= E)
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void k_check_layout(pdp_graph g, int32_t* errs, uint32_t* seen /* [E/32+1] x 2, zeroed */) {
    GS(c, g.E) {
        const int p = g.c_pos[c];
        if (cvpos(g, c) != g.p_vpos[p] || cqpos(g, c) != g.p_qpos[p] || (int)(g.v_cedge[p] & PDP_IDX_MASK) != (int)c) atomicAdd(&errs[0], 1);
    }
    if (!g.blocked_ok) return;
    uint32_t* seen_c = seen + g.E / 32 + 1;
    GS(blk, g.nvb) {
        const int v0 = g.vb_ptr[blk], v1 = g.vb_ptr[blk + 1];
        const int e0 = g.var_ptr[v0], e1 = g.var_ptr[v1];
        if (e1 - e0 > PDP_BLK_V / g.ctas) atomicAdd(&errs[1], 1);
        for (int p = e0; p < e1; ++p) {
            const int x = g.p_vpos[p];
            if (x < e0 || x >= e1 || (int)(g.vinv[x] & 0x7fff) != p - e0 ||
                ((g.vinv[x] & PDP_VINV_NEG) != 0) != ((g.v_cedge[p] & PDP_SIGN_BIT) != 0)) atomicAdd(&errs[1], 1);
        }
        for (int w = e0; w < e1; ++w) {
            const int p = e0 + (int)g.vsrc[w];
            if (p < e0 || p >= e1 || g.p_qpos[p] != g.vdst[w]) { atomicAdd(&errs[3], 1); continue; }
            if (w > e0 && g.vdst[w] <= g.vdst[w - 1]) atomicAdd(&errs[3], 1);
            if (!(atomicOr(&seen[p >> 5], 1u << (p & 31)) & (1u << (p & 31)))) atomicAdd(&errs[6], 1);
        }
        for (int p = e0; p < e1 && PDP_TMA; ++p) {
            const int x = (int)(g.vperm[p] & 0x7fff) + (e0 & ~3);
            if (x != g.p_vpos[p] || ((g.vperm[p] & PDP_VINV_NEG) != 0) != ((g.v_cedge[p] & PDP_SIGN_BIT) != 0)) atomicAdd(&errs[1], 1);
        }
        for (int w = e0; w < e1 && PDP_TMA; ++w) {
            const int p = e0 + (int)g.vsrc[w];
            if ((int)g.vsrc2[w] + (e0 & ~3) != g.p_vpos[p]) atomicAdd(&errs[3], 1);
        }
        long long sum = 0;
        for (int t = v0; t < v1; ++t) {
            const int2 e = g.vsort[t];
            const int v = e.x, lo = e.y & 0xffff, deg = e.y >> 16;
            if (v < v0 || v >= v1 || lo != g.var_ptr[v] - e0 || deg != g.var_ptr[v + 1] - g.var_ptr[v]) atomicAdd(&errs[5], 1);
            if (t > v0 && deg > (g.vsort[t - 1].y >> 16) && g.V <= g.E) atomicAdd(&errs[5], 1);
            sum += v;
        }
        if (sum != ((long long)v0 + v1 - 1) * (v1 - v0) / 2) atomicAdd(&errs[5], 1);
    }
    GS(blk, g.ncb) {
        const int a0 = g.cb_ptr[blk], a1 = g.cb_ptr[blk + 1];
        const int e0 = g.cl_ptr[a0], e1 = g.cl_ptr[a1];
        if (e1 - e0 > PDP_BLK_C / g.ctas) atomicAdd(&errs[2], 1);
        for (int c = e0; c < e1; ++c) {
            const int x = cqpos(g, c);
            if (x < e0 || x >= e1 || (int)g.cinv[x] != c - e0) atomicAdd(&errs[2], 1);
        }
        for (int w = e0; w < e1; ++w) {
            const int c = e0 + (int)g.csrc[w];
            if (c < e0 || c >= e1 || cvpos(g, c) != g.cdst[w]) { atomicAdd(&errs[4], 1); continue; }
            if (w > e0 && g.cdst[w] <= g.cdst[w - 1]) atomicAdd(&errs[4], 1);
            if (!(atomicOr(&seen_c[c >> 5], 1u << (c & 31)) & (1u << (c & 31)))) atomicAdd(&errs[7], 1);
        }
        for (int c = e0; c < e1 && PDP_TMA; ++c) if ((int)g.cperm[c] + (e0 & ~3) != cqpos(g, c)) atomicAdd(&errs[2], 1);
        for (int w = e0; w < e1 && PDP_TMA; ++w) if ((int)g.csrc2[w] + (e0 & ~3) != cqpos(g, e0 + (int)g.csrc[w])) atomicAdd(&errs[4], 1);
        const int k = g.cb_k[blk];
        bool uni = (a1 > a0);
        for (int a = a0; a < a1; ++a) if (g.cl_ptr[a + 1] - g.cl_ptr[a] != g.cl_ptr[a0 + 1] - g.cl_ptr[a0]) uni = false;
        const int kk = uni ? g.cl_ptr[a0 + 1] - g.cl_ptr[a0] : 0;
        if (k != ((kk >= 1 && kk <= 8) ? kk : 0)) atomicAdd(&errs[5], 1);
    }
}
}  // namespace

// d_errs: device int32[8], zeroed here.  host_info (nullable): {blocked_ok, nvb, ncb, sv, sc}.
// Uses the eta[1] message buffer as scratch: call it before pdp_load_state.
extern "C" int pdp_debug_check_layout(pdp_ctx* c, int32_t* d_errs, int32_t* host_info, void* stream_) {
    if (!c || !d_errs) { pdp_set_error("pdp_debug_check_layout: null argument"); return PDP_ERR_ARG; }
    cudaStream_t stream = (cudaStream_t)stream_;
    const int nsm = c->num_sms;
    LCK(cudaMemsetAsync(d_errs, 0, 8 * sizeof(int32_t), stream));
    int64_t n = c->g.E;
    if (c->g.nvb > n) n = c->g.nvb;
    if (n > 0) {
        uint32_t* seen = reinterpret_cast<uint32_t*>(c->s.eta[1]);   // 2 * (E/32 + 1) words <= E floats for E >= 3
        if (c->g.E < 64) seen = reinterpret_cast<uint32_t*>(c->cub_tmp);
        LCK(cudaMemsetAsync(seen, 0, 2 * (size_t)(c->g.E / 32 + 1) * sizeof(uint32_t), stream));
        k_check_layout<<<G1(n)>>>(c->g, d_errs, seen);
        LLK();
    }
    if (host_info) { host_info[0] = c->g.blocked_ok + 16 * c->g.ctas; host_info[1] = c->g.nvb; host_info[2] = c->g.ncb; host_info[3] = c->g.sv; host_info[4] = c->g.sc; }
    return PDP_OK;
}
