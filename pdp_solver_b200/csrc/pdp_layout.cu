// pdp_layout.cu -- builds the blocked message layout of the SP sweep (pdp_common.cuh):
// block partitions of both edge orders, the V-layout / C-layout positions of every edge, the 16-bit
// node-order -> region-position tables of the two passes (variable blocks: transposed by warp groups, pdp_sweep.cuh) and
// the run-length coded destinations of their write-outs.  Runs once per batch inside pdp_create, entirely on the
// device (five stable radix sorts, three scans); the message arrays are not live yet and serve as scratch.
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
// block of a slot of the node-major order: the stride region the slot lies in, or an earlier one when the slot's node
// started before that region (block t begins at the first node whose first slot is >= t * stride, k_block_ptr).  Two
// loads from the small block tables instead of three dependent random gathers through the adjacency.
__device__ __forceinline__ int block_of_slot(const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ blk_ptr, int stride, int slot) {
    int b = slot / stride;
    while (b > 0 && slot < node_ptr[blk_ptr[b]]) --b;
    return b;
}

// x = V-layout position, lv[x] = clause-major slot stored there
// tslot[p] = local (transposed) slot of variable-major slot p inside its variable block
__global__ void k_fill_vlayout(pdp_graph g, const int32_t* __restrict__ lv, const uint16_t* __restrict__ tslot) {
    GS(x, g.E) {
        const int c = lv[x];
        const int p = g.c_pos[c];
        g.p_vpos[p] = (int32_t)x;
        const uint32_t cv = g.c_var[c];
        const int blk = g.var_ptr[cv & PDP_IDX_MASK] / g.sv;
        const int e0 = g.var_ptr[g.vb_ptr[blk]];          // the block's region starts at its first variable-major slot
        g.vfwd[(int64_t)g.vb_t0[blk] + tslot[p]] = (uint16_t)(((int)x - e0) | ((cv & PDP_SIGN_BIT) ? PDP_VINV_NEG : 0u));
    }
}
// keys of the C-layout sort, taken in V-layout order (variable block, clause-major slot): the clause block of the
// edge.  The stable sort leaves every clause block's edges ordered by (variable block, clause-major slot): the edges
// between one clause block and one variable block -- a RUN -- appear in the same order in both layouts.
__global__ void k_key_cblock_of_v(pdp_graph g, const int32_t* __restrict__ lv, int32_t* key, int32_t* val) {
    GS(x, g.E) {
        key[x] = block_of_slot(g.cl_ptr, g.cb_ptr, g.sc, lv[x]);
        val[x] = lv[x];
    }
}
// x = C-layout position, lq[x] = clause-major slot stored there, kj[x] = its clause block
__global__ void k_fill_clayout(pdp_graph g, const int32_t* __restrict__ kj, const int32_t* __restrict__ lq) {
    GS(x, g.E) {
        const int c = lq[x];
        g.p_qpos[g.c_pos[c]] = (int32_t)x;
        g.cfwd[c] = (uint16_t)((int)x - g.cl_ptr[g.cb_ptr[kj[x]]]);
    }
}
// Write-out order = load order: a pass writes the result of the edge it loaded from position x of its own layout to
// that edge's position in the other layout, which ascends with x inside a block.  dst[x], and the block of x.
__global__ void k_wo_seq_var(pdp_graph g, const int32_t* __restrict__ lv, int32_t* dst, int32_t* blk) {
    GS(x, g.E) {
        const int p = g.c_pos[lv[x]];
        dst[x] = g.p_qpos[p];
        blk[x] = block_of_slot(g.var_ptr, g.vb_ptr, g.sv, p);
    }
}
__global__ void k_wo_seq_clause(pdp_graph g, const int32_t* __restrict__ lq, const int32_t* __restrict__ kj, int32_t* dst, int32_t* blk) {
    GS(x, g.E) {
        dst[x] = g.p_vpos[g.c_pos[lq[x]]];
        blk[x] = kj[x];
    }
}

// run-length coding of the destinations: slot w starts a run unless it continues the previous slot's block and
// destination.  One warp per 32 slots (coalesced reads, the word is a ballot): bits[word], cnt[word] = popc(bits).
__global__ void k_wo_bits(int64_t E, const int32_t* __restrict__ kb, const int32_t* __restrict__ dst, int64_t nwords,
                          uint32_t* bits, int32_t* cnt) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t word = warp0; word < nwords; word += nwarps) {
        const int64_t w = word * 32 + lane;
        bool start = false;
        if (w < E) start = (w == 0) || kb[w] != kb[w - 1] || dst[w] != dst[w - 1] + 1;
        const uint32_t b = __ballot_sync(0xffffffffu, start);
        if (lane == 0) { bits[word] = b; cnt[word] = __popc(b); }
    }
}
// rank[word] = run starts before the word (exclusive scan of cnt): wrun[word] = {bits, rank - 1}, wadj[run] = dst - slot
__global__ void k_wo_pack(int64_t E, const int32_t* __restrict__ dst, int64_t nwords, const uint32_t* __restrict__ bits,
                          const int32_t* __restrict__ rank, uint2* wrun, int32_t* wadj) {
    GS(word, nwords) {
        uint32_t b = bits[word];
        int r = rank[word];
        wrun[word] = make_uint2(b, (uint32_t)(r - 1));
        while (b) {
            const int i = __ffs(b) - 1;
            b &= b - 1;
            const int64_t w = word * 32 + i;
            wadj[r++] = dst[w] - (int32_t)w;
        }
    }
}

// degree-sorted variable order: key = block << 14 | (16383 - degree)
__global__ void k_key_vsort(pdp_graph g, int32_t* key, int32_t* val) {
    GS(v, g.V) {
        const int deg = g.var_ptr[v + 1] - g.var_ptr[v];
        key[v] = ((g.var_ptr[v] / g.sv) << 14) | (16383 - (deg > 16383 ? 16383 : deg));
        val[v] = (int32_t)v;
    }
}
// pad[t] = slots of the group that starts at rank t (64 x the row PAIRS of its first, largest member), 0 for other ranks.
// Ranks are cut into groups of 32 from the first rank of every block.
__global__ void k_group_slots(pdp_graph g, const int32_t* __restrict__ order, int32_t* pad) {
    GS(t, g.V + 1) {
        if (t == g.V) { pad[t] = 0; continue; }      // (the scan then leaves the padded total in psum[V])
        const int v = order[t];
        const int t0 = g.vb_ptr[g.var_ptr[v] / g.sv];
        pad[t] = ((((int)t - t0) & 31) == 0) ? 64 * ((g.var_ptr[v + 1] - g.var_ptr[v] + 1) >> 1) : 0;   // rows go in pairs
    }
}
// psum = exclusive scan of pad: psum[first rank of a group] - psum[first rank of its block] is the group's base slot
__global__ void k_fill_vsort(pdp_graph g, const int32_t* __restrict__ order, const int32_t* __restrict__ psum, int32_t* max_slots) {
    int worst = 0;
    GS(t, g.V) {
        const int v = order[t];
        const int blk = g.var_ptr[v] / g.sv;
        const int t0 = g.vb_ptr[blk];
        const int tg = t0 + (((int)t - t0) & ~31);
        const int deg = g.var_ptr[v + 1] - g.var_ptr[v];
        const int base = psum[tg] - psum[t0];
        g.vsort[t] = make_int2(v, (base & 0xffff) | (deg << 16));
        if (tg == (int)t) worst = max(worst, base + 64 * ((deg + 1) >> 1));     // end of this group's slots = padded size so far
    }
    if (worst) atomicMax(max_slots, worst);
}
// t0 of a variable block in g.vfwd = padded slots of all earlier blocks
__global__ void k_block_t0(pdp_graph g, const int32_t* __restrict__ psum) {
    GS(blk, (int64_t)g.nvb + 1) g.vb_t0[blk] = psum[blk == g.nvb ? g.V : g.vb_ptr[blk]];
}
// padded transposed slots: lane l = rank within the group, row j = the j-th edge; the entries of rows 2k and 2k+1 of a lane sit
// side by side (one 32-bit load per row pair): slot = base + 64 k + 2 l + (j & 1)  (pdp_sweep.cuh var_group)
__global__ void k_fill_tslot(pdp_graph g, uint16_t* tslot) {
    GS(t, g.V) {
        const int2 e = g.vsort[t];
        const int v = e.x, base = e.y & 0xffff, deg = (int)((unsigned)e.y >> 16);
        const int lane = ((int)t - g.vb_ptr[g.var_ptr[v] / g.sv]) & 31;
        const int p0 = g.var_ptr[v];
        for (int j = 0; j < deg; ++j) tslot[p0 + j] = (uint16_t)(base + 64 * (j >> 1) + 2 * lane + (j & 1));
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

// block descriptors (after the run tables exist)
__device__ __forceinline__ int run_of(const uint2* __restrict__ wrun, int w) {
    const uint2 rb = wrun[w >> 5];
    return (int)rb.y + __popc(rb.x & (0xffffffffu >> (31 - (w & 31))));
}
__global__ void k_fill_desc(pdp_graph g) {
    GS(t, (int64_t)g.nvb + g.ncb) {
        const bool var_side = t < g.nvb;
        const int blk = var_side ? (int)t : (int)t - g.nvb;
        const int32_t* bp = var_side ? g.vb_ptr : g.cb_ptr;
        const int32_t* np = var_side ? g.var_ptr : g.cl_ptr;
        const int32_t* bm = var_side ? g.bvm : g.bfm;
        const uint2* wr = var_side ? g.v_wrun : g.c_wrun;
        pdp_blk d;
        d.n0 = bp[blk]; d.n1 = bp[blk + 1];
        d.e0 = 0; d.ne = 0; d.b0 = 0; d.b1 = 0; d.run0 = 0; d.nruns = 0; d.t0 = 0; d.tn = 0; d.pad_[0] = d.pad_[1] = 0;
        if (d.n1 > d.n0) {
            d.e0 = np[d.n0]; d.ne = np[d.n1] - d.e0;
            d.t0 = var_side ? g.vb_t0[blk] : d.e0;
            d.tn = var_side ? g.vb_t0[blk + 1] - g.vb_t0[blk] : d.ne;
            d.b0 = bm[d.n0]; d.b1 = bm[d.n1 - 1];
            if (d.ne > 0) { d.run0 = run_of(wr, d.e0); d.nruns = run_of(wr, d.e0 + d.ne - 1) - d.run0 + 1; }
        }
        (var_side ? g.vb_desc : g.cb_desc)[blk] = d;
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
    const char* forced = getenv("PDP_B200_CTAS");
    if (forced) g.ctas = (atoi(forced) == 2) ? 2 : 1;
    else if (g.B > 0 && E / g.B >= PDP_CTAS_EDGES_PER_PROBLEM) g.ctas = 2;
    const int blk_v = PDP_BLK_V / g.ctas, blk_c = PDP_BLK_C / g.ctas;
    // (V <= E: the degree sort below borrows E-sized scratch; batches of mostly isolated variables take the generic passes)
    const bool ok = monotone_maps && g.max_var_degree <= blk_v / 2 && g.max_clause_degree <= blk_c / 2 &&
                    g.V > 0 && g.F > 0 && g.V < E && getenv("PDP_B200_NO_BLOCKED") == nullptr;
    if (!ok) {
        k_identity_layout<<<G1(E)>>>(g);
        LLK();
        return PDP_OK;
    }
    g.sc = pick_stride(E, blk_c - 8, g.max_clause_degree, nsm * g.ctas);   // (8 words of alignment slack in the plane)
    g.ncb = (int32_t)(E / g.sc + 1);

    // scratch: the message arrays
    int32_t* S0 = reinterpret_cast<int32_t*>(c->s.eta[0]);
    int32_t* S1 = reinterpret_cast<int32_t*>(c->s.eta[1]);
    int32_t* S2 = reinterpret_cast<int32_t*>(c->s.qu);
    int32_t* S3 = reinterpret_cast<int32_t*>(c->s.qs);
    int32_t* LV = reinterpret_cast<int32_t*>(c->s.qd);    // V-layout position -> clause-major slot
    int32_t* LQ = reinterpret_cast<int32_t*>(c->s.ext);   // C-layout position -> variable-major slot
    uint16_t* TSLOT = reinterpret_cast<uint16_t*>(g.v_wadj);   // free until the variable side's write-out is coded (last step)

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
    auto exclusive_scan = [&](int32_t* data, int64_t count) -> int {   // in place
        size_t q = 0;
        if (cub::DeviceScan::ExclusiveSum(nullptr, q, data, data, (int)count, stream) != cudaSuccess) return PDP_ERR_CUDA;
        if (q > c->cub_tmp_bytes) { pdp_set_error("pdp_create: scan scratch %zu > reserve %zu", q, c->cub_tmp_bytes); return PDP_ERR_WORKSPACE; }
        if (cub::DeviceScan::ExclusiveSum(c->cub_tmp, q, data, data, (int)count, stream) != cudaSuccess) return PDP_ERR_CUDA;
        c->launches++;
        return PDP_OK;
    };
    int32_t *K, *L, *KF, *LF;
    int rc;

    // ---- variable blocks.  Per block: variables by descending degree, cut into groups of 32 (one warp each, nearly equal
    //      trip counts); a group's entries of the position table g.vfwd are stored transposed and padded to its largest
    //      degree (pdp_sweep.cuh; 2-5 % padding on random k-SAT).  The padded slots of a block are addressed with 16 bits
    //      (vsort.y): the edge budget of a block shrinks until the largest padded block stays below that.
    int32_t* d_max_slots = g.wo_tmp;
    int budget = blk_v;
    const int slot_limit = 65535 - 32;
    bool fits = false;
    for (int attempt = 0; attempt < 8 && !fits; ++attempt) {
        if (budget / 2 < g.max_var_degree) break;
        g.sv = pick_stride(E, budget, g.max_var_degree, nsm * g.ctas);
        g.nvb = (int32_t)(E / g.sv + 1);
        if (nsm > PDP_MAX_SMS || g.nvb > E / (PDP_BLK_V / 4) + 2 * PDP_MAX_SMS + 2 || g.ncb > E / (PDP_BLK_C / 4) + 2 * PDP_MAX_SMS + 2 ||
            g.nvb >= (1 << 17)) {
            pdp_set_error("pdp_create: block tables too small (nvb=%d ncb=%d)", g.nvb, g.ncb);
            return PDP_ERR_WORKSPACE;
        }
        k_block_ptr<<<G1(g.nvb + 1)>>>(g.var_ptr, g.V, g.sv, g.nvb, g.vb_ptr);
        LLK();
        k_key_vsort<<<G1(g.V)>>>(g, S0, S2);
        LLK();
        if ((rc = sort_pairs(g.V, (int64_t)g.nvb << 14, &K, &L, &KF, &LF)) != PDP_OK) return rc;
        k_group_slots<<<G1(g.V + 1)>>>(g, L, KF);
        LLK();
        if ((rc = exclusive_scan(KF, g.V + 1)) != PDP_OK) return rc;
        LCK(cudaMemsetAsync(d_max_slots, 0, sizeof(int32_t), stream));
        k_fill_vsort<<<G1(g.V)>>>(g, L, KF, d_max_slots);
        LLK();
        k_block_t0<<<G1(g.nvb + 1)>>>(g, KF);
        LLK();
        int32_t max_slots = 0, total_slots = 0;
        LCK(cudaMemcpyAsync(&max_slots, d_max_slots, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        LCK(cudaMemcpyAsync(&total_slots, KF + g.V, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        LCK(cudaStreamSynchronize(stream));
        if (total_slots > g.vfwd_cap) break;    // pads too badly for the position table: generic passes
        fits = max_slots <= slot_limit;
        if (!fits) budget = (int)((int64_t)budget * slot_limit / max_slots) - 64;   // scale by the overshoot, plus a margin
    }
    if (!fits) {   // degree distributions that do not pad well (a few huge variables among tiny ones): generic passes
        g.nvb = 0; g.ncb = 0; g.sv = 1; g.sc = 1;
        k_identity_layout<<<G1(E)>>>(g);
        LLK();
        return PDP_OK;
    }
    k_block_ptr<<<G1(g.ncb + 1)>>>(g.cl_ptr, g.F, g.sc, g.ncb, g.cb_ptr);
    LLK();
    k_fill_tslot<<<G1(g.V)>>>(g, TSLOT);
    LLK();

    // ---- V-layout: edges sorted by (variable block, clause-major slot)
    k_key_vblock_of_c<<<G1(E)>>>(g, S0, S2);
    LLK();
    if ((rc = sort_pairs(E, g.nvb, &K, &L, &KF, &LF)) != PDP_OK) return rc;
    k_fill_vlayout<<<G1(E)>>>(g, L, TSLOT);
    LLK();
    LCK(cudaMemcpyAsync(LV, L, sizeof(int32_t) * (size_t)E, cudaMemcpyDeviceToDevice, stream));
    // ---- C-layout: the V-layout sequence sorted (stable) by clause block = (clause block, variable block, clause-major slot)
    k_key_cblock_of_v<<<G1(E)>>>(g, LV, S0, S2);
    LLK();
    if ((rc = sort_pairs(E, g.ncb, &K, &L, &KF, &LF)) != PDP_OK) return rc;
    k_fill_clayout<<<G1(E)>>>(g, K, L);
    LLK();

    // ---- destinations of the write-outs (write-out order = load order), run-length coded.  K / L still hold the clause
    //      block and the clause-major slot of every C-layout position; KF / LF are free.
    const int64_t nwords = E / 32 + 2;
    uint32_t* wbits = reinterpret_cast<uint32_t*>(g.wo_tmp);
    int32_t* wcnt = g.wo_tmp + nwords;
    for (int side = 0; side < 2; ++side) {
        const bool var_side = (side == 1);     // clause side first: it reads K / L
        if (var_side) k_wo_seq_var<<<G1(E)>>>(g, LV, KF, LF);
        else k_wo_seq_clause<<<G1(E)>>>(g, L, K, KF, LF);
        LLK();
        k_wo_bits<<<G1(nwords * 32)>>>(E, LF, KF, nwords, wbits, wcnt);
        LLK();
        if ((rc = exclusive_scan(wcnt, nwords)) != PDP_OK) return rc;
        k_wo_pack<<<G1(nwords)>>>(E, KF, nwords, wbits, wcnt, var_side ? g.v_wrun : g.c_wrun, var_side ? g.v_wadj : g.c_wadj);
        LLK();
    }
    k_fill_desc<<<G1(g.nvb + g.ncb)>>>(g);
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
//   0 position maps inconsistent between the two edge orders     1 V-layout position outside its block / vfwd is not its position
//   2 C-layout position outside its block / cfwd wrong           3 variable write-out does not land on p_qpos
//   4 clause write-out does not land on c_vpos                   5 vsort / cb_k wrong
//   6 distinct write-out sources of the variable blocks (= E)    7 ... of the clause blocks (= E)
// ------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ int wo_dest(const uint2* __restrict__ wrun, const int32_t* __restrict__ wadj, int w) {
    const uint2 rb = wrun[w >> 5];
    return wadj[(int)rb.y + __popc(rb.x & (0xffffffffu >> (31 - (w & 31))))] + w;
}
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
        const int cap = 65535;
        if (e1 - e0 > PDP_BLK_V / g.ctas || g.vb_desc[blk].t0 != g.vb_t0[blk] || g.vb_desc[blk].e0 != e0 || g.vb_desc[blk].ne != e1 - e0) atomicAdd(&errs[1], 1);
        const uint16_t* fw = g.vfwd + g.vb_t0[blk];
        // vsort: a permutation of the block's variables by descending degree; groups of 32 ranks, rows of 32 slots
        long long sum = 0;
        int run_base = 0;
        for (int tg = v0; tg < v1; tg += 32) {
            const int gn = min(32, v1 - tg);
            const int base = g.vsort[tg].y & 0xffff;
            const int maxdeg = (int)((unsigned)g.vsort[tg].y >> 16);
            if (base != run_base || (base & 63) || base + 64 * ((maxdeg + 1) >> 1) > cap) atomicAdd(&errs[5], 1);
            for (int l = 0; l < gn; ++l) {
                const int2 e = g.vsort[tg + l];
                const int v = e.x, deg = (int)((unsigned)e.y >> 16);
                if (v < v0 || v >= v1 || deg != g.var_ptr[v + 1] - g.var_ptr[v] || (e.y & 0xffff) != base) { atomicAdd(&errs[5], 1); continue; }
                if (tg + l > v0 && deg > (int)((unsigned)g.vsort[tg + l - 1].y >> 16)) atomicAdd(&errs[5], 1);
                sum += v;
                for (int j = 0; j < deg; ++j) {
                    const int p = g.var_ptr[v] + j;
                    const int x = g.p_vpos[p];
                    const int slot = base + 64 * (j >> 1) + 2 * l + (j & 1);
                    if (x < e0 || x >= e1 || (int)(fw[slot] & 0x7fff) != x - e0 ||
                        ((fw[slot] & PDP_VINV_NEG) != 0) != ((g.v_cedge[p] & PDP_SIGN_BIT) != 0)) { atomicAdd(&errs[1], 1); continue; }
                    // write-out slot = load position: its destination must be this edge's C-layout position
                    if (wo_dest(g.v_wrun, g.v_wadj, x) != g.p_qpos[p]) { atomicAdd(&errs[3], 1); continue; }
                    if (!(atomicOr(&seen[x >> 5], 1u << (x & 31)) & (1u << (x & 31)))) atomicAdd(&errs[6], 1);
                }
            }
            run_base += 64 * ((maxdeg + 1) >> 1);
        }
        if (sum != ((long long)v0 + v1 - 1) * (v1 - v0) / 2) atomicAdd(&errs[5], 1);
        for (int w = e0 + 1; w < e1; ++w)
            if (wo_dest(g.v_wrun, g.v_wadj, w) <= wo_dest(g.v_wrun, g.v_wadj, w - 1)) atomicAdd(&errs[3], 1);
    }
    GS(blk, g.ncb) {
        const int a0 = g.cb_ptr[blk], a1 = g.cb_ptr[blk + 1];
        const int e0 = g.cl_ptr[a0], e1 = g.cl_ptr[a1];
        if (e1 - e0 > PDP_BLK_C / g.ctas - 8 || g.cb_desc[blk].t0 != e0 || g.cb_desc[blk].e0 != e0 || g.cb_desc[blk].ne != e1 - e0) atomicAdd(&errs[2], 1);
        for (int c = e0; c < e1; ++c) {
            const int x = cqpos(g, c);
            if (x < e0 || x >= e1 || (int)g.cfwd[c] != x - e0) { atomicAdd(&errs[2], 1); continue; }
            // write-out slot = region position: its destination must be this edge's V-layout position
            if (cvpos(g, c) != wo_dest(g.c_wrun, g.c_wadj, x)) { atomicAdd(&errs[4], 1); continue; }
            if (!(atomicOr(&seen_c[x >> 5], 1u << (x & 31)) & (1u << (x & 31)))) atomicAdd(&errs[7], 1);
        }
        for (int w = e0 + 1; w < e1; ++w)
            if (wo_dest(g.c_wrun, g.c_wadj, w) <= wo_dest(g.c_wrun, g.c_wadj, w - 1)) atomicAdd(&errs[4], 1);
        const int k = g.cb_k[blk];
        bool uni = (a1 > a0);
        for (int a = a0; a < a1; ++a) if (g.cl_ptr[a + 1] - g.cl_ptr[a] != g.cl_ptr[a0 + 1] - g.cl_ptr[a0]) uni = false;
        const int kk = uni ? g.cl_ptr[a0 + 1] - g.cl_ptr[a0] : 0;
        if (k != ((kk >= 1 && kk <= 8) ? kk : 0)) atomicAdd(&errs[5], 1);
    }
}
}  // namespace

// d_errs: device int32[8], zeroed here.  host_info (nullable): {blocked_ok, nvb, ncb, sv, sc}.
// Uses the eta message buffers as scratch: call it before pdp_load_state.
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
