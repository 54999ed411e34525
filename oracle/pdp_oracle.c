/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the SATYR message-passing hot path.
 *
 * A plain-C restatement of the reference's (microsoft/PDP-Solver) PyTorch implementation of
 *   - SurveyPropagator.forward            reference src/pdp/nn/pdp_propagate.py:139-221
 *   - SequentialDecimator.forward         reference src/pdp/nn/pdp_decimate.py:122-177
 *   - SurveyScorer.forward                reference src/pdp/nn/pdp_predict.py:155-192
 *   - IdentityPredictor.forward           reference src/pdp/nn/pdp_predict.py:118-128
 *   - sparse_smooth_max/max/argmax        reference src/pdp/nn/util.py:257-286
 *   - SatCNFEvaluator.forward             reference src/pdp/nn/util.py:210-236
 *   - SATProblem._peel/_set_variable_core/_propagate_single_clauses/simplify
 *                                         reference src/pdp/nn/solver.py:180-285
 *   - _forward_core/_update_solution/_local_search/_compute_energy(_diff)/_deduplicate
 *                                         reference src/pdp/nn/solver.py:355-496
 *   - _check_recurrence_termination       reference src/pdp/trainer.py:150-162
 *
 * Nothing here is shipped or measured as the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load this library.
 *
 * PARITY PIN: the reference ships no tests or golden vectors.  This oracle is
 * pinned against outputs of the reference itself, run in the authoring container through
 * oracle/compat.py, committed as tests/golden/ *.npz by oracle/make_golden.py.
 *
 * Arithmetic conventions reproduced from PyTorch-CPU: fp32 everywhere; every segmented sum
 * (torch.mm(sparse_COO, dense)) accumulates from 0 in ascending edge (storage) order; torch.max/min
 * against a constant propagate NaN; subnormals are kept (eps = 1e-40 is subnormal in fp32).
 *
 * `strict` = 1 reproduces the reference's accidental cross-problem coupling (global x.min() in
 * sparse_max/sparse_argmax, batch-global guards, NaN poisoning of the whole batch,
 * SURVEY.md section 8c).  `strict` = 0 gives every problem the semantics it would have alone in a
 * batch of one, which is what the CUDA path implements.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#define PAR_FOR _Pragma("omp parallel for schedule(static)")
#else
#define PAR_FOR
#endif

typedef struct {
    int64_t E, V, F, B;
    int strict;
    int32_t *evar, *ecls;   /* [E] endpoints of edge e (original edge order)            */
    float *esgn;            /* [E] literal sign +-1                                      */
    int32_t *bvm, *bfm;     /* [V], [F] problem ids                                      */
    int64_t *vptr, *cptr;   /* CSC / CSR pointers                                        */
    int32_t *vedge, *cedge; /* edges of each variable / clause, ascending edge index     */
    /* solver state (SATProblem) */
    float *av, *af, *sol, *is_sat;
    float *em;              /* sat_problem._edge_mask [E]; valid iff em_set               */
    int em_set;             /* _edge_mask is not None                                     */
    int em_in_state;        /* edge mask was appended to the decimator state              */
    /* message state: propagator state == decimator state after iteration 1 (p-d-p)      */
    float *q3, *fs2;        /* decimator state [E,3], [E,2]                               */
    float *pq3, *pfs2;      /* propagator state (only read for frozen problems)           */
    /* SequentialDecimator module state */
    float *prev;            /* _previous_function_state [E]                               */
    int has_prev;
    float *counters;        /* [B], valid iff has_counters                                */
    int has_counters;
    uint8_t *active;        /* active_mask [B]                                            */
    int64_t iters_done;
    /* trace of decimation events */
    int64_t trace_cap, trace_len;
    int64_t *trace;         /* triples (iteration, variable, sign)                        */
    /* scratch */
    float *tE0, *tE1, *tE2, *tV0, *tV1, *tV2, *tV3, *tF0, *tF1, *tF2, *tB0, *tB1;
} ora_t;

/* ---- torch.max(x, c) / torch.min(x, c): NaN propagates ------------------------------------ */
static inline float tmax(float x, float c) { return (x != x) ? x : (x > c ? x : c); }
static inline float tmin(float x, float c) { return (x != x) ? x : (x < c ? x : c); }
static inline float sgnf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : (x == 0.f ? 0.f : x)); }

#define EPS40 1e-40f
#define EPS10 1e-10f
/* math mode 0: libm logf/expf (default).  mode 1: fp32 results rounded from fp64 log/exp -- the
 * definition the CUDA library's strict-math TEST build uses, so trajectories can be compared bit for
 * bit (two correctly rounded implementations agree except ~1e-8 of the time). */
static int g_math_mode = 0;
void ora_set_math_mode(int m) { g_math_mode = m; }
static inline float o_logf(float x) { return g_math_mode ? (float)log((double)x) : logf(x); }
static inline float o_expf(float x) { return g_math_mode ? (float)exp((double)x) : expf(x); }
static inline float L40(float x) { return o_logf(tmax(x, EPS40)); }   /* pdp_propagate.py:133-134 */
static inline float L10(float x) { return o_logf(tmax(x, EPS10)); }   /* pdp_predict.py:149-150   */
static inline float X30(float x) { return o_expf(tmin(x, 30.0f)); }   /* pdp_propagate.py:136-137 */

static void build_adj(int64_t n_nodes, int64_t E, const int32_t* key, int64_t* ptr, int32_t* adj) {
    memset(ptr, 0, sizeof(int64_t) * (size_t)(n_nodes + 1));
    for (int64_t e = 0; e < E; ++e) ptr[key[e] + 1]++;
    for (int64_t i = 0; i < n_nodes; ++i) ptr[i + 1] += ptr[i];
    int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_nodes + 1));
    memcpy(cur, ptr, sizeof(int64_t) * (size_t)(n_nodes + 1));
    for (int64_t e = 0; e < E; ++e) adj[cur[key[e]]++] = (int32_t)e;   /* stable: ascending e */
    free(cur);
}

#define ALLOC(T, n) ((T*)calloc((size_t)((n) > 0 ? (n) : 1), sizeof(T)))

ora_t* ora_create(int64_t E, int64_t V, int64_t F, int64_t B, const int32_t* graph_map,
                  const float* sign, const int32_t* bvm, const int32_t* bfm, int strict) {
    ora_t* o = ALLOC(ora_t, 1);
    o->E = E; o->V = V; o->F = F; o->B = B; o->strict = strict;
    o->evar = ALLOC(int32_t, E); o->ecls = ALLOC(int32_t, E); o->esgn = ALLOC(float, E);
    memcpy(o->evar, graph_map, sizeof(int32_t) * (size_t)E);
    memcpy(o->ecls, graph_map + E, sizeof(int32_t) * (size_t)E);
    memcpy(o->esgn, sign, sizeof(float) * (size_t)E);
    o->bvm = ALLOC(int32_t, V); o->bfm = ALLOC(int32_t, F);
    memcpy(o->bvm, bvm, sizeof(int32_t) * (size_t)V);
    memcpy(o->bfm, bfm, sizeof(int32_t) * (size_t)F);
    o->vptr = ALLOC(int64_t, V + 1); o->cptr = ALLOC(int64_t, F + 1);
    o->vedge = ALLOC(int32_t, E); o->cedge = ALLOC(int32_t, E);
    build_adj(V, E, o->evar, o->vptr, o->vedge);
    build_adj(F, E, o->ecls, o->cptr, o->cedge);
    /* solver.py:49-54 */
    o->av = ALLOC(float, V); o->af = ALLOC(float, F); o->sol = ALLOC(float, V); o->is_sat = ALLOC(float, B);
    for (int64_t i = 0; i < V; ++i) { o->av[i] = 1.f; o->sol[i] = 0.5f; }
    for (int64_t a = 0; a < F; ++a) o->af[a] = 1.f;
    for (int64_t b = 0; b < B; ++b) o->is_sat[b] = 0.5f;
    o->em = ALLOC(float, E); o->em_set = 0; o->em_in_state = 0;
    o->q3 = ALLOC(float, 3 * E); o->fs2 = ALLOC(float, 2 * E);
    o->pq3 = ALLOC(float, 3 * E); o->pfs2 = ALLOC(float, 2 * E);
    o->prev = ALLOC(float, E); o->has_prev = 0;
    o->counters = ALLOC(float, B); o->has_counters = 0;
    o->active = ALLOC(uint8_t, B);
    for (int64_t b = 0; b < B; ++b) o->active[b] = 1;   /* solver.py:363 */
    o->trace_cap = 1024; o->trace_len = 0; o->trace = ALLOC(int64_t, 3 * o->trace_cap);
    o->tE0 = ALLOC(float, E); o->tE1 = ALLOC(float, E); o->tE2 = ALLOC(float, E);
    o->tV0 = ALLOC(float, V); o->tV1 = ALLOC(float, V); o->tV2 = ALLOC(float, V); o->tV3 = ALLOC(float, V);
    o->tF0 = ALLOC(float, F); o->tF1 = ALLOC(float, F); o->tF2 = ALLOC(float, F);
    o->tB0 = ALLOC(float, B); o->tB1 = ALLOC(float, B);
    return o;
}

void ora_destroy(ora_t* o) {
    if (!o) return;
    free(o->evar); free(o->ecls); free(o->esgn); free(o->bvm); free(o->bfm); free(o->vptr); free(o->cptr);
    free(o->vedge); free(o->cedge); free(o->av); free(o->af); free(o->sol); free(o->is_sat); free(o->em);
    free(o->q3); free(o->fs2); free(o->pq3); free(o->pfs2); free(o->prev); free(o->counters); free(o->active);
    free(o->trace); free(o->tE0); free(o->tE1); free(o->tE2); free(o->tV0); free(o->tV1); free(o->tV2);
    free(o->tV3); free(o->tF0); free(o->tF1); free(o->tF2); free(o->tB0); free(o->tB1);
    free(o);
}

/* ------------------------------------------------------------------------------------------ */
/* accessors (copy in / out)                                                                    */
/* ------------------------------------------------------------------------------------------ */
void ora_set_state(ora_t* o, const float* prop_q3, const float* prop_fs2, const float* dec_q3, const float* dec_fs2) {
    memcpy(o->pq3, prop_q3, sizeof(float) * 3 * (size_t)o->E);
    memcpy(o->pfs2, prop_fs2, sizeof(float) * 2 * (size_t)o->E);
    memcpy(o->q3, dec_q3, sizeof(float) * 3 * (size_t)o->E);
    memcpy(o->fs2, dec_fs2, sizeof(float) * 2 * (size_t)o->E);
}
void ora_set_masks(ora_t* o, const float* av, const float* af, const float* sol) {
    if (av) memcpy(o->av, av, sizeof(float) * (size_t)o->V);
    if (af) memcpy(o->af, af, sizeof(float) * (size_t)o->F);
    if (sol) memcpy(o->sol, sol, sizeof(float) * (size_t)o->V);
}
void ora_get_state(const ora_t* o, float* q3, float* fs2) {
    if (q3) memcpy(q3, o->q3, sizeof(float) * 3 * (size_t)o->E);
    if (fs2) memcpy(fs2, o->fs2, sizeof(float) * 2 * (size_t)o->E);
}
void ora_get_masks(const ora_t* o, float* av, float* af, float* sol, float* is_sat, uint8_t* active, float* em) {
    if (av) memcpy(av, o->av, sizeof(float) * (size_t)o->V);
    if (af) memcpy(af, o->af, sizeof(float) * (size_t)o->F);
    if (sol) memcpy(sol, o->sol, sizeof(float) * (size_t)o->V);
    if (is_sat) memcpy(is_sat, o->is_sat, sizeof(float) * (size_t)o->B);
    if (active) memcpy(active, o->active, (size_t)o->B);
    if (em) memcpy(em, o->em, sizeof(float) * (size_t)o->E);
}
void ora_get_counters(const ora_t* o, float* counters) { memcpy(counters, o->counters, sizeof(float) * (size_t)o->B); }
int64_t ora_iters_done(const ora_t* o) { return o->iters_done; }
int64_t ora_trace_len(const ora_t* o) { return o->trace_len; }
void ora_get_trace(const ora_t* o, int64_t* out) { memcpy(out, o->trace, sizeof(int64_t) * 3 * (size_t)o->trace_len); }
int ora_edge_mask_set(const ora_t* o) { return o->em_set; }

static void trace_push(ora_t* o, int64_t it, int64_t var, int64_t sg) {
    if (o->trace_len == o->trace_cap) {
        o->trace_cap *= 2;
        o->trace = (int64_t*)realloc(o->trace, sizeof(int64_t) * 3 * (size_t)o->trace_cap);
    }
    o->trace[3 * o->trace_len + 0] = it; o->trace[3 * o->trace_len + 1] = var; o->trace[3 * o->trace_len + 2] = sg;
    o->trace_len++;
}

/* ------------------------------------------------------------------------------------------ */
/* SATProblem: solver.py:180-285                                                                */
/* ------------------------------------------------------------------------------------------ */

/* solver.py:205-226.  `asg` [V] is modified in place (assignment *= active_variables). */
static void set_variable_core(ora_t* o, float* asg) {
    const int64_t V = o->V, F = o->F;
    PAR_FOR
    for (int64_t i = 0; i < V; ++i) asg[i] *= o->av[i];
    PAR_FOR
    for (int64_t a = 0; a < F; ++a) {
        float input_num = 0.f, feval = 0.f;
        for (int64_t p = o->cptr[a]; p < o->cptr[a + 1]; ++p) {
            int32_t e = o->cedge[p];
            float x = asg[o->evar[e]];
            input_num += fabsf(x);
            feval += o->esgn[e] * x;
        }
        if (((feval > -input_num) ? 1.f : 0.f) * o->af[a] == 1.f) o->af[a] = 0.f;
    }
    PAR_FOR
    for (int64_t i = 0; i < V; ++i)
        if (fabsf(asg[i]) == 1.f) { o->av[i] = 0.f; o->sol[i] = (asg[i] + 1.f) / 2.0f; }
}

/* solver.py:228-273 */
static void propagate_single_clauses(ora_t* o) {
    const int64_t V = o->V, F = o->F, B = o->B;
    float* single = o->tF0; float* input_num = o->tV0; float* veval = o->tV1; float* asg = o->tV2;
    float* unsat_ex = o->tB0;
    for (;;) {
        double n_single = 0;
        for (int64_t a = 0; a < F; ++a) {
            float deg = 0.f;
            for (int64_t p = o->cptr[a]; p < o->cptr[a + 1]; ++p) deg += o->av[o->evar[o->cedge[p]]];
            single[a] = ((deg == 1.f) ? 1.f : 0.f) * o->af[a];
            n_single += single[a];
        }
        if (n_single <= 0) break;
        double n_conf = 0;
        for (int64_t i = 0; i < V; ++i) {
            float in = 0.f, ev = 0.f;
            for (int64_t p = o->vptr[i]; p < o->vptr[i + 1]; ++p) {
                int32_t e = o->vedge[p];
                in += single[o->ecls[e]];
                ev += o->esgn[e] * single[o->ecls[e]];
            }
            input_num[i] = in; veval[i] = ev;
            n_conf += ((fabsf(ev) != in) ? 1.f : 0.f) * o->av[i];
        }
        if (n_conf > 0) {
            for (int64_t b = 0; b < B; ++b) unsat_ex[b] = 0.f;
            for (int64_t i = 0; i < V; ++i)
                unsat_ex[o->bvm[i]] += ((fabsf(veval[i]) != input_num[i]) ? 1.f : 0.f) * o->av[i];
            for (int64_t b = 0; b < B; ++b) if (unsat_ex[b] >= 1.f) o->is_sat[b] = 0.f;
            /* quirk (solver.py:257,261): the `== 1` test is on the per-problem conflict COUNT */
            for (int64_t a = 0; a < F; ++a) if (unsat_ex[o->bfm[a]] * o->af[a] == 1.f) o->af[a] = 0.f;
            for (int64_t i = 0; i < V; ++i) if (unsat_ex[o->bvm[i]] * o->av[i] == 1.f) o->av[i] = 0.f;
        }
        for (int64_t i = 0; i < V; ++i) {
            float assigned = ((fabsf(veval[i]) == input_num[i]) ? 1.f : 0.f) * o->av[i];
            asg[i] = sgnf(veval[i]) * assigned;
        }
        for (int64_t a = 0; a < F; ++a) if (single[a] == 1.f) o->af[a] = 0.f;
        set_variable_core(o, asg);
    }
}

/* solver.py:180-203 */
static void peel(ora_t* o) {
    const int64_t V = o->V, F = o->F;
    float* deg = o->tV0; float* sdeg = o->tV1; float* sv = o->tV2; float* sf = o->tF0;
    PAR_FOR
    for (int64_t i = 0; i < V; ++i) {
        float d = 0.f, s = 0.f;
        for (int64_t p = o->vptr[i]; p < o->vptr[i + 1]; ++p) {
            int32_t e = o->vedge[p];
            d += o->af[o->ecls[e]];
            s += o->esgn[e] * o->af[o->ecls[e]];
        }
        deg[i] = d; sdeg[i] = s;
    }
    for (;;) {
        double n = 0;
        for (int64_t i = 0; i < V; ++i) { sv[i] = ((deg[i] == fabsf(sdeg[i])) ? 1.f : 0.f) * o->av[i]; n += sv[i]; }
        if (n <= 0) break;
        PAR_FOR
        for (int64_t a = 0; a < F; ++a) {
            float c = 0.f;
            for (int64_t p = o->cptr[a]; p < o->cptr[a + 1]; ++p) c += sv[o->evar[o->cedge[p]]];
            sf[a] = ((c > 0.f) ? 1.f : 0.f) * o->af[a];
        }
        PAR_FOR
        for (int64_t i = 0; i < V; ++i) {
            float dd = 0.f, ds = 0.f;
            for (int64_t p = o->vptr[i]; p < o->vptr[i + 1]; ++p) {
                int32_t e = o->vedge[p];
                dd += sf[o->ecls[e]];
                ds += o->esgn[e] * sf[o->ecls[e]];
            }
            if (sv[i] == 1.f) o->sol[i] = (sgnf(sdeg[i]) + 1.f) / 2.0f;
            deg[i] -= dd * o->av[i];
            sdeg[i] -= ds * o->av[i];
        }
        for (int64_t i = 0; i < V; ++i) if (sv[i] == 1.f) o->av[i] = 0.f;
        for (int64_t a = 0; a < F; ++a) if (sf[a] == 1.f) o->af[a] = 0.f;
    }
}

/* solver.py:281-285 */
void ora_simplify(ora_t* o) { propagate_single_clauses(o); peel(o); }

/* solver.py:275-279.  `assignment` [V] float in {-1,0,1}. */
void ora_set_variables(ora_t* o, const float* assignment) {
    memcpy(o->tV3, assignment, sizeof(float) * (size_t)o->V);
    set_variable_core(o, o->tV3);
    ora_simplify(o);
}

/* ------------------------------------------------------------------------------------------ */
/* SurveyPropagator.forward, adaptors off: pdp_propagate.py:139-221                             */
/* ------------------------------------------------------------------------------------------ */
/* Stateless form.  dec_q3/dec_fs2: decimator state (inputs of the update); edge_mask: NULL or [E];
 * prop_q3/prop_fs2: propagator state (only read through the frozen-problem blend); active: NULL or
 * uint8 [B].  Outputs out_q3 [E,3], out_fs2 [E,2]. */
static void sp_step_core(const ora_t* o, const float* dec_q3, const float* dec_fs2, const float* edge_mask,
                         const float* prop_q3, const float* prop_fs2, const uint8_t* active, float pi,
                         float* out_q3, float* out_fs2, float* x, float* y, float* ftot, float* pos, float* neg) {
    const int64_t E = o->E, V = o->V, F = o->F;
    /* functions --> variables half: pdp_propagate.py:166-175 */
    PAR_FOR
    for (int64_t e = 0; e < E; ++e) {
        float v = L40(dec_q3[3 * e + 0]);
        if (edge_mask) v = v * edge_mask[e];
        x[e] = v;
        float w = L40(1.f - dec_fs2[2 * e + 0]);   /* :185 */
        if (edge_mask) w = w * edge_mask[e];
        y[e] = w;
    }
    PAR_FOR
    for (int64_t a = 0; a < F; ++a) {
        float s = 0.f;
        for (int64_t p = o->cptr[a]; p < o->cptr[a + 1]; ++p) s += x[o->cedge[p]];
        ftot[a] = s;
    }
    /* variables half sums: pdp_propagate.py:190-193 (pos/neg masks hold explicit zeros) */
    PAR_FOR
    for (int64_t i = 0; i < V; ++i) {
        float ps = 0.f, ns = 0.f;
        for (int64_t p = o->vptr[i]; p < o->vptr[i + 1]; ++p) {
            int32_t e = o->vedge[p];
            float pm = (o->esgn[e] == 1.f) ? 1.f : 0.f, nm = (o->esgn[e] == -1.f) ? 1.f : 0.f;
            ps += pm * y[e];
            ns += nm * y[e];
        }
        pos[i] = ps; neg[i] = ns;
    }
    PAR_FOR
    for (int64_t e = 0; e < E; ++e) {
        const float mask = active ? (float)active[o->bvm[o->evar[e]]] : 1.f;   /* :146-151 */
        const float s = o->esgn[e];
        /* :173-175 */
        float agg = ftot[o->ecls[e]] - x[e];
        float fstate = mask * X30(agg) + (1.f - mask) * prop_fs2[2 * e + 0];
        /* :184-218 */
        float ext = dec_fs2[2 * e + 1];
        float P = pos[o->evar[e]], N = neg[o->evar[e]];
        float same = 0.5f * (1.f + s) * P + 0.5f * (1.f - s) * N;
        same = same - y[e];
        same += L40(1.0f - pi * ((ext == s) ? 1.f : 0.f));
        float opp = 0.5f * (1.f - s) * P + 0.5f * (1.f + s) * N;
        opp += L40(1.0f - pi * ((ext == -s) ? 1.f : 0.f));
        float dc = same + opp;
        dc = X30(dc);
        float S = X30(same), O = X30(opp);
        float qu = S * (1.f - O), qs = O * (1.f - S);
        float total = qu + qs + dc;
        out_q3[3 * e + 0] = mask * (qu / total) + (1.f - mask) * prop_q3[3 * e + 0];
        out_q3[3 * e + 1] = mask * (qs / total) + (1.f - mask) * prop_q3[3 * e + 1];
        out_q3[3 * e + 2] = mask * (dc / total) + (1.f - mask) * prop_q3[3 * e + 2];
        out_fs2[2 * e + 0] = fstate;
        out_fs2[2 * e + 1] = ext;
    }
}

void ora_sp_step(const ora_t* o, const float* dec_q3, const float* dec_fs2, const float* edge_mask,
                 const float* prop_q3, const float* prop_fs2, const uint8_t* active, float pi,
                 float* out_q3, float* out_fs2) {
    float* x = ALLOC(float, o->E); float* y = ALLOC(float, o->E); float* ft = ALLOC(float, o->F);
    float* ps = ALLOC(float, o->V); float* ns = ALLOC(float, o->V);
    sp_step_core(o, dec_q3, dec_fs2, edge_mask, prop_q3, prop_fs2, active, pi, out_q3, out_fs2, x, y, ft, ps, ns);
    free(x); free(y); free(ft); free(ps); free(ns);
}

/* ------------------------------------------------------------------------------------------ */
/* SurveyScorer.forward, adaptors off: pdp_predict.py:155-192                                   */
/* ------------------------------------------------------------------------------------------ */
void ora_score(const ora_t* o, const float* fs2, const float* af, float pi, float* score) {
    const int64_t V = o->V;
    PAR_FOR
    for (int64_t i = 0; i < V; ++i) {
        float extsum = 0.f, ps = 0.f, ns = 0.f, as = 0.f;
        for (int64_t p = o->vptr[i]; p < o->vptr[i + 1]; ++p) {
            int32_t e = o->vedge[p];
            extsum += fs2[2 * e + 1];
            float f = L10(1.f - fs2[2 * e + 0]) * af[o->ecls[e]];
            float pm = (o->esgn[e] == 1.f) ? 1.f : 0.f, nm = (o->esgn[e] == -1.f) ? 1.f : 0.f;
            ps += pm * f; ns += nm * f; as += f;
        }
        float ext = sgnf(extsum);
        float pos = ps + L10(1.0f - pi * ((ext == 1.f) ? 1.f : 0.f));
        float neg = ns + L10(1.0f - pi * ((ext == -1.f) ? 1.f : 0.f));
        float pn = pos + neg;
        float dc = as + L10(1.0f - pi);
        float bias = (2.f * pn + dc) / 4.0f;
        pos = pos - bias; neg = neg - bias; pn = pn - bias;
        dc = X30(dc - bias);
        float q0 = X30(pos) - X30(pn);
        float q1 = X30(neg) - X30(pn);
        float total = L10(q0 + q1 + dc);
        score[i] = X30(L10(q1) - total) - X30(L10(q0) - total);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* util.py:257-286                                                                              */
/* ------------------------------------------------------------------------------------------ */
/* sparse_smooth_max(x[E], variable_mask) -> out[V] */
static void smooth_max_var(const ora_t* o, const float* x, float* out) {
    const int64_t V = o->V;
    PAR_FOR
    for (int64_t i = 0; i < V; ++i) {
        float num = 0.f, den = 0.f;
        for (int64_t p = o->vptr[i]; p < o->vptr[i + 1]; ++p) {
            float v = x[o->vedge[p]];
            float c = X30(30.f * v);
            num += v * c; den += c;
        }
        out[i] = num / tmax(den, 1.0f);
    }
}

/* torch.min over a vector with NaN propagation */
static float vmin_range(const float* x, const int32_t* key, int64_t n, int32_t want, int all) {
    float m = INFINITY; int seen = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!all && key[i] != want) continue;
        float v = x[i];
        if (v != v) return v;
        if (!seen || v < m) { m = v; seen = 1; }
    }
    return seen ? m : 0.f;
}

/* sparse_max(x[V], batch_variable_mask) -> out[B]  (util.py:267-275) */
static void sparse_max_batch(const ora_t* o, const float* x, float* out) {
    const int64_t V = o->V, B = o->B;
    float gmin = o->strict ? vmin_range(x, NULL, V, 0, 1) : 0.f;
    float* colmax = (float*)malloc(sizeof(float) * (size_t)(B > 0 ? B : 1));
    float* mins = (float*)malloc(sizeof(float) * (size_t)(B > 0 ? B : 1));
    int* nanf_ = (int*)calloc((size_t)(B > 0 ? B : 1), sizeof(int));
    if (!o->strict) {
        int* seen = (int*)calloc((size_t)(B > 0 ? B : 1), sizeof(int));
        for (int64_t b = 0; b < B; ++b) mins[b] = 0.f;
        for (int64_t i = 0; i < V; ++i) {
            int32_t b = o->bvm[i]; float v = x[i];
            if (v != v) { nanf_[b] = 1; continue; }
            if (!seen[b] || v < mins[b]) { mins[b] = v; seen[b] = 1; }
        }
        for (int64_t b = 0; b < B; ++b) if (nanf_[b]) mins[b] = NAN;
        free(seen);
    } else {
        for (int64_t b = 0; b < B; ++b) mins[b] = gmin;
    }
    for (int64_t b = 0; b < B; ++b) { colmax[b] = 0.f; nanf_[b] = 0; }   /* dense zeros off-block */
    for (int64_t i = 0; i < V; ++i) {
        int32_t b = o->bvm[i];
        float d = (x[i] - mins[b]) + 1.f;
        if (d != d) nanf_[b] = 1;
        else if (d > colmax[b]) colmax[b] = d;
    }
    for (int64_t b = 0; b < B; ++b) {
        float c = nanf_[b] ? NAN : colmax[b];
        out[b] = (c + mins[b]) - 1.f;
    }
    free(colmax); free(mins); free(nanf_);
}

/* sparse_argmax(x[V], batch_variable_mask) -> out[B] (first maximal index; util.py:257-265).
 * Problems without variables get index 0 like torch.argmax over an all-zero column. */
static void sparse_argmax_batch(const ora_t* o, const float* x, int64_t* out) {
    const int64_t V = o->V, B = o->B;
    float* mins = (float*)malloc(sizeof(float) * (size_t)(B > 0 ? B : 1));
    float* best = (float*)malloc(sizeof(float) * (size_t)(B > 0 ? B : 1));
    if (o->strict) {
        float g = vmin_range(x, NULL, V, 0, 1);
        for (int64_t b = 0; b < B; ++b) mins[b] = g;
    } else {
        int* seen = (int*)calloc((size_t)(B > 0 ? B : 1), sizeof(int));
        int* nn = (int*)calloc((size_t)(B > 0 ? B : 1), sizeof(int));
        for (int64_t b = 0; b < B; ++b) mins[b] = 0.f;
        for (int64_t i = 0; i < V; ++i) {
            int32_t b = o->bvm[i]; float v = x[i];
            if (v != v) { nn[b] = 1; continue; }
            if (!seen[b] || v < mins[b]) { mins[b] = v; seen[b] = 1; }
        }
        for (int64_t b = 0; b < B; ++b) if (nn[b]) mins[b] = NAN;
        free(seen); free(nn);
    }
    for (int64_t b = 0; b < B; ++b) { out[b] = 0; best[b] = 0.f; }
    /* column b of the dense matrix: zeros for rows outside the problem.  torch.argmax treats NaN as
     * the maximum and returns its first occurrence. */
    int* isnan_best = (int*)calloc((size_t)(B > 0 ? B : 1), sizeof(int));
    int64_t* first_zero_row_checked = NULL; (void)first_zero_row_checked;
    for (int64_t i = 0; i < V; ++i) {
        int32_t b = o->bvm[i];
        if (isnan_best[b]) continue;
        float d = (x[i] - mins[b]) + 1.f;
        if (d != d) { isnan_best[b] = 1; out[b] = i; continue; }
        /* rows before the first member of b are zeros: index 0 wins only if d <= 0 everywhere */
        if (d > best[b]) { best[b] = d; out[b] = i; }
    }
    free(mins); free(best); free(isnan_best);
}

/* ------------------------------------------------------------------------------------------ */
/* SatCNFEvaluator.forward: util.py:210-236                                                     */
/* ------------------------------------------------------------------------------------------ */
void ora_cnf_eval(const ora_t* o, const float* pred, float* solved, float* n_unsat) {
    const int64_t F = o->F, B = o->B;
    float* max_sat = ALLOC(float, B); float* bval = ALLOC(float, B);
    for (int64_t a = 0; a < F; ++a) {
        float cv = 0.f;
        for (int64_t p = o->cptr[a]; p < o->cptr[a + 1]; ++p) {
            int32_t e = o->cedge[p];
            float ev = o->esgn[e] * pred[o->evar[e]];
            ev = ev + (1.f - o->esgn[e]) / 2.f;
            cv += (ev > 0.5f) ? 1.f : 0.f;
        }
        max_sat[o->bfm[a]] += 1.f;
        bval[o->bfm[a]] += (cv > 0.f) ? 1.f : 0.f;
    }
    for (int64_t b = 0; b < B; ++b) {
        solved[b] = (max_sat[b] == bval[b]) ? 1.f : 0.f;
        n_unsat[b] = max_sat[b] - bval[b];
    }
    free(max_sat); free(bval);
}

/* ------------------------------------------------------------------------------------------ */
/* _compute_energy / _compute_energy_diff: solver.py:469-496                                    */
/* ------------------------------------------------------------------------------------------ */
void ora_energy(const ora_t* o, const float* a_in, const float* av, const float* af, float* energy, float* unsat_fn) {
    const int64_t F = o->F, B = o->B;
    for (int64_t b = 0; b < B; ++b) energy[b] = 0.f;
    PAR_FOR
    for (int64_t a = 0; a < F; ++a) {
        float agg = 0.f, deg = 0.f;
        for (int64_t p = o->cptr[a]; p < o->cptr[a + 1]; ++p) {
            int32_t e = o->cedge[p]; int32_t i = o->evar[e];
            agg += o->esgn[e] * (a_in[i] * av[i]);
            deg += av[i];
        }
        unsat_fn[a] = ((agg == -deg) ? 1.f : 0.f) * af[a];
    }
    for (int64_t a = 0; a < F; ++a) energy[o->bfm[a]] += unsat_fn[a];
}

void ora_energy_diff(const ora_t* o, const float* a_in, const float* av, const float* edge_mask, float* delta) {
    const int64_t V = o->V, F = o->F;
    float* agg = ALLOC(float, F); float* deg = ALLOC(float, F);
    PAR_FOR
    for (int64_t a = 0; a < F; ++a) {
        float g = 0.f, d = 0.f;
        for (int64_t p = o->cptr[a]; p < o->cptr[a + 1]; ++p) {
            int32_t e = o->cedge[p]; int32_t i = o->evar[e];
            g += o->esgn[e] * (a_in[i] * av[i]);
            d += av[i];
        }
        agg[a] = g; deg[a] = d;
    }
    PAR_FOR
    for (int64_t i = 0; i < V; ++i) {
        float s = 0.f;
        for (int64_t p = o->vptr[i]; p < o->vptr[i + 1]; ++p) {
            int32_t e = o->vedge[p];
            float dist = o->esgn[e] * (a_in[i] * av[i]);
            float others = agg[o->ecls[e]] - dist;
            float crit = ((others == (1.f - deg[o->ecls[e]])) ? 1.f : 0.f) * edge_mask[e];
            s += crit * dist;
        }
        delta[i] = s;
    }
    free(agg); free(deg);
}

static void compute_edge_mask(ora_t* o) {   /* solver.py:370-371 / 439-440 */
    PAR_FOR
    for (int64_t e = 0; e < o->E; ++e) o->em[e] = o->av[o->evar[e]] * o->af[o->ecls[e]];
    o->em_set = 1;
}
void ora_compute_edge_mask(ora_t* o) { compute_edge_mask(o); }

/* ------------------------------------------------------------------------------------------ */
/* SequentialDecimator.forward: pdp_decimate.py:122-177                                         */
/* ------------------------------------------------------------------------------------------ */
static void decimate(ora_t* o, float tol, float t_max, int has_active_mask, float pi) {
    const int64_t E = o->E, V = o->V, B = o->B;
    float* eta = o->tE0;                 /* message_state[1][:, 0] */
    for (int64_t e = 0; e < E; ++e) eta[e] = o->fs2[2 * e];
    if (!o->has_counters) { for (int64_t b = 0; b < B; ++b) o->counters[b] = 0.f; o->has_counters = 1; }

    if (has_active_mask) {               /* :127-133 */
        float* sv = o->tV0; float* mb = o->tB0;
        smooth_max_var(o, eta, sv);
        for (int64_t i = 0; i < V; ++i) sv[i] = sv[i] * o->av[i];
        sparse_max_batch(o, sv, mb);
        for (int64_t b = 0; b < B; ++b) if (mb[b] <= 1e-10f) o->active[b] = 0;
    }

    /* :135 -- `_active_variables.sum() > 0` is batch-global in the reference */
    double nav = 0; for (int64_t i = 0; i < V; ++i) nav += o->av[i];
    float* navb = o->tB1;
    for (int64_t b = 0; b < B; ++b) navb[b] = 0.f;
    for (int64_t i = 0; i < V; ++i) navb[o->bvm[i]] += o->av[i];

    if (o->has_prev && (o->strict ? (nav > 0) : 1)) {
        float* d = o->tE1; float* sd = o->tV0; float* db = o->tB0;
        for (int64_t e = 0; e < E; ++e) {
            d[e] = fabsf(o->prev[e] - eta[e]);
            if (o->em_set) d[e] = d[e] * o->em[e];          /* :138-139 */
        }
        smooth_max_var(o, d, sd);
        for (int64_t i = 0; i < V; ++i) sd[i] = sd[i] * o->av[i];
        sparse_max_batch(o, sd, db);
        float* conv = o->tB0;   /* reuse: conv[b] written after db[b] is consumed */
        int* in_block = (int*)malloc(sizeof(int) * (size_t)(B > 0 ? B : 1));
        for (int64_t b = 0; b < B; ++b) {
            in_block[b] = o->strict ? 1 : (navb[b] > 0.f);
            if (!in_block[b]) { conv[b] = 0.f; continue; }
            float v = db[b];
            if (v < tol) o->counters[b] = 0.f;                        /* :145 */
            float c = (v < tol) ? 1.f : 0.f;                          /* :146 */
            if (o->counters[b] >= t_max) { c = 1.f; o->counters[b] = 0.f; }   /* :147-148 */
            conv[b] = c;
        }
        float* convv = o->tV1;
        double nconv = 0;
        for (int64_t i = 0; i < V; ++i) { convv[i] = conv[o->bvm[i]]; nconv += convv[i]; }   /* :150 */
        if (nconv > 0) {                                              /* :152 */
            float* score = o->tV2; float* coeff = o->tV3;
            ora_score(o, o->fs2, o->af, pi, score);
            for (int64_t i = 0; i < V; ++i) coeff[i] = fabsf(score[i]) * o->av[i] * convv[i];
            float* norm = o->tB1;
            for (int64_t b = 0; b < B; ++b) norm[b] = 0.f;
            for (int64_t i = 0; i < V; ++i) norm[o->bvm[i]] += coeff[i];      /* :160 */
            float csum = 0.f; for (int64_t i = 0; i < V; ++i) csum += coeff[i];
            if (o->strict ? (csum > 0.f) : 1) {                       /* :158 */
                int64_t* mi = (int64_t*)malloc(sizeof(int64_t) * (size_t)(B > 0 ? B : 1));
                sparse_argmax_batch(o, coeff, mi);
                float* asg = o->tV0;
                for (int64_t i = 0; i < V; ++i) asg[i] = 0.f;
                int any = 0;
                for (int64_t b = 0; b < B; ++b) {
                    int sel;
                    if (o->strict) sel = (has_active_mask ? o->active[b] : 1) && (norm[b] != 0.f);
                    else           sel = (has_active_mask ? o->active[b] : 1) && in_block[b] && (norm[b] > 0.f);
                    if (!sel) continue;
                    asg[mi[b]] = sgnf(score[mi[b]]);                  /* :168-169 */
                    trace_push(o, o->iters_done, mi[b], (int64_t)sgnf(score[mi[b]]));
                    any = 1;
                }
                if (any) ora_set_variables(o, asg);                   /* :171 */
                free(mi);
            }
        }
        for (int64_t b = 0; b < B; ++b) if (in_block[b]) o->counters[b] = o->counters[b] + 1.f;   /* :173 */
        free(in_block);
    }
    for (int64_t e = 0; e < E; ++e) o->prev[e] = eta[e];              /* :175 */
    o->has_prev = 1;
}

/* ------------------------------------------------------------------------------------------ */
/* one trip of the _forward_core loop: solver.py:365-384 + trainer.py:150-162                   */
/* ------------------------------------------------------------------------------------------ */
/* returns the number of still-active problems (or B when there is no termination check) */
int64_t ora_iterate(ora_t* o, float tol, float t_max, int check_termination, int batch_replication, float pi) {
    const int64_t E = o->E, B = o->B;
    o->iters_done++;
    float* nq = ALLOC(float, 3 * E); float* nf = ALLOC(float, 2 * E);
    float* ft = o->tF1; float* ps = ALLOC(float, o->V); float* ns = ALLOC(float, o->V);
    sp_step_core(o, o->q3, o->fs2, o->em_in_state ? o->em : NULL, o->pq3, o->pfs2,
                 check_termination ? o->active : NULL, pi, nq, nf, o->tE1, o->tE2, ft, ps, ns);
    /* p-d-p: the decimator returns the propagator state unchanged, so both states are the new one */
    memcpy(o->q3, nq, sizeof(float) * 3 * (size_t)E); memcpy(o->fs2, nf, sizeof(float) * 2 * (size_t)E);
    memcpy(o->pq3, nq, sizeof(float) * 3 * (size_t)E); memcpy(o->pfs2, nf, sizeof(float) * 2 * (size_t)E);
    free(nq); free(nf); free(ps); free(ns);

    decimate(o, tol, t_max, check_termination, pi);

    compute_edge_mask(o);                                             /* solver.py:370-371 */
    double s = 0; for (int64_t e = 0; e < E; ++e) s += o->em[e];
    o->em_in_state = (s < (double)E);                                 /* solver.py:373-374 */

    if (!check_termination) return B;
    /* predictor (IdentityPredictor) + _update_solution leave _solution unchanged; trainer.py:150-162 */
    float* solved = ALLOC(float, B); float* nun = ALLOC(float, B);
    ora_cnf_eval(o, o->sol, solved, nun);
    if (batch_replication > 1) {
        const int64_t B0 = B / batch_replication;
        for (int64_t j = 0; j < B0; ++j) {
            float any = 0.f;
            for (int r = 0; r < batch_replication; ++r) any += (solved[r * B0 + j] > 0.5f) ? 1.f : 0.f;
            for (int r = 0; r < batch_replication; ++r)
                if (o->active[r * B0 + j]) o->active[r * B0 + j] = (any == 0.f) ? 1 : 0;
        }
    } else {
        for (int64_t b = 0; b < B; ++b) if (o->active[b]) o->active[b] = (solved[b] <= 0.5f) ? 1 : 0;
    }
    free(solved); free(nun);
    int64_t na = 0; for (int64_t b = 0; b < B; ++b) na += o->active[b];
    return na;
}

/* solver.py:355-386: runs up to T iterations, returns the number executed */
int64_t ora_run(ora_t* o, int64_t T, float tol, float t_max, int check_termination, int batch_replication, float pi) {
    int64_t t = 0;
    for (; t < T; ++t) {
        int64_t na = ora_iterate(o, tol, t_max, check_termination, batch_replication, pi);
        if (check_termination && na <= 0) { ++t; break; }
    }
    return t;
}

/* ------------------------------------------------------------------------------------------ */
/* IdentityPredictor last call (pdp_predict.py:121-126): draws are consumed by ACTIVE variables   */
/* in batch-global variable order; `draws` holds at least n_active values.                        */
/* ------------------------------------------------------------------------------------------ */
int64_t ora_count_active_variables(const ora_t* o) {
    int64_t n = 0; for (int64_t i = 0; i < o->V; ++i) n += (o->av[i] > 0.f); return n;
}
void ora_random_fill(ora_t* o, const float* draws) {
    int64_t k = 0;
    for (int64_t i = 0; i < o->V; ++i) if (o->av[i] > 0.f) o->sol[i] = draws[k++];
}

/* ------------------------------------------------------------------------------------------ */
/* _local_search (WalkSAT): solver.py:433-467, followed by _update_solution: solver.py:388-399    */
/* rand_var [W,V] and rand_coin [W,B] are the torch.rand draws of iterations 0..W-1 in the        */
/* reference's order.  Writes the merged prediction [V]; returns the iterations executed.         */
/* ------------------------------------------------------------------------------------------ */
int64_t ora_local_search(ora_t* o, int64_t W, float epsilon, int batch_replication,
                         const float* rand_var, const float* rand_coin, float* prediction) {
    const int64_t V = o->V, F = o->F, B = o->B;
    float* a = ALLOC(float, V);
    for (int64_t i = 0; i < V; ++i) a[i] = o->av[i] * (2.f * ((o->sol[i] > 0.5f) ? 1.f : 0.f) - 1.0f);   /* :436-437 */
    compute_edge_mask(o);                                             /* :439-440 */
    float* energy = ALLOC(float, B); float* unsat_fn = ALLOC(float, F); float* delta = ALLOC(float, V);
    float* x = ALLOC(float, V); int64_t* gi = ALLOC(int64_t, B); int64_t* ri = ALLOC(int64_t, B);
    int64_t it = 0;
    for (; it < W; ++it) {
        ora_energy(o, a, o->av, o->af, energy, unsat_fn);             /* :443-444 */
        if (batch_replication > 1) {                                  /* :446-449 */
            const int64_t B0 = B / batch_replication; double s = 0;
            for (int64_t j = 0; j < B0; ++j) {
                float sat_reps = 0.f;
                for (int r = 0; r < batch_replication; ++r) sat_reps += 1.f - ((energy[r * B0 + j] > 0.f) ? 1.f : 0.f);
                s += 1.f - ((sat_reps > 0.f) ? 1.f : 0.f);
            }
            if (s == 0) break;
        } else {
            double s = 0; for (int64_t b = 0; b < B; ++b) s += (energy[b] > 0.f) ? 1.f : 0.f;
            if (s == 0) break;                                        /* :450-451 */
        }
        ora_energy_diff(o, a, o->av, o->em, delta);                   /* :453 */
        for (int64_t i = 0; i < V; ++i) x[i] = -delta[i];
        sparse_argmax_batch(o, x, gi);                                /* :454 */
        for (int64_t i = 0; i < V; ++i) {                             /* :456-457 */
            float uv = 0.f;
            for (int64_t p = o->vptr[i]; p < o->vptr[i + 1]; ++p) uv += unsat_fn[o->ecls[o->vedge[p]]];
            uv = uv * o->av[i];
            x[i] = ((uv > 0.f) ? 1.f : 0.f) * rand_var[it * V + i];
        }
        sparse_argmax_batch(o, x, ri);                                /* :458 */
        for (int64_t b = 0; b < B; ++b) {                             /* :460-465 */
            if (!(energy[b] > 0.f)) continue;
            int64_t coin = (rand_coin[it * B + b] > epsilon) ? 1 : 0;
            int64_t ind = coin * gi[b] + (1 - coin) * ri[b];
            a[ind] = -a[ind];
        }
    }
    for (int64_t i = 0; i < V; ++i) {                                 /* :467 + :388-399 */
        float walk = (a[i] + 1.f) / 2.0f;
        float merged = o->av[i] * walk + (1.0f - o->av[i]) * o->sol[i];
        if (o->av[i] == 1.f) o->sol[i] = merged;
        prediction[i] = merged;
    }
    free(a); free(energy); free(unsat_fn); free(delta); free(x); free(gi); free(ri);
    return it;
}

/* ------------------------------------------------------------------------------------------ */
/* _deduplicate: solver.py:401-431 (replica r of problem j has problem id r*B0+j, its variables  */
/* are the r-th block of V/b variables).  Writes the winning replica's prediction [V/b].         */
/* ------------------------------------------------------------------------------------------ */
void ora_deduplicate(ora_t* o, int batch_replication, const float* prediction, float* out_pred, int64_t* winner) {
    const int64_t V = o->V, F = o->F, B = o->B, B0 = B / batch_replication, V0 = V / batch_replication;
    float* asg = ALLOC(float, V); float* energy = ALLOC(float, B); float* uf = ALLOC(float, F);
    for (int64_t i = 0; i < V; ++i) asg[i] = 2.f * prediction[i] - 1.0f;       /* :407 */
    ora_energy(o, asg, o->av, o->af, energy, uf);                               /* :408 */
    for (int64_t j = 0; j < B0; ++j) {                                          /* :409 argmax(-energy), first */
        int64_t best = 0; float be = energy[j];
        for (int r = 1; r < batch_replication; ++r) if (energy[r * B0 + j] < be) { be = energy[r * B0 + j]; best = r; }
        winner[j] = best;
    }
    for (int64_t i = 0; i < V0; ++i) {                                          /* :411-415 */
        int64_t j = o->bvm[i];
        out_pred[i] = prediction[winner[j] * V0 + i];
    }
    free(asg); free(energy); free(uf);
}

int ora_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* pins the OpenMP team size (torchrun exports OMP_NUM_THREADS=1 into every rank) */
void ora_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
