/*
 * mtfjsp_oracle.c -- CPU restatement of the reference MT-FJSP disjunctive-graph environment.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA path in
 * e2e-mappo-for-mt-fjsp_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against
 *   (1) the reference's shipped result CSVs (results/test_results/Real_{MK,PT,TT,IT}_J6_M6_E2_Seed3_Weight442.csv,
 *       rows FIFO/LWKR/MWKR x SPT/SEC, 100 test instances) and
 *   (2) per-step dumps of the reference itself (trainer/parallel_env.py Parallel_env replayed by
 *       oracle/ref_harness.py in the build container; fixtures + generator under tests/golden/).
 *
 * It follows the reference literally, with the networkx graph replaced by the arc set it encodes.
 * Reference locations ("SS" = graph-jsp-env/src/graph_jsp_env/disjunctive_graph_jsp_env_singlestep.py):
 *   step                      SS:716-974              _schedule_task            SS:1476-1685
 *   _append_at_the_end        SS:1689-1775            _insert_at_index_0        SS:1777-1809
 *   job-arc refresh           SS:1356-1434            estimator                 SS:1920-1999
 *   _state_array (obs)        SS:2001-2515            wrk_reward_function       SS:1051-1171
 *   arrival / transport / idle   trainer/DGenv_func.py:46-66, 108-128, 144-170
 *   reward scaling            algorithm/ppo_trick.py:54-122
 *   candidate-machine feats   trainer/parallel_env.py:152-214
 *   job mask / candidates     algorithm/ppo_algorithm.py:202-317
 *
 * Arc set.  The reference keeps a DiGraph; every quantity it reads from it is a function of
 *   - the machine routes (ordered op lists),  - per-op machine/duration/start/finish,
 *   - two transients that live for exactly one observation:
 *       removed_head : op v whose job arc (v-1 -> v) was deleted by an in-between insertion
 *                      (SS:1660) and is re-created by the next step's job-arc refresh (SS:1502);
 *       fresh_co     : op a whose machine arc (a-1 -> a) coincides with its job arc; it carries
 *                      the machine-arc weight (SS:1764 / SS:1644) until the next refresh overwrites it.
 * In-arcs of op v = {source if first of job} U {v-1 unless v == removed_head} U {route predecessor}.
 *
 * Floating point: FP64 throughout, same operand order as the reference; np.sum is restated as
 * numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, 8 accumulators, block 128).
 * Compile WITHOUT -ffast-math and with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    /* running reward scaler, algorithm/ppo_trick.py:54-122 (shape = 4) */
    double R[4], mean[4], S[4], std[4];
    long n;
} scaler_t;

typedef struct {
    int *mach;           /* [N] -1 = unassigned                              SS:579,1496 */
    double *dur;         /* [N] 0 until assigned                              SS:580,1498 */
    double *st, *ft;     /* [N]                                               SS:1621-1622 */
    uint8_t *sched;      /* [N]                                               SS:1623 */
    int *route;          /* [M][N] op ids in processing order                 SS:484 */
    int *rlen;           /* [M] */
    int *rpos;           /* [N] index of op inside its route (derived, kept for speed) */
    int removed_head, fresh_co;
    int nsched, last_op, last_m;
    double mk_prev, e_prev, trans, trans_prev, idle, idle_prev;   /* SS:215-250 */
    double *mfea;        /* [M][8]                                            SS:427, 2315-2354 */
    double w[3];         /* per-env reward weights                            SS:1253-1270 */
    double *est_st, *est_ft, *est_pt; /* [N] last estimator output            SS:1920-1999 */
    scaler_t sc;
} env_t;

typedef struct {
    int B, J, M, N, E, left_shift;
    double cfg_w[3], divisor, gamma;
    double *t, *p, *tt;  /* [B][N][M], [B][N][M], [B][M][M] */
    int *edge_id;        /* [B][M] 1-based edge group of each machine        parallel_env.py:212 */
    double *mind, *minpt;/* [B][N] min feasible t / min feasible t*|p|        SS:1932-1950 */
    env_t *env;
    int nthreads;
} oracle_t;

/* ---- numpy pairwise sum, loops_utils.h.src ---- */
static double np_pairwise_sum(const double *a, long n) {
    if (n < 8) {
        double res = 0.;
        for (long i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8], res;
        long i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; k++) r[k] += a[i + k];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}
double oracle_np_sum(const double *a, long n) { return np_pairwise_sum(a, n); }

/* ---- helpers on one env ---- */
#define T_(o, b, i, m) ((o)->t[((size_t)(b) * (o)->N + (i)) * (o)->M + (m)])
#define P_(o, b, i, m) ((o)->p[((size_t)(b) * (o)->N + (i)) * (o)->M + (m)])
#define TT_(o, b, k, l) ((o)->tt[((size_t)(b) * (o)->M + (k)) * (o)->M + (l)])

/* find_transportT, DGenv_func.py:108-128: tt[m(u)][m(v)] iff both assigned and same job */
static double tr(const oracle_t *o, int b, const env_t *e, int u, int v) {
    if (e->mach[v] < 0 || e->mach[u] < 0) return 0.0;
    if (u / o->M != v / o->M) return 0.0;
    return TT_(o, b, e->mach[u], e->mach[v]);
}

static int route_pred(const oracle_t *o, const env_t *e, int v) {
    int m = e->mach[v];
    if (m < 0 || !e->sched[v]) return -1;
    int k = e->rpos[v];
    return k > 0 ? e->route[m * o->N + k - 1] : -1;
}

/* find_max_arrivaTime_for_currentNode, DGenv_func.py:46-66 */
static double arrival(const oracle_t *o, int b, const env_t *e, int v) {
    double best = -INFINITY;
    int first = (v % o->M) == 0;
    if (first) best = 0.0; /* source arc: ft 0 + 0 */
    if (!first && v != e->removed_head) {
        double val = e->ft[v - 1] + tr(o, b, e, v - 1, v);
        if (val > best) best = val;
    }
    int rp = route_pred(o, e, v);
    if (rp >= 0) {
        double val = e->ft[rp] + tr(o, b, e, rp, v);
        if (val > best) best = val;
    }
    return best;
}

static void route_insert(const oracle_t *o, env_t *e, int m, int k, int a) {
    int *r = e->route + m * o->N;
    for (int i = e->rlen[m]; i > k; i--) {
        r[i] = r[i - 1];
        e->rpos[r[i]] = i;
    }
    r[k] = a;
    e->rpos[a] = k;
    e->rlen[m]++;
}

/* estiamte_st_ft_pt_eachStep_noTransT, SS:1920-1999 */
static void estimate(const oracle_t *o, int b, env_t *e) {
    int M = o->M, N = o->N;
    const double *mind = o->mind + (size_t)b * N, *minpt = o->minpt + (size_t)b * N;
    for (int i = 0; i < N; i++) {
        double bft = e->sched[i] ? e->ft[i] : 0.0; /* current_ft * if_schedule */
        if (bft == 0) {
            if (i % M != 0) e->est_ft[i] = e->est_ft[i - 1] + mind[i];
            else e->est_ft[i] = 0 + mind[i];
        } else e->est_ft[i] = bft;
    }
    for (int i = 0; i < N; i++) {
        if (!e->sched[i]) {
            e->est_st[i] = (i % M == 0) ? 0.0 : e->est_ft[i - 1];
            e->est_pt[i] = minpt[i];
        } else {
            e->est_st[i] = e->st[i];
            e->est_pt[i] = T_(o, b, i, e->mach[i]) * P_(o, b, i, e->mach[i]); /* SS:356, 2175 */
        }
    }
}

/* calculate_idle_t_for_each_machine, DGenv_func.py:144-170 (standby power p2 == 1, SS:371) */
static double idle_total(const oracle_t *o, const env_t *e) {
    double sum_idle = 0;
    for (int m = 0; m < o->M; m++) {
        const int *r = e->route + m * o->N;
        int len = e->rlen[m];
        if (len >= 1) {
            double blank = e->st[r[0]] - 0;
            blank = blank * 1.0;
            sum_idle = sum_idle + blank;
            for (int i = 0; i < len - 1; i++) {
                blank = e->st[r[i + 1]] - e->ft[r[i]];
                blank = blank * 1.0;
                sum_idle = sum_idle + blank;
            }
        }
    }
    return sum_idle;
}

static void scaler_call(scaler_t *s, double gamma, const double x[4], double out[4]) {
    /* RewardScaling.__call__ + RunningMeanStd.update, ppo_trick.py:73-88,115-119 */
    for (int k = 0; k < 4; k++) s->R[k] = gamma * s->R[k] + x[k];
    s->n += 1;
    if (s->n == 1) {
        for (int k = 0; k < 4; k++) { s->mean[k] = s->R[k]; s->std[k] = fabs(s->R[k]); }
    } else {
        for (int k = 0; k < 4; k++) {
            double old = s->mean[k];
            s->mean[k] = old + (s->R[k] - old) / (double)s->n;
            s->S[k] = s->S[k] + (s->R[k] - old) * (s->R[k] - s->mean[k]);
            s->std[k] = sqrt(s->S[k] / (double)s->n);
        }
    }
    for (int k = 0; k < 4; k++) out[k] = x[k] / (s->std[k] + 1e-8);
}

/* one transition; returns 0 ok, 1 invalid action (state untouched: documented deviation, the
 * reference corrupts its state instead, SS:1495-1528) */
static int env_step(const oracle_t *o, int b, env_t *e, int a, int m, double r5[5], double sc4[4], uint8_t *done) {
    int M = o->M, N = o->N;
    if (a < 0 || a >= N || m < 0 || m >= M || e->sched[a] || ((a % M) != 0 && !e->sched[a - 1]) ||
        T_(o, b, a, m) < 0 || e->nsched >= N) {
        for (int k = 0; k < 5; k++) r5[k] = 0.0;
        for (int k = 0; k < 4; k++) sc4[k] = 0.0;
        *done = (e->nsched == N);
        return 1;
    }
    double d = T_(o, b, a, m); /* SS:741 */
    e->mach[a] = m;            /* SS:1496-1498 */
    e->dur[a] = d;
    /* SS:1502 job-arc refresh: re-creates a removed job arc, overwrites a coincident machine arc */
    e->removed_head = -1;
    e->fresh_co = -1;
    int *r = e->route + m * N;
    int len = e->rlen[m];
    double st;
    int where; /* insertion index */
    if (len == 0) { /* SS:1684-1685 -> _insert_at_index_0 */
        st = arrival(o, b, e, a);
        where = 0;
    } else {
        where = -1;
        if (o->left_shift) {
            double lb = arrival(o, b, e, a); /* SS:1538 */
            double lbft = lb + d;
            if (lbft <= arrival(o, b, e, r[0])) { /* SS:1548 */
                st = lb;
                where = 0;
            } else if (len > 1) { /* SS:1588-1675 */
                for (int k = 0; k + 1 < len; k++) {
                    int prev = r[k], next = r[k + 1];
                    double nst = arrival(o, b, e, next);
                    if (lbft > nst) continue;
                    double gap = nst - e->ft[prev];
                    if (gap < d) continue;
                    double x = arrival(o, b, e, a);
                    double y = e->ft[prev] + tr(o, b, e, prev, a);
                    st = (y > x) ? y : x; /* python max(x, y): returns x unless y > x */
                    where = k + 1;
                    if (next == prev + 1 && (next % M) != 0) e->removed_head = next; /* SS:1660 */
                    break;
                }
            }
        }
        if (where < 0) { /* _append_at_the_end, SS:1689-1775 */
            int last = r[len - 1];
            double x = arrival(o, b, e, a);
            double y = e->ft[last] + tr(o, b, e, last, a);
            st = (y > x) ? y : x;
            where = len;
        }
    }
    route_insert(o, e, m, where, a);
    e->st[a] = st;
    e->ft[a] = st + d;
    e->sched[a] = 1;
    e->nsched++;
    e->last_op = a;
    e->last_m = m;
    if (where > 0 && r[where - 1] == a - 1 && (a % M) != 0) e->fresh_co = a;

    *done = (e->nsched == N); /* SS:797-800 */
    e->idle = idle_total(o, e); /* SS:857 */
    double nt = ((a % M) == 0) ? 0.0 : tr(o, b, e, a - 1, a); /* SS:872-877 */
    e->trans += nt;

    /* _state_array side effects, SS:2315-2338 (obs itself is built by env_obs) */
    estimate(o, b, e);
    double mk = -INFINITY;
    for (int i = 0; i < N; i++) if (e->est_ft[i] > mk) mk = e->est_ft[i]; /* SS:894 */
    double en = np_pairwise_sum(e->est_pt, N);                              /* SS:896 */
    double *mf = e->mfea + m * 8;
    mf[0] = e->ft[r[e->rlen[m] - 1]];
    mf[1] += (P_(o, b, a, m) * T_(o, b, a, m)) / (double)N;
    mf[2] += nt;
    mf[3] += e->idle - e->idle_prev;
    mf[4] += 1;

    /* wrk_reward_function, SS:1066-1132 */
    double r_t = 1.0 * e->mk_prev - mk;
    double r_pt = 1.0 * e->e_prev - en;
    r_pt = r_pt / (double)N;
    double r_tt = 1.0 * e->trans_prev - e->trans;
    double r_idle = 1.0 * e->idle_prev - e->idle;
    double total = o->cfg_w[0] * r_t + o->cfg_w[1] * (r_pt + 1 * r_idle) + o->cfg_w[2] * r_tt * 1;
    r5[0] = total / o->divisor;
    r5[1] = r_t; r5[2] = r_idle; r5[3] = r_pt; r5[4] = r_tt;
    e->mk_prev = mk; e->e_prev = en; e->trans_prev = e->trans; e->idle_prev = e->idle; /* SS:932-936 */
    if (*done) { e->trans = 0; e->idle = 0; } /* SS:956-958; *_prev survive */

    double x4[4] = {r_t, r_idle, r_pt, r_tt}; /* parallel_env.py:255-256 */
    scaler_call(&e->sc, o->gamma, x4, sc4);
    return 0;
}

/* weight of arc (u -> v), or a negative number if the arc is not in the graph */
static double machine_arc_w(const oracle_t *o, int b, const env_t *e, int u, int v) {
    double blank = e->st[v] - e->ft[u];
    return e->dur[u] + tr(o, b, e, u, v) + blank; /* SS:1573, 1644, 1657, 1764 */
}
static double job_arc_w(const oracle_t *o, int b, const env_t *e, int v) {
    int u = v - 1;
    if (e->fresh_co == v) return machine_arc_w(o, b, e, u, v);
    if (e->dur[u] != 0) return e->dur[u] + tr(o, b, e, u, v); /* SS:1392-1422 */
    return 1.0;                                                 /* SS:625, 642 */
}
/* adjacency value, SS:2019, 2050-2064; returns 0 when the arc vanishes under int truncation */
static double adj_val(const env_t *e, int u, double w) {
    long long wi = (long long)w; /* astype(int) */
    if (wi == 0) return 0.0;
    double nd = (e->mach[u] < 0) ? 1.0 : e->dur[u];
    long long v = (long long)((double)wi - nd);
    return (double)(v + 1);
}

static void env_obs(const oracle_t *o, int b, const env_t *e, double *tfea, double *mfea, int32_t *ell_idx,
                    double *ell_w, uint8_t *job_mask, int32_t *cand, int mask_mode) {
    int M = o->M, N = o->N, J = o->J;
    for (int v = 0; v < N; v++) {
        int first = (v % M) == 0;
        int rp = route_pred(o, e, v);
        int has_job = (!first && v != e->removed_head);
        int indeg = (first ? 1 : 0) + (has_job ? 1 : 0) + ((rp >= 0 && !(has_job && rp == v - 1)) ? 1 : 0);
        if (tfea) { /* SS:2246-2277 */
            double *f = tfea + (size_t)v * 12;
            f[0] = e->est_st[v]; f[1] = e->est_ft[v]; f[2] = e->est_pt[v];
            f[3] = e->sched[v] ? 1.0 : 0.0;
            f[4] = (double)indeg;
            if (e->sched[v]) {
                f[5] = (double)(e->mach[v] + 1);
                f[6] = T_(o, b, v, e->mach[v]);
                f[7] = P_(o, b, v, e->mach[v]);
            } else { f[5] = 0; f[6] = 0; f[7] = 0; }
            f[8] = (double)(v / M + 1);
            f[9] = e->w[0]; f[10] = e->w[1]; f[11] = e->w[2];
        }
        if (ell_idx) { /* slot 0 self, slot 1 job predecessor, slot 2 machine predecessor */
            int32_t *ix = ell_idx + (size_t)v * 3;
            double *wv = ell_w + (size_t)v * 3;
            ix[0] = v; wv[0] = 1.0;
            ix[1] = -1; wv[1] = 0.0; ix[2] = -1; wv[2] = 0.0;
            if (has_job) {
                double val = adj_val(e, v - 1, job_arc_w(o, b, e, v));
                if (val != 0.0) { ix[1] = v - 1; wv[1] = val; }
            }
            if (rp >= 0 && !(has_job && rp == v - 1)) {
                double val = adj_val(e, rp, machine_arc_w(o, b, e, rp, v));
                if (val != 0.0) { ix[2] = rp; wv[2] = val; }
            }
        }
    }
    if (mfea) memcpy(mfea, e->mfea, sizeof(double) * M * 8);
    if (job_mask) { /* ppo_algorithm.py:202-317; SURVEY Appendix C */
        int any_first_missing = 0, any_unfinished = 0;
        for (int j = 0; j < J; j++) {
            int nx = 0;
            while (nx < M && e->sched[j * M + nx]) nx++;
            cand[j] = j * M + (nx < M - 1 ? nx : M - 1);
            if (nx == 0) any_first_missing = 1;
            if (nx < M) any_unfinished = 1;
        }
        for (int j = 0; j < J; j++) {
            int nx = 0;
            while (nx < M && e->sched[j * M + nx]) nx++;
            job_mask[j] = (nx == M);
        }
        if (mask_mode == 1) {
            if (any_first_missing) {
                for (int j = 0; j < J; j++) job_mask[j] = e->sched[j * M];
            } else if (any_unfinished) {
                double mn = INFINITY;
                double v[J];
                for (int j = 0; j < J; j++) {
                    double rm = 0; /* max(row) over ft with unscheduled = 0 */
                    for (int c = 0; c < M; c++) if (e->sched[j * M + c] && e->ft[j * M + c] > rm) rm = e->ft[j * M + c];
                    v[j] = job_mask[j] ? INFINITY : rm;
                    if (v[j] < mn) mn = v[j];
                }
                for (int j = 0; j < J; j++) job_mask[j] = (v[j] != mn);
            }
        }
    }
}

/* cal_cur_task_machine_feature, parallel_env.py:152-214 */
static void env_mfea1(const oracle_t *o, int b, const env_t *e, int a, double *out, uint8_t *mmask) {
    int M = o->M;
    double tv[M], ptv[M], pv[M];
    int nt = 0, npt = 0, np_ = 0;
    for (int m = 0; m < M; m++) {
        double t = T_(o, b, a, m), p = P_(o, b, a, m), pt = t * fabs(p);
        if (t > 0) tv[nt++] = t;
        if (pt > 0) ptv[npt++] = pt;
        if (p > 0) pv[np_++] = p;
    }
    double mean_t = np_pairwise_sum(tv, nt) / (double)nt;
    double mean_pt = np_pairwise_sum(ptv, npt) / (double)npt;
    double mean_p = np_pairwise_sum(pv, np_) / (double)np_;
    for (int m = 0; m < M; m++) {
        double t = T_(o, b, a, m), p = P_(o, b, a, m), pt = t * fabs(p);
        double *f = out + m * 6;
        f[0] = t > 0 ? t : mean_t;
        f[1] = pt > 0 ? pt : mean_pt;
        if (a % M == 0) f[2] = 0;
        else {
            /* int(tfea[a-1][5]) - 1: machine of the job predecessor; -1 wraps to the last row */
            int pm = e->sched[a - 1] ? e->mach[a - 1] : M - 1;
            f[2] = TT_(o, b, pm, m);
        }
        int infeasible = !(t >= 0); /* Run.py:262-266 */
        f[3] = (double)(1 - infeasible);
        f[4] = p > 0 ? p : mean_p;
        f[5] = (double)o->edge_id[b * M + m];
        if (mmask) mmask[m] = (uint8_t)infeasible;
    }
}

/* ------------------------------------------------------------------ batch API (ctypes) */
oracle_t *oracle_create(int B, int J, int M, int E, int left_shift) {
    oracle_t *o = (oracle_t *)calloc(1, sizeof(oracle_t));
    int N = J * M;
    o->B = B; o->J = J; o->M = M; o->N = N; o->E = E; o->left_shift = left_shift;
    o->cfg_w[0] = 0.4; o->cfg_w[1] = 0.4; o->cfg_w[2] = 0.2; o->divisor = 1.0; o->gamma = 0.99;
    o->t = (double *)calloc((size_t)B * N * M, 8);
    o->p = (double *)calloc((size_t)B * N * M, 8);
    o->tt = (double *)calloc((size_t)B * M * M, 8);
    o->edge_id = (int *)calloc((size_t)B * M, sizeof(int));
    o->mind = (double *)calloc((size_t)B * N, 8);
    o->minpt = (double *)calloc((size_t)B * N, 8);
    o->env = (env_t *)calloc(B, sizeof(env_t));
    o->nthreads = 1;
    for (int b = 0; b < B; b++) {
        env_t *e = &o->env[b];
        e->mach = (int *)calloc(N, sizeof(int));
        e->dur = (double *)calloc(N, 8); e->st = (double *)calloc(N, 8); e->ft = (double *)calloc(N, 8);
        e->sched = (uint8_t *)calloc(N, 1);
        e->route = (int *)calloc((size_t)M * N, sizeof(int)); e->rlen = (int *)calloc(M, sizeof(int));
        e->rpos = (int *)calloc(N, sizeof(int));
        e->mfea = (double *)calloc(M * 8, 8);
        e->est_st = (double *)calloc(N, 8); e->est_ft = (double *)calloc(N, 8); e->est_pt = (double *)calloc(N, 8);
    }
    return o;
}

void oracle_destroy(oracle_t *o) {
    if (!o) return;
    for (int b = 0; b < o->B; b++) {
        env_t *e = &o->env[b];
        free(e->mach); free(e->dur); free(e->st); free(e->ft); free(e->sched); free(e->route); free(e->rlen);
        free(e->rpos); free(e->mfea); free(e->est_st); free(e->est_ft); free(e->est_pt);
    }
    free(o->env); free(o->t); free(o->p); free(o->tt); free(o->edge_id); free(o->mind); free(o->minpt);
    free(o);
}

void oracle_set_params(oracle_t *o, double w_mk, double w_ec, double w_tt, double divisor, double gamma, int nthreads) {
    o->cfg_w[0] = w_mk; o->cfg_w[1] = w_ec; o->cfg_w[2] = w_tt; o->divisor = divisor; o->gamma = gamma;
    o->nthreads = nthreads > 0 ? nthreads : 1;
}

/* edge: [B][E][W] machine ids, padded with -1 */
void oracle_load(oracle_t *o, const double *t, const double *p, const double *tt, const int32_t *edge, int W) {
    int B = o->B, N = o->N, M = o->M;
    memcpy(o->t, t, (size_t)B * N * M * 8);
    memcpy(o->p, p, (size_t)B * N * M * 8);
    memcpy(o->tt, tt, (size_t)B * M * M * 8);
    for (int b = 0; b < B; b++) {
        for (int m = 0; m < M; m++) o->edge_id[b * M + m] = 0;
        for (int g = 0; g < o->E; g++)
            for (int k = 0; k < W; k++) {
                int m = edge[((size_t)b * o->E + g) * W + k];
                if (m >= 0 && m < M && o->edge_id[b * M + m] == 0) o->edge_id[b * M + m] = g + 1;
            }
        for (int i = 0; i < N; i++) {
            double mn = INFINITY, mnpt = INFINITY;
            for (int m = 0; m < M; m++) {
                double tv = T_(o, b, i, m), ptv = tv * fabs(P_(o, b, i, m)); /* SS:1933 */
                if (!(tv < 0) && tv < mn) mn = tv;
                if (!(ptv < 0) && ptv < mnpt) mnpt = ptv;
            }
            o->mind[(size_t)b * N + i] = mn;
            o->minpt[(size_t)b * N + i] = mnpt;
        }
    }
}

void oracle_scaler_init(oracle_t *o) { /* parallel_env.py:70-83 */
    for (int b = 0; b < o->B; b++) memset(&o->env[b].sc, 0, sizeof(scaler_t));
}
void oracle_scaler_reset(oracle_t *o) { /* RewardScaling.reset, Run.py:283-284 */
    for (int b = 0; b < o->B; b++) memset(o->env[b].sc.R, 0, sizeof(double) * 4);
}

/* load_instance + reset, SS:397-714, 1183-1245; weights [B][3] */
void oracle_reset(oracle_t *o, const double *weights) {
    int N = o->N, M = o->M;
#pragma omp parallel for num_threads(o->nthreads) schedule(static)
    for (int b = 0; b < o->B; b++) {
        env_t *e = &o->env[b];
        for (int i = 0; i < N; i++) { e->mach[i] = -1; e->dur[i] = 0; e->st[i] = 0; e->ft[i] = 0; e->sched[i] = 0; e->rpos[i] = -1; }
        for (int m = 0; m < M; m++) e->rlen[m] = 0;
        e->removed_head = -1; e->fresh_co = -1; e->nsched = 0; e->last_op = -1; e->last_m = -1;
        for (int k = 0; k < 3; k++) e->w[k] = weights[b * 3 + k];
        estimate(o, b, e);
        double mk = -INFINITY;
        for (int i = 0; i < N; i++) if (e->est_ft[i] > mk) mk = e->est_ft[i];
        e->mk_prev = mk;                              /* SS:697-700 */
        e->e_prev = np_pairwise_sum(e->est_pt, N);    /* SS:702 */
        e->trans = 0; e->trans_prev = 0; e->idle = 0; e->idle_prev = 0;
        for (int m = 0; m < M; m++) {                 /* SS:2343-2354 */
            double *mf = e->mfea + m * 8;
            mf[0] = mf[1] = mf[2] = mf[3] = mf[4] = 0;
            mf[5] = e->w[0]; mf[6] = e->w[1]; mf[7] = e->w[2];
        }
    }
}

void oracle_step(oracle_t *o, const int32_t *op, const int32_t *mach, double *reward5, double *scaled4,
                 uint8_t *done, uint8_t *invalid) {
#pragma omp parallel for num_threads(o->nthreads) schedule(static)
    for (int b = 0; b < o->B; b++) {
        double r5[5], s4[4];
        uint8_t dn = 0;
        int inv = env_step(o, b, &o->env[b], op[b], mach[b], r5, s4, &dn);
        if (reward5) memcpy(reward5 + (size_t)b * 5, r5, sizeof r5);
        if (scaled4) memcpy(scaled4 + (size_t)b * 4, s4, sizeof s4);
        if (done) done[b] = dn;
        if (invalid) invalid[b] = (uint8_t)inv;
    }
}

void oracle_obs(oracle_t *o, double *tfea, double *mfea, int32_t *ell_idx, double *ell_w, uint8_t *job_mask,
                int32_t *cand, int mask_mode) {
    int N = o->N, M = o->M, J = o->J;
#pragma omp parallel for num_threads(o->nthreads) schedule(static)
    for (int b = 0; b < o->B; b++)
        env_obs(o, b, &o->env[b], tfea ? tfea + (size_t)b * N * 12 : 0, mfea ? mfea + (size_t)b * M * 8 : 0,
                ell_idx ? ell_idx + (size_t)b * N * 3 : 0, ell_w ? ell_w + (size_t)b * N * 3 : 0,
                job_mask ? job_mask + (size_t)b * J : 0, cand ? cand + (size_t)b * J : 0, mask_mode);
}

void oracle_mfea1(oracle_t *o, const int32_t *op, double *out, uint8_t *mmask) {
    int M = o->M;
#pragma omp parallel for num_threads(o->nthreads) schedule(static)
    for (int b = 0; b < o->B; b++)
        env_mfea1(o, b, &o->env[b], op[b], out + (size_t)b * M * 6, mmask ? mmask + (size_t)b * M : 0);
}

/* dense [B][N][N], adj[dst][src], diagonal 1 (SS:2068-2073) */
void oracle_dense_adj(oracle_t *o, double *adj) {
    int N = o->N;
#pragma omp parallel for num_threads(o->nthreads) schedule(static)
    for (int b = 0; b < o->B; b++) {
        int32_t *ix = (int32_t *)malloc(sizeof(int32_t) * N * 3);
        double *wv = (double *)malloc(sizeof(double) * N * 3);
        env_obs(o, b, &o->env[b], 0, 0, ix, wv, 0, 0, 0);
        double *A = adj + (size_t)b * N * N;
        memset(A, 0, sizeof(double) * N * N);
        for (int v = 0; v < N; v++)
            for (int s = 0; s < 3; s++)
                if (ix[v * 3 + s] >= 0) A[(size_t)v * N + ix[v * 3 + s]] = wv[v * 3 + s];
        free(ix); free(wv);
    }
}

void oracle_costs(oracle_t *o, double *cost4) { /* Run.py:632-633, validate.py:273-277 */
    for (int b = 0; b < o->B; b++) {
        env_t *e = &o->env[b];
        cost4[b * 4 + 0] = e->mk_prev;
        cost4[b * 4 + 1] = e->e_prev / (double)o->N;
        cost4[b * 4 + 2] = e->trans_prev;
        cost4[b * 4 + 3] = e->idle_prev;
    }
}

void oracle_export_state(oracle_t *o, int32_t *mach, double *st, double *ft, int32_t *routes) {
    int N = o->N, M = o->M;
    for (int b = 0; b < o->B; b++) {
        env_t *e = &o->env[b];
        for (int i = 0; i < N; i++) {
            mach[(size_t)b * N + i] = e->mach[i];
            st[(size_t)b * N + i] = e->sched[i] ? e->st[i] : 0.0;
            ft[(size_t)b * N + i] = e->sched[i] ? e->ft[i] : 0.0;
        }
        for (int m = 0; m < M; m++)
            for (int k = 0; k < N; k++)
                routes[((size_t)b * M + m) * N + k] = k < e->rlen[m] ? e->route[m * N + k] : -1;
    }
}

void oracle_export_scaler(oracle_t *o, double *R, double *mean, double *S, int64_t *n) {
    for (int b = 0; b < o->B; b++) {
        memcpy(R + b * 4, o->env[b].sc.R, 32); memcpy(mean + b * 4, o->env[b].sc.mean, 32);
        memcpy(S + b * 4, o->env[b].sc.S, 32); n[b] = o->env[b].sc.n;
    }
}

/* ---- random-policy rollout used as the CPU baseline: one full pass of the hot path per step
 * (job mask + candidates, candidate-machine features, transition + reward + scaling, observation),
 * i.e. what Run.py's while-loop asks of the environment side each step. ---- */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
/* the same counter-based action draw the CUDA policy kernels use: one 64-bit value per (seed, env, step); stream 0 (the
 * job draw) is its high half, stream 1 (the machine draw) its low half */
uint32_t oracle_rand_u32(uint64_t seed, uint64_t env, uint64_t step, uint64_t stream) {
    uint64_t x = splitmix64(seed ^ splitmix64(env * 0x100000001B3ULL + step * 0x9E3779B1ULL));
    return stream == 0 ? (uint32_t)(x >> 32) : (uint32_t)x;
}

/* draws a uniformly random allowed job and feasible machine for env b from its current masks */
static void random_action(const oracle_t *o, int b, const uint8_t *job_mask, const int32_t *cand, uint64_t seed,
                          uint64_t genv, uint64_t step, int32_t *op, int32_t *mc) {
    int J = o->J, M = o->M, n = 0;
    for (int j = 0; j < J; j++) n += !job_mask[j];
    if (n == 0) { *op = -1; *mc = -1; return; }
    int k = (int)(((uint64_t)oracle_rand_u32(seed, genv, step, 0) * (uint64_t)n) >> 32);
    int a = -1;
    for (int j = 0; j < J; j++) if (!job_mask[j]) { if (k == 0) { a = cand[j]; break; } k--; }
    int nf = 0;
    for (int m = 0; m < M; m++) nf += (T_(o, b, a, m) >= 0);
    int km = (int)(((uint64_t)oracle_rand_u32(seed, genv, step, 1) * (uint64_t)nf) >> 32);
    int mm = -1;
    for (int m = 0; m < M; m++) if (T_(o, b, a, m) >= 0) { if (km == 0) { mm = m; break; } km--; }
    *op = a; *mc = mm;
}

/* runs `steps` env-steps per env with the random policy; env_offset makes the action stream
 * identical to a sharded GPU run.  If actions_out != NULL the drawn actions are recorded
 * [steps][B][2].  Outputs of the last step are left in the caller-provided buffers. */
void oracle_rollout_random(oracle_t *o, int steps, uint64_t seed, uint64_t env_offset, int mask_mode,
                           int32_t *actions_out, double *tfea, double *mfea, int32_t *ell_idx, double *ell_w,
                           double *mfea1, double *reward5, double *scaled4, uint8_t *done) {
    int N = o->N, M = o->M, J = o->J;
#pragma omp parallel for num_threads(o->nthreads) schedule(static)
    for (int b = 0; b < o->B; b++) {
        env_t *e = &o->env[b];
        uint8_t jm[J], mm[M], dn = 0;
        int32_t cd[J];
        double r5[5], s4[4];
        env_obs(o, b, e, 0, 0, 0, 0, jm, cd, mask_mode);
        for (int s = 0; s < steps; s++) {
            int32_t a, m;
            random_action(o, b, jm, cd, seed, env_offset + (uint64_t)b, (uint64_t)e->nsched, &a, &m);
            if (actions_out) { actions_out[((size_t)s * o->B + b) * 2] = a; actions_out[((size_t)s * o->B + b) * 2 + 1] = m; }
            if (a < 0) break;
            env_mfea1(o, b, e, a, mfea1 + (size_t)b * M * 6, mm);
            env_step(o, b, e, a, m, r5, s4, &dn);
            env_obs(o, b, e, tfea + (size_t)b * N * 12, mfea + (size_t)b * M * 8, ell_idx + (size_t)b * N * 3,
                    ell_w + (size_t)b * N * 3, jm, cd, mask_mode);
        }
        memcpy(reward5 + (size_t)b * 5, r5, sizeof r5);
        memcpy(scaled4 + (size_t)b * 4, s4, sizeof s4);
        done[b] = dn;
    }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
