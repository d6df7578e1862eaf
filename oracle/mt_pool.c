/* TEST INFRASTRUCTURE ONLY (see go1_oracle.h): native thread pool that times the CPU restatement on the host cores.
 *
 * BASELINE.md section 3 / SURVEY.md 8(d): the CPU baseline is the restated reference solver run over the same synthetic
 * batch (i) single-threaded and (ii) with a static split of the batch over all host cores by native threads, timed
 * with a monotonic clock around the whole batch.  bench.py's `cpu_baseline` leg and `--impl reference` arm call this
 * (Python threads under the GIL understated the 32-core figure in round 1).  One "pass" = one step of the bench:
 * per robot one body-inclination MPC tick (orc_body_step_batch: PRMPCClass::body_theta_mpc) and one step-timing SQP
 * tick (orc_step_timing_batch: NLPClass::step_timing_opti_loop), each thread on its own contiguous slice, every pass
 * on the same inputs (the slice's state is restored from the pristine copy first, as the GPU leg does).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "go1_oracle.h"

#define ORC_STEP_STATE 202   /* planner state doubles (go1_oracle.h: orc_step_timing_batch layouts) */
#define ORC_STEP_IN 20

typedef struct orc_pool orc_pool;
typedef struct {
    orc_pool *pool;
    int id, lo, hi;
    pthread_t th;
    /* per-thread scratch, allocated before any timing */
    double *theta, *out14, *x, *states, *out38;
    int *active, *nactive, *iters, *status;
} orc_worker;

struct orc_pool {
    int threads, B, nh, passes, quit;
    orc_body_cfg bc;
    orc_step_cfg sc;
    const int *btick, *stick;
    const double *tx, *theta0, *bstate, *refs, *states0, *ins;
    pthread_barrier_t start, end;
    orc_worker *w;
};

static void run_slice(orc_worker *w)
{
    orc_pool *p = w->pool;
    const int n = w->hi - w->lo, nh = p->nh;
    if (n <= 0) return;
    memcpy(w->theta, p->theta0 + (size_t)w->lo * 4, sizeof(double) * 4 * (size_t)n);
    memset(w->x, 0, sizeof(double) * 2 * nh * (size_t)n);
    orc_body_step_batch(&p->bc, n, p->btick + w->lo, p->tx + (size_t)w->lo * 27, w->theta, p->bstate + (size_t)w->lo * 4,
                        p->refs + (size_t)w->lo * 9 * nh, w->out14, w->x, w->active, w->nactive, w->iters, w->status);
    memcpy(w->states, p->states0 + (size_t)w->lo * ORC_STEP_STATE, sizeof(double) * ORC_STEP_STATE * (size_t)n);
    orc_step_timing_batch(&p->sc, n, p->stick + w->lo, w->states, p->ins + (size_t)w->lo * ORC_STEP_IN, w->out38, NULL);
}

static void *worker_main(void *arg)
{
    orc_worker *w = (orc_worker *)arg;
    orc_pool *p = w->pool;
    for (;;) {
        pthread_barrier_wait(&p->start);
        if (p->quit) break;
        for (int k = 0; k < p->passes; k++) run_slice(w);
        pthread_barrier_wait(&p->end);
    }
    return NULL;
}

/* All input arrays are instance-major and must outlive the pool:
 * btick [B], tx [B][27], theta0 [B][4], bstate [B][4], refs [B][9*nh]; stick [B], states0 [B][202], ins [B][20]. */
orc_pool *orc_pool_create(int threads, const orc_body_cfg *bc, const orc_step_cfg *sc, int B,
                          const int *btick, const double *tx, const double *theta0, const double *bstate, const double *refs,
                          const int *stick, const double *states0, const double *ins)
{
    if (threads < 1 || B < 1) return NULL;
    orc_pool *p = (orc_pool *)calloc(1, sizeof *p);
    p->threads = threads; p->B = B; p->nh = bc->nh; p->bc = *bc; p->sc = *sc;
    p->btick = btick; p->tx = tx; p->theta0 = theta0; p->bstate = bstate; p->refs = refs;
    p->stick = stick; p->states0 = states0; p->ins = ins;
    pthread_barrier_init(&p->start, NULL, (unsigned)threads + 1);
    pthread_barrier_init(&p->end, NULL, (unsigned)threads + 1);
    p->w = (orc_worker *)calloc((size_t)threads, sizeof(orc_worker));
    const int nh = p->nh;
    for (int t = 0; t < threads; t++) {
        orc_worker *w = &p->w[t];
        w->pool = p; w->id = t;
        w->lo = (int)((long long)B * t / threads); w->hi = (int)((long long)B * (t + 1) / threads);
        const size_t n = (size_t)(w->hi - w->lo) + 1;
        w->theta = (double *)malloc(sizeof(double) * 4 * n); w->out14 = (double *)malloc(sizeof(double) * 14 * n);
        w->x = (double *)malloc(sizeof(double) * 2 * nh * n); w->states = (double *)malloc(sizeof(double) * ORC_STEP_STATE * n);
        w->out38 = (double *)malloc(sizeof(double) * 38 * n);
        w->active = (int *)malloc(sizeof(int) * 12 * nh * n); w->nactive = (int *)malloc(sizeof(int) * n);
        w->iters = (int *)malloc(sizeof(int) * 4 * n); w->status = (int *)malloc(sizeof(int) * n);
        pthread_create(&w->th, NULL, worker_main, w);
    }
    return p;
}

/* Runs `passes` passes over the batch on all pool threads; returns the elapsed seconds (CLOCK_MONOTONIC, from the
 * release of the workers to the arrival of the last one). */
double orc_pool_run(orc_pool *p, int passes)
{
    struct timespec t0, t1;
    p->passes = passes;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_barrier_wait(&p->start);
    pthread_barrier_wait(&p->end);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* Thread t's slice results of the last pass (for the bit-identity test against the single-threaded drivers). */
void orc_pool_results(orc_pool *p, double *out14 /*B*14*/, double *out38 /*B*38*/)
{
    for (int t = 0; t < p->threads; t++) {
        orc_worker *w = &p->w[t];
        const size_t n = (size_t)(w->hi - w->lo);
        if (out14) memcpy(out14 + (size_t)w->lo * 14, w->out14, sizeof(double) * 14 * n);
        if (out38) memcpy(out38 + (size_t)w->lo * 38, w->out38, sizeof(double) * 38 * n);
    }
}

void orc_pool_destroy(orc_pool *p)
{
    if (!p) return;
    p->quit = 1;
    pthread_barrier_wait(&p->start);
    for (int t = 0; t < p->threads; t++) {
        orc_worker *w = &p->w[t];
        pthread_join(w->th, NULL);
        free(w->theta); free(w->out14); free(w->x); free(w->states); free(w->out38);
        free(w->active); free(w->nactive); free(w->iters); free(w->status);
    }
    pthread_barrier_destroy(&p->start); pthread_barrier_destroy(&p->end);
    free(p->w); free(p);
}
