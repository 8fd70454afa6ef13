/* probe_oracle.c -- C restatement of the health-probe arithmetic (TEST INFRASTRUCTURE ONLY).
 *
 * Same definition as oracle/probe.py (which is the specification):
 *     word(i, seed) = (uint32(i) * 2654435761) ^ seed
 *     pass: checksum = sum(src) mod 2^64; mismatches = #{src[i] != word(i, seed)};
 *           first_bad = min such i else UINT64_MAX; dst[i] = src[i] ^ delta
 * Used (a) as a second, independent implementation to cross-check probe.py, and (b) as the
 * `cpu_baseline` of bench.py: the identical algorithm streamed through host DRAM on the GPU
 * box's CPU cores (pthreads; the image has no libgomp).  Never linked into or called by the
 * product.
 *
 *   gcc -O3 -march=x86-64-v3 -pthread -shared -fPIC oracle/probe_oracle.c -o oracle/_build/libprobe_oracle.so
 */
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>

#define PATTERN_MUL 2654435761u
#define MAX_THREADS 256

typedef struct {
    const uint32_t *src;
    uint32_t *dst;
    uint64_t begin, end;
    uint32_t seed, delta;
    int mode; /* 0 fill, 1 pass, 2 expected checksum */
    uint64_t sum, bad, first;
} job_t;

static void *worker(void *p) {
    job_t *j = (job_t *)p;
    uint64_t sum = 0, bad = 0, first = UINT64_MAX;
    if (j->mode == 0) {
        for (uint64_t i = j->begin; i < j->end; ++i) j->dst[i] = ((uint32_t)i * PATTERN_MUL) ^ j->seed;
    } else if (j->mode == 1) {
        for (uint64_t i = j->begin; i < j->end; ++i) {
            const uint32_t w = j->src[i];
            sum += w;
            if (w != (((uint32_t)i * PATTERN_MUL) ^ j->seed)) {
                bad++;
                if (i < first) first = i;
            }
            j->dst[i] = w ^ j->delta;
        }
    } else {
        for (uint64_t i = j->begin; i < j->end; ++i) sum += ((uint32_t)i * PATTERN_MUL) ^ j->seed;
    }
    j->sum = sum; j->bad = bad; j->first = first;
    return NULL;
}

static void run(job_t *proto, uint64_t n_words, int threads, uint64_t *out) {
    if (threads < 1) threads = 1;
    if (threads > MAX_THREADS) threads = MAX_THREADS;
    pthread_t th[MAX_THREADS];
    job_t jobs[MAX_THREADS];
    const uint64_t per = (n_words + (uint64_t)threads - 1) / (uint64_t)threads;
    for (int t = 0; t < threads; ++t) {
        jobs[t] = *proto;
        jobs[t].begin = per * (uint64_t)t < n_words ? per * (uint64_t)t : n_words;
        jobs[t].end = jobs[t].begin + per < n_words ? jobs[t].begin + per : n_words;
        if (t) pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    worker(&jobs[0]);
    uint64_t sum = jobs[0].sum, bad = jobs[0].bad, first = jobs[0].first;
    for (int t = 1; t < threads; ++t) {
        pthread_join(th[t], NULL);
        sum += jobs[t].sum; bad += jobs[t].bad;
        if (jobs[t].first < first) first = jobs[t].first;
    }
    if (out) { out[0] = sum; out[1] = bad; out[2] = first; }
}

void oracle_fill(uint32_t *dst, uint64_t n_words, uint32_t seed, int threads) {
    job_t j = {0};
    j.dst = dst; j.seed = seed; j.mode = 0;
    run(&j, n_words, threads, NULL);
}

/* out[0] = checksum, out[1] = mismatches, out[2] = first_bad */
void oracle_probe_pass(const uint32_t *src, uint32_t *dst, uint64_t n_words, uint32_t seed, uint32_t delta,
                       int threads, uint64_t *out) {
    job_t j = {0};
    j.src = src; j.dst = dst; j.seed = seed; j.delta = delta; j.mode = 1;
    run(&j, n_words, threads, out);
}

uint64_t oracle_expected_checksum(uint64_t n_words, uint32_t seed, int threads) {
    job_t j = {0};
    uint64_t out[3];
    j.seed = seed; j.mode = 2;
    run(&j, n_words, threads, out);
    return out[0];
}
