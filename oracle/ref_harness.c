/*
 * ref_harness.c — thin C entry points around the UNMODIFIED reference sources.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (same rule as vs_oracle.c).  This file contains no
 * reference code: it #includes /root/reference/src/sqlite-vector.c where it lies (the include
 * path is given by oracle/Makefile) so that the reference's *static* scan functions
 * (vQuantRunMemory, vFullScanFindMaxIndex, vFullScanSortSlots, quantize_*, nearly_zero_float32)
 * can be called on flat buffers without going through SQL.  Outputs go to oracle/_ref/ only.
 *
 * Built twice by oracle/Makefile:
 *   _ref/libref_cpu.so   stock flags            -> distance_backend_name "CPU"  (parity oracle)
 *   _ref/libref_avx2.so  + -mavx2 -mfma         -> distance_backend_name "AVX2" (timed CPU baseline)
 */
#define SQLITE_CORE 1
#include "sqlite-vector.c" /* resolved through -I$(REF)/src, never copied */

#include <pthread.h>
#include <time.h>

#define REFH_API __attribute__((visibility("default")))

REFH_API const char *refh_backend(int force_cpu) {
    init_distance_functions(force_cpu != 0);
    return distance_backend_name;
}

REFH_API float refh_distance(int metric, int vtype, const void *a, const void *b, int n) {
    return dispatch_distance_table[metric][vtype](a, b, n);
}

REFH_API float refh_clamp_tiny(float d) { return nearly_zero_float32(d) ? 0.0f : d; }

REFH_API void refh_quantize(int vtype, const void *v, uint8_t *q, float offset, float scale, int dim, int qtype) {
    switch (vtype) {
    case VECTOR_TYPE_F32: quantize_float32((const float *)v, q, offset, scale, dim, (vector_qtype)qtype); break;
    case VECTOR_TYPE_F16: quantize_float16((const uint16_t *)v, q, offset, scale, dim, (vector_qtype)qtype); break;
    case VECTOR_TYPE_BF16: quantize_bfloat16((const uint16_t *)v, q, offset, scale, dim, (vector_qtype)qtype); break;
    case VECTOR_TYPE_U8: quantize_u8((const uint8_t *)v, q, offset, scale, dim, (vector_qtype)qtype); break;
    case VECTOR_TYPE_I8: quantize_i8((const int8_t *)v, q, offset, scale, dim, (vector_qtype)qtype); break;
    }
}

REFH_API uint16_t refh_f32_to_f16(float f) { return float32_to_float16(f); }
REFH_API uint16_t refh_f32_to_bf16(float f) { return float32_to_bfloat16(f); }
REFH_API float refh_f16_to_f32(uint16_t h) { return float16_to_float32(h); }

static void cursor_prepare(vFullScanCursor *c, table_context *t, int k, int start_max_index, int64_t *ids, double *dist) {
    memset(c, 0, sizeof(*c));
    c->table = t;
    c->rowids = ids;
    c->distance = dist;
    c->row_count = k;
    c->max_index = start_max_index;
    memset(ids, 0, (size_t)k * sizeof(int64_t));
    for (int i = 0; i < k; ++i) dist[i] = INFINITY;
}

/* the reference's own preloaded-scan loop + sort; returns valid rows */
REFH_API int refh_quant_scan(const uint8_t *preloaded, int counter, int dim, int qtype, int metric, int k,
                             const uint8_t *qvec, int start_max_index, int64_t *ids, double *dist, int *out_max_index) {
    table_context t;
    memset(&t, 0, sizeof(t));
    t.options.v_dim = dim;
    t.options.v_distance = (vector_distance)metric;
    t.options.q_type = (vector_qtype)qtype;
    t.preloaded = (void *)preloaded;
    t.precounter = counter;
    vFullScanCursor c;
    cursor_prepare(&c, &t, k, start_max_index, ids, dist);
    vQuantRunMemory(&c, (uint8_t *)qvec, (vector_qtype)qtype, dim);
    if (out_max_index) *out_max_index = c.max_index;
    int unused = vFullScanSortSlots(&c);
    return k - unused;
}

/* the arithmetic of vFullScanRun over a flat column: reference kernel, clamp, slot update and sort */
REFH_API int refh_flat_scan(int metric, int vtype, const void *query, const uint8_t *data, int64_t n, int dim,
                            size_t stride, size_t vec_off, const int64_t *rowids, int k, int start_max_index,
                            int64_t *ids, double *dist) {
    table_context t;
    memset(&t, 0, sizeof(t));
    vFullScanCursor c;
    cursor_prepare(&c, &t, k, start_max_index, ids, dist);
    distance_function_t fn = dispatch_distance_table[metric][vtype];
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t *row = data + (size_t)i * stride;
        float d = fn(query, row + vec_off, dim);
        if (nearly_zero_float32(d)) d = 0.0;
        if (d < c.distance[c.max_index]) {
            c.distance[c.max_index] = d;
            c.rowids[c.max_index] = rowids ? rowids[i] : INT64_FROM_INT8PTR(row);
            c.max_index = vFullScanFindMaxIndex(c.distance, c.row_count);
        }
    }
    int unused = vFullScanSortSlots(&c);
    return k - unused;
}

/* ---- timing: T threads, each running `reps` independent queries over the shared read-only buffer ---- */
typedef struct {
    const uint8_t *data; int64_t n; int dim; size_t stride; size_t vec_off;
    int metric; int vtype; int k; const uint8_t *queries; size_t qbytes; int reps; int quant; int qtype;
    int64_t checksum;
} refh_job;

static void *refh_worker(void *arg) {
    refh_job *j = (refh_job *)arg;
    int64_t *ids = (int64_t *)malloc((size_t)j->k * sizeof(int64_t));
    double *dist = (double *)malloc((size_t)j->k * sizeof(double));
    for (int r = 0; r < j->reps; ++r) {
        const uint8_t *q = j->queries + (size_t)r * j->qbytes;
        if (j->quant) refh_quant_scan(j->data, (int)j->n, j->dim, j->qtype, j->metric, j->k, q, 0, ids, dist, NULL);
        else refh_flat_scan(j->metric, j->vtype, q, j->data, j->n, j->dim, j->stride, j->vec_off, NULL, j->k, 0, ids, dist);
        j->checksum += ids[0];
    }
    free(ids); free(dist);
    return NULL;
}

/* returns elapsed wall seconds for threads*reps queries; queries = threads*reps vectors of qbytes each */
REFH_API double refh_time_queries(const uint8_t *data, int64_t n, int dim, size_t stride, size_t vec_off, int metric,
                                  int vtype, int quant, int qtype, int k, const uint8_t *queries, size_t qbytes,
                                  int threads, int reps, int64_t *checksum) {
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    refh_job *jobs = (refh_job *)calloc((size_t)threads, sizeof(refh_job));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; ++t) {
        refh_job j = {data, n, dim, stride, vec_off, metric, vtype, k, queries + (size_t)t * reps * qbytes, qbytes, reps, quant, qtype, 0};
        jobs[t] = j;
        pthread_create(&tid[t], NULL, refh_worker, &jobs[t]);
    }
    int64_t cs = 0;
    for (int t = 0; t < threads; ++t) { pthread_join(tid[t], NULL); cs += jobs[t].checksum; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (checksum) *checksum = cs;
    free(tid); free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
