/*
 * vs_oracle.c — CPU restatement of sqlite-vector's brute-force scan path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (sqlite_vector_b200/) may
 * include, link or call this file; it is the checker used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py compares every
 * function here, bit for bit, against the unmodified reference compiled into
 * oracle/_ref/ (see oracle/Makefile) and tests/golden/ holds vectors generated
 * from that reference build (tests/golden/make_golden.py).
 *
 * What is restated (reference = /root/reference, sqliteai/sqlite-vector 0.9.23):
 *   - the 25 scalar distance kernels           src/distance-cpu.c:39-693
 *   - bf16 / f16 bit helpers                    src/distance-cpu.h:69-128, libs/fp16/fp16.h
 *   - nearly-zero clamp                         src/sqlite-vector.c:994-996
 *   - element quantizers + rounding             src/sqlite-vector.c:495-757
 *   - quantization parameters (min/max pass)    src/sqlite-vector.c:1199-1272
 *   - the k-slot top-k scan and exchange sort   src/sqlite-vector.c:1808-1817, 2022-2069, 2121-2157
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off (no -march flags): the stock
 * reference Makefile builds for baseline x86-64, i.e. SSE2 scalar math with no
 * FMA contraction, and that is what the float evaluation order below assumes.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VSO_API __attribute__((visibility("default")))

/* enums mirror src/distance-cpu.h:36-58 (1-based) */
enum { T_F32 = 1, T_F16 = 2, T_BF16 = 3, T_U8 = 4, T_I8 = 5 };
enum { M_L2 = 1, M_L2SQ = 2, M_COS = 3, M_DOT = 4, M_L1 = 5 };
enum { Q_AUTO = 0, Q_U8 = 1, Q_S8 = 2 };

/* ---------------------------------------------------------------- bit helpers */

static inline uint32_t bits_of(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float float_of(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* bf16 <-> f32, RNE without NaN special-casing: src/distance-cpu.h:100-108 */
static inline float bf16_to_f32(uint16_t h) { return float_of((uint32_t)h << 16); }
VSO_API uint16_t vso_f32_to_bf16(float f) {
    uint32_t x = bits_of(f);
    return (uint16_t)((x + 0x7FFFu + ((x >> 16) & 1u)) >> 16);
}

/* IEEE binary16 <-> binary32 (value-exact; the reference delegates to libs/fp16/fp16.h:115,256) */
static inline float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t man = h & 0x3FFu;
    if (exp == 0x1Fu) return float_of(sign | 0x7F800000u | (man << 13));      /* inf / nan */
    if (exp != 0) return float_of(sign | ((exp + 112u) << 23) | (man << 13)); /* normal */
    if (man == 0) return float_of(sign);                                       /* +-0 */
    int sh = 0;                                                                /* subnormal: renormalise */
    while (!(man & 0x400u)) { man <<= 1; ++sh; }
    man &= 0x3FFu;
    return float_of(sign | ((uint32_t)(113 - sh) << 23) | (man << 13));
}
VSO_API float vso_f16_to_f32(uint16_t h) { return f16_to_f32(h); }

VSO_API uint16_t vso_f32_to_f16(float f) {
    uint32_t x = bits_of(f);
    uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
    uint32_t ax = x & 0x7FFFFFFFu;
    if (ax > 0x7F800000u) return (uint16_t)(sign | 0x7E00u);           /* NaN -> canonical quiet NaN (fp16.h) */
    if (ax >= 0x47800000u) return (uint16_t)(sign | 0x7C00u);          /* >= 65536 (incl. inf) -> inf */
    if (ax >= 0x38800000u) {                                           /* normal half range */
        uint32_t mant = ax & 0x7FFFFFu;
        uint32_t e = (ax >> 23) - 112u;
        uint32_t h = (e << 10) | (mant >> 13);
        uint32_t rem = mant & 0x1FFFu;
        if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;        /* RNE, may carry into exponent / inf */
        return (uint16_t)(sign | h);
    }
    if (ax < 0x33000000u) return sign;                                 /* < 2^-25 -> +-0 */
    /* subnormal half: value = mant24 * 2^(e-150), half ulp = 2^-24 */
    uint32_t mant = (ax & 0x7FFFFFu) | 0x800000u;
    int shift = 126 - (int)(ax >> 23);                                 /* 14..24 */
    uint32_t h = mant >> shift;
    uint32_t rem = mant & ((1u << shift) - 1u);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1u))) ++h;
    return (uint16_t)(sign | h);
}

static inline int f16_nan(uint16_t h) { return (h & 0x7C00u) == 0x7C00u && (h & 0x3FFu); }
static inline int f16_inf(uint16_t h) { return (h & 0x7C00u) == 0x7C00u && !(h & 0x3FFu); }
static inline int f16_neg(uint16_t h) { return (h >> 15) & 1; }

/* ------------------------------------------------------------ f32 kernels */
/* src/distance-cpu.c:39-64: 4-wide groups folded into one float accumulator */
static float f32_l2(const float *a, const float *b, int n, int root) {
    float acc = 0.0f;
    int i = 0;
    for (; i + 4 <= n; i += 4) {
        float d0 = a[i] - b[i], d1 = a[i + 1] - b[i + 1], d2 = a[i + 2] - b[i + 2], d3 = a[i + 3] - b[i + 3];
        acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    for (; i < n; ++i) { float d = a[i] - b[i]; acc += d * d; }
    return root ? sqrtf(acc) : acc;
}
/* src/distance-cpu.c:74-110 */
static float f32_cos(const float *a, const float *b, int n) {
    float dot = 0.0f, na = 0.0f, nb = 0.0f;
    int i = 0;
    for (; i + 4 <= n; i += 4) {
        float x0 = a[i], x1 = a[i + 1], x2 = a[i + 2], x3 = a[i + 3];
        float y0 = b[i], y1 = b[i + 1], y2 = b[i + 2], y3 = b[i + 3];
        dot += x0 * y0 + x1 * y1 + x2 * y2 + x3 * y3;
        na += x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3;
        nb += y0 * y0 + y1 * y1 + y2 * y2 + y3 * y3;
    }
    for (; i < n; ++i) { dot += a[i] * b[i]; na += a[i] * a[i]; nb += b[i] * b[i]; }
    if (na == 0.0f || nb == 0.0f) return 1.0f;
    return 1.0f - (dot / (sqrtf(na) * sqrtf(nb)));
}
/* src/distance-cpu.c:112-136 */
static float f32_dot(const float *a, const float *b, int n) {
    float dot = 0.0f;
    int i = 0;
    for (; i + 4 <= n; i += 4)
        dot += a[i] * b[i] + a[i + 1] * b[i + 1] + a[i + 2] * b[i + 2] + a[i + 3] * b[i + 3];
    for (; i < n; ++i) dot += a[i] * b[i];
    return -dot;
}
/* src/distance-cpu.c:138-159: every |a-b| is added on its own, so a plain loop is the same order */
static float f32_l1(const float *a, const float *b, int n) {
    float acc = 0.0f;
    for (int i = 0; i < n; ++i) acc += fabsf(a[i] - b[i]);
    return acc;
}

/* ------------------------------------------------------------ scaled sum of squares */
/* LASSQ_UPDATE, src/distance-cpu.c:23-35 */
typedef struct { double scale, ssq; } lassq_t;
static inline void lassq_add(lassq_t *s, double mag) {
    if (mag == 0.0) return;
    if (s->scale < mag) {
        double r = s->scale / mag;
        s->ssq = 1.0 + s->ssq * (r * r);
        s->scale = mag;
    } else {
        double r = mag / s->scale;
        s->ssq += r * r;
    }
}
static inline float lassq_result(const lassq_t *s, int root) {
    double sum_sq = (s->scale == 0.0) ? 0.0 : (s->scale * s->scale * s->ssq);
    return (float)(root ? sqrt(sum_sq) : sum_sq);
}

/* ------------------------------------------------------------ bf16 kernels */
/* src/distance-cpu.c:164-197: difference in f32, inf -> +inf, NaN lane skipped */
static float bf16_l2(const uint16_t *a, const uint16_t *b, int n, int root) {
    lassq_t s = {0.0, 1.0};
    for (int i = 0; i < n; ++i) {
        float d = bf16_to_f32(a[i]) - bf16_to_f32(b[i]);
        if (isinf(d)) return INFINITY;
        if (!isnan(d)) lassq_add(&s, fabs((double)d));
    }
    return lassq_result(&s, root);
}
/* src/distance-cpu.c:207-253: three fmaf chains */
static float bf16_cos(const uint16_t *a, const uint16_t *b, int n) {
    float dot = 0.0f, na = 0.0f, nb = 0.0f;
    for (int i = 0; i < n; ++i) {
        float x = bf16_to_f32(a[i]), y = bf16_to_f32(b[i]);
        dot = fmaf(x, y, dot);
        na = fmaf(x, x, na);
        nb = fmaf(y, y, nb);
    }
    if (na == 0.0f || nb == 0.0f) return 1.0f;
    return 1.0f - (dot / (sqrtf(na) * sqrtf(nb)));
}
/* src/distance-cpu.c:255-284 */
static float bf16_dot(const uint16_t *a, const uint16_t *b, int n) {
    float dot = 0.0f;
    for (int i = 0; i < n; ++i) dot = fmaf(bf16_to_f32(a[i]), bf16_to_f32(b[i]), dot);
    return -dot;
}
/* src/distance-cpu.c:286-314 */
static float bf16_l1(const uint16_t *a, const uint16_t *b, int n) {
    float acc = 0.0f;
    for (int i = 0; i < n; ++i) acc += fabsf(bf16_to_f32(a[i]) - bf16_to_f32(b[i]));
    return acc;
}

/* ------------------------------------------------------------ f16 kernels */
static inline int f16_inf_mismatch(uint16_t x, uint16_t y) {
    /* an infinity not paired with a same-signed infinity (src/distance-cpu.c:332) */
    return (f16_inf(x) || f16_inf(y)) && !(f16_inf(x) && f16_inf(y) && f16_neg(x) == f16_neg(y));
}
/* src/distance-cpu.c:318-356.  The reference checks the four lanes of a group for
 * infinities before touching any of them; because an infinity anywhere makes the
 * function return +inf, lane-by-lane evaluation gives the same value. */
static float f16_l2(const uint16_t *a, const uint16_t *b, int n, int root) {
    lassq_t s = {0.0, 1.0};
    for (int i = 0; i < n; ++i)
        if (f16_inf_mismatch(a[i], b[i])) return INFINITY;
    for (int i = 0; i < n; ++i) {
        if (f16_nan(a[i]) || f16_nan(b[i])) continue;
        double d = (double)f16_to_f32(a[i]) - (double)f16_to_f32(b[i]);
        lassq_add(&s, fabs(d));
    }
    return lassq_result(&s, root);
}
/* src/distance-cpu.c:366-397 */
static float f16_l1(const uint16_t *a, const uint16_t *b, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i)
        if (f16_inf_mismatch(a[i], b[i])) return INFINITY;
    for (int i = 0; i < n; ++i) {
        if (f16_nan(a[i]) || f16_nan(b[i])) continue;
        acc += fabs((double)f16_to_f32(a[i]) - (double)f16_to_f32(b[i]));
    }
    return (float)acc;
}
/* src/distance-cpu.c:399-429: first infinite product (in index order) decides the sign */
static float f16_dot(const uint16_t *a, const uint16_t *b, int n) {
    double dot = 0.0;
    for (int i = 0; i < n; ++i) {
        float x = f16_to_f32(a[i]), y = f16_to_f32(b[i]);
        if (isnan(x) || isnan(y)) continue;
        double p = (double)x * (double)y;
        if (isinf(p)) return (p > 0) ? -INFINITY : INFINITY;
        dot += p;
    }
    return (float)(-dot);
}
/* src/distance-cpu.c:431-466 */
static float f16_cos(const uint16_t *a, const uint16_t *b, int n) {
    double dot = 0.0, na = 0.0, nb = 0.0;
    for (int i = 0; i < n; ++i) {
        float x = f16_to_f32(a[i]), y = f16_to_f32(b[i]);
        if (isnan(x) || isnan(y)) continue;
        if (isinf(x) || isinf(y)) return 1.0f;
        double xd = x, yd = y;
        dot += xd * yd;
        na += xd * xd;
        nb += yd * yd;
    }
    double denom = sqrt(na) * sqrt(nb);
    if (!(denom > 0.0) || !isfinite(denom) || !isfinite(dot)) return 1.0f;
    double c = dot / denom;
    if (c > 1.0) c = 1.0;
    if (c < -1.0) c = -1.0;
    return (float)(1.0 - c);
}

/* ------------------------------------------------------------ 8-bit kernels */
/* The scalar reference folds integer work into FLOAT accumulators for L2 / dot / L1
 * (src/distance-cpu.c:470-502, 541-578, 582-614, 653-693); its AVX2 twin keeps int32
 * and converts once (src/distance-avx2.c:586-950).  `exact` selects the latter. */
#define INT8_KERNELS(NAME, ELT, ACC_T)                                                        \
    static float NAME##_l2(const ELT *a, const ELT *b, int n, int root, int exact) {          \
        if (exact) {                                                                          \
            int64_t s = 0;                                                                    \
            for (int i = 0; i < n; ++i) { int d = (int)a[i] - (int)b[i]; s += d * d; }        \
            float f = (float)(int32_t)s;                                                      \
            return root ? sqrtf(f) : f;                                                       \
        }                                                                                     \
        float acc = 0.0f;                                                                     \
        int i = 0;                                                                            \
        for (; i + 4 <= n; i += 4) {                                                          \
            int d0 = (int)a[i] - (int)b[i], d1 = (int)a[i + 1] - (int)b[i + 1];               \
            int d2 = (int)a[i + 2] - (int)b[i + 2], d3 = (int)a[i + 3] - (int)b[i + 3];       \
            acc += (float)(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);                            \
        }                                                                                     \
        for (; i < n; ++i) { int d = (int)a[i] - (int)b[i]; acc += (float)(d * d); }          \
        return root ? sqrtf(acc) : acc;                                                       \
    }                                                                                         \
    static float NAME##_cos(const ELT *a, const ELT *b, int n) {                              \
        ACC_T dot = 0, na = 0, nb = 0;                                                        \
        for (int i = 0; i < n; ++i) {                                                         \
            ACC_T x = a[i], y = b[i];                                                         \
            dot += x * y; na += x * x; nb += y * y;                                           \
        }                                                                                     \
        if (na == 0 || nb == 0) return 1.0f;                                                  \
        float c = dot / (sqrtf((float)na) * sqrtf((float)nb));                                \
        return 1.0f - c;                                                                      \
    }                                                                                         \
    static float NAME##_dot(const ELT *a, const ELT *b, int n, int exact) {                   \
        if (exact) {                                                                          \
            int64_t s = 0;                                                                    \
            for (int i = 0; i < n; ++i) s += (int)a[i] * (int)b[i];                           \
            return -(float)(int32_t)s;                                                        \
        }                                                                                     \
        float dot = 0.0f;                                                                     \
        for (int i = 0; i < n; ++i) dot += (float)a[i] * b[i];                                \
        return -dot;                                                                          \
    }                                                                                         \
    static float NAME##_l1(const ELT *a, const ELT *b, int n) {                               \
        float acc = 0.0f;                                                                     \
        for (int i = 0; i < n; ++i) acc += fabsf((float)a[i] - (float)b[i]);                  \
        return acc;                                                                           \
    }

INT8_KERNELS(u8, uint8_t, uint32_t) /* src/distance-cpu.c:470-578 */
INT8_KERNELS(i8, int8_t, int32_t)   /* src/distance-cpu.c:582-693 */

/* ------------------------------------------------------------ dispatch */
/* dispatch_distance_table[metric][type], src/distance-cpu.c:755-795 */
VSO_API float vso_distance(int metric, int vtype, const void *a, const void *b, int n, int int_exact) {
    switch (vtype) {
    case T_F32:
        switch (metric) {
        case M_L2: return f32_l2(a, b, n, 1);
        case M_L2SQ: return f32_l2(a, b, n, 0);
        case M_COS: return f32_cos(a, b, n);
        case M_DOT: return f32_dot(a, b, n);
        case M_L1: return f32_l1(a, b, n);
        }
        break;
    case T_F16:
        switch (metric) {
        case M_L2: return f16_l2(a, b, n, 1);
        case M_L2SQ: return f16_l2(a, b, n, 0);
        case M_COS: return f16_cos(a, b, n);
        case M_DOT: return f16_dot(a, b, n);
        case M_L1: return f16_l1(a, b, n);
        }
        break;
    case T_BF16:
        switch (metric) {
        case M_L2: return bf16_l2(a, b, n, 1);
        case M_L2SQ: return bf16_l2(a, b, n, 0);
        case M_COS: return bf16_cos(a, b, n);
        case M_DOT: return bf16_dot(a, b, n);
        case M_L1: return bf16_l1(a, b, n);
        }
        break;
    case T_U8:
        switch (metric) {
        case M_L2: return u8_l2(a, b, n, 1, int_exact);
        case M_L2SQ: return u8_l2(a, b, n, 0, int_exact);
        case M_COS: return u8_cos(a, b, n);
        case M_DOT: return u8_dot(a, b, n, int_exact);
        case M_L1: return u8_l1(a, b, n);
        }
        break;
    case T_I8:
        switch (metric) {
        case M_L2: return i8_l2(a, b, n, 1, int_exact);
        case M_L2SQ: return i8_l2(a, b, n, 0, int_exact);
        case M_COS: return i8_cos(a, b, n);
        case M_DOT: return i8_dot(a, b, n, int_exact);
        case M_L1: return i8_l1(a, b, n);
        }
        break;
    }
    return NAN;
}

VSO_API int vso_elem_size(int vtype) {
    switch (vtype) {
    case T_F32: return 4;
    case T_F16: case T_BF16: return 2;
    case T_U8: case T_I8: return 1;
    }
    return 0;
}

/* |x| <= 8*FLT_EPSILON -> 0 (also maps -0.0 to +0.0): src/sqlite-vector.c:994-996 */
static inline float clamp_tiny(float d) { return (fabsf(d) <= 8.0f * FLT_EPSILON) ? 0.0f : d; }
VSO_API float vso_clamp_tiny(float d) { return clamp_tiny(d); }

/* ------------------------------------------------------------ quantizers */
/* half-away-from-zero with saturation: src/sqlite-vector.c:495-515 */
static inline uint8_t round_u8(float s) {
    if (!isfinite(s)) return (s > 0.0f) ? 255u : 0u;
    float r = s + 0.5f * (1.0f - 2.0f * (s < 0.0f));
    if (r >= 255.0f) return 255u;
    if (r <= 0.0f) return 0u;
    return (uint8_t)(int)r;
}
static inline int8_t round_s8(float s) {
    if (!isfinite(s)) return (s > 0.0f) ? 127 : (s < 0.0f ? -128 : 0);
    float r = s + 0.5f * (1.0f - 2.0f * (s < 0.0f));
    if (r >= 127.0f) return 127;
    if (r <= -128.0f) return -128;
    return (int8_t)(int)r;
}
static inline float elem_as_f32(int vtype, const void *v, int i) {
    switch (vtype) {
    case T_F32: return ((const float *)v)[i];
    case T_F16: return f16_to_f32(((const uint16_t *)v)[i]);
    case T_BF16: return bf16_to_f32(((const uint16_t *)v)[i]);
    case T_U8: return (float)((const uint8_t *)v)[i];
    case T_I8: return (float)((const int8_t *)v)[i];
    }
    return 0.0f;
}
/* quantize_<type>(): src/sqlite-vector.c:517-757.  The f32 source path truncates an
 * int cast then clamps (:525-533, :634-642); every other source type goes through the
 * NaN/Inf-safe rounders above.  For finite in-range input the two agree. */
VSO_API void vso_quantize(int vtype, const void *v, uint8_t *q, float offset, float scale, int dim, int qtype) {
    for (int i = 0; i < dim; ++i) {
        float s = (elem_as_f32(vtype, v, i) - offset) * scale;
        if (vtype == T_F32) {
            int r = (int)(s + 0.5f * (1.0f - 2.0f * (s < 0.0f)));
            if (qtype == Q_U8) q[i] = (uint8_t)(r > 255 ? 255 : (r < 0 ? 0 : r));
            else ((int8_t *)q)[i] = (int8_t)(r > 127 ? 127 : (r < -128 ? -128 : r));
        } else {
            if (qtype == Q_U8) q[i] = round_u8(s);
            else ((int8_t *)q)[i] = round_s8(s);
        }
    }
}

/* global min/max pass and the derived scale/offset: src/sqlite-vector.c:1199-1272 */
VSO_API int vso_quant_params(int vtype, const void *vectors, int64_t nrows, int dim, int qtype_in,
                             float *scale, float *offset, int *qtype_out) {
    float lo = FLT_MAX, hi = -FLT_MAX;
    int neg = 0;
    size_t row_bytes = (size_t)dim * (size_t)vso_elem_size(vtype);
    for (int64_t r = 0; r < nrows; ++r) {
        const void *row = (const uint8_t *)vectors + (size_t)r * row_bytes;
        for (int i = 0; i < dim; ++i) {
            float val = elem_as_f32(vtype, row, i);
            if (val < lo) lo = val;
            if (val > hi) hi = val;
            if (val < 0.0) neg = 1;
        }
    }
    int qt = qtype_in;
    if (qt == Q_AUTO) qt = neg ? Q_S8 : Q_U8;
    float abs_max = fmaxf(fabsf(lo), fabsf(hi));
    *scale = (qt == Q_U8) ? (255.0f / (hi - lo)) : (127.0f / abs_max);
    *offset = (qt == Q_U8) ? lo : 0.0f;
    *qtype_out = qt;
    return 0;
}

/* ------------------------------------------------------------ k-slot top-k */
/* first index holding the maximum: src/sqlite-vector.c:2022-2049 (both code paths use strict >) */
static int first_argmax(const double *v, int n) {
    int best = 0;
    for (int i = 1; i < n; ++i)
        if (v[i] > v[best]) best = i;
    return best;
}

typedef struct {
    int k;
    int max_index;
    double *dist;
    int64_t *ids;
} slots_t;

/* slot initialisation: src/sqlite-vector.c:1808-1813 (max_index deliberately NOT reset there) */
static void slots_reset(slots_t *s) {
    for (int i = 0; i < s->k; ++i) { s->dist[i] = INFINITY; s->ids[i] = 0; }
}
/* slot replacement: src/sqlite-vector.c:2102-2106 and :2145-2152 */
static inline void slots_offer(slots_t *s, float d, int64_t id) {
    if ((double)d < s->dist[s->max_index]) {
        s->dist[s->max_index] = d;
        s->ids[s->max_index] = id;
        s->max_index = first_argmax(s->dist, s->k);
    }
}
/* exchange sort + count of unused slots: src/sqlite-vector.c:2051-2069 */
static int slots_sort(slots_t *s) {
    int unused = 0, k = s->k;
    for (int i = 0; i < k - 1; ++i) {
        if (s->dist[i] == INFINITY) ++unused;
        for (int j = i + 1; j < k; ++j) {
            if (s->dist[j] < s->dist[i]) {
                double td = s->dist[i]; s->dist[i] = s->dist[j]; s->dist[j] = td;
                int64_t ti = s->ids[i]; s->ids[i] = s->ids[j]; s->ids[j] = ti;
            }
        }
    }
    if (s->dist[k - 1] == INFINITY) ++unused;
    return unused;
}

static inline int64_t le64(const uint8_t *p) { /* INT64_FROM_INT8PTR, src/sqlite-vector.c:86-94 */
    uint64_t v = 0;
    for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
    return (int64_t)v;
}

/*
 * One query against n rows laid out `stride` bytes apart; the vector of row i starts at
 * data + i*stride + vec_off.  rowids == NULL means "little-endian int64 at the row start",
 * i.e. the vector_quantize_preload buffer (stride = 8 + dim, vec_off = 8):
 * vQuantRunMemory, src/sqlite-vector.c:2121-2157.  With rowids given and vec_off = 0 it is
 * the arithmetic of vFullScanRun (src/sqlite-vector.c:2089-2107) over a dense column.
 * Returns the number of valid result rows (k minus unused slots, :1816-1817).
 */
VSO_API int vso_scan_topk(int metric, int vtype, const void *query, const uint8_t *data, int64_t n, int dim,
                          size_t stride, size_t vec_off, const int64_t *rowids, int k, int start_max_index,
                          int int_exact, int64_t *out_ids, double *out_dist, int *out_max_index) {
    if (k <= 0) return 0;
    slots_t s = {k, start_max_index, out_dist, out_ids};
    slots_reset(&s);
    if (s.max_index < 0 || s.max_index >= k) s.max_index = 0;
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t *row = data + (size_t)i * stride;
        float d = clamp_tiny(vso_distance(metric, vtype, query, row + vec_off, dim, int_exact));
        slots_offer(&s, d, rowids ? rowids[i] : le64(row));
    }
    if (out_max_index) *out_max_index = s.max_index;
    return k - slots_sort(&s);
}

/* The same slot algorithm over precomputed (distance, id) pairs in scan order. */
VSO_API int vso_topk_from_distances(const float *dist, const int64_t *ids, int64_t n, int k, int start_max_index,
                                    int64_t *out_ids, double *out_dist) {
    if (k <= 0) return 0;
    slots_t s = {k, start_max_index, out_dist, out_ids};
    slots_reset(&s);
    if (s.max_index < 0 || s.max_index >= k) s.max_index = 0;
    for (int64_t i = 0; i < n; ++i) slots_offer(&s, clamp_tiny(dist[i]), ids ? ids[i] : i);
    return k - slots_sort(&s);
}

/* all n clamped distances, as the *_stream modules emit them (src/sqlite-vector.c:1926-1927, 1958-1959) */
VSO_API void vso_distances_all(int metric, int vtype, const void *query, const uint8_t *data, int64_t n, int dim,
                               size_t stride, size_t vec_off, int int_exact, float *out) {
    for (int64_t i = 0; i < n; ++i)
        out[i] = clamp_tiny(vso_distance(metric, vtype, query, data + (size_t)i * stride + vec_off, dim, int_exact));
}

/* build the preload buffer [int64 LE rowid][dim x q8] from a dense source column: src/sqlite-vector.c:1295-1311 */
VSO_API void vso_build_quant_buffer(int vtype, const void *vectors, const int64_t *rowids, int64_t nrows, int dim,
                                    float offset, float scale, int qtype, uint8_t *out) {
    size_t row_bytes = (size_t)dim * (size_t)vso_elem_size(vtype);
    size_t stride = 8 + (size_t)dim;
    for (int64_t r = 0; r < nrows; ++r) {
        uint8_t *dst = out + (size_t)r * stride;
        uint64_t id = (uint64_t)(rowids ? rowids[r] : r + 1);
        for (int b = 0; b < 8; ++b) dst[b] = (uint8_t)(id >> (8 * b));
        vso_quantize(vtype, (const uint8_t *)vectors + (size_t)r * row_bytes, dst + 8, offset, scale, dim, qtype);
    }
}
