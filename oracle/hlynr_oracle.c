/*
 * hlynr_oracle.c -- TEST INFRASTRUCTURE.  CPU restatement of the reference's per-step hot path.
 *
 * This file is the parity oracle for the CUDA path.  It is NOT part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or call
 * it.  The product path (hlynr_intercept_b200/) never links or imports anything under oracle/.
 *
 * It restates, function by function, the reference algorithm (file:line relative to the reference
 * tree RomanSlack/Hlynr_Intercept):
 *     rl_system/environment.py   reset :353-603, step :605-859, _update_interceptor :861-963,
 *                                _update_missile_state :1069-1117, _update_wind :1119-1129,
 *                                _calculate_reward :1273-1320, _quaternion_multiply :1322-1331
 *     rl_system/core.py          SimpleKalmanFilter :12-133, SensorDelayBuffer :147-223,
 *                                _compute_ground_radar_detection :368-438, _compute_datalink_quality :440-474,
 *                                _compute_fusion_confidence :476-509, compute_radar_detection :511-691,
 *                                compute :693-1032, SafetyClamp.apply :1069-1100, quaternion_to_euler :1103-1121,
 *                                get_forward_vector :1143-1152
 *     rl_system/physics_models.py  AtmosphericModel :56-177, MachDragModel :197-264, EnhancedWindModel :304-387
 *     rl_system/physics_randomizer.py  randomize_for_episode :137-221, apply_to_* :243-297
 *
 * Precision model.  The reference is NumPy-2 code whose dtypes follow NEP-50 promotion: float32 state
 * arrays with float64 islands (SURVEY 8a "precision map").  Every arithmetic operation below is carried
 * out in double and then rounded to the dtype NumPy would have produced (`rn(x, p)`, p = F32 or F64);
 * for + - * / sqrt this yields exactly the correctly-rounded float32 result.  `S` is the dtype of the
 * integrator state: F32 for the native reference, F64 for the "up-cast" float64 reference
 * (SURVEY Appendix B.5).  Python-float constants are "weak": they are rounded to the dtype of the
 * array/scalar they meet (`wk`).  Small dot products use sequential accumulation without FMA.
 *
 * Structure deliberately follows the reference (6x6 Kalman matrices with a general 3x3 inverse, FIFO
 * delay buffers) rather than the optimised forms used by the CUDA kernels, so that the two are
 * independent restatements.
 *
 * Pinning: validated against the unmodified reference run in the build container through
 * oracle/ref_harness.py (identical seeds, actions and injected draws) and against the golden fixtures
 * it generated under tests/golden/ (tests/test_oracle_vs_golden.py).  The draws are the site-keyed
 * Philox4x32-10 contract of include/hlynr_rng.h (third-party published algorithm: Salmon et al. SC'11,
 * Random123 v1.09), pinned by the Random123 known-answer vectors.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/hlynr.h"
#include "../include/hlynr_rng.h"

typedef int Prec;
enum { F32 = 0, F64 = 1 };
#define PMAX(a, b) ((a) | (b))

static inline double rn(double x, Prec p) { return p ? x : (double)(float)x; }
static inline double wk(double c, Prec p) { return rn(c, p); } /* weak Python float meeting dtype p */
static inline double add(double a, double b, Prec p) { return rn(a + b, p); }
static inline double sub(double a, double b, Prec p) { return rn(a - b, p); }
static inline double mul(double a, double b, Prec p) { return rn(a * b, p); }
static inline double dvd(double a, double b, Prec p) { return rn(a / b, p); }
static inline double sqr(double a, Prec p) { return rn(sqrt(a), p); }
/* np.dot / np.linalg.norm on float32 vectors go through OpenBLAS sdot, which (measured against the NumPy
 * 2.3.5 build of the reference container, 20000/20000 random cases) forms float32 products and sums them
 * in a double accumulator, rounding once at the end.  float64 operands: plain sequential ddot. */
static inline double dotn(const double* a, const double* b, int n, Prec p) {
    if (p == F32) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += (double)(float)(a[i] * b[i]);
        return (double)(float)s;
    }
    double s = a[0] * b[0];
    for (int i = 1; i < n; ++i) s += a[i] * b[i];
    return s;
}
static inline double dot3(const double* a, const double* b, Prec p) { return dotn(a, b, 3, p); }
static inline double norm3(const double* a, Prec p) { return sqr(dot3(a, a, p), p); }
static inline double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* ---------------------------------------------------------------------------------------------- */
/* Philox4x32-10 and the draw contract (include/hlynr_rng.h)                                        */
/* ---------------------------------------------------------------------------------------------- */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)HLYNR_PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)HLYNR_PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += HLYNR_PHILOX_W0;
        k1 += HLYNR_PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }

typedef struct {
    uint64_t seed;
    uint64_t env_id;
    uint32_t episode;
} DrawCtx;

static void draw_raw(const DrawCtx* c, uint32_t step, uint32_t blk, uint32_t out[4]) {
    uint32_t ctr[4] = {(uint32_t)c->env_id, c->episode, step, blk | ((uint32_t)(c->env_id >> 32) << 16)};
    uint32_t key[2] = {(uint32_t)c->seed, (uint32_t)(c->seed >> 32)};
    philox4x32_10(ctr, key, out);
}
static inline float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }
static inline float u01_open(uint32_t x) { return (float)((x >> 8) + 1u) * 5.9604644775390625e-8f; }
static void draw_uniform4(const DrawCtx* c, uint32_t step, uint32_t blk, double u[4]) {
    uint32_t r[4];
    draw_raw(c, step, blk, r);
    for (int i = 0; i < 4; ++i) u[i] = (double)u01(r[i]);
}
static void draw_normal4(const DrawCtx* c, uint32_t step, uint32_t blk, double z[4]) {
    uint32_t r[4];
    draw_raw(c, step, blk, r);
    for (int a = 0; a < 4; a += 2) {
        float rad = sqrtf(-2.0f * logf(u01_open(r[a])));
        double t = 2.0 * (double)u01(r[a + 1]); /* exact */
        z[a] = (double)(rad * (float)cos(M_PI * t));
        z[a + 1] = (double)(rad * (float)sin(M_PI * t));
    }
}
static double draw_exp(const DrawCtx* c, uint32_t step, uint32_t blk) {
    uint32_t r[4];
    draw_raw(c, step, blk, r);
    return (double)(-logf(u01_open(r[0])));
}

void oracle_draws(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t step, uint32_t blk, uint32_t raw[4],
                  float uni[4], float nrm[4]) {
    DrawCtx c = {seed, env_id, episode};
    double u[4], z[4];
    draw_raw(&c, step, blk, raw);
    draw_uniform4(&c, step, blk, u);
    draw_normal4(&c, step, blk, z);
    for (int i = 0; i < 4; ++i) { uni[i] = (float)u[i]; nrm[i] = (float)z[i]; }
}

/* ---------------------------------------------------------------------------------------------- */
/* Reference objects                                                                              */
/* ---------------------------------------------------------------------------------------------- */
#define DB_MAX 33
typedef struct { /* core.py:147 SensorDelayBuffer: deque(maxlen = delay+1) of (measurement, detected) */
    double meas[DB_MAX][8];
    int meas_f64[DB_MAX]; /* dtype of the stored ground measurement arrays (float64 if noise was added) */
    int det[DB_MAX];
    int head, len, maxlen, delay, samples_received;
} DelayBuffer;

static void db_init(DelayBuffer* b, int delay_samples) {
    memset(b, 0, sizeof(*b));
    b->delay = delay_samples < 1 ? 1 : delay_samples; /* core.py:164 */
    b->maxlen = b->delay + 1;                           /* core.py:165 */
}
static void db_reset(DelayBuffer* b) { b->head = b->len = b->samples_received = 0; } /* core.py:210-215 */
/* core.py:175-208.  Returns 1 and the delayed sample, or 0 for (None, False). */
static int db_add(DelayBuffer* b, const double m[8], int m_f64, int detected, double out[8], int* out_f64,
                  int* out_det) {
    int slot;
    if (b->len == b->maxlen) { /* deque(maxlen): drop the oldest */
        slot = b->head;
        b->head = (b->head + 1) % b->maxlen;
    } else {
        slot = (b->head + b->len) % b->maxlen;
        b->len++;
    }
    memcpy(b->meas[slot], m, sizeof(double) * 8);
    b->meas_f64[slot] = m_f64;
    b->det[slot] = detected;
    b->samples_received++;
    if (b->samples_received < b->delay) return 0;
    if (b->len > b->delay) {
        memcpy(out, b->meas[b->head], sizeof(double) * 8);
        *out_f64 = b->meas_f64[b->head];
        *out_det = b->det[b->head];
        return 1;
    }
    return 0;
}

typedef struct { /* core.py:12 SimpleKalmanFilter, q = 5^2, r = 20^2 (core.py:331-335) */
    double x[6];
    int x_f64;       /* dtype of self.state: float32 until a float64 measurement is folded in */
    float P[6][6];   /* float32 always */
    int initialized;
} Kalman;

static float KF_F[6][6], KF_Q[6][6], KF_H[3][6], KF_R[3][3];
static double KF_DT = -1.0;

static void kf_build(double dt) { /* core.py:33-63, float32 matrices */
    if (KF_DT == dt) return;
    double q = 5.0 * 5.0, r = 20.0 * 20.0;
    memset(KF_F, 0, sizeof(KF_F)); memset(KF_Q, 0, sizeof(KF_Q)); memset(KF_H, 0, sizeof(KF_H)); memset(KF_R, 0, sizeof(KF_R));
    for (int a = 0; a < 3; ++a) {
        KF_F[a][a] = 1.f; KF_F[a + 3][a + 3] = 1.f; KF_F[a][a + 3] = (float)dt;
        KF_Q[a][a] = (float)(q * pow(dt, 4) / 4); KF_Q[a][a + 3] = (float)(q * pow(dt, 3) / 2);
        KF_Q[a + 3][a] = (float)(q * pow(dt, 3) / 2); KF_Q[a + 3][a + 3] = (float)(q * dt * dt);
        KF_H[a][a] = 1.f; KF_R[a][a] = (float)r;
    }
    KF_DT = dt;
}
static void kf_reset(Kalman* k) { /* core.py:65-69 */
    memset(k, 0, sizeof(*k));
    for (int i = 0; i < 6; ++i) k->P[i][i] = 1000.f;
}
static void matmul66(float A[6][6], float B[6][6], float C[6][6]) {
    float T[6][6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            float s = 0.f; /* OpenBLAS sgemm accumulates with FMA (measured) */
            for (int k = 0; k < 6; ++k) s = fmaf(A[i][k], B[k][j], s);
            T[i][j] = s;
        }
    memcpy(C, T, sizeof(T));
}
static int inv33(float A[3][3], float X[3][3]) { /* Gauss-Jordan with partial pivoting, float32 */
    float M[3][6];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { M[i][j] = A[i][j]; M[i][j + 3] = (i == j) ? 1.f : 0.f; }
    for (int c = 0; c < 3; ++c) {
        int piv = c;
        for (int r = c + 1; r < 3; ++r) if (fabsf(M[r][c]) > fabsf(M[piv][c])) piv = r;
        if (M[piv][c] == 0.f) return 0;
        if (piv != c) for (int j = 0; j < 6; ++j) { float t = M[c][j]; M[c][j] = M[piv][j]; M[piv][j] = t; }
        float inv = 1.f / M[c][c];
        for (int j = 0; j < 6; ++j) M[c][j] = M[c][j] * inv;
        for (int r = 0; r < 3; ++r) if (r != c && M[r][c] != 0.f) {
            float f = M[r][c];
            for (int j = 0; j < 6; ++j) M[r][j] = M[r][j] - f * M[c][j];
        }
    }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) X[i][j] = M[i][j + 3];
    return 1;
}
static void kf_predict(Kalman* k) { /* core.py:80-89 */
    if (!k->initialized) return;
    Prec p = k->x_f64 ? F64 : F32;
    double nx[6];
    for (int i = 0; i < 6; ++i) { /* state = F @ state (sequential accumulation) */
        double s = 0.0;
        for (int j = 0; j < 6; ++j) s = add(s, mul((double)KF_F[i][j], k->x[j], p), p);
        nx[i] = s;
    }
    memcpy(k->x, nx, sizeof(nx));
    float FP[6][6], Ft[6][6];
    matmul66(KF_F, k->P, FP);
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Ft[i][j] = KF_F[j][i];
    matmul66(FP, Ft, FP);
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) k->P[i][j] = FP[i][j] + KF_Q[i][j];
}
static void kf_update(Kalman* k, const double z[3], int z_f64) { /* core.py:91-116 */
    if (!k->initialized) { /* initialize(): assignment into the float32 state array */
        for (int a = 0; a < 3; ++a) { k->x[a] = rn(z[a], F32); k->x[a + 3] = 0.0; }
        k->initialized = 1;
        return;
    }
    Prec p = (k->x_f64 || z_f64) ? F64 : F32;
    double y[3];
    for (int a = 0; a < 3; ++a) y[a] = sub(z[a], k->x[a], p); /* H @ state selects state[0:3] */
    float S[3][3], Si[3][3], K[6][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) S[i][j] = k->P[i][j] + KF_R[i][j]; /* H P H^T + R */
    if (!inv33(S, Si)) return;
    for (int i = 0; i < 6; ++i) /* K = (P @ H^T) @ inv(S); P @ H^T = P[:, 0:3] */
        for (int j = 0; j < 3; ++j) {
            float s = 0.f;
            for (int m = 0; m < 3; ++m) s = fmaf(k->P[i][m], Si[m][j], s);
            K[i][j] = s;
        }
    for (int i = 0; i < 6; ++i) { /* state = state + K @ y */
        double s = 0.0;
        for (int j = 0; j < 3; ++j) s = add(s, mul((double)K[i][j], y[j], p), p);
        k->x[i] = add(k->x[i], s, p);
    }
    k->x_f64 = (p == F64);
    float IKH[6][6];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) IKH[i][j] = ((i == j) ? 1.f : 0.f) - (j < 3 ? K[i][j] : 0.f);
    matmul66(IKH, k->P, k->P);
}

typedef struct {
    /* InterceptEnvironment state */
    double ipos[3], ivel[3], quat[4], fuel, fuel_used;
    double mpos[3], mvel[3];
    double wind[3];
    int wind_f64; /* AR(1) wind turns current_wind into a float64 array (environment.py:1129) */
    double thrust[3];
    int steps, worsen, crossed;
    double prev_d, last_d, min_d;
    /* volley mode (environment.py:42-44, 370-373): missile_states[], missile_min_distances[], intercepted indices;
     * mpos/mvel above mirror self.missile_state (the priority missile, an alias of one list entry) */
    double vpos[HLYNR_MAX_VOLLEY][3], vvel[HLYNR_MAX_VOLLEY][3], vmin[HLYNR_MAX_VOLLEY];
    int vact[HLYNR_MAX_VOLLEY], vcur, vcount;
    /* DR-mutable model constants */
    double T0, base_cd, peak;
    int cd_strong; /* base_cd / peak are np.float64 scalars after DR (physics_randomizer.py:273,278) */
    /* observation generator */
    DelayBuffer onb, gnd;
    int has_onb;
    Kalman kf;
    int last_onb_det, last_gnd_det;
    int last_onb_fill; /* the delayed onboard sample was None (buffer still filling): last_detection_info['radar_quality'] = 0.0 */
    /* Monitor-like accounting */
    double ep_return;
    int ep_length;
    DrawCtx rng;
    int started;
    double margin; /* smallest relative distance to a decision threshold in the current tick */
} Env;

typedef struct {
    HlynrParams P;
    HlynrCurriculum C;
    int64_t n;
    Prec S;
    Env* env;
    HlynrStats stats;
} Oracle;

/* ---------------------------------------------------------------------------------------------- */
/* decisions with margins                                                                          */
/* ---------------------------------------------------------------------------------------------- */
static inline void note_margin(Env* e, double a, double b) {
    double s = fabs(b) > 1e-3 ? fabs(b) : 1e-3;
    double m = fabs(a - b) / s;
    if (m < e->margin) e->margin = m;
}
#define GT(e, a, b) (note_margin(e, a, b), (a) > (b))
#define LT(e, a, b) (note_margin(e, a, b), (a) < (b))
#define GE(e, a, b) (note_margin(e, a, b), (a) >= (b))
#define LE(e, a, b) (note_margin(e, a, b), (a) <= (b))

/* ---------------------------------------------------------------------------------------------- */
/* physics_models.py                                                                               */
/* ---------------------------------------------------------------------------------------------- */
/* AtmosphericModel.get_atmospheric_properties :154-177.  altitude is an S-typed scalar. */
static void isa_props(Env* e, double alt, Prec S, double* rho, double* cs) {
    const double R = 287.05, g0 = 9.80665, L = 0.0065;
    double T, Pr;
    if (alt <= 11000.0) {
        T = sub(wk(e->T0, S), mul(wk(L, S), alt, S), S);                   /* :70-71 */
        double ratio = dvd(T, wk(e->T0, S), S);                            /* :95 */
        double expo = g0 / (R * L);                                         /* :99 python float */
        Pr = mul(wk(101325.0, S), rn(pow(ratio, wk(expo, S)), S), S);       /* :100 */
    } else if (alt <= 20000.0) {
        T = 216.65; /* python float (weak) */
        double ex = sub(alt, wk(11000.0, S), S);
        double arg = dvd(mul(wk(-g0, S), ex, S), wk(R * 216.65, S), S);
        Pr = mul(wk(22632.0, S), rn(exp(arg), S), S);                       /* :105-108 */
    } else {
        double ex = sub(alt, wk(20000.0, S), S);
        T = mul(wk(216.65, S), rn(exp(dvd(-ex, wk(10000.0, S), S)), S), S); /* :77-78 */
        double base = 22632.0 * exp(-g0 * 9000.0 / (R * 216.65));           /* get_pressure(20000.0): python floats */
        Pr = mul(wk(base, S), rn(exp(dvd(-ex, wk(6000.0, S), S)), S), S);   /* :111-113 */
    }
    *rho = dvd(Pr, mul(wk(R, S), rn(T, S), S), S);                          /* :166 */
    *cs = sqr(mul(wk(1.4 * R, S), rn(T, S), S), S);                         /* :167-169 */
}

/* MachDragModel.get_drag_force :236-264 (+ get_drag_coefficient :197-220).
 * v: velocity vector at precision pv; rho/cs: scalars at precision pa (or weak if pa_weak).
 * Output force at precision *pf. */
static void mach_drag_force(Env* e, const HlynrParams* P, const double v[3], Prec pv, double rho, double cs,
                            int atm_weak, Prec S, double area, double F[3], Prec* pf) {
    double vm = norm3(v, pv);
    Prec pa = atm_weak ? pv : PMAX(pv, S);
    double mach = dvd(vm, atm_weak ? wk(cs, pv) : cs, pa);
    Prec pm = pa;
    double cd;
    Prec pc;     /* dtype of cd */
    int cd_weak; /* cd is a Python float */
    double base_cd = e->base_cd, peak = e->peak;
    /* thresholds are Python floats: the comparison happens at the dtype of mach */
    if (LT(e, mach, wk(P->sub_mach, pm))) {
        cd = base_cd; cd_weak = !e->cd_strong; pc = F64;
    } else if (LT(e, mach, wk(P->sup_mach, pm))) {
        double frac = dvd(sub(mach, wk(P->sub_mach, pm), pm), wk(P->sup_mach - P->sub_mach, pm), pm);
        if (e->cd_strong) {
            double mult = 1.0 + (peak - 1.0) * frac; /* np.float64 * pm-scalar -> float64 */
            cd = base_cd * mult; pc = F64;
        } else {
            double mult = add(wk(1.0, pm), mul(wk(peak - 1.0, pm), frac, pm), pm);
            cd = mul(wk(base_cd, pm), mult, pm); pc = pm;
        }
        cd_weak = 0;
    } else {
        cd = base_cd * P->sup_mult; cd_weak = !e->cd_strong; pc = F64;
    }
    /* drag_magnitude = 0.5 * air_density * velocity_magnitude**2 * drag_coefficient * reference_area  :258-259 */
    Prec p1 = pa;
    double t = atm_weak ? wk(0.5 * rho, pv) : mul(wk(0.5, S), rho, S);
    t = mul(t, mul(vm, vm, pv), p1);
    Prec p2 = cd_weak ? p1 : PMAX(p1, pc);
    t = mul(t, cd_weak ? wk(cd, p1) : cd, p2);
    t = mul(t, wk(area, p2), p2);
    /* drag_direction = -velocity / velocity_magnitude ; return drag_direction * drag_magnitude */
    Prec p3 = PMAX(pv, p2);
    for (int i = 0; i < 3; ++i) F[i] = mul(dvd(-v[i], vm, pv), t, p3);
    *pf = p3;
}

/* ---------------------------------------------------------------------------------------------- */
/* core.py helpers (float32)                                                                       */
/* ---------------------------------------------------------------------------------------------- */
static void forward_vector(const double q[4], double f[3]) { /* :1143-1152, all float32 */
    double w = q[0], x = q[1], y = q[2], z = q[3];
    f[0] = mul(2.0, add(mul(x, z, F32), mul(w, y, F32), F32), F32);
    f[1] = mul(2.0, sub(mul(y, z, F32), mul(w, x, F32), F32), F32);
    f[2] = sub(1.0, mul(2.0, add(mul(x, x, F32), mul(y, y, F32), F32), F32), F32);
    double n = add(norm3(f, F32), wk(1e-6, F32), F32);
    for (int i = 0; i < 3; ++i) f[i] = dvd(f[i], n, F32);
}
static void quat_to_euler(const double q[4], double eul[3]) { /* :1103-1121, float32 scalars */
    double w = q[0], x = q[1], y = q[2], z = q[3];
    double sinr = mul(2.0, add(mul(w, x, F32), mul(y, z, F32), F32), F32);
    double cosr = sub(1.0, mul(2.0, add(mul(x, x, F32), mul(y, y, F32), F32), F32), F32);
    eul[0] = rn(atan2(sinr, cosr), F32);
    double sinp = mul(2.0, sub(mul(w, y, F32), mul(z, x, F32), F32), F32);
    eul[1] = rn(asin(clipd(sinp, -1.0, 1.0)), F32);
    double siny = mul(2.0, add(mul(w, z, F32), mul(x, y, F32), F32), F32);
    double cosy = sub(1.0, mul(2.0, add(mul(y, y, F32), mul(z, z, F32), F32), F32), F32);
    eul[2] = rn(atan2(siny, cosy), F32);
}

static void right_vector(const double q[4], double r[3]) { /* core.py:1155-1164, all float32 */
    double w = q[0], x = q[1], y = q[2], z = q[3];
    r[0] = sub(1.0, mul(2.0, add(mul(y, y, F32), mul(z, z, F32), F32), F32), F32);
    r[1] = mul(2.0, add(mul(x, y, F32), mul(w, z, F32), F32), F32);
    r[2] = mul(2.0, sub(mul(x, z, F32), mul(w, y, F32), F32), F32);
    double n = add(norm3(r, F32), wk(1e-6, F32), F32);
    for (int i = 0; i < 3; ++i) r[i] = dvd(r[i], n, F32);
}
static void up_vector(const double q[4], double u[3]) { /* core.py:1167-1176, all float32 */
    double w = q[0], x = q[1], y = q[2], z = q[3];
    u[0] = mul(2.0, sub(mul(x, y, F32), mul(w, z, F32), F32), F32);
    u[1] = sub(1.0, mul(2.0, add(mul(x, x, F32), mul(z, z, F32), F32), F32), F32);
    u[2] = mul(2.0, add(mul(y, z, F32), mul(w, x, F32), F32), F32);
    double n = add(norm3(u, F32), wk(1e-6, F32), F32);
    for (int i = 0; i < 3; ++i) u[i] = dvd(u[i], n, F32);
}
/* np.cross for 3-vectors: cp0 = a1*b2 - a2*b1, cp1 = a2*b0 - a0*b2, cp2 = a0*b1 - a1*b0, each product and each
 * difference rounded to the result dtype */
static void cross3(const double a[3], const double b[3], double c[3], Prec p) {
    c[0] = sub(mul(a[1], b[2], p), mul(a[2], b[1], p), p);
    c[1] = sub(mul(a[2], b[0], p), mul(a[0], b[2], p), p);
    c[2] = sub(mul(a[0], b[1], p), mul(a[1], b[0], p), p);
}
/* The LOS basis both the observation (core.py:803-845, :929-945) and the action transform (environment.py:965-1020)
 * build from a relative position: los_unit, los_horizontal = normalize(los_unit x world_up), los_vertical =
 * los_unit x los_horizontal. */
static void los_basis(const double rel[3], double range, Prec p, double lu[3], double lh[3], double lv[3]) {
    const double up[3] = {0.0, 0.0, 1.0};
    if (range > 1e-6) for (int i = 0; i < 3; ++i) lu[i] = dvd(rel[i], range, p);
    else { lu[0] = 1.0; lu[1] = 0.0; lu[2] = 0.0; }
    double lr[3];
    cross3(lu, up, lr, p);
    double n = norm3(lr, p);
    if (n > 1e-6) for (int i = 0; i < 3; ++i) lh[i] = dvd(lr[i], n, p);
    else { lh[0] = 1.0; lh[1] = 0.0; lh[2] = 0.0; }
    cross3(lu, lh, lv, p);
}
/* world_to_body_frame core.py:1179-1205: np.dot against the float32 body axes, result array forced to float32 */
static void to_body(const double v[3], Prec pv, const double fwd[3], const double right[3], const double up[3], double out[3]) {
    out[0] = rn(dot3(v, fwd, pv), F32);
    out[1] = rn(dot3(v, right, pv), F32);
    out[2] = rn(dot3(v, up, pv), F32);
}

/* ---------------------------------------------------------------------------------------------- */
/* Radar26DObservation.compute_radar_detection :511-691 + compute :693-1032 (all observation modes) */
/* ---------------------------------------------------------------------------------------------- */
static void observe(Oracle* o, Env* e, uint32_t step, float obs[26]) {
    const HlynrParams* P = &o->P;
    const HlynrCurriculum* C = &o->C;
    Prec S = o->S;
    double ip[3], iv[3], q[4], mp[3], mv[3];
    for (int i = 0; i < 3; ++i) { ip[i] = rn(e->ipos[i], F32); iv[i] = rn(e->ivel[i], F32); mp[i] = rn(e->mpos[i], F32); mv[i] = rn(e->mvel[i], F32); }
    for (int i = 0; i < 4; ++i) q[i] = rn(e->quat[i], F32);
    double u[4];
    draw_uniform4(&e->rng, step, HLYNR_BLK_UNI, u);

    /* === onboard radar :531-593 === */
    double rel[3];
    for (int i = 0; i < 3; ++i) rel[i] = sub(mp[i], ip[i], F32);
    double range = norm3(rel, F32);
    int onb = 1;
    if (GT(e, range, wk(P->radar_range, F32))) onb = 0;
    double fwd[3], tom[3], right[3] = {0, 0, 0}, up[3] = {0, 0, 0};
    double lu[3] = {1, 0, 0}, lh[3] = {1, 0, 0}, lv[3] = {0, 0, 0};
    int have_los = 0;
    forward_vector(q, fwd);
    if (P->obs_mode == HLYNR_OBS_BODY) { right_vector(q, right); up_vector(q, up); }
    double rden = add(range, wk(1e-6, F32), F32);
    for (int i = 0; i < 3; ++i) tom[i] = dvd(rel[i], rden, F32);
    double beam = rn(acos(clipd(dot3(fwd, tom, F32), -1.0, 1.0)), F32);
    double half_beam = (C->beam_width_deg / 2.0) * (M_PI / 180.0); /* np.radians(python float) -> float64 */
    if (onb) { if (GT(e, beam, half_beam)) onb = 0; }
    if (onb) {
        double rf = sub(1.0, mul(dvd(range, wk(P->radar_range, F32), F32), wk(0.5, F32), F32), F32);
        double qual = mul(mul(wk(P->radar_quality, F32), rf, F32), wk(C->onboard_reliability, F32), F32);
        if (GT(e, u[1], qual)) onb = 0;
    }
    double o_rel[3], o_mv[3];
    int o_det;
    int o_zeros64 = 0; /* np.zeros(3) float64 placeholder while the buffer fills (:581-582) */
    if (e->has_onb) {
        double m[8] = {rel[0], rel[1], rel[2], mv[0], mv[1], mv[2], 0, 0}, out[8];
        int of64, odet;
        if (db_add(&e->onb, m, 0, onb, out, &of64, &odet)) {
            for (int i = 0; i < 3; ++i) { o_rel[i] = out[i]; o_mv[i] = out[3 + i]; }
            o_det = odet;
        } else {
            for (int i = 0; i < 3; ++i) { o_rel[i] = 0.0; o_mv[i] = 0.0; }
            o_det = 0; o_zeros64 = 1;
        }
    } else {
        for (int i = 0; i < 3; ++i) { o_rel[i] = rel[i]; o_mv[i] = mv[i]; }
        o_det = onb;
    }
    (void)o_mv;

    /* === ground radar :368-438 === */
    int gdet = 0;
    double g_rel[3] = {0, 0, 0}, g_vel[3] = {0, 0, 0}, g_q = 0.0;
    int g_f64 = 0;   /* measurement arrays are float64 once noise has been added */
    if (P->ground_enabled) {
        double gp[3] = {P->ground_pos[0], P->ground_pos[1], P->ground_pos[2]}, g2m[3];
        for (int i = 0; i < 3; ++i) g2m[i] = sub(mp[i], gp[i], F32);
        double gr = norm3(g2m, F32);
        int ok = 1;
        if (GT(e, gr, wk(P->g_max_range, F32))) ok = 0;
        if (ok && gr > 1e-6) {
            double el = rn(asin(clipd(dvd(g2m[2], gr, F32), -1.0, 1.0)), F32);
            if (LT(e, el, P->g_min_el)) ok = 0;          /* np.radians() -> float64 threshold */
            else if (GT(e, el, P->g_max_el)) ok = 0;
        }
        if (ok && LT(e, mp[2], wk(50.0, F32))) ok = 0;
        if (ok) {
            double prob = mul(wk(P->g_base_quality, F32),
                              sub(1.0, mul(dvd(gr, wk(P->g_max_range, F32), F32), wk(0.4, F32), F32), F32), F32);
            prob = mul(prob, wk(1.0, F32), F32);                       /* weather_factor */
            prob = mul(prob, wk(C->ground_reliability, F32), F32);
            if (GT(e, u[2], prob)) ok = 0;
            else {
                double zp[4], zv[4];
                draw_normal4(&e->rng, step, HLYNR_BLK_GPOS, zp);
                draw_normal4(&e->rng, step, HLYNR_BLK_GVEL, zv);
                for (int i = 0; i < 3; ++i) {
                    double tr = sub(mp[i], ip[i], F32);
                    g_rel[i] = tr + (0.0 + P->g_sigma_r * zp[i]);            /* float32 + float64 noise */
                    g_vel[i] = sub(mv[i], iv[i], F32) + (0.0 + P->g_sigma_v * zv[i]);
                }
                g_q = prob; g_f64 = 1; gdet = 1;
            }
        }
    }
    double dg_rel[3], dg_vel[3], dg_q;
    int dg_det, dg_f64;
    if (P->ground_enabled && P->ground_delay > 0) { /* :609-627: delayed values, CURRENT flag (quirk Q3) */
        double m[8] = {g_rel[0], g_rel[1], g_rel[2], g_vel[0], g_vel[1], g_vel[2], g_q, 0}, out[8];
        int of64, odet;
        if (db_add(&e->gnd, m, g_f64, gdet, out, &of64, &odet)) {
            for (int i = 0; i < 3; ++i) { dg_rel[i] = out[i]; dg_vel[i] = out[3 + i]; }
            dg_q = out[6]; dg_det = gdet; dg_f64 = of64;
        } else {
            for (int i = 0; i < 3; ++i) { dg_rel[i] = 0.0; dg_vel[i] = 0.0; }
            dg_q = 0.0; dg_det = 0; dg_f64 = 1;
        }
    } else {
        for (int i = 0; i < 3; ++i) { dg_rel[i] = g_rel[i]; dg_vel[i] = g_vel[i]; }
        dg_q = g_q; dg_det = gdet; dg_f64 = g_f64;
    }

    /* === datalink :440-474 === */
    double link = 0.0;
    if (P->ground_enabled) {
        double d[3];
        for (int i = 0; i < 3; ++i) d[i] = sub(ip[i], P->ground_pos[i], F32);
        double lr = norm3(d, F32);
        if (!GT(e, lr, wk(P->max_datalink_range, F32))) {
            double r1 = dvd(lr, wk(P->max_datalink_range, F32), F32);
            double rf = sub(1.0, mul(r1, r1, F32), F32);
            double vm = norm3(iv, F32);
            double dv = dvd(vm, wk(1000.0, F32), F32);
            double dop = sub(1.0, (dv < wk(0.3, F32)) ? dv : wk(0.3, F32), F32);
            if (LT(e, u[3], wk(P->datalink_packet_loss, F32))) link = 0.0;
            else link = clipd(mul(mul(rf, dop, F32), wk(0.95, F32), F32), 0.0, 1.0);
        }
    }

    /* === fusion :476-509 === */
    double o_q = o_det ? P->radar_quality : 0.0;
    double fus;
    if (!o_det && !dg_det) fus = 0.0;
    else if (o_det && !dg_det) fus = o_q * 0.5;
    else if (!o_det) fus = mul(dg_q, wk(0.6, F32), F32);
    else {
        Prec pe = dg_f64 ? F64 : F32;
        double dd[3];
        for (int i = 0; i < 3; ++i) dd[i] = sub(o_rel[i], dg_rel[i], pe);
        double perr = norm3(dd, pe);
        double r = dvd(perr, wk(200.0, pe), pe);
        double agr = sub(1.0, r < 1.0 ? r : 1.0, pe);
        double t = add(wk(0.35 * o_q, F32), mul(wk(0.5, F32), dg_q, F32), F32);
        fus = clipd(add(t, mul(wk(0.15, pe), agr, pe), pe), 0.0, 1.0);
    }
    e->last_onb_det = o_det;
    e->last_onb_fill = o_zeros64;
    e->last_gnd_det = dg_det;

    /* === compute() :693-1032 === */
    for (int i = 0; i < 26; ++i) obs[i] = 0.f;
    Kalman* kf = &e->kf;
    int meas_avail = 0;
    double frp[3] = {0, 0, 0}, frv[3] = {0, 0, 0};
    Prec pk;
    if (o_det || dg_det) {
        double z[3];
        int z_f64;
        if (o_det && dg_det) {
            /* fused = (onboard*ow + ground*gw) / total  :736-739; ow python float, gw np.float32 (or python 0.0) */
            double ow = P->radar_quality, gw = dg_q;
            double tw = add(wk(ow, F32), gw, F32);
            Prec pg = dg_f64 ? F64 : F32;
            for (int i = 0; i < 3; ++i) {
                double a = mul(o_rel[i], wk(ow, F32), F32);
                double b = mul(dg_rel[i], gw, pg);
                z[i] = add(ip[i], dvd(add(a, b, pg), tw, pg), pg);
            }
            z_f64 = dg_f64;
        } else if (o_det) {
            for (int i = 0; i < 3; ++i) z[i] = add(ip[i], o_rel[i], F32);
            z_f64 = 0;
        } else {
            Prec pg = dg_f64 ? F64 : F32;
            for (int i = 0; i < 3; ++i) z[i] = add(ip[i], dg_rel[i], pg);
            z_f64 = dg_f64;
        }
        kf_update(kf, z, z_f64);
        meas_avail = 1;
    } else {
        kf_predict(kf);
    }
    pk = kf->x_f64 ? F64 : F32;
    if (meas_avail || kf->initialized) {
        for (int i = 0; i < 3; ++i) { frp[i] = sub(kf->x[i], ip[i], pk); frv[i] = sub(kf->x[3 + i], iv[i], pk); }
        double rr = norm3(frp, pk);
        double cl = dvd(-dot3(frp, frv, pk), add(rr, wk(1e-6, pk), pk), pk);
        if (P->obs_mode == HLYNR_OBS_LOS) { /* core.py:791-872 */
            have_los = 1;
            obs[0] = (float)clipd(dvd(rr, wk(P->max_range, pk), pk), 0.0, 1.0);
            obs[1] = (float)clipd(dvd(cl, wk(P->max_velocity, pk), pk), -1.0, 1.0);
            los_basis(frp, rr, pk, lu, lh, lv);
            double rate[3];
            double rden2 = add(rr, wk(1e-6, pk), pk);
            for (int i = 0; i < 3; ++i) rate[i] = dvd(sub(frv[i], mul(cl, lu[i], pk), pk), rden2, pk);
            obs[2] = (float)clipd(dvd(dot3(rate, lh, pk), wk(0.5, pk), pk), -1.0, 1.0);
            obs[3] = (float)clipd(dvd(dot3(rate, lv, pk), wk(0.5, pk), pk), -1.0, 1.0);
            double ivm = norm3(iv, F32);
            if (ivm > 1e-6) {
                double un[3];
                for (int i = 0; i < 3; ++i) un[i] = dvd(iv[i], ivm, F32);
                obs[4] = (float)dot3(un, lu, pk);
            } else obs[4] = 0.f;
            double tv[3];
            for (int i = 0; i < 3; ++i) tv[i] = add(frv[i], iv[i], pk);
            double tvm = norm3(tv, pk);
            if (tvm > 1e-6) {
                double un[3], nl[3];
                for (int i = 0; i < 3; ++i) { un[i] = dvd(tv[i], tvm, pk); nl[i] = -lu[i]; }
                obs[5] = (float)dot3(un, nl, pk);
            } else obs[5] = 0.f;
        } else if (P->obs_mode == HLYNR_OBS_BODY) { /* core.py:874-880 */
            double bp[3], bv[3];
            to_body(frp, pk, fwd, right, up, bp);
            to_body(frv, pk, fwd, right, up, bv);
            for (int i = 0; i < 3; ++i) {
                obs[i] = (float)clipd(dvd(bp[i], wk(P->max_range, F32), F32), -1.0, 1.0);
                obs[3 + i] = (float)clipd(dvd(bv[i], wk(P->max_velocity, F32), F32), -1.0, 1.0);
            }
        } else
        for (int i = 0; i < 3; ++i) {
            obs[i] = (float)clipd(dvd(frp[i], wk(P->max_range, pk), pk), -1.0, 1.0);
            obs[3 + i] = (float)clipd(dvd(frv[i], wk(P->max_velocity, pk), pk), -1.0, 1.0);
        }
        if (GT(e, cl, 0.0)) {
            double tti = dvd(rr, cl, pk);
            obs[13] = (float)clipd(sub(1.0, dvd(tti, wk(100.0, pk), pk), pk), -1.0, 1.0);
        } else obs[13] = -1.f;
        float tr = (kf->P[0][0] + kf->P[1][1]) + kf->P[2][2];       /* np.trace float32 -> python float */
        double tq = clipd(1.0 - (double)tr / 10000.0, 0.0, 1.0);   /* float64 */
        if (o_det) tq *= P->radar_quality;
        obs[14] = (float)tq;
        obs[15] = (float)clipd(dvd(cl, wk(P->max_velocity, pk), pk), -1.0, 1.0);
        if (rr > 1e-6) {
            double tt[3];
            for (int i = 0; i < 3; ++i) tt[i] = dvd(frp[i], rr, pk);
            obs[16] = (float)dot3(fwd, tt, pk);
        } else obs[16] = 1.f;
    } else {
        for (int i = 0; i < 6; ++i) obs[i] = -2.f;
        obs[13] = -1.f; obs[14] = 0.f; obs[15] = 0.f; obs[16] = 0.f;
    }
    if (P->obs_mode == HLYNR_OBS_LOS) { /* core.py:920-958 */
        double sp = norm3(iv, F32);
        obs[6] = (float)clipd(dvd(sp, wk(P->max_velocity, F32), F32), 0.0, 1.0);
        if (have_los) {
            obs[7] = (float)clipd(dvd(dot3(iv, lh, pk), wk(P->max_velocity, pk), pk), -1.0, 1.0);
            obs[8] = (float)clipd(dvd(dot3(iv, lv, pk), wk(P->max_velocity, pk), pk), -1.0, 1.0);
        } else { obs[7] = 0.f; obs[8] = 0.f; }
    } else if (P->obs_mode == HLYNR_OBS_BODY) { /* :959-962 */
        double bv[3];
        to_body(iv, F32, fwd, right, up, bv);
        for (int i = 0; i < 3; ++i) obs[6 + i] = (float)clipd(dvd(bv[i], wk(P->max_velocity, F32), F32), -1.0, 1.0);
    } else
    for (int i = 0; i < 3; ++i) obs[6 + i] = (float)clipd(dvd(iv[i], wk(P->max_velocity, F32), F32), -1.0, 1.0);
    if (P->obs_mode == HLYNR_OBS_WORLD) { /* :967-974; zeros in the other modes */
        double eul[3];
        quat_to_euler(q, eul);
        for (int i = 0; i < 3; ++i) obs[9 + i] = (float)dvd(eul[i], wk(M_PI, F32), F32);
    }
    {
        Prec pf = e->steps == 0 && !e->started ? F64 : S; /* fuel is a Python float until the first tick */
        obs[12] = (float)clipd(dvd(e->fuel, wk(100.0, pf), pf), 0.0, 1.0);
    }
    if (dg_det && GT(e, link, wk(0.1, F32))) {
        Prec pg = dg_f64 ? F64 : F32;
        if (P->obs_mode == HLYNR_OBS_LOS) { /* core.py:985-1006 */
            double gr = norm3(dg_rel, pg);
            double gc = dvd(-dot3(dg_rel, dg_vel, pg), add(gr, wk(1e-6, pg), pg), pg);
            obs[17] = (float)clipd(dvd(gr, wk(P->max_range, pg), pg), 0.0, 1.0);
            obs[18] = (float)clipd(dvd(gc, wk(P->max_velocity, pg), pg), -1.0, 1.0);
            if (gr > 1e-6) {
                double tv[3];
                for (int i = 0; i < 3; ++i) tv[i] = sub(dg_vel[i], mul(gc, dvd(dg_rel[i], gr, pg), pg), pg);
                obs[19] = (float)clipd(dvd(dvd(norm3(tv, pg), gr, pg), wk(0.5, pg), pg), 0.0, 1.0);
            } else obs[19] = 0.f;
            obs[20] = obs[21] = obs[22] = 0.f;
        } else if (P->obs_mode == HLYNR_OBS_BODY) { /* :1008-1012 */
            double bp[3], bv[3];
            to_body(dg_rel, pg, fwd, right, up, bp);
            to_body(dg_vel, pg, fwd, right, up, bv);
            for (int i = 0; i < 3; ++i) {
                obs[17 + i] = (float)clipd(dvd(bp[i], wk(P->max_range, F32), F32), -1.0, 1.0);
                obs[20 + i] = (float)clipd(dvd(bv[i], wk(P->max_velocity, F32), F32), -1.0, 1.0);
            }
        } else
        for (int i = 0; i < 3; ++i) {
            obs[17 + i] = (float)clipd(dvd(dg_rel[i], wk(P->max_range, pg), pg), -1.0, 1.0);
            obs[20 + i] = (float)clipd(dvd(dg_vel[i], wk(P->max_velocity, pg), pg), -1.0, 1.0);
        }
        obs[23] = (float)dg_q;
    } else {
        for (int i = 17; i < 23; ++i) obs[i] = -2.f;
        obs[23] = 0.f;
    }
    obs[24] = (float)link;
    obs[25] = (float)fus;
}

/* ---------------------------------------------------------------------------------------------- */
/* reset :353-603                                                                                  */
/* ---------------------------------------------------------------------------------------------- */
static void env_reset(Oracle* o, Env* e, float obs[26]) {
    const HlynrParams* P = &o->P;
    if (e->started) e->rng.episode++;
    /* observation_generator.reset_sensor_delays() :362 */
    if (e->has_onb) db_reset(&e->onb);
    db_reset(&e->gnd);
    kf_reset(&e->kf);
    double u0[4], u1[4], u2[4];
    draw_uniform4(&e->rng, 0, HLYNR_BLK_SPAWN0, u0);
    draw_uniform4(&e->rng, 0, HLYNR_BLK_SPAWN1, u1);
    draw_uniform4(&e->rng, 0, HLYNR_BLK_SPAWN2, u2);
    double tgt[3] = {P->target[0], P->target[1], P->target[2]};
    /* missile(s) :386-435 */
    const int K = P->volley_size > 0 ? P->volley_size : 1;
    for (int m = 0; m < K; ++m) {
        double um[4], mp[3], mv[3];
        if (m == 0) memcpy(um, u0, sizeof(um));
        else draw_uniform4(&e->rng, 0, HLYNR_BLK_VSPAWN(m), um);
        if (P->m_spawn_spherical) {
            double radius = P->m_radius_lo + (P->m_radius_hi - P->m_radius_lo) * um[0];
            double az = (P->m_az_lo + (P->m_az_hi - P->m_az_lo) * um[1]) * M_PI / 180.0;
            double el = (P->m_el_lo + (P->m_el_hi - P->m_el_lo) * um[2]) * M_PI / 180.0;
            double x = radius * cos(el) * cos(az), y = radius * cos(el) * sin(az), z = radius * sin(el);
            mp[0] = rn(tgt[0] + x, F32); mp[1] = rn(tgt[1] + y, F32); mp[2] = rn(tgt[2] + z, F32);
        } else {
            for (int i = 0; i < 3; ++i) mp[i] = rn(P->m_pos_lo[i] + (P->m_pos_hi[i] - P->m_pos_lo[i]) * um[i], F32);
        }
        double speed = P->m_speed_lo + (P->m_speed_hi - P->m_speed_lo) * um[3];
        double tt[3];
        for (int i = 0; i < 3; ++i) tt[i] = sub(tgt[i], mp[i], F32);
        double td = norm3(tt, F32);
        for (int i = 0; i < 3; ++i) mv[i] = mul(dvd(tt[i], td, F32), wk(speed, F32), F32); /* td > 1e-6 always */
        if (P->volley_size > 0) for (int i = 0; i < 3; ++i) { e->vpos[m][i] = mp[i]; e->vvel[m][i] = mv[i]; }
        if (m == 0) for (int i = 0; i < 3; ++i) { e->mpos[i] = mp[i]; e->mvel[i] = mv[i]; } /* self.missile_state = missile_states[0] :438 */
    }
    /* interceptor :442-467 */
    for (int i = 0; i < 3; ++i) e->ipos[i] = rn(P->i_pos_lo[i] + (P->i_pos_hi[i] - P->i_pos_lo[i]) * u1[i], F32);
    if (P->i_vel_toward_missile) {
        double tm[3];
        for (int i = 0; i < 3; ++i) tm[i] = sub(e->mpos[i], e->ipos[i], F32);
        double d = norm3(tm, F32);
        double sp = P->i_speed_lo + (P->i_speed_hi - P->i_speed_lo) * u1[3];
        for (int i = 0; i < 3; ++i) e->ivel[i] = mul(dvd(tm[i], d, F32), wk(sp, F32), F32);
    } else {
        double uu[3] = {u1[3], u2[0], u2[1]};
        for (int i = 0; i < 3; ++i) e->ivel[i] = rn(P->i_vel_lo[i] + (P->i_vel_hi[i] - P->i_vel_lo[i]) * uu[i], F32);
    }
    /* initial orientation :475-530: rotate +Z onto the line of sight; float64 math on float32 inputs */
    double rv[3];
    for (int i = 0; i < 3; ++i) rv[i] = sub(e->mpos[i], e->ipos[i], F32);
    if (P->volley_size > 0) { /* :470-486: min distances of every missile, point at the closest one (first wins ties) */
        double best = INFINITY;
        int bm = 0;
        for (int m = 0; m < K; ++m) {
            double d3[3];
            for (int i = 0; i < 3; ++i) d3[i] = sub(e->vpos[m][i], e->ipos[i], F32);
            double d = norm3(d3, F32);
            e->vmin[m] = d; e->vact[m] = 1;
            if (d < best) { best = d; bm = m; }
        }
        for (int i = 0; i < 3; ++i) rv[i] = sub(e->vpos[bm][i], e->ipos[i], F32);
        e->vcur = 0; e->vcount = 0;
    }
    double rd = norm3(rv, F32);
    if (rd > 1e-6) {
        double f[3];
        for (int i = 0; i < 3; ++i) f[i] = dvd(rv[i], rd, F32);
        double ax[3] = {0.0 * f[2] - 1.0 * f[1], 1.0 * f[0] - 0.0 * f[2], 0.0};
        double al = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
        double ca = f[2];
        if (al > 1e-6) {
            for (int i = 0; i < 3; ++i) ax[i] /= al;
            double ang = acos(clipd(ca, -1.0, 1.0)), h = ang / 2.0, sh = sin(h);
            e->quat[0] = rn(cos(h), F32); e->quat[1] = rn(ax[0] * sh, F32); e->quat[2] = rn(ax[1] * sh, F32); e->quat[3] = rn(ax[2] * sh, F32);
        } else if (ca > 0) { /* quirk Q9: the reference raises UnboundLocalError here; aligned -> identity */
            e->quat[0] = 1; e->quat[1] = e->quat[2] = e->quat[3] = 0;
        } else {
            e->quat[0] = 0; e->quat[1] = 1; e->quat[2] = e->quat[3] = 0;
        }
    } else {
        e->quat[0] = 1; e->quat[1] = e->quat[2] = e->quat[3] = 0;
    }
    e->fuel = 100.0;
    for (int i = 0; i < 3; ++i) { e->wind[i] = P->base_wind[i]; e->thrust[i] = 0.0; }
    e->wind_f64 = 0;
    /* domain randomization :552-562 */
    if (P->dr_enabled) {
        double z[16];
        for (int b = 0; b < 4; ++b) draw_normal4(&e->rng, 0, HLYNR_BLK_DR0 + b, z + 4 * b);
        double val[HLYNR_N_DR];
        for (int i = 0; i < HLYNR_N_DR; ++i) {
            if (i == 1) { val[i] = 0.0 + P->dr_variation[1] * z[1]; continue; }  /* temperature_offset :169 */
            double lo = 0.1, hi = 3.0;
            if (i == 6) { lo = 0.5; hi = 1.0; }
            if (i == 7) { lo = 0.5; hi = 2.0; }
            val[i] = clipd(1.0 + P->dr_variation[i] * z[i], lo, hi);             /* :238-241 */
        }
        if (P->isa_enabled) e->T0 = e->T0 + val[1];                              /* :258-261 compounding */
        if (P->mach_enabled) { e->base_cd = 0.3 * val[2]; e->peak = 3.0 * val[3]; e->cd_strong = 1; } /* :271-280 */
        if (e->has_onb) {                                                        /* :290-297 */
            int nd = (int)(3 * val[4]);
            if (nd > 10) nd = 10;
            if (nd < 1) nd = 1;
            db_init(&e->onb, nd);
        }
    }
    e->steps = 0; e->fuel_used = 0.0;
    e->started = 0; /* fuel is a python float for the initial observation */
    observe(o, e, 0, obs);
    e->started = 1;
    double dd[3];
    for (int i = 0; i < 3; ++i) dd[i] = sub(e->mpos[i], e->ipos[i], F32);
    e->prev_d = norm3(dd, F32); /* :579-581, state arrays are float32 at this point in both modes */
    e->worsen = 0; e->last_d = e->prev_d; e->min_d = e->prev_d; e->crossed = 0;
    e->ep_return = 0.0; e->ep_length = 0;
}

/* ---------------------------------------------------------------------------------------------- */
/* step :605-859                                                                                   */
/* ---------------------------------------------------------------------------------------------- */
/* _update_missile_state :1069-1117 for one missile (same current_wind for all of them) */
static void missile_physics(Oracle* o, Env* e, uint32_t step, uint32_t evade_blk, double mpos[3], double mvel[3]) {
    const HlynrParams* P = &o->P;
    Prec S = o->S;
    double dt = P->dt;
        double alt = mpos[2] > 0.0 ? mpos[2] : 0.0;
        double rho = 1.225, cs = 343.0;
        int atm_weak = 1;
        if (P->isa_enabled) { isa_props(e, alt, S, &rho, &cs); atm_weak = 0; }
        Prec pw = e->wind_f64 ? F64 : F32;
        Prec pv = PMAX(S, pw);
        double va[3];
        for (int i = 0; i < 3; ++i) va[i] = sub(mvel[i], e->wind[i], pv);
        double vmag = norm3(va, pv);
        double da[3];
        Prec pd;
        if (P->mach_enabled && vmag > 1e-6) {
            double F[3];
            mach_drag_force(e, P, va, pv, rho, cs, atm_weak, S, 2.0, F, &pd);
            double ratio = (0.3 * 1.5) / 0.3; /* python floats :1090,1095 */
            for (int i = 0; i < 3; ++i) da[i] = dvd(mul(F[i], wk(ratio, pd), pd), wk(1000.0, pd), pd);
        } else {
            pd = pv;
            double c = atm_weak ? wk(-0.5 * 0.3 * rho, pv) : mul(wk(-0.5 * 0.3, S), rho, S);
            c = mul(c, vmag, pv);
            for (int i = 0; i < 3; ++i) da[i] = dvd(mul(c, va[i], pv), wk(1000.0, pv), pv);
        }
        double ev[3] = {0, 0, 0};
        if (P->evasion_enabled) {
            double z[4];
            draw_normal4(&e->rng, step, evade_blk, z);
            for (int i = 0; i < 3; ++i) ev[i] = z[i] * 2.0;
        }
        double g[3] = {0.0, 0.0, (double)(-9.81f)};
        double acc[3];
        Prec p1 = PMAX(pd, F32);
        for (int i = 0; i < 3; ++i) acc[i] = add(add(da[i], g[i], p1), ev[i], F64); /* evasion = np.zeros(3): float64 */
        if (P->validate_enabled && !(isfinite(acc[0]) && isfinite(acc[1]) && isfinite(acc[2])))
            for (int i = 0; i < 3; ++i) acc[i] = isnan(acc[i]) ? 0.0 : (isinf(acc[i]) ? (acc[i] > 0 ? 20.0 : -20.0) : acc[i]);
        for (int i = 0; i < 3; ++i) {
            mvel[i] = rn(mvel[i] + acc[i] * dt, S);
            mpos[i] = add(mpos[i], mul(mvel[i], wk(dt, S), S), S);
        }
}

typedef struct {
    float obs[26];
    double reward;
    int terminated, truncated;
    double distance;
    int intercepted, hit_target, clamped, fuze;
} StepOut;

static void env_step(Oracle* o, Env* e, const float act[6], StepOut* out) {
    const HlynrParams* P = &o->P;
    const HlynrCurriculum* C = &o->C;
    Prec S = o->S;
    e->margin = 1e30;
    e->steps += 1;
    uint32_t step = (uint32_t)e->steps;
    double a[6];
    for (int i = 0; i < 6; ++i) a[i] = (double)act[i];
    if (P->obs_mode == HLYNR_OBS_LOS) { /* _update_los_frame + _transform_los_action_to_world, environment.py:965-1061 */
        double rel[3], lu[3], lh[3], lv[3];
        for (int i = 0; i < 3; ++i) rel[i] = sub(rn(e->mpos[i], F32), rn(e->ipos[i], F32), F32); /* float32 copies of the state */
        los_basis(rel, norm3(rel, F32), F32, lu, lh, lv);
        double w[3]; /* action scalar (float32, or float64 in the up-cast oracle) times float32 basis vectors */
        for (int i = 0; i < 3; ++i)
            w[i] = add(add(mul(a[0], lu[i], S), mul(a[1], lh[i], S), S), mul(a[2], lv[i], S), S);
        for (int i = 0; i < 3; ++i) a[i] = w[i];
    }
    /* SafetyClamp.apply core.py:1069-1100 */
    int clamped = 0;
    if (e->fuel <= 0) { a[0] = a[1] = a[2] = 0.0; clamped = 1; }
    double am = norm3(a, S);
    if (am > wk(50.0, S)) { double f = dvd(wk(50.0, S), am, S); for (int i = 0; i < 3; ++i) a[i] = mul(a[i], f, S); clamped = 1; }
    double gm = norm3(a + 3, S);
    if (gm > wk(5.0, S)) { double f = dvd(wk(5.0, S), gm, S); for (int i = 3; i < 6; ++i) a[i] = mul(a[i], f, S); clamped = 1; }

    double dt = P->dt;
    /* ---- _update_interceptor :861-963 ---- */
    {
        double Tdes[3], w[3], T[3];
        for (int i = 0; i < 3; ++i) { Tdes[i] = mul(a[i], wk(10000.0, S), S); w[i] = mul(a[3 + i], wk(20.0, S), S); }
        if (P->thrust_dyn_enabled) {
            for (int i = 0; i < 3; ++i) {
                double err = sub(Tdes[i], e->thrust[i], S);
                e->thrust[i] = add(e->thrust[i], dvd(mul(err, wk(dt, S), S), wk(P->thrust_tau, S), S), S);
                T[i] = e->thrust[i];
            }
        } else for (int i = 0; i < 3; ++i) T[i] = Tdes[i];
        double tm = norm3(T, S);
        Prec pf = S; /* python float 100.0 - np scalar -> S */
        double burn = mul(mul(dvd(tm, wk(500.0, S), S), wk(0.1, S), S), wk(dt, S), S);
        e->fuel = sub(e->fuel, burn, pf);
        e->fuel_used = add(e->fuel_used, burn, S);
        if (LE(e, e->fuel, 0.0)) {
            e->fuel = 0.0;
            for (int i = 0; i < 3; ++i) { T[i] = mul(T[i], 0.0, S); if (P->thrust_dyn_enabled) e->thrust[i] = mul(e->thrust[i], 0.0, S); }
        }
        double ta[3];
        for (int i = 0; i < 3; ++i) ta[i] = dvd(T[i], wk(500.0, S), S);
        double alt = e->ipos[2] > 0.0 ? e->ipos[2] : 0.0;
        double rho = 1.225, cs = 343.0;
        int atm_weak = 1;
        if (P->isa_enabled) { isa_props(e, alt, S, &rho, &cs); atm_weak = 0; }
        Prec pw = e->wind_f64 ? F64 : F32;
        Prec pv = PMAX(S, pw);
        double va[3];
        for (int i = 0; i < 3; ++i) va[i] = sub(e->ivel[i], e->wind[i], pv);
        double da[3];
        Prec pd;
        double vmag = norm3(va, pv);
        if (P->mach_enabled && vmag > 1e-6) {
            double F[3];
            mach_drag_force(e, P, va, pv, rho, cs, atm_weak, S, 1.0, F, &pd);
            for (int i = 0; i < 3; ++i) da[i] = dvd(F[i], wk(500.0, pd), pd);
        } else { /* :920-921  -0.5 * cd * rho * |v| * v / mass */
            pd = pv;
            double c = atm_weak ? wk(-0.5 * 0.3 * rho, pv) : mul(wk(-0.5 * 0.3, S), rho, S);
            c = mul(c, vmag, pv);
            for (int i = 0; i < 3; ++i) da[i] = dvd(mul(c, va[i], pv), wk(500.0, pv), pv);
        }
        Prec pt = PMAX(S, pd);
        double g[3] = {0.0, 0.0, (double)(-9.81f)};
        double acc[3];
        for (int i = 0; i < 3; ++i) acc[i] = add(add(ta[i], da[i], pt), g[i], pt);
        if (P->validate_enabled && !(isfinite(acc[0]) && isfinite(acc[1]) && isfinite(acc[2])))
            for (int i = 0; i < 3; ++i) acc[i] = isnan(acc[i]) ? 0.0 : (isinf(acc[i]) ? (acc[i] > 0 ? 50.0 : -50.0) : acc[i]);
        for (int i = 0; i < 3; ++i) {
            e->ivel[i] = rn(add(e->ivel[i], mul(acc[i], wk(dt, pt), pt), pt), S);
            e->ipos[i] = add(e->ipos[i], mul(e->ivel[i], wk(dt, S), S), S);
        }
        double wn = norm3(w, S);
        double ang = mul(wn, wk(dt, S), S);
        if (ang > wk(1e-6, S)) {
            double h = dvd(ang, 2.0, S);
            double ch = rn(cos(h), S), sh = rn(sin(h), S);
            double dq[4] = {ch, mul(dvd(w[0], wn, S), sh, S), mul(dvd(w[1], wn, S), sh, S), mul(dvd(w[2], wn, S), sh, S)};
            double* q2 = e->quat;
            double w1 = dq[0], x1 = dq[1], y1 = dq[2], z1 = dq[3], w2 = q2[0], x2 = q2[1], y2 = q2[2], z2 = q2[3];
            double nq[4];
            nq[0] = rn(sub(sub(sub(mul(w1, w2, S), mul(x1, x2, S), S), mul(y1, y2, S), S), mul(z1, z2, S), S), F32);
            nq[1] = rn(sub(add(add(mul(w1, x2, S), mul(x1, w2, S), S), mul(y1, z2, S), S), mul(z1, y2, S), S), F32);
            nq[2] = rn(add(add(sub(mul(w1, y2, S), mul(x1, z2, S), S), mul(y1, w2, S), S), mul(z1, x2, S), S), F32);
            nq[3] = rn(add(sub(add(mul(w1, z2, S), mul(x1, y2, S), S), mul(y1, x2, S), S), mul(z1, w2, S), S), F32);
            double nn = sqr(dotn(nq, nq, 4, F32), F32);
            for (int i = 0; i < 4; ++i) e->quat[i] = dvd(nq[i], nn, F32);
        }
    }
    /* ---- missiles :631-638 ---- */
    if (P->volley_size > 0) {
        for (int m = 0; m < P->volley_size; ++m)
            if (e->vact[m]) missile_physics(o, e, step, m == 0 ? HLYNR_BLK_EVADE : HLYNR_BLK_VEVADE(m), e->vpos[m], e->vvel[m]);
    } else missile_physics(o, e, step, HLYNR_BLK_EVADE, e->mpos, e->mvel);
    /* ---- _update_wind :1119-1129 ---- */
    if (P->enh_wind_enabled) { /* EnhancedWindModel.get_wind_vector physics_models.py:351-387 */
        double alt = e->ipos[2] > 0.0 ? e->ipos[2] : 0.0;
        double u[4];
        draw_uniform4(&e->rng, step, HLYNR_BLK_UNI, u);
        double pf;
        int pf_weak = 0;
        double a2 = alt;
        int alt_weak = !(e->ipos[2] > 0.0);
        if (LE(e, a2, 10.0)) { a2 = 10.0; alt_weak = 1; }
        if (LE(e, a2, P->blh)) {
            if (alt_weak) { pf = pow(a2 / 10.0, 0.143); pf_weak = 1; }
            else pf = rn(pow(dvd(a2, wk(10.0, S), S), wk(0.143, S)), S);
        } else { pf = pow(P->blh / 10.0, 0.143); pf_weak = 1; }
        Prec pwv = pf_weak ? F32 : PMAX(F32, S);
        double wv[3];
        for (int i = 0; i < 3; ++i) wv[i] = mul(P->base_wind[i], pf_weak ? wk(pf, F32) : pf, pwv);
        double ti;
        int ti_weak = 0;
        if (alt <= 10.0) { ti = P->turb_intensity * 2.0; ti_weak = 1; }
        else if (alt <= P->blh) {
            double hf = sub(1.0, mul(dvd(alt, wk(P->blh, S), S), wk(0.7, S), S), S);
            ti = mul(wk(P->turb_intensity, S), hf, S);
        } else { ti = P->turb_intensity * 0.3; ti_weak = 1; }
        if (ti > 0) {
            double wn = norm3(wv, pwv);
            Prec ps = ti_weak ? pwv : PMAX(S, pwv);
            double scale = mul(ti_weak ? wk(ti, pwv) : ti, wn, ps);
            double z[4];
            draw_normal4(&e->rng, step, HLYNR_BLK_WIND, z);
            double lp = 1.0 - exp(-dt / 0.1);
            for (int i = 0; i < 3; ++i) {
                double tb = (0.0 + scale * z[i]) * lp;
                wv[i] = rn(wv[i] + tb, pwv);
            }
        }
        if (LT(e, u[0], 0.001)) {
            double z[4];
            draw_normal4(&e->rng, step, HLYNR_BLK_GUST_DIR, z);
            double nn = sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]) + 1e-6;
            double mag = P->gust_scale * draw_exp(&e->rng, step, HLYNR_BLK_GUST_MAG);
            for (int i = 0; i < 3; ++i) wv[i] = rn(wv[i] + (z[i] / nn) * mag, pwv);
        }
        for (int i = 0; i < 3; ++i) e->wind[i] = rn(wv[i], F32);
        e->wind_f64 = 0;
    } else if (P->wind_variability > 0) {
        double z[4];
        draw_normal4(&e->rng, step, HLYNR_BLK_WIND, z);
        Prec pw = e->wind_f64 ? F64 : F32;
        for (int i = 0; i < 3; ++i) {
            double ch = z[i] * P->wind_variability;
            e->wind[i] = mul(wk(0.95, pw), e->wind[i], pw) + 0.05 * (P->base_wind[i] + ch);
        }
        e->wind_f64 = 1;
    }
    /* ---- distance / intercept / termination :640-814 ---- */
    double dist;
    int intercepted = 0, term = 0, trunc = 0, hit = 0, fuze = 0;
    double radius = P->fuze_enabled ? P->kill_radius : C->intercept_radius;
    if (P->volley_size > 0) {
        const int K = P->volley_size;
        /* _select_priority_missile :236-267: closest ACTIVE missile (strict <, first wins), else missile_states[0] */
        double dm[HLYNR_MAX_VOLLEY];
        int pri = -1;
        double pbest = 0.0;
        for (int m = 0; m < K; ++m) {
            double d3[3];
            for (int i = 0; i < 3; ++i) d3[i] = sub(e->vpos[m][i], e->ipos[i], S);
            dm[m] = norm3(d3, S);
            if (e->vact[m] && (pri < 0 || dm[m] < pbest)) { pri = m; pbest = dm[m]; }
        }
        if (pri < 0) pri = 0;
        e->vcur = pri;
        /* :661-692: per-missile min distance, interception (missile becomes inactive), distance = closest still active */
        int any_active = 0;
        dist = 0.0;
        for (int m = 0; m < K; ++m) {
            if (!e->vact[m]) continue;
            if (dm[m] < e->vmin[m]) e->vmin[m] = dm[m];
            if (LT(e, dm[m], wk(radius, S))) { intercepted = 1; e->vcount += 1; e->vact[m] = 0; }
        }
        for (int m = 0; m < K; ++m)
            if (e->vact[m] && (!any_active || dm[m] < dist)) { dist = dm[m]; any_active = 1; }
        if (dist < e->min_d) e->min_d = dist;
        if (intercepted && !e->crossed) e->crossed = 1;
        if (P->fuze_enabled && LT(e, e->min_d, wk(P->kill_radius, S))) { fuze = 1; intercepted = 1; }
        /* :724-748: every missile at or below the ground becomes inactive; near the target = mission failure */
        int all_inactive = 1;
        for (int m = 0; m < K; ++m) {
            if (LE(e, e->vpos[m][2], 0.0)) {
                e->vact[m] = 0;
                double gx = sub(e->vpos[m][0], P->target[0], S), gy = sub(e->vpos[m][1], P->target[1], S);
                double gxy[2] = {gx, gy};
                if (LT(e, sqr(dotn(gxy, gxy, 2, S), S), wk(500.0, S))) hit = 1;
            }
            if (e->vact[m]) all_inactive = 0;
        }
        if (all_inactive) term = 1;
        else if (fuze) term = 1;
        for (int i = 0; i < 3; ++i) { e->mpos[i] = e->vpos[pri][i]; e->mvel[i] = e->vvel[pri][i]; } /* self.missile_state = priority */
    } else {
    double dd[3];
    for (int i = 0; i < 3; ++i) dd[i] = sub(e->mpos[i], e->ipos[i], S);
    dist = norm3(dd, S);
    intercepted = LT(e, dist, wk(radius, S));
    if (dist < e->min_d) e->min_d = dist;
    if (intercepted && !e->crossed) e->crossed = 1;
    if (P->fuze_enabled && LT(e, e->min_d, wk(P->kill_radius, S))) { fuze = 1; intercepted = 1; }
    if (P->precision_mode) {
        if (LE(e, e->mpos[2], 0.0)) {
            double gx = sub(e->mpos[0], P->target[0], S), gy = sub(e->mpos[1], P->target[1], S);
            double gxy[2] = {gx, gy};
            double gd = sqr(dotn(gxy, gxy, 2, S), S);
            if (LT(e, gd, wk(500.0, S))) hit = 1;
            term = 1;
        }
    } else {
        if (intercepted) term = 1;
        else if (fuze) term = 1;
        else if (LE(e, e->mpos[2], 0.0)) {
            double gx = sub(e->mpos[0], P->target[0], S), gy = sub(e->mpos[1], P->target[1], S);
            double gxy[2] = {gx, gy};
            double gd = sqr(dotn(gxy, gxy, 2, S), S);
            if (LT(e, gd, wk(500.0, S))) hit = 1;
            term = 1;
        }
    }
    }
    if (LT(e, e->ipos[2], 0.0)) term = 1;
    else if (e->fuel <= 0.0) term = 1;
    else if (e->steps > 1000) {
        if (GT(e, dist, e->last_d)) e->worsen += 1;
        else e->worsen = e->worsen - 5 > 0 ? e->worsen - 5 : 0;
        e->last_d = dist;
        if (e->worsen > 500 && GT(e, dist, wk(2500.0, S))) term = 1;
    }
    if (e->steps >= P->max_steps) trunc = 1;

    observe(o, e, step, out->obs);

    /* ---- _calculate_reward :1131-1320 ---- */
    double r;
    if (P->precision_mode) {
        if (term) {
            double md = e->min_d;
            if (e->crossed) {
                r = 3000.0;
                double cr = C->intercept_radius;
                if (md < wk(cr, S)) r = add(r, mul(dvd(sub(wk(cr, S), md, S), wk(cr, S), S), wk(1000.0, S), S), S);
                r = add(r, mul(rn(exp(dvd(-md, wk(25.0, S), S)), S), wk(500.0, S), S), S);
                r = add(r, mul(rn(exp(dvd(-md, wk(10.0, S), S)), S), wk(1000.0, S), S), S);
                r = add(r, mul(rn(exp(dvd(-md, wk(3.0, S), S)), S), wk(500.0, S), S), S);
                r = add(r, wk((P->max_steps - e->steps) * 0.3, S), S);
            } else {
                r = mul(-md, wk(0.5, S), S);
                if (r < -2000.0) r = -2000.0;
                if (hit) r = sub(r, wk(1000.0, S), S);
                else if (e->ipos[2] < 0) r = sub(r, wk(500.0, S), S);
                else if (e->fuel <= 0) r = sub(r, wk(300.0, S), S);
            }
        } else {
            double dl = sub(e->prev_d, dist, S);
            double cv = dvd(dl, wk(dt, S), S);
            r = mul(clipd(dvd(cv, wk(100.0, S), S), wk(-0.5, S), wk(2.0, S)), wk(0.5, S), S);
            if (LT(e, dist, wk(50.0, S))) {
                r = add(r, mul(dl, wk(5.0, S), S), S);
                r = add(r, mul(rn(exp(dvd(-dist, wk(10.0, S), S)), S), wk(1.0, S), S), S);
            } else if (LT(e, dist, wk(150.0, S))) r = add(r, mul(dl, wk(3.0, S), S), S);
            else if (LT(e, dist, wk(500.0, S))) r = add(r, mul(dl, wk(1.5, S), S), S);
            else r = add(r, mul(dl, wk(0.8, S), S), S);
            double sp = norm3(e->ivel, S);
            if (sp > 1.0 && dist > 10.0) {
                double al = 0.0;
                for (int i = 0; i < 3; ++i) al = add(al, mul(dvd(e->ivel[i], sp, S), dvd(sub(e->mpos[i], e->ipos[i], S), dist, S), S), S);
                r = add(r, mul(al, wk(0.3, S), S), S);
            }
            /* forward-thrust shaping on the ORIGINAL LOS-frame action, environment.py:1252-1264 */
            if (P->obs_mode == HLYNR_OBS_LOS) r = add(r, mul((double)act[0], wk(0.4, S), S), S);
            r = sub(r, wk(0.2, S), S);
            e->prev_d = dist;
        }
    } else if (intercepted) {
        r = 5000.0 + (P->max_steps - e->steps) * 0.5;
    } else if (term) {
        r = mul(-dist, wk(0.5, S), S);
        if (r < -2000.0) r = -2000.0;
        if (hit) r = sub(r, wk(1000.0, S), S);
        else if (e->ipos[2] < 0) r = sub(r, wk(500.0, S), S);
        else if (e->fuel <= 0) r = sub(r, wk(300.0, S), S);
    } else {
        double dl = sub(e->prev_d, dist, S);
        double cv = dvd(dl, wk(dt, S), S);
        r = mul(clipd(dvd(cv, wk(100.0, S), S), wk(-0.5, S), wk(2.0, S)), wk(0.3, S), S);
        if (LT(e, dist, wk(200.0, S))) r = add(r, mul(dl, wk(2.0, S), S), S);
        else if (LT(e, dist, wk(500.0, S))) r = add(r, mul(dl, wk(1.0, S), S), S);
        else r = add(r, mul(dl, wk(0.5, S), S), S);
        r = sub(r, wk(0.5, S), S);
        e->prev_d = dist;
    }
    out->reward = r;
    out->terminated = term; out->truncated = trunc;
    out->distance = dist; out->intercepted = intercepted; out->hit_target = hit; out->clamped = clamped; out->fuze = fuze;
    e->ep_return += r;
    e->ep_length += 1;
}

/* ---------------------------------------------------------------------------------------------- */
/* batch API (ctypes)                                                                              */
/* ---------------------------------------------------------------------------------------------- */
void* oracle_create(const HlynrParams* P, int64_t n, uint64_t seed, int64_t env_id_offset, int f64) {
    if (!P || P->abi_version != HLYNR_ABI_VERSION || n <= 0) return NULL;
    Oracle* o = (Oracle*)calloc(1, sizeof(Oracle));
    o->P = *P;
    o->n = n;
    o->S = f64 ? F64 : F32;
    o->C.intercept_radius = 200.0; o->C.beam_width_deg = 60.0; o->C.onboard_reliability = 1.0; o->C.ground_reliability = 1.0;
    o->env = (Env*)calloc((size_t)n, sizeof(Env));
    kf_build(P->dt);
    for (int64_t i = 0; i < n; ++i) {
        Env* e = &o->env[i];
        e->rng.seed = seed; e->rng.env_id = (uint64_t)(env_id_offset + i); e->rng.episode = 0;
        e->T0 = 288.15; e->base_cd = 0.3; e->peak = P->peak_mult; e->cd_strong = 0;
        e->has_onb = P->onboard_delay > 0;
        if (e->has_onb) db_init(&e->onb, P->onboard_delay);
        db_init(&e->gnd, P->ground_delay > 0 ? P->ground_delay : 1);
        kf_reset(&e->kf);
    }
    return o;
}
void oracle_destroy(void* h) { if (h) { Oracle* o = (Oracle*)h; free(o->env); free(o); } }
void oracle_set_curriculum(void* h, const HlynrCurriculum* c) { ((Oracle*)h)->C = *c; }
void oracle_seed(void* h, uint64_t seed) { Oracle* o = (Oracle*)h; for (int64_t i = 0; i < o->n; ++i) o->env[i].rng.seed = seed; }

void oracle_reset(void* h, const uint8_t* mask, float* obs) {
    Oracle* o = (Oracle*)h;
    for (int64_t i = 0; i < o->n; ++i)
        if (!mask || mask[i]) env_reset(o, &o->env[i], obs + 26 * i);
}

static pthread_mutex_t STATS_MU = PTHREAD_MUTEX_INITIALIZER;
static void account(HlynrStats* st, Env* e, const StepOut* s) {
    st->env_steps += 1;
    if (e->last_onb_det) st->onboard_locks += 1;
    if (!(s->terminated || s->truncated)) return;
    st->episodes += 1;
    st->successes += s->intercepted ? 1 : 0;
    st->return_sum += e->ep_return;
    st->length_sum += e->ep_length;
    st->min_distance_sum += e->min_d;
    st->final_distance_sum += s->distance;
    if (s->terminated) {
        if (s->intercepted) {}
        else if (s->hit_target) st->hit_target += 1;
        else if (e->ipos[2] < 0) st->interceptor_crash += 1;
        else if (e->fuel <= 0) st->fuel_out += 1;
        else if (e->mpos[2] <= 0) st->missile_ground += 1;
        else st->worsening += 1;
    } else st->timeouts += 1;
}

/* Steps envs [i0, i1).  All output pointers except obs/reward/terminated/truncated are optional. */
void oracle_step_range(void* h, int64_t i0, int64_t i1, const float* actions, float* obs, double* reward,
                       uint8_t* terminated, uint8_t* truncated, float* terminal_obs, const HlynrInfoSoA* info,
                       double* margin, int auto_reset) {
    Oracle* o = (Oracle*)h;
    HlynrStats local;
    memset(&local, 0, sizeof(local));
    for (int64_t i = i0; i < i1; ++i) {
        Env* e = &o->env[i];
        StepOut s;
        env_step(o, e, actions + 6 * i, &s);
        reward[i] = s.reward; terminated[i] = (uint8_t)s.terminated; truncated[i] = (uint8_t)s.truncated;
        if (margin) margin[i] = e->margin;
        if (info) {
            if (info->distance) info->distance[i] = (float)s.distance;
            if (info->min_distance) info->min_distance[i] = (float)e->min_d;
            if (info->fuel_remaining) info->fuel_remaining[i] = (float)e->fuel;
            if (info->fuel_used) info->fuel_used[i] = (float)e->fuel_used;
            if (info->steps) info->steps[i] = e->steps;
            if (info->flags)
                info->flags[i] = (uint8_t)((s.intercepted ? 1 : 0) | (s.hit_target ? 2 : 0) | (s.clamped ? 4 : 0) |
                                           (e->last_onb_det ? 8 : 0) | (e->last_gnd_det ? 16 : 0) | (e->crossed ? 32 : 0) |
                                           (s.fuze ? 64 : 0) | (e->kf.initialized ? 128 : 0));
            if (info->interceptor_pos) for (int k = 0; k < 3; ++k) info->interceptor_pos[3 * i + k] = (float)e->ipos[k];
            if (info->missile_pos) for (int k = 0; k < 3; ++k) info->missile_pos[3 * i + k] = (float)e->mpos[k];
            if (info->episode_return) info->episode_return[i] = (float)e->ep_return;
            if (info->episode_length) info->episode_length[i] = e->ep_length;
            const int vk = o->P.volley_size;
            if (info->missiles_intercepted) info->missiles_intercepted[i] = vk > 0 ? e->vcount : (s.intercepted ? 1 : 0);
            if (info->missiles_remaining) {
                int rem = 0;
                for (int m = 0; m < vk; ++m) rem += e->vact[m];
                info->missiles_remaining[i] = vk > 0 ? rem : (s.intercepted ? 0 : 1);
            }
            if (info->radar_quality) info->radar_quality[i] = e->last_onb_fill ? 0.f : (float)o->P.radar_quality; /* environment.py:840 */
            if (info->missile_min_distances)
                for (int m = 0; m < HLYNR_MAX_VOLLEY; ++m)
                    info->missile_min_distances[HLYNR_MAX_VOLLEY * i + m] =
                        vk > 0 ? (m < vk ? (float)e->vmin[m] : 0.f) : (m == 0 ? (float)s.distance : 0.f);
        }
        account(&local, e, &s);
        if ((s.terminated || s.truncated) && auto_reset) {
            if (terminal_obs) memcpy(terminal_obs + 26 * i, s.obs, sizeof(float) * 26);
            env_reset(o, e, obs + 26 * i);
        } else memcpy(obs + 26 * i, s.obs, sizeof(float) * 26);
    }
    /* callers may step disjoint ranges from several threads (ctypes releases the GIL) */
    pthread_mutex_lock(&STATS_MU);
    {
        double* dst = (double*)&o->stats;
        const double* src = (const double*)&local;
        for (int k = 0; k < HLYNR_STATS_WORDS; ++k) dst[k] += src[k];
    }
    pthread_mutex_unlock(&STATS_MU);
}
void oracle_step(void* h, const float* actions, float* obs, double* reward, uint8_t* terminated, uint8_t* truncated,
                 float* terminal_obs, const HlynrInfoSoA* info, double* margin, int auto_reset) {
    oracle_step_range(h, 0, ((Oracle*)h)->n, actions, obs, reward, terminated, truncated, terminal_obs, info, margin,
                      auto_reset);
}
void oracle_get_stats(void* h, HlynrStats* out, int zero_after) {
    Oracle* o = (Oracle*)h;
    *out = o->stats;
    if (zero_after) memset(&o->stats, 0, sizeof(o->stats));
}
void oracle_export_state(void* h, int64_t first, int64_t count, HlynrEnvState* out) {
    Oracle* o = (Oracle*)h;
    for (int64_t j = 0; j < count; ++j) {
        Env* e = &o->env[first + j];
        HlynrEnvState* s = &out[j];
        memset(s, 0, sizeof(*s));
        for (int k = 0; k < 3; ++k) { s->ipos[k] = e->ipos[k]; s->ivel[k] = e->ivel[k]; s->mpos[k] = e->mpos[k]; s->mvel[k] = e->mvel[k]; s->wind[k] = e->wind[k]; s->thrust[k] = e->thrust[k]; }
        for (int k = 0; k < 4; ++k) s->quat[k] = e->quat[k];
        s->fuel = e->fuel; s->fuel_used = e->fuel_used;
        s->prev_d = e->prev_d; s->last_d = e->last_d; s->min_d = e->min_d; s->episode_return = e->ep_return;
        for (int k = 0; k < 6; ++k) s->kf_x[k] = e->kf.x[k];
        s->kf_P[0] = e->kf.P[0][0]; s->kf_P[1] = e->kf.P[0][3]; s->kf_P[2] = e->kf.P[3][0]; s->kf_P[3] = e->kf.P[3][3];
        s->T0 = e->T0; s->base_cd = e->base_cd; s->peak = e->peak;
        s->steps = e->steps; s->worsen_count = e->worsen; s->crossed = e->crossed; s->kf_init = e->kf.initialized;
        s->onboard_delay = e->has_onb ? e->onb.delay : 0; s->episode = (int32_t)e->rng.episode;
        for (int m = 0; m < o->P.volley_size; ++m) {
            for (int k = 0; k < 3; ++k) { s->vpos[3 * m + k] = e->vpos[m][k]; s->vvel[3 * m + k] = e->vvel[m][k]; }
            s->vmin[m] = e->vmin[m]; s->vactive[m] = e->vact[m];
        }
        s->vcur = o->P.volley_size > 0 ? e->vcur : 0; s->vcount = o->P.volley_size > 0 ? e->vcount : 0;
        s->kf_f64 = e->kf.x_f64;
    }
}
/* max |P - blockdiag(2x2)| over all envs: checks the decoupling claim the CUDA Kalman relies on */
double oracle_kalman_decoupling_error(void* h) {
    Oracle* o = (Oracle*)h;
    double worst = 0.0;
    for (int64_t i = 0; i < o->n; ++i) {
        float(*P)[6] = o->env[i].kf.P;
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) {
                double want = 0.0;
                if (r % 3 == c % 3) want = P[r / 3 * 3][c / 3 * 3];
                double d = fabs((double)P[r][c] - want);
                if (d > worst) worst = d;
            }
    }
    return worst;
}
size_t oracle_params_size(void) { return sizeof(HlynrParams); }
size_t oracle_env_state_size(void) { return sizeof(HlynrEnvState); }
