/*
 * oracle/fx8010_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's per-sample interpreter (easypx/FX8010-Emulator-Core,
 * FX8010::process and its helpers).  It is the CHECKER for the CUDA path: tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it; nothing
 * under fx8010-emulator-core_b200/ may.  It is written for obviousness, not speed: one
 * instance at a time, one sample at a time, literal operation order.
 *
 * PARITY PIN: this restatement is pinned against the unmodified reference compiled from
 * /root/reference (oracle/_ref/libfx8010_ref.so, see oracle/Makefile) by
 * tests/test_oracle_vs_reference.py and by the golden vectors under tests/golden/ that
 * tests/golden/make_golden.py generated from that binary (the reference ships no test vectors
 * of its own — SURVEY.md §4).
 *
 * Must be compiled with -ffp-contract=off: every float/double operation below
 * is meant to round individually, as the reference's x86-64 SSE2 build does.
 *
 * Where the reference's behaviour is undefined this file implements the rule DESIGN.md states
 * ("UB ledger"); those spots are marked [U*].
 */
#include "fx8010_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

struct fx_oracle {
    int n, c, n_regs, n_instrs;
    int itram, xtram;          /* ring sizes (0 = not allocated)                              */
    fx8010_instr* instrs;
    fx8010_reg* regs;
    double log_t[FX8010_TABLE_COUNT][FX8010_TABLE_ENTRIES + 1]; /* [64] = 0.0 pad, rule [U5] */
    double exp_t[FX8010_TABLE_COUNT][FX8010_TABLE_ENTRIES + 1];
    float* gpr;                /* [n_regs][n]                                                 */
    double* acc;               /* [n]                                                         */
    uint32_t* lfsr;            /* [2][n]                                                      */
    float* latch;              /* [c][n]                                                      */
    int32_t* ptrs;             /* [4][n]                                                      */
    float* itram_buf;          /* [n][itram]  (instance-major here; accessors hide the layout) */
    float* xtram_buf;          /* [n][xtram]                                                  */
    unsigned long long* counts;/* [n]                                                         */
    unsigned int flags;
};

/* ---- scalar helpers ------------------------------------------------------------------------ */

/* static_cast<int32_t>(float) as the reference's x86-64 build executes it (cvttss2si):
 * out-of-range and NaN give 0x80000000.  [U7]  (source/FX8010.cpp:343-345,1019,1115,1177,…) */
static int32_t cvt_f32_i32(float f) {
    if (!(f < 2147483648.0f) || f < -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}
/* static_cast<int>(double), cvttsd2si (source/FX8010.cpp:286) */
static int32_t cvt_f64_i32(double d) {
    if (!(d < 2147483648.0) || d <= -2147483649.0) return INT32_MIN;
    return (int32_t)d;
}

/* saturate(input, 1.0f): source/FX8010.cpp:275-279.  NaN passes through. */
static float saturate1(float v) { return (v >= 1.0f) ? 1.0f : ((v <= -1.0f) ? -1.0f : v); }

/* setCCR: source/FX8010.cpp:211-232 */
static float ccr_of(float r) {
    if (r == 0) return 8.0f;                  /* 0b01000 zero                  */
    else if (r < 0 && r > -1.0) return 6.0f;  /* 0b00110 normalized negative   */
    else if (r > 0 && r < 1.0) return 2.0f;   /* 0b00010 normalized positive   */
    else if (r == 1.0) return 16.0f;          /* 0b10000 positive saturation   */
    else if (r == -1.0) return 20.0f;         /* 0b10100 negative saturation   */
    return 0.0f;
}

/* floatToInt / intToFloat: source/FX8010.cpp:1009-1020; (float)INT32_MAX == 2^31 */
static int32_t float_to_q31(float f) { return cvt_f32_i32(f * 2147483648.0f); }
static float q31_to_float(int32_t i) { return (float)i / 2147483648.0f; }

/* wrapAround value path: source/FX8010.cpp:299-328.  Its CCR side effect is restated
 * dead: every caller runs setCCR right after (see the MACW case below). */
static float wrap_value(float a) {
    if (a >= 1.0f) return a - 2.0f;
    else if (a < -1.0f) return a + 2.0f;
    return a;
}
/* logicOps: source/FX8010.cpp:330-360 */
static int32_t logic_ops(float fa, float fx, float fy) {
    const int32_t A = cvt_f32_i32(fa), X = cvt_f32_i32(fx), Y = cvt_f32_i32(fy);
    if (Y == 0) return A & X;
    else if (X == 0xFFFFFF) return A ^ Y;
    else if (X == 0xFFFFFFF && Y == 0xFFFFFF) return ~A;
    else if (Y == ~X) return A | Y;
    else if (Y == 0xFFFFFF) return ~A & X;
    return (A & X) ^ Y;
}

/* linearInterpolate(x, table, -1.0, 1.0): source/FX8010.cpp:283-296, called from :1115/:1121.
 * Rules for the reference's undefined corners: selector outside 0..31 and index outside 0..63
 * are clamped and flagged [U6]; table[64] reads as 0.0 [U5]. */
static float table_eval(const double (*tables)[FX8010_TABLE_ENTRIES + 1], float sel_f, float a, unsigned int* flags) {
    int32_t sel = cvt_f32_i32(sel_f);
    if (sel < 0 || sel > FX8010_TABLE_COUNT - 1) { *flags |= FX8010_RT_TABLE_RANGE; sel = sel < 0 ? 0 : FX8010_TABLE_COUNT - 1; }
    if (!(a >= -1.0f && a <= 1.0f)) *flags |= FX8010_RT_TABLE_RANGE;
    const double* t = tables[sel];
    const double x = (double)a, x_min = -1.0, x_max = 1.0;
    const double step = (x_max - x_min) / (double)(FX8010_TABLE_ENTRIES - 1);
    int32_t index = cvt_f64_i32((x - x_min) / step);
    if (index < 0) index = 0;
    if (index > FX8010_TABLE_ENTRIES - 1) index = FX8010_TABLE_ENTRIES - 1;
    const double x1 = x_min + index * step;
    const double x2 = x_min + (index + 1) * step;
    const double y1 = t[index];
    const double y2 = t[index + 1];
    const double y = (y2 - y1) / (x2 - x1) * (x - x1) + y1;
    return (float)y;
}

float fx_oracle_table_eval(const double* tables, int selector, float a, unsigned int* flag) {
    double padded[1][FX8010_TABLE_ENTRIES + 1];
    memcpy(padded[0], tables + (size_t)selector * FX8010_TABLE_ENTRIES, sizeof(double) * FX8010_TABLE_ENTRIES);
    padded[0][FX8010_TABLE_ENTRIES] = 0.0;
    unsigned int f = 0;
    float r = table_eval((const double (*)[FX8010_TABLE_ENTRIES + 1])padded, 0.0f, a, &f);
    if (flag) *flag = f;
    return r;
}

/* Table construction: source/FX8010.cpp:63-105 with createLog/ExpLookupTable :129-163,
 * mirrorYVector :167-175, negateVector :190-199 (the negation runs through a `float` loop
 * variable, so the mirrored half is float-rounded), concatenateVectors :179-186. */
void fx_oracle_build_tables(double* log_tables, double* exp_tables) {
    const int entries = 32;
    for (int e = 0; e < FX8010_TABLE_COUNT; ++e) {
        double lg[32], ex[32];
        const double step = (1.0 - 0.0) / (entries - 1);
        for (int i = 0; i < entries; ++i) {
            const double x = 0.0 + (i * step);
            lg[i] = pow(x, 1.0 / (float)e);   /* e == 0: 1.0/0.0f = +inf */
            ex[i] = pow(x, (float)e);
        }
        double* L = log_tables + (size_t)e * FX8010_TABLE_ENTRIES;
        double* X = exp_tables + (size_t)e * FX8010_TABLE_ENTRIES;
        for (int i = 0; i < entries; ++i) {
            const float ml = (float)lg[entries - 1 - i];
            const float mx = (float)ex[entries - 1 - i];
            L[i] = -ml;            /* float negation promoted to double */
            X[i] = -mx;
            L[entries + i] = lg[i];
            X[entries + i] = ex[i];
        }
    }
}

/* ---- lifecycle --------------------------------------------------------------------------------- */

static int uses_tram(const fx8010_program_image* im, int opcode) {
    for (int i = 0; i < im->n_instrs; ++i) {
        const fx8010_instr* in = &im->instrs[i];
        if (in->opcode != opcode) continue;
        const int t = im->regs[in->r].type;
        if (t == FX_REG_READ || t == FX_REG_WRITE) return 1;
    }
    return 0;
}

fx_oracle* fx_oracle_create(const fx8010_program_image* im, int n, int c) {
    if (!im || n <= 0 || c <= 0 || im->n_instrs <= 0 || im->n_regs <= 0) return NULL;
    for (int i = 0; i < im->n_instrs; ++i) {
        const fx8010_instr* in = &im->instrs[i];
        if (in->opcode < 0 || in->opcode >= FX_NUM_OPCODES) return NULL;
        if (in->r < 0 || in->r >= im->n_regs || in->a < 0 || in->a >= im->n_regs ||
            in->x < 0 || in->x >= im->n_regs || in->y < 0 || in->y >= im->n_regs) return NULL;
    }
    for (int i = 0; i < im->n_regs; ++i)
        if (im->regs[i].io_index < 0 || im->regs[i].io_index >= c) return NULL;
    const int use_i = uses_tram(im, FX_IDELAY), use_x = uses_tram(im, FX_XDELAY);
    if ((use_i && im->itram_size <= 0) || (use_x && im->xtram_size <= 0)) return NULL; /* [U4] */

    fx_oracle* o = (fx_oracle*)calloc(1, sizeof(fx_oracle));
    o->n = n; o->c = c; o->n_regs = im->n_regs; o->n_instrs = im->n_instrs;
    o->itram = use_i ? im->itram_size : 0;
    o->xtram = use_x ? im->xtram_size : 0;
    o->instrs = (fx8010_instr*)malloc(sizeof(fx8010_instr) * im->n_instrs);
    memcpy(o->instrs, im->instrs, sizeof(fx8010_instr) * im->n_instrs);
    o->regs = (fx8010_reg*)malloc(sizeof(fx8010_reg) * im->n_regs);
    memcpy(o->regs, im->regs, sizeof(fx8010_reg) * im->n_regs);
    for (int t = 0; t < FX8010_TABLE_COUNT; ++t) {
        memcpy(o->log_t[t], im->log_tables + (size_t)t * FX8010_TABLE_ENTRIES, sizeof(double) * FX8010_TABLE_ENTRIES);
        memcpy(o->exp_t[t], im->exp_tables + (size_t)t * FX8010_TABLE_ENTRIES, sizeof(double) * FX8010_TABLE_ENTRIES);
        o->log_t[t][FX8010_TABLE_ENTRIES] = 0.0;
        o->exp_t[t][FX8010_TABLE_ENTRIES] = 0.0;
    }
    o->gpr = (float*)malloc(sizeof(float) * (size_t)im->n_regs * n);
    for (int r = 0; r < im->n_regs; ++r)
        for (int i = 0; i < n; ++i) o->gpr[(size_t)r * n + i] = im->regs[r].init_value;
    o->acc = (double*)calloc(n, sizeof(double));
    o->lfsr = (uint32_t*)malloc(sizeof(uint32_t) * 2 * n);
    for (int i = 0; i < n; ++i) { o->lfsr[i] = FX8010_LFSR_SEED1; o->lfsr[n + i] = FX8010_LFSR_SEED2; }
    o->latch = (float*)calloc((size_t)c * n, sizeof(float));
    o->ptrs = (int32_t*)calloc((size_t)4 * n, sizeof(int32_t));
    o->itram_buf = o->itram ? (float*)calloc((size_t)o->itram * n, sizeof(float)) : NULL; /* zeros: [U2] */
    o->xtram_buf = o->xtram ? (float*)calloc((size_t)o->xtram * n, sizeof(float)) : NULL;
    o->counts = (unsigned long long*)calloc(n, sizeof(unsigned long long));
    return o;
}

void fx_oracle_destroy(fx_oracle* o) {
    if (!o) return;
    free(o->instrs); free(o->regs); free(o->gpr); free(o->acc); free(o->lfsr); free(o->latch);
    free(o->ptrs); free(o->itram_buf); free(o->xtram_buf); free(o->counts); free(o);
}

/* ---- TRAM: source/FX8010.cpp:909-967 ----------------------------------------------------------- */

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* writeSmallDelay / writeLargeDelay.  The reference stores at wp+p WITHOUT wrapping; slots at
 * or beyond `size` are never read back through the ring, so they are dropped here. */
static void tram_write(float* ring, int size, int32_t* wp, float sample, int32_t pos) {
    pos = clampi(pos, 0, size - 1);
    const int idx = *wp + pos;
    if (idx < size) ring[idx] = sample;
    *wp = (*wp + 1) % size;
}
/* readSmallDelay / readLargeDelay.  (rp - p) % size is negative in the reference when rp < p
 * (reads before the array — undefined); rule [U1]: mathematical modulo. */
static float tram_read(const float* ring, int size, int32_t* rp, int32_t pos) {
    pos = clampi(pos, 0, size - 1);
    int idx = (*rp - pos) % size;
    if (idx < 0) idx += size;
    const float out = ring[idx];
    *rp = (*rp + 1) % size;
    return out;
}

/* ---- the interpreter: FX8010::process, source/FX8010.cpp:1023-1249 ----------------------------- */

static void run_instance(fx_oracle* o, int inst, const float* in, float* out, int n_samples, unsigned int* flags_out) {
    const int n = o->n, C = o->c;
    float* g = o->gpr + inst;                    /* register r lives at g[r * n]   */
#define REG(r) g[(size_t)(r) * n]
    double acc = o->acc[inst];
    uint32_t x1 = o->lfsr[inst], x2 = o->lfsr[n + inst];
    int32_t* iw = &o->ptrs[0 * n + inst]; int32_t* ir = &o->ptrs[1 * n + inst];
    int32_t* xw = &o->ptrs[2 * n + inst]; int32_t* xr = &o->ptrs[3 * n + inst];
    float* iring = o->itram_buf ? o->itram_buf + (size_t)inst * o->itram : NULL;
    float* xring = o->xtram_buf ? o->xtram_buf + (size_t)inst * o->xtram : NULL;
    unsigned long long count = o->counts[inst];
    unsigned int flags = 0;

    for (int s = 0; s < n_samples; ++s) {
        int is_end = 0;
        int num_skip = 0;                                             /* :1030 */
        int passes = 0;
        do {
            for (int pc = 0; pc < o->n_instrs; ++pc) {
                const fx8010_instr* ins = &o->instrs[pc];
                if (num_skip == 0) {                                  /* :1037 */
                    const int R = ins->r, A = ins->a, X = ins->x, Y = ins->y;
                    const fx8010_reg* rR = &o->regs[R];
                    const fx8010_reg* rA = &o->regs[A];
                    const fx8010_reg* rX = &o->regs[X];
                    const fx8010_reg* rY = &o->regs[Y];
                    if (ins->has_input) {                             /* :1053-1061, X and Y use A's IOIndex */
                        const float v = in ? in[((size_t)rA->io_index * n_samples + s) * n + inst] : 0.0f;
                        if (rA->type == FX_REG_INPUT) REG(A) = v;
                        if (rX->type == FX_REG_INPUT) REG(X) = v;
                        if (rY->type == FX_REG_INPUT) REG(Y) = v;
                    }
                    if (ins->has_noise) {                             /* :1063-1071, whitenoise :993-1000 */
                        const int tgt = rA->is_noise ? A : (rX->is_noise ? X : (rY->is_noise ? Y : -1));
                        if (tgt >= 0) {
                            x1 ^= x2;
                            const float nz = (float)(int32_t)x2 * 4.656612873077392578125e-10f; /* 2^-31 */
                            x2 += x1;                                 /* wrapping, [U8] */
                            REG(tgt) = nz;
                        }
                    }
                    const float a = REG(A), x = REG(X), y = REG(Y);
                    float r;
                    switch (ins->opcode) {
                    case FX_MACS: case FX_MACINTS:                    /* :1077-1085, :1095-1103 */
                        r = a + x * y; acc = r; r = saturate1(r); REG(R) = r; REG(0) = ccr_of(r); break;
                    case FX_MACSN:                                    /* :1086-1094 */
                        r = a - x * y; acc = r; r = saturate1(r); REG(R) = r; REG(0) = ccr_of(r); break;
                    case FX_ACC3:                                     /* :1104-1112 */
                        r = a + x + y; acc = r; r = saturate1(r); REG(R) = r; REG(0) = ccr_of(r); break;
                    case FX_LOG:                                      /* :1113-1119 */
                        r = table_eval(o->log_t, x, a, &flags); acc = r; REG(R) = r; REG(0) = ccr_of(r); break;
                    case FX_EXP:                                      /* :1120-1125 */
                        r = table_eval(o->exp_t, x, a, &flags); acc = r; REG(R) = r; REG(0) = ccr_of(r); break;
                    case FX_MACW: case FX_MACWN: {                    /* :1126-1137 */
                        /* A.registerValue +/- wrapAround(X*Y).  wrapAround also rewrites CCR (:302-320),
                         * but setCCR overwrites that right after; it could only be seen if A is the ccr
                         * register AND the compiler called wrapAround before reading A (unspecified in
                         * C++).  The g++ 13.3 -O2 reference build reads A first (checked by
                         * tests/test_oracle_vs_reference.py::test_macw_with_ccr_operand), so does this. */
                        const float w = wrap_value(x * y);
                        r = (ins->opcode == FX_MACW) ? a + w : a - w;
                        REG(R) = r; acc = r; REG(0) = ccr_of(r); break;
                    }
                    case FX_MACINTW:                                  /* :1138-1143 */
                        r = wrap_value(a + x * y); REG(R) = r; acc = r; REG(0) = ccr_of(r); break;
                    case FX_MACMV:                                    /* :1144-1149 */
                        acc = acc + (double)(x * y); r = a; REG(R) = r; REG(0) = ccr_of(r); break;
                    case FX_ANDXOR:                                   /* :1150-1154 */
                        r = (float)logic_ops(a, x, y); REG(R) = r; REG(0) = ccr_of(r); break;
                    case FX_TSTNEG:                                   /* :1155-1162 */
                        r = (a >= y) ? x : q31_to_float(~float_to_q31(x));
                        REG(R) = r; acc = r; REG(0) = ccr_of(r); break;
                    case FX_LIMIT:                                    /* :1163-1168 */
                        r = (a >= y) ? x : y; REG(R) = r; acc = r; REG(0) = ccr_of(r); break;
                    case FX_LIMITN:                                   /* :1169-1174 */
                        r = (a < y) ? x : y; REG(R) = r; acc = r; REG(0) = ccr_of(r); break;
                    case FX_SKIP:                                     /* :1175-1179 */
                        if ((float)cvt_f32_i32(x) == REG(0)) num_skip = cvt_f32_i32(y);
                        break;
                    case FX_INTERP: {                                 /* :1180-1187 */
                        const double d = (1.0 - (double)x) * (double)a + (double)(x * y);
                        r = (float)d; acc = r; r = saturate1(r); REG(R) = r; REG(0) = ccr_of(r); break;
                    }
                    case FX_IDELAY:                                   /* :1188-1199 */
                        if (rR->type == FX_REG_READ) REG(A) = tram_read(iring, o->itram, ir, cvt_f32_i32(y));
                        else if (rR->type == FX_REG_WRITE) tram_write(iring, o->itram, iw, a, cvt_f32_i32(y));
                        break;
                    case FX_XDELAY:                                   /* :1200-1211 */
                        if (rR->type == FX_REG_READ) REG(A) = tram_read(xring, o->xtram, xr, cvt_f32_i32(y));
                        else if (rR->type == FX_REG_WRITE) tram_write(xring, o->xtram, xw, a, cvt_f32_i32(y));
                        break;
                    case FX_END:                                      /* :1212-1215 */
                        is_end = 1; break;
                    default: break;
                    }
                    count++;                                          /* :1222 */
                    if (rR->type == FX_REG_OUTPUT)                    /* :1229-1233 */
                        o->latch[(size_t)rR->io_index * n + inst] = REG(R);
                } else {
                    num_skip = (num_skip > 0) ? num_skip - 1 : 0;     /* :1238; a negative count skips one */
                }
            }
            ++passes;
            if (!is_end && passes >= FX8010_MAX_PASSES) { flags |= FX8010_RT_END_SKIPPED_CAP; break; } /* [U9] */
        } while (!is_end);                                            /* :1243 */
        for (int ch = 0; ch < C; ++ch)                                /* :1248 */
            out[((size_t)ch * n_samples + s) * n + inst] = o->latch[(size_t)ch * n + inst];
    }
#undef REG
    o->acc[inst] = acc;
    o->lfsr[inst] = x1; o->lfsr[n + inst] = x2;
    o->counts[inst] = count;
    *flags_out |= flags;
}

struct worker { fx_oracle* o; int i0, i1; const float* in; float* out; int n_samples; unsigned int flags; };

static void* worker_main(void* p) {
    struct worker* w = (struct worker*)p;
    for (int i = w->i0; i < w->i1; ++i) run_instance(w->o, i, w->in, w->out, w->n_samples, &w->flags);
    return NULL;
}

int fx_oracle_process(fx_oracle* o, const float* in, float* out, int n_samples, int n_threads) {
    if (!o || !out || n_samples < 0) return FX8010_ERR_ARG;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > o->n) n_threads = o->n;
    if (n_threads > 256) n_threads = 256;
    struct worker w[256];
    pthread_t th[256];
    for (int t = 0; t < n_threads; ++t) {
        w[t].o = o; w[t].in = in; w[t].out = out; w[t].n_samples = n_samples; w[t].flags = 0;
        w[t].i0 = (int)((long long)o->n * t / n_threads);
        w[t].i1 = (int)((long long)o->n * (t + 1) / n_threads);
    }
    if (n_threads == 1) worker_main(&w[0]);
    else {
        for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, worker_main, &w[t]);
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    }
    for (int t = 0; t < n_threads; ++t) o->flags |= w[t].flags;
    return FX8010_OK;
}

/* ---- state views ------------------------------------------------------------------------------- */
float* fx_oracle_registers(fx_oracle* o) { return o->gpr; }
double* fx_oracle_acc(fx_oracle* o) { return o->acc; }
uint32_t* fx_oracle_lfsr(fx_oracle* o) { return o->lfsr; }
float* fx_oracle_out_latch(fx_oracle* o) { return o->latch; }
int32_t* fx_oracle_tram_ptrs(fx_oracle* o) { return o->ptrs; }
unsigned long long* fx_oracle_counts(fx_oracle* o) { return o->counts; }
unsigned int fx_oracle_runtime_flags(fx_oracle* o) { return o->flags; }

int fx_oracle_get_tram(fx_oracle* o, int which, int instance, float* out) {
    const int size = which == 0 ? o->itram : o->xtram;
    const float* buf = which == 0 ? o->itram_buf : o->xtram_buf;
    if (!buf || instance < 0 || instance >= o->n) return FX8010_ERR_ARG;
    memcpy(out, buf + (size_t)instance * size, sizeof(float) * size);
    return FX8010_OK;
}
int fx_oracle_set_tram(fx_oracle* o, int which, int instance, const float* in) {
    const int size = which == 0 ? o->itram : o->xtram;
    float* buf = which == 0 ? o->itram_buf : o->xtram_buf;
    if (!buf || instance < 0 || instance >= o->n) return FX8010_ERR_ARG;
    memcpy(buf + (size_t)instance * size, in, sizeof(float) * size);
    return FX8010_OK;
}
