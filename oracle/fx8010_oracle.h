/*
 * oracle/fx8010_oracle.h — TEST INFRASTRUCTURE ONLY (see fx8010_oracle.c).
 *
 * CPU restatement of the reference's per-sample interpreter, used as the checker for the CUDA
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this; the product (fx8010-emulator-core_b200/) never does.
 */
#ifndef FX8010_ORACLE_H
#define FX8010_ORACLE_H

#include "fx8010_gpu.h" /* shares the decoded-image structs and enums with the C ABI */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fx_oracle fx_oracle;

/* LOG / EXP lookup tables, [32][64] doubles each (reference source/FX8010.cpp:63-105,129-199). */
void fx_oracle_build_tables(double* log_tables, double* exp_tables);

/* Copies the image; state = freshly constructed + loaded reference objects.  NULL on a bad image. */
fx_oracle* fx_oracle_create(const fx8010_program_image* image, int n_instances, int n_channels);
void fx_oracle_destroy(fx_oracle* o);

/* in/out: [n_channels][n_samples][n_instances] float32 (same layout as the C ABI).
 * n_threads > 1 splits the instance range over pthreads. */
int fx_oracle_process(fx_oracle* o, const float* in, float* out, int n_samples, int n_threads);

/* state views (owned by the oracle, valid until destroy) */
float* fx_oracle_registers(fx_oracle* o);              /* [n_regs][N]                           */
double* fx_oracle_acc(fx_oracle* o);                   /* [N]                                   */
uint32_t* fx_oracle_lfsr(fx_oracle* o);                /* [2][N]                                */
float* fx_oracle_out_latch(fx_oracle* o);              /* [C][N]                                */
int32_t* fx_oracle_tram_ptrs(fx_oracle* o);            /* [4][N]: iw, ir, xw, xr                */
unsigned long long* fx_oracle_counts(fx_oracle* o);    /* [N]                                   */
/* copies ring positions 0..size-1 of one instance; which: 0 iTRAM, 1 xTRAM */
int fx_oracle_get_tram(fx_oracle* o, int which, int instance, float* out);
int fx_oracle_set_tram(fx_oracle* o, int which, int instance, const float* in);
unsigned int fx_oracle_runtime_flags(fx_oracle* o);

/* single LOG/EXP evaluation exactly as the interpreter does it (for exhaustive sweeps):
 * which 0 = LOG, 1 = EXP; returns the float result, *flag gets FX8010_RT_TABLE_RANGE or 0. */
float fx_oracle_table_eval(const double* tables, int selector, float a, unsigned int* flag);

#ifdef __cplusplus
}
#endif
#endif
