// fx8010_stateless.cuh — the INSTRUCTION-MAJOR kernel (sm_100a): stateless programs, self recurrences, delay lines.
//
// The interpreter runs a batch of M sample periods instruction by instruction: each DSP instruction is fetched from
// __constant__ memory and decoded once, then executed for the M samples of the batch (and K adjacent instances each),
// which amortises the interpretive overhead M-fold.  That is loop distribution of the sample loop over the program, and
// it keeps the reference's results (source/FX8010.cpp:1023-1249 runs sample by sample) exactly when no value flows from
// a LATER instruction of one sample period to an EARLIER instruction of the next.  fx8010_gpu.cu::analyse decides that
// at load time and sorts every operand read into
//   * same period  — the producer comes earlier in program order: registers the program writes and reads back get one
//                    shared-memory row PER SAMPLE of the batch; registers it only reads keep one row; registers it only
//                    writes (outputs, an unobserved CCR) are stored on the batch-final sample only; INPUT registers are
//                    not copied at all: the cp.async input stage rows ARE their rows (every read of an INPUT register
//                    is preceded by its preload in the same instruction, :1053-1061);
//   * self-carried — the operand is the instruction's OWN result of the previous sample period (`interp out, out, c,
//                    in`, `macs a, a, x, y`; the register has no other writer): forwarded in a hardware register from
//                    sample to sample, so a recurrence's dependency chain holds arithmetic only; row M - 1 carries the
//                    value between batches and calls;
//   * TRAM         — (TRAM = true) one READ and one WRITE per ring, offsets from registers the program never writes:
//                    both pointers advance by one per sample period (:909-967), so the slot a READ fetches was written a
//                    constant number of periods earlier.  The READs of batch b + 1 are prefetched with cp.async while
//                    batch b is computed (their stage rows stand in for the target register) as long as that distance
//                    exceeds two batches — checked per thread at kernel start, voted per warp; a warp with a shorter
//                    delay runs the same code one sample at a time with synchronous reads, which is the sequential order.
//                    When the ring is the program's only state across sample periods, the samples of a batch are
//                    independent: P threads then share one instance column and take M / P samples of every batch each
//                    (P times the warps for the same instances), meeting at one __syncthreads per batch.
// Programs with anything else carried across instructions, SKIP, noise or MACMV take the sample-major kernels.
//
// Stateless programs (nothing carried, no TRAM) are additionally cut along time into segments across blockIdx.y; only
// the thread that owns the call's last sample writes state back, after re-running that one sample in a cold FINAL copy
// of the code that stores everything the write-back needs.  Recurrences and delay lines run one segment with batches of
// up to 32 samples (the batch is also how far the input stage runs ahead of the arithmetic).
//
// Arithmetic is the same code as the generic kernel (fx8010_kernel.cuh): bit-exact with the reference.
#pragma once

#include <type_traits>

#include "fx8010_kernel.cuh"

namespace fxk {

// operand word: where a register lives in the thread's column for sample m of the batch
constexpr uint32_t SL_OFF_MASK = 0xfffffu;   // [0:20)  byte offset
constexpr int SL_STRIDE_SHIFT = 20;          // [20:31) bytes between consecutive samples of the batch, in 16-byte units
constexpr uint32_t SL_BUF = 1u << 31;        // input-stage row: add the offset of the current stage buffer
__host__ __device__ inline uint32_t sl_stride(uint32_t w) { return ((w >> SL_STRIDE_SHIFT) & 0x7ffu) << 4; }
constexpr uint32_t F_ST_LAST = 1u << 17;     // R is never read by the program: store it on the batch-final sample only
constexpr int SL_MAX_M = 64;             // a serial (recurrent) launch has few warps: its batch is also how far the input stage runs ahead of the arithmetic
constexpr uint32_t F_FUSE = 1u << 21;        // the NEXT instruction is a MACS/MACSN whose only varying operand is this instruction's result, and nobody
                                             // else reads that result: both run in one sample loop, the value is forwarded in a hardware register and
                                             // never touches shared memory (wB.w: bit 0 = forwarded into the product (X or Y) rather than the addend A,
                                             // bit 1 = MACSN).  The cold FINAL / CCR-live copies ignore the flag and run the two instructions one by one.
constexpr int MAX_FUSED_BLOCKS = 32;         // sample blocks one launch of a time-split program can cover
constexpr int SL_CARRY_SHIFT = 18;           // w0 bits 18..20: operand A / X / Y is this instruction's OWN result of the previous sample
                                             // (a self recurrence): row (m - 1) mod M on the first sample of a batch, then forwarded
                                             // in a hardware register — the recurrence never waits for shared memory

// (the encoded instruction format is described at sl_exec below; load / write-back lists still use the packed operand word above)

// A CUtensorMap (128 bytes, 64-byte aligned), opaque here: the host encodes it (fx8010_gpu.cu::make_input_map), the kernel only
// passes its address to cp.async.bulk.tensor.
struct alignas(64) SLTensorMap { unsigned char bytes[128]; };

struct alignas(64) SLParams {
    SLTensorMap in_map;         // use_tma: the launch's input block as a 2-D tensor [rows][N], box = [M rows][B * K instances]
    int use_tma;                // the input stage is filled by ONE bulk tensor copy per batch, channel and thread block (TMA, completion on an
                                // mbarrier) instead of one cp.async per thread and sample row
    int tma_rows_per_channel;   // rows between two channels of the input block (in_cstride / N)
    float* gpr;                 // [n_regs][N]
    double* acc;                // [N]
    float* latch;               // [C][N]
    unsigned long long* counts; // [N]
    unsigned int* rt_flags;
    const TableEntry* tabs;
    const uint2* load_list;     // (operand word, register index): read-only rows fetched at start
    const uint2* wb_list;       // (operand word, register index): rows written back by the owner of the last sample
    // I/O: a launch covers n_blk consecutive sample blocks of n_samples each (more than one only for time-split
    // programs: block, time segment and instance group are then all independent work items; blockIdx.y = blk * n_seg + seg)
    const float* blk_in[MAX_FUSED_BLOCKS];    // element (c, s, i) of block b at blk_in[b][c * in_cstride + s * N + i]; all NULL or none
    float* blk_out[MAX_FUSED_BLOCKS];
    int n_blk;
    size_t in_cstride, out_cstride;
    int n_samples, seg_len, n_seg;
    int N, C, n_instrs, n_exec, prog_off, n_load, n_wb;
    int M;                      // samples per batch (power of two <= SL_MAX_M)
    uint32_t stage0;            // byte offset of the input stage rows: [C][2 buffers][M]
    int n_smem_tabs;
    int smem_tab_id[MAX_SMEM_TABLES];
    int acc_writer;             // some instruction sets the accumulator (else it keeps its value)
    int ccr_live;               // somebody reads `ccr`: every setCCR is materialised per sample (else only the call's last one)
    // TRAM (programs with IDELAY/XDELAY): pointers, rings, and the READ streams prefetched into stage rows
    int32_t* ptrs;              // [4][N]  iw, ir, xw, xr
    float* itram; float* xtram; // [size][N]
    int itram_size, xtram_size;
    int has_tram, n_tr;
    int P;                      // threads per instance column: each takes M / P samples of every batch (1 unless the program's only state is TRAM)
    int tr_ops[4];              // executed TRAM ops per sample period that move pointer iw, ir, xw, xr (0 or 1 each)
    int tr_on[2];               // TRAM t (0 = iTRAM, 1 = xTRAM) has a READ stream
    uint32_t tr_stage[2];       // byte offset of the stream's stage rows [2 buffers][M]
    uint32_t tr_y[2];           // byte offset of the row holding the READ's offset operand
    uint32_t tr_wy[2];          // ... of the same TRAM's WRITE (0xffffffff: no WRITE instruction)
    int tr_wfirst[2];           // the WRITE comes before the READ in program order
    int tr_base_valid;          // a launch cut along time: the ring pointers at its first period come from the host (tr_base) — the owner of the
    int tr_base[4];             // last period writes the new ones to `ptrs`, and a thread block scheduled after it must not start from those
    int pdl_late_wait;
};

// Per-thread context of the stateless kernel, handed to the batch executor.
__device__ __forceinline__ uint64_t pin64s(uint64_t v) { uint64_t o; asm volatile("mov.b64 %0, %1;" : "=l"(o) : "l"(v)); return o; }

template <int K> struct SLCtx {
    uint32_t col_s;                 // this thread's shared-memory column (32-bit shared address)
    uint32_t tab_s;                 // staged tables + this lane's replica (32-bit shared address)
    const uint4* prog;
    float* out_b;                   // output row of the batch's first sample, this thread's instances
    int N, inst0, n_exec;
    uint64_t Nl;                    // N, pinned in a register (ptxas otherwise re-reads the parameter bank inside the sample loops)
    size_t out_cstride;
    bool valid;
    uint32_t boff;                  // byte offset of the current input-stage buffer
    unsigned int flags;
    Vec<K> acc_last;
    // TRAM
    int32_t tp[4][K];               // iw, ir, xw, xr of this thread's instances
    bool tram_fast;                 // the READ streams are prefetched (else: synchronous reads, one sample at a time)
    bool split;                     // the kernel sets the TRAM pointers before every sl_exec call (several threads share the column, or the READs are prefetched)
    float* ring[2];                 // iTRAM / xTRAM at this thread's first instance
    int rsize[2];
};

// Runs the whole program for samples [m_lo, m_hi) of the current batch, instruction-major.
//   FINAL = false: the bulk path — stores only what a later instruction or the caller can see
//                  (live registers, a CCR somebody reads, the output block).
//   FINAL = true : re-run of the call's last sample by its owner, storing EVERYTHING the state
//                  write-back needs (all result registers, CCR, output latch, accumulator) and nothing
//                  to the output block.  Stateless programs are idempotent per sample, so the re-run
//                  reproduces the same values.
// All operand addresses are running 32-bit shared addresses (one add per operand and sample).
// One decoded instruction, ready to run over the samples [m_lo, m_lo + n_m) of the batch.
struct SLInstr {
    uint32_t w0, uop, aux;              // flags word, micro-op, table word (wB.y)
    uint32_t qr, qa, qx, qy, qccr;      // running 32-bit shared addresses of R, A, X, Y, CCR for the current sample
    uint32_t sr, sa, sx, sy, sccr;      // bytes between consecutive samples (0: the same row for the whole batch)
    bool ca, cx, cy;                    // operand is this instruction's own previous result (self recurrence)
    bool st_r, st_c, st_o;              // store R / CCR / the output block
    float* qo;                          // output block slot of the current sample
    int n_m;
    // fused consumer (F_FUSE): R2 = sat(fc0 + r * fc1) (result forwarded into the product) or sat(r + fc0) (into the addend; fc0 = X2 * Y2)
    uint32_t qr2, sr2;                  // running shared address / per-sample stride of the consumer's R
    bool st_r2, st_o2;
    float* qo2;
    float fc0[4], fc1[4];
};

// Runs ONE instruction over the batch.  CM = how results are carried from sample to sample: 0 nothing (stateless
// instruction), 1 operand A only (the common recurrence: `interp out, out, c, in`, `macs a, a, x, y`), 2 any mix.
// The sample loop is software-pipelined over TWO operand register sets (unrolled by two, no register moves): the
// operands of sample m + 1 go in flight before the arithmetic of sample m, operands whose row does not change are
// read once, and a carried operand is written straight into the next set — a recurrence's chain holds arithmetic only.
template <int K, bool FINAL, bool CCRV, int CM, bool TRAM, int FUSE = 0>     // FUSE: 0 none, 1 result forwarded into the consumer's addend, 2 into its product,
                                                                            //       3 / 4 the consumer is the iTRAM / xTRAM WRITE of the result
__device__ __forceinline__ void sl_run(const SLParams& p, SLCtx<K>& cx, SLInstr& I) {
    const uint64_t Nl = cx.Nl;
    const uint32_t w0 = I.w0, uop = I.uop;
    const int n_m = I.n_m;
    const bool la = I.sa != 0u, lx = I.sx != 0u, ly = I.sy != 0u;   // the operand's row changes from sample to sample
#define SL_EACH _Pragma("unroll") for (int k = 0; k < K; ++k)
    // R store (:1079-1082 etc.), setCCR (:211-232), output (:1229-1233, :1248) for the current sample
#define SL_WRITE(SETS_ACC)                                                                                       \
        if (I.st_r) sts<K>(I.qr, r);                                                                             \
        if (FINAL || CCRV) { if (I.st_c) { Vec<K> c; SL_EACH { c[k] = ccr_of(r[k]); } sts<K>(I.qccr, c); } }      \
        if (!FINAL) { if (I.st_o) vstore<K>(I.qo, r); }                                                          \
        else {                                                                                                   \
            if (I.st_o) vstore<K>(p.latch + (size_t)(w0 >> 24) * cx.N + cx.inst0, r);                             \
            if (SETS_ACC) cx.acc_last = accv;                                                                    \
        }
    // what one sample's result does: this instruction's own stores, or its fused consumer; then every running address moves on
#define SL_EMIT(SETS_ACC)                                                                                        \
        if (FUSE == 0) { SL_WRITE(SETS_ACC) }                                                                    \
        else if (FUSE <= 2) {   /* the consumer MACS / MACSN (:1077-1094) on the forwarded result; this instruction's own R is dead in the bulk path */ \
            Vec<K> r2;                                                                                           \
            SL_EACH { r2[k] = sat1(FUSE == 1 ? __fadd_rn(r[k], I.fc0[k]) : __fadd_rn(I.fc0[k], __fmul_rn(r[k], I.fc1[k]))); } \
            if (I.st_r2) sts<K>(I.qr2, r2);                                                                      \
            if (I.st_o2) vstore<K>(I.qo2, r2);                                                                   \
            I.qr2 += I.sr2; I.qo2 += Nl;                                                                         \
        } else {                /* the consumer IDELAY / XDELAY WRITE (:1195-1198 / :1207-1210, writeSmallDelay :909-917): the result goes straight to the ring */ \
            if (K > 1 && fw_same) { if (cx.tp[2 * FT][0] + fw_pos[0] < fw_size && cx.valid) vstore<K>(fw_q[0], r); } \
            else { SL_EACH { if (cx.tp[2 * FT][k] + fw_pos[k] < fw_size && cx.valid) *fw_q[k] = r[k]; } }         \
            SL_EACH {                                                                                            \
                int32_t& wp = cx.tp[2 * FT][k];                                                                  \
                const bool wrap = (++wp == fw_size);                                                             \
                wp = wrap ? 0 : wp;                                                                              \
                fw_q[k] = wrap ? fw_ring + (uint64_t)fw_pos[k] * Nl + k : fw_q[k] + Nl;                          \
            }                                                                                                    \
        }                                                                                                        \
        if (FUSE == 0) { I.qr += I.sr; I.qccr += I.sccr; I.qo += Nl; }      /* (a fused producer stores nothing of its own in the bulk path) */
    // one sample: operands in set CUR, the next sample's go to set NXT
#define SL_HALF(SETS_ACC, LOADS_XY, CUR, NXT, ...)                                                               \
    {                                                                                                            \
        Vec<K>&a = A##CUR, &x = X##CUR, &y = Y##CUR; Vec<K> r, accv;                                             \
        (void)x; (void)y;                                                                                        \
        if (!FINAL && m + 1 < n_m) {                                                                             \
            if (la) { I.qa += I.sa; A##NXT = lds<K>(I.qa); }                                                     \
            if (LOADS_XY && lx) { I.qx += I.sx; X##NXT = lds<K>(I.qx); }                                         \
            if (LOADS_XY && ly) { I.qy += I.sy; Y##NXT = lds<K>(I.qy); }                                         \
        }                                                                                                        \
        __VA_ARGS__                                                                                              \
        SL_EMIT(SETS_ACC)                                                                                        \
        if (CM == 1) A##NXT = r;                                                                                 \
        else if (CM == 2) { SL_EACH { if (I.ca) A##NXT[k] = r[k]; if (I.cx) X##NXT[k] = r[k]; if (I.cy) Y##NXT[k] = r[k]; } } \
        ++m;                                                                                                     \
    }
#define SL_LOOP(SETS_ACC, LOADS_XY, ...)                                                                         \
    {                                                                                                            \
        int m = 0;                                                                                               \
        if (!FINAL) {                                                                                            \
            _Pragma("unroll 1") while (m + 1 < n_m) {                                                            \
                SL_HALF(SETS_ACC, LOADS_XY, 0, 1, __VA_ARGS__) SL_HALF(SETS_ACC, LOADS_XY, 1, 0, __VA_ARGS__)     \
            }                                                                                                    \
        }                                                                                                        \
        if (m < n_m) SL_HALF(SETS_ACC, LOADS_XY, 0, 1, __VA_ARGS__)                                              \
    }
    const bool tab_op = (uop == U_LOG || uop == U_EXP);
    const bool dyn_sel = tab_op && !(w0 & (F_TAB_SMEM | F_TAB_IMM));     // LOG/EXP reading its selector from X every sample
    Vec<K> A0 = lds<K>(I.qa), X0, Y0;                                    // the first sample's operands (in both sets: rows that never change stay put)
    if (!tab_op) { X0 = lds<K>(I.qx); Y0 = lds<K>(I.qy); }
    else {
        SL_EACH { X0[k] = 0.0f; Y0[k] = 0.0f; }                          // LOG/EXP never read Y (:1114 TODO), and X only as a dynamic selector
        if (dyn_sel) X0 = lds<K>(I.qx);
    }
    Vec<K> A1 = A0, X1 = X0, Y1 = Y0;
    // fused TRAM WRITE consumer: running address of slot wp + pos per context (slots at or beyond the ring are dropped, as in SL_TRAM_WRITE)
    constexpr int FT = FUSE == 4 ? 1 : 0;
    int fw_pos[K];
    float* fw_q[K];
    bool fw_same = true;
    const int fw_size = cx.rsize[FT];
    float* const fw_ring = cx.ring[FT];
    if (FUSE >= 3) {
        const Vec<K> wy = lds<K>(I.qr2);                                 // the WRITE's offset operand (one value for the whole batch)
        SL_EACH {
            fw_pos[k] = min(max(cvt_x86(wy[k]), 0), fw_size - 1);
            fw_q[k] = fw_ring + (uint64_t)(cx.tp[2 * FT][k] + fw_pos[k]) * Nl + k;
            fw_same = fw_same && (cx.tp[2 * FT][k] + fw_pos[k] == cx.tp[2 * FT][0] + fw_pos[0]);
        }
    }
    switch (uop) {
    case U_MACS: SL_LOOP(true, true,
        SL_EACH { accv[k] = __fadd_rn(a[k], __fmul_rn(x[k], y[k])); r[k] = sat1(accv[k]); }) break;
    case U_MACSN: SL_LOOP(true, true,
        SL_EACH { accv[k] = __fsub_rn(a[k], __fmul_rn(x[k], y[k])); r[k] = sat1(accv[k]); }) break;
    case U_ACC3: SL_LOOP(true, true,
        SL_EACH { accv[k] = __fadd_rn(__fadd_rn(a[k], x[k]), y[k]); r[k] = sat1(accv[k]); }) break;
    case U_MACW: SL_LOOP(true, true,
        SL_EACH { r[k] = __fadd_rn(a[k], wrap1(__fmul_rn(x[k], y[k]))); accv[k] = r[k]; }) break;
    case U_MACWN: SL_LOOP(true, true,
        SL_EACH { r[k] = __fsub_rn(a[k], wrap1(__fmul_rn(x[k], y[k]))); accv[k] = r[k]; }) break;
    case U_MACINTW: SL_LOOP(true, true,
        SL_EACH { r[k] = wrap1(__fadd_rn(a[k], __fmul_rn(x[k], y[k]))); accv[k] = r[k]; }) break;
    case U_ANDXOR: SL_LOOP(false, true,
        SL_EACH { r[k] = __int2float_rn(logic_ops(a[k], x[k], y[k])); accv[k] = 0.0f; }) break;
    case U_TSTNEG: SL_LOOP(true, true,
        SL_EACH {
            const int32_t q = cvt_x86(__fmul_rn(x[k], 2147483648.0f));
            r[k] = (a[k] >= y[k]) ? x[k] : __fmul_rn(__int2float_rn(~q), 4.656612873077392578125e-10f); accv[k] = r[k];
        }) break;
    case U_LIMIT: SL_LOOP(true, true,
        SL_EACH { r[k] = (a[k] >= y[k]) ? x[k] : y[k]; accv[k] = r[k]; }) break;
    case U_LIMITN: SL_LOOP(true, true,
        SL_EACH { r[k] = (a[k] < y[k]) ? x[k] : y[k]; accv[k] = r[k]; }) break;
    case U_INTERP: {
        const bool x_varies = lx || I.cx;         // a constant coefficient: 1.0 - X is formed once per batch
        double omx[K];
        SL_EACH { omx[k] = __dsub_rn(1.0, (double)X0[k]); }
        // (two copies of the loop: with `if (x_varies)` inside one, ptxas converts and selects 1 - X every sample period
        //  — a fourth conversion-unit instruction per context on the path of a one-pole recurrence)
        if (x_varies) {
            SL_LOOP(true, true,
                one_minus<K>(x.v, omx);
                interp_core<K>(omx, a.v, x.v, y.v, accv.v);
                SL_EACH { r[k] = sat1(accv[k]); })
        } else {
            SL_LOOP(true, true,
                interp_core<K>(omx, a.v, x.v, y.v, accv.v);
                SL_EACH { r[k] = sat1(accv[k]); })
        }
        break; }
    case U_LOG:
    case U_EXP: {
        const uint32_t tb_s = cx.tab_s + (I.aux >> 24) * (uint32_t)TAB_SMEM_BYTES;
        if (!FINAL && CM == 0 && (w0 & F_TAB_SMEM)) {
            // A literal table staged in shared memory, nothing carried (the waveshaper of cfg2): TWO samples per iteration with one
            // range test for both, so that the 2 K dependency chains (F2F -> 3 x FP64 -> gather -> 3 x FP64 -> F2F, about 110 cycles
            // each) sit in ONE basic block and ptxas interleaves them; the operands of the next pair are fetched meanwhile.
#define SL_TAB_FAST(src, dst) SL_EACH { double di; const double xd = (double)src[k]; const int ix = table_index_inrange(xd, di); double y1, sl; \
                                        lds_f64x2(tb_s + (uint32_t)ix * (TAB_REPL * 16u), y1, sl); dst[k] = table_finish(xd, di, y1, sl); }
#define SL_TAB_SLOW(src, dst) SL_EACH { const bool in_range = fabsf(src[k]) <= 1.0f; double di; const double xd = (double)src[k]; int ix; \
                                        if (in_range) ix = table_index_inrange(xd, di); else { ix = table_index_wild(src[k]); di = (double)ix; cx.flags |= FX8010_RT_TABLE_RANGE; } /* rule U6 */ \
                                        double y1, sl; lds_f64x2(tb_s + (uint32_t)ix * (TAB_REPL * 16u), y1, sl); dst[k] = table_finish(xd, di, y1, sl); }
            int m = 0;
            Vec<K> B0 = A0, B1 = A0;
            if (n_m > 1 && la) { I.qa += I.sa; B1 = lds<K>(I.qa); }
            _Pragma("unroll 1") while (m + 1 < n_m) {
                Vec<K> N0 = B0, N1 = B1;
                if (la && m + 2 < n_m) { I.qa += I.sa; N0 = lds<K>(I.qa); }
                if (la && m + 3 < n_m) { I.qa += I.sa; N1 = lds<K>(I.qa); }
                bool wild = false;
                SL_EACH { wild |= !(fabsf(B0[k]) <= 1.0f) || !(fabsf(B1[k]) <= 1.0f); }
                Vec<K> r0, r1;
                if (!wild) { SL_TAB_FAST(B0, r0) SL_TAB_FAST(B1, r1) }
                else { SL_TAB_SLOW(B0, r0) SL_TAB_SLOW(B1, r1) }
                { Vec<K>& r = r0; Vec<K>& accv = r0; SL_EMIT(true) }
                { Vec<K>& r = r1; Vec<K>& accv = r1; SL_EMIT(true) }
                B0 = N0; B1 = N1; m += 2;
            }
            if (m < n_m) { Vec<K> r; SL_TAB_SLOW(B0, r) Vec<K>& accv = r; SL_EMIT(true) }
#undef SL_TAB_FAST
#undef SL_TAB_SLOW
            break;
        }
        SL_LOOP(true, false,
            if (dyn_sel && lx && m + 1 < n_m) { I.qx += I.sx; if (m & 1) X0 = lds<K>(I.qx); else X1 = lds<K>(I.qx); }
            int idx[K];
            double xd[K], di[K];
            bool wild = false;
            SL_EACH { wild |= !(fabsf(a[k]) <= 1.0f); xd[k] = (double)a[k]; }
            if (!wild) { SL_EACH { idx[k] = table_index_inrange(xd[k], di[k]); } }
            else { SL_EACH { idx[k] = table_index_wild(a[k]); di[k] = (double)idx[k]; if (!(fabsf(a[k]) <= 1.0f)) cx.flags |= FX8010_RT_TABLE_RANGE; } }   // rule U6
            if (w0 & F_TAB_SMEM) {
                SL_EACH { double y1; double slope; lds_f64x2(tb_s + (uint32_t)idx[k] * (TAB_REPL * 16u), y1, slope); r[k] = table_finish(xd[k], di[k], y1, slope); }
            } else {
                SL_EACH {
                    int tsel;
                    if (w0 & F_TAB_IMM) tsel = (int)(I.aux >> 24);
                    else {
                        int32_t sel = cvt_x86(x[k]);
                        if (sel < 0 || sel > FX8010_TABLE_COUNT - 1) { cx.flags |= FX8010_RT_TABLE_RANGE; sel = sel < 0 ? 0 : FX8010_TABLE_COUNT - 1; }
                        tsel = (uop == U_EXP ? FX8010_TABLE_COUNT : 0) + sel;
                    }
                    const double2 e = __ldg(reinterpret_cast<const double2*>(p.tabs + tsel * FX8010_TABLE_ENTRIES + idx[k]));
                    r[k] = table_finish(xd[k], di[k], e.x, e.y);
                }
            }
            SL_EACH { accv[k] = r[k]; })
        break; }
        // TRAM READ (:1190-1193 / :1202-1205, readSmallDelay :934-956) and WRITE (:1195-1198 / :1207-1210,
        // writeSmallDelay :909-917); T = 0 iTRAM, 1 xTRAM (a constant, so that the pointers stay in registers).
        // The final-state pass re-runs arithmetic only: TRAM state is already final.
#define SL_TRAM_READ(T)                                                                                          \
        if (TRAM && !FINAL && FUSE == 0) {                                                                                    \
            const int size = cx.rsize[T];                                                                        \
            if (!cx.tram_fast) {   /* (prefetched READs never get here: sl_exec skips them, the kernel moves the pointers per batch) */ \
                const float* const ring = cx.ring[T];                                                            \
                for (int m = 0; m < n_m; ++m, I.qa += I.sa) {                                                    \
                    Vec<K> v;                                                                                    \
                    SL_EACH {                                                                                    \
                        int32_t& rp = cx.tp[2 * T + 1][k];                                                       \
                        const int pos = min(max(cvt_x86(Y0[k]), 0), size - 1);                                   \
                        int idx = rp - pos;                                                                      \
                        idx += (idx < 0) ? size : 0;   /* rule U1: mathematical modulo */                        \
                        rp = (rp + 1 == size) ? 0 : rp + 1;                                                      \
                        v[k] = ring[(uint64_t)idx * Nl + k];                                                     \
                    }                                                                                            \
                    sts<K>(I.qa, v);                                                                             \
                }                                                                                                \
            }                                                                                                    \
        }
#define SL_TRAM_WRITE(T)                                                                                         \
        if (TRAM && !FINAL && FUSE == 0) {                                                                                    \
            const int size = cx.rsize[T];                                                                        \
            float* const ring = cx.ring[T];                                                                      \
            int pos[K];                                                                                          \
            float* wq[K];            /* running address of slot wp + pos (not wrapped in the reference: slots at or */ \
            bool same = true;        /* beyond the ring are never read back -> dropped) */                       \
            SL_EACH {                                                                                            \
                pos[k] = min(max(cvt_x86(Y0[k]), 0), size - 1);                                                  \
                wq[k] = ring + (uint64_t)(cx.tp[2 * T][k] + pos[k]) * Nl + k;                                    \
                same = same && (cx.tp[2 * T][k] + pos[k] == cx.tp[2 * T][0] + pos[0]);                           \
            }                                                                                                    \
            Vec<K> a = A0;                                                                                       \
            for (int m = 0; m < n_m; ++m) {                                                                      \
                Vec<K> an = a;                                                                                   \
                if (la && m + 1 < n_m) { I.qa += I.sa; an = lds<K>(I.qa); }                                      \
                if (K > 1 && same) { if (cx.tp[2 * T][0] + pos[0] < size && cx.valid) vstore<K>(wq[0], a); }     \
                else { SL_EACH { if (cx.tp[2 * T][k] + pos[k] < size && cx.valid) *wq[k] = a[k]; } }             \
                SL_EACH {                                                                                        \
                    int32_t& wp = cx.tp[2 * T][k];                                                               \
                    const bool wrap = (++wp == size);                                                            \
                    wp = wrap ? 0 : wp;                                                                          \
                    wq[k] = wrap ? ring + (uint64_t)pos[k] * Nl + k : wq[k] + Nl;                                \
                }                                                                                                \
                a = an;                                                                                          \
            }                                                                                                    \
        }
    case U_IREAD: SL_TRAM_READ(0) break;
    case U_XREAD: SL_TRAM_READ(1) break;
    case U_IWRITE: SL_TRAM_WRITE(0) break;
    case U_XWRITE: SL_TRAM_WRITE(1) break;
#undef SL_TRAM_READ
#undef SL_TRAM_WRITE
    default: break;      // nothing else can appear in this kernel's encoded stream
    }
#undef SL_EACH
#undef SL_WRITE
#undef SL_EMIT
#undef SL_HALF
#undef SL_LOOP
}

// Runs the whole program for samples [m_lo, m_hi) of the current batch, instruction-major.
// The encoded stream holds four words per instruction with everything the host can work out at load time already
// worked out (fx8010_gpu.cu::encode_stateless) — per batch and instruction the thread only adds its column, the stage
// buffer in use and, in the delay-line kernel, the first sample of its share:
//   Q0 { uop | flags | carry | out channel << 24,  table slot/id << 24,  pair bits (F_FUSE),  0 }
//   Q1 { byte offset of R, A, X, Y at sample 0 (a self-carried operand: its row M - 1);  bit 31: a stage row, add the buffer in use }
//   Q2 { bytes between consecutive samples for R, A, X, Y (0: one row for the whole batch, or self-carried) }
//   Q3 { byte offset of CCR, its bytes per sample, 0, 0 }
template <int K, bool FINAL, bool CCRV, bool TRAM>
__device__ __forceinline__ void sl_exec(const SLParams& p, SLCtx<K>& cx, const int m_lo, const int m_hi) {
    uint4 n0 = cx.prog[0], n1 = cx.prog[1], n2 = cx.prog[2], n3 = cx.prog[3];
    const int n_exec = cx.n_exec;
    const uint32_t mlo = (uint32_t)m_lo;
    for (int pc = 0; pc < n_exec; ++pc) {
        const uint4 q0 = n0, q1 = n1, q2 = n2, q3 = n3;
        n0 = cx.prog[4 * pc + 4]; n1 = cx.prog[4 * pc + 5]; n2 = cx.prog[4 * pc + 6]; n3 = cx.prog[4 * pc + 7];   // (the stream ends with one pad record)
        const uint32_t w0 = q0.x;
        if (TRAM) {     // a prefetched READ has nothing to execute (its rows arrive with the batch, the kernel moves the pointers); the final-state pass re-runs arithmetic only
            const uint32_t u0 = w0 & 0xffu;
            if ((u0 == U_IREAD || u0 == U_XREAD) && (FINAL || cx.tram_fast)) continue;
        }
        // address of an operand at sample m_lo: column + offset (+ the stage buffer in use) + m_lo * stride
#define SL_Q(off, str) (cx.col_s + ((off) & 0x7fffffffu) + ((uint32_t)((int32_t)(off) >> 31) & cx.boff) + mlo * (str))
        // a self-carried operand (the register is this instruction's own R): row M - 1 before sample 0, else the row of sample m_lo - 1
#define SL_QC(off) (cx.col_s + (off) + (m_lo == 0 ? 0u : (uint32_t)(m_lo - p.M) * q2.x))
        SLInstr I;
        I.w0 = w0; I.uop = w0 & 0xffu; I.aux = q0.y; I.n_m = m_hi - m_lo;
        const uint32_t cbits = (w0 >> SL_CARRY_SHIFT) & 7u;
        I.ca = cbits & 1u; I.cx = cbits & 2u; I.cy = cbits & 4u;
        I.sr = q2.x; I.sa = q2.y; I.sx = q2.z; I.sy = q2.w; I.sccr = q3.y;
        I.qr = cx.col_s + q1.x + mlo * q2.x;
        I.qa = I.ca ? SL_QC(q1.y) : SL_Q(q1.y, q2.y); I.qx = I.cx ? SL_QC(q1.z) : SL_Q(q1.z, q2.z); I.qy = I.cy ? SL_QC(q1.w) : SL_Q(q1.w, q2.w);
        I.qccr = cx.col_s + q3.x + mlo * q3.y;
        I.st_r = FINAL || !(w0 & F_ST_LAST);
        I.st_c = FINAL || (CCRV && (w0 & F_CCR));
        I.st_o = (w0 & F_OUT_DIRECT) && cx.valid;
        I.qo = cx.out_b + (size_t)(w0 >> 24) * cx.out_cstride + (size_t)m_lo * cx.Nl;
        if (FINAL || CCRV) sl_run<K, FINAL, CCRV, 2, TRAM>(p, cx, I);          // cold paths: one general copy
        else if ((w0 & F_FUSE) && (q0.z & 8u)) {
            // producer + TRAM WRITE: the result is stored straight into the ring (n1.w = the WRITE's offset operand, one row)
            if constexpr (TRAM) {
                I.qr2 = cx.col_s + n1.w;
                if (q0.z & 16u) sl_run<K, false, false, 0, TRAM, 4>(p, cx, I); else sl_run<K, false, false, 0, TRAM, 3>(p, cx, I);
            }
            ++pc;
            n0 = cx.prog[4 * pc + 4]; n1 = cx.prog[4 * pc + 5]; n2 = cx.prog[4 * pc + 6]; n3 = cx.prog[4 * pc + 7];
        }
        else if (w0 & F_FUSE) {
            // producer + consumer in one sample loop: n0..n2 hold the consumer (a MACS / MACSN with two batch-constant operands)
            const uint32_t v0 = n0.x;
            const bool into_product = q0.z & 1u, negate = q0.z & 2u;
            const Vec<K> ca = lds<K>(cx.col_s + n1.y), cxx = lds<K>(cx.col_s + n1.z), cy = lds<K>(cx.col_s + n1.w);   // (constant rows: plain offsets; the forwarded one is not used)
            const bool fwd_x = into_product && (q0.z & 4u);                    // which factor of the product is the forwarded one
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (into_product) { const float m = fwd_x ? cy[k] : cxx[k]; I.fc0[k] = ca[k]; I.fc1[k] = negate ? -m : m; }
                else { const float pr = __fmul_rn(cxx[k], cy[k]); I.fc0[k] = negate ? -pr : pr; I.fc1[k] = 0.0f; }
            }
            I.sr2 = n2.x; I.qr2 = cx.col_s + n1.x + mlo * n2.x;
            I.st_r2 = !(v0 & F_ST_LAST);
            I.st_o2 = (v0 & F_OUT_DIRECT) && cx.valid;
            I.qo2 = cx.out_b + (size_t)(v0 >> 24) * cx.out_cstride + (size_t)m_lo * cx.Nl;
            if (into_product) sl_run<K, false, false, 0, TRAM, 2>(p, cx, I); else sl_run<K, false, false, 0, TRAM, 1>(p, cx, I);
            ++pc;                                                              // the consumer is done
            n0 = cx.prog[4 * pc + 4]; n1 = cx.prog[4 * pc + 5]; n2 = cx.prog[4 * pc + 6]; n3 = cx.prog[4 * pc + 7];
        }
        else if (cbits == 0u) sl_run<K, false, false, 0, TRAM>(p, cx, I);
        else if (cbits == 1u) sl_run<K, false, false, 1, TRAM>(p, cx, I);
        else sl_run<K, false, false, 2, TRAM>(p, cx, I);
#undef SL_Q
#undef SL_QC
    }
}

// ---- bulk tensor copies (TMA) into the input stage: mbarrier + cp.async.bulk.tensor ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
    asm volatile("{\n\t.reg .pred P1;\n\tFXK_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra FXK_DONE;\n\tbra FXK_WAIT;\n\tFXK_DONE:\n\t}" ::"r"(bar), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, uint32_t bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y) : "memory");
}

// (x + inc) mod size for 0 <= x < size, inc >= 0; rings longer than the increment (the common case) need no division
__device__ __forceinline__ int ring_add(int x, int inc, int size) {
    x += inc;
    if (size > inc) return x - ((x >= size) ? size : 0);
    return size > 0 ? x % size : 0;
}

template <int K, bool TRAM>
__global__ void __launch_bounds__(128, 3) fx_stateless_kernel(const __grid_constant__ SLParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int P = TRAM ? p.P : 1;                      // threads per instance column; they split each batch's samples
    const int B = blockDim.x / P;                      // instance threads (columns) per block
    const int part = TRAM ? (int)threadIdx.x / B : 0;  // this thread's share of every batch: samples [part, part + 1) * M / P
    const int tid = TRAM ? (int)threadIdx.x - part * B : (int)threadIdx.x;
    const int N = p.N, C = p.C, M = p.M;
    const int tslot_raw = blockIdx.x * B + tid;
    const bool valid = tslot_raw * K < N;
    const int inst0 = valid ? tslot_raw * K : N - K;
    const int blk = (p.n_blk > 1) ? (int)blockIdx.y / p.n_seg : 0;       // sample block of the call this thread block works on
    const int seg = (int)blockIdx.y - blk * p.n_seg;
    const float* const in_base = p.blk_in[blk];
    float* const out_base = p.blk_out[blk];
    const int s_begin = seg * p.seg_len;
    const int s_end = min(p.n_samples, s_begin + p.seg_len);
    const uint32_t row_bytes = (uint32_t)B * K * 4u;
    const uint32_t buf_bytes = (uint32_t)M * row_bytes;
    const int sub = M / P;                             // samples of a batch per thread (P divides M)

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (!p.pdl_late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");

    TableEntry* const s_tab = reinterpret_cast<TableEntry*>(smem_raw);
    unsigned char* const col = smem_raw + (size_t)p.n_smem_tabs * TAB_SMEM_BYTES + (size_t)tid * K * 4;
    auto at = [&](uint32_t byte_off) { return reinterpret_cast<float*>(col + byte_off); };

    // this thread's samples [lo, hi) of a batch of nm samples
    auto my_lo = [&](int nm) { return min(nm, part * sub); };
    auto my_hi = [&](int nm) { return min(nm, part * sub + sub); };

    // input stage: batch starting at sample s0 -> buffer at byte offset boff (channel 0 unrolled)
    const bool has_in = (in_base != nullptr);
    auto fetch_batch = [&](int s0, uint32_t boff) {
        if (has_in && s0 < s_end) {
            const int nm = min(M, s_end - s0);
            const int lo = my_lo(nm), hi = my_hi(nm);
            const float* g = in_base + (size_t)(s0 + lo) * N + inst0;
            unsigned char* d = reinterpret_cast<unsigned char*>(at(p.stage0 + boff)) + (uint32_t)lo * row_bytes;
#pragma unroll 4
            for (int m = lo; m < hi; ++m, d += row_bytes, g += N) cp_async<4 * K>(d, g);
            for (int c = 1; c < C; ++c) {
                const float* gc = in_base + (size_t)c * p.in_cstride + (size_t)(s0 + lo) * N + inst0;
                const uint32_t base = p.stage0 + (uint32_t)c * 2u * buf_bytes + boff;
                for (int m = lo; m < hi; ++m, gc += N) cp_async<4 * K>(at(base + (uint32_t)m * row_bytes), gc);
            }
        }
    };
    // TMA variant of the input stage (delay-free kernel only): thread 0 arms one mbarrier per stage buffer with the bytes of a
    // batch and issues ONE bulk tensor copy per channel ([M rows][B * K instances], rows past the block's range or the
    // tensor's end are harmless / zero-filled); everybody waits on the barrier's phase instead of cp.async.wait_group.
    const bool tma = !TRAM && p.use_tma && has_in;
    __shared__ __align__(8) unsigned long long s_mbar[2];
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(s_mbar);
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(smem_raw) + (uint32_t)p.n_smem_tabs * TAB_SMEM_BYTES + p.stage0;   // column 0 of the stage rows
    auto fetch_tma = [&](int s0, uint32_t boff) {
        if (threadIdx.x == 0 && s0 < s_end) {
            const uint32_t bar = bar_s + (boff ? 8u : 0u);
            mbar_expect_tx(bar, (uint32_t)C * buf_bytes);
            for (int c = 0; c < C; ++c)
                tma_load_2d(stage_s + (uint32_t)c * 2u * buf_bytes + boff, &p.in_map, bar, blockIdx.x * B * K, c * p.tma_rows_per_channel + s0);
        }
    };
    if (tma) {
        if (threadIdx.x == 0) { mbar_init(bar_s, 1); mbar_init(bar_s + 8u, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    } else
    fetch_batch(s_begin, 0);                            // (its group is committed after the TRAM streams joined it, below)
    if (!has_in && part == 0) {                         // no input block: INPUT operands read silence (their rows are the stage rows)
        Vec<K> z;
#pragma unroll
        for (int k = 0; k < K; ++k) z[k] = 0.0f;
        for (int j = 0; j < 2 * C * M; ++j) vstore<K>(at(p.stage0 + (uint32_t)j * row_bytes), z);
    }

    for (int t = 0; t < p.n_smem_tabs; ++t) {
        const TableEntry* src = p.tabs + (size_t)p.smem_tab_id[t] * FX8010_TABLE_ENTRIES;
#pragma unroll 4
        for (int i = (int)threadIdx.x; i < FX8010_TABLE_ENTRIES * TAB_REPL; i += (int)blockDim.x)
            s_tab[t * FX8010_TABLE_ENTRIES * TAB_REPL + i] = src[i / TAB_REPL];
    }
    if (part == 0) {   // read-only rows (controls, literals) and carried-in rows: the only state such a program can see
        int j = 0;
        for (; j + 4 <= p.n_load; j += 4) {
            Vec<K> t[4];
            uint2 e[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { e[q] = p.load_list[j + q]; t[q] = vload<K>(p.gpr + (size_t)e[q].y * N + inst0); }
#pragma unroll
            for (int q = 0; q < 4; ++q) vstore<K>(at(e[q].x & SL_OFF_MASK), t[q]);
        }
        for (; j < p.n_load; ++j) { const uint2 e = p.load_list[j]; vstore<K>(at(e.x & SL_OFF_MASK), vload<K>(p.gpr + (size_t)e.y * N + inst0)); }
    }
    __syncthreads();

    SLCtx<K> cx;
    cx.col_s = (uint32_t)__cvta_generic_to_shared(col);
    cx.tab_s = (uint32_t)__cvta_generic_to_shared(s_tab) + (uint32_t)(tid & (TAB_REPL - 1)) * 16u;
    cx.prog = c_prog + p.prog_off;
    cx.N = N; cx.Nl = pin64s((uint64_t)N); cx.inst0 = inst0; cx.valid = valid; cx.boff = 0; cx.flags = 0;
    cx.n_exec = p.n_exec; cx.out_cstride = p.out_cstride;
#pragma unroll
    for (int k = 0; k < K; ++k) cx.acc_last[k] = 0.0f;
    cx.out_b = out_base + (size_t)s_begin * N + inst0;
    const size_t out_step = (size_t)M * N;

    // ---- TRAM: pointers, and the READ streams (source/FX8010.cpp:934-967) prefetched like input channels ----
    // Every executed READ / WRITE moves its pointer by one, so with one READ and one WRITE per TRAM the slot a READ
    // fetches was written a constant number of sample periods earlier (`dist`).  A batch's READs are fetched while the
    // previous batch is still being computed: that is the same data as long as dist > 2 M.
    cx.tram_fast = true;
    cx.split = TRAM && P > 1;
    cx.ring[0] = p.itram + inst0; cx.ring[1] = p.xtram + inst0; cx.rsize[0] = p.itram_size; cx.rsize[1] = p.xtram_size;
    int32_t tp0[4][K];                                  // iw, ir, xw, xr at the start of the current batch
    int tr_next[2][K];                                  // ring slot of the streams' first sample of the NEXT batch to fetch
    bool tr_same[2] = {true, true};
    if (TRAM && p.has_tram) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (p.tr_base_valid) { tp0[0][k] = p.tr_base[0]; tp0[1][k] = p.tr_base[1]; tp0[2][k] = p.tr_base[2]; tp0[3][k] = p.tr_base[3]; }
            else {
                tp0[0][k] = p.ptrs[inst0 + k]; tp0[1][k] = p.ptrs[N + inst0 + k];
                tp0[2][k] = p.ptrs[2 * N + inst0 + k]; tp0[3][k] = p.ptrs[3 * N + inst0 + k];
            }
            if (s_begin) {          // a later time segment of a launch whose periods are independent (fx8010_gpu.cu::known_tram_span): pointers at its first period
#pragma unroll
                for (int j = 0; j < 4; ++j) if (p.tr_ops[j]) tp0[j][k] = ring_add(tp0[j][k], s_begin * p.tr_ops[j], cx.rsize[j >> 1]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) cx.tp[j][k] = tp0[j][k];    // (a column with one thread just lets them run)
        }
        bool safe = true;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            if (!p.tr_on[t]) continue;
            const int size = cx.rsize[t];
            const Vec<K> yv = lds<K>(cx.col_s + p.tr_y[t]);
            Vec<K> wv = yv;
            const bool has_w = p.tr_wy[t] != 0xffffffffu;
            if (has_w) wv = lds<K>(cx.col_s + p.tr_wy[t]);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int pos = min(max(cvt_x86(yv[k]), 0), size - 1);
                int nx = tp0[2 * t + 1][k] - pos;
                nx += (nx < 0) ? size : 0;
                tr_next[t][k] = nx;
                tr_same[t] = tr_same[t] && (nx == tr_next[t][0]);
                if (has_w) {
                    const int wpos = min(max(cvt_x86(wv[k]), 0), size - 1);
                    int d = (tp0[2 * t][k] + wpos - nx) % size;      // write slot minus read slot of the same sample
                    d += (d < 0) ? size : 0;
                    const int dist = (d == 0 && !p.tr_wfirst[t]) ? size : d;
                    safe = safe && (dist > 2 * M);
                }
            }
        }
        // threads sharing a column write their samples' slots in no particular order: a written ring must not be
        // shorter than a batch (two samples of one batch would meet in one slot)
        if (P > 1 && ((p.tr_ops[0] > 0 && cx.rsize[0] < M) || (p.tr_ops[2] > 0 && cx.rsize[1] < M))) safe = false;
        if (P > 1) {                                    // threads sharing a column must agree: different warps, the same instances
            __shared__ int s_unsafe;
            if (threadIdx.x == 0) s_unsafe = 0;
            __syncthreads();
            if (!safe) s_unsafe = 1;
            __syncthreads();
            cx.tram_fast = (s_unsafe == 0);
        } else cx.tram_fast = __all_sync(0xffffffffu, safe);
        // With prefetched READs the pointers are set from the batch-start values before every call (a READ then has
        // nothing left to do and is not even decoded); only a lone thread with a short delay lets them run.
        cx.split = (P > 1) || cx.tram_fast;
    }
    // Running global address of each stream's next row for THIS thread (one add per row, a reset where the ring wraps);
    // tr_skip = rows between it and the thread's first row of the next batch (the other threads' shares).
    const float* tr_ptr[2][K];
    int tr_skip[2] = {0, 0};
#pragma unroll
    for (int k = 0; k < K; ++k) {
        tr_ptr[0][k] = cx.ring[0] + (size_t)(TRAM && p.tr_on[0] ? tr_next[0][k] : 0) * N + k;
        tr_ptr[1][k] = cx.ring[1] + (size_t)(TRAM && p.tr_on[1] ? tr_next[1][k] : 0) * N + k;
    }
    auto fetch_stream = [&](auto tc, const int nm, const uint32_t boff) {
        constexpr int t = decltype(tc)::value;
        const float* const ring = cx.ring[t];
        const size_t ring_len = (size_t)cx.rsize[t] * N;
        const float* const ring_end = ring + ring_len;
        const int lo = my_lo(nm), hi = my_hi(nm);
        const int skip = tr_skip[t] + lo;               // (prefetching implies size > 2 M >= skip)
        if (skip) {
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (k == 0 || !tr_same[t]) {
                    const float* q = tr_ptr[t][k] + (size_t)skip * N;
                    tr_ptr[t][k] = (q >= ring_end + k) ? q - ring_len : q;
                }
        }
        tr_skip[t] = nm - hi;
        unsigned char* d = reinterpret_cast<unsigned char*>(at(p.tr_stage[t] + boff)) + (uint32_t)lo * row_bytes;
        for (int m = lo; m < hi; ++m, d += row_bytes) {
            if (K > 1 && tr_same[t]) cp_async<4 * K>(d, tr_ptr[t][0]);
            else {
#pragma unroll
                for (int k = 0; k < K; ++k) cp_async<4>(d + 4 * k, tr_ptr[t][k]);
            }
#pragma unroll
            for (int k = 0; k < K; ++k)                    // (the address itself tells where the ring ends)
                if (k == 0 || !tr_same[t]) {
                    const float* const nx = tr_ptr[t][k] + N;
                    tr_ptr[t][k] = (nx == ring_end + k) ? ring + k : nx;
                }
        }
    };
    auto fetch_tram = [&](int s0, uint32_t boff) {
        if (TRAM && p.n_tr > 0 && cx.tram_fast && s0 < s_end) {
            const int nm = min(M, s_end - s0);
            if (p.tr_on[0]) fetch_stream(std::integral_constant<int, 0>{}, nm, boff);
            if (p.tr_on[1]) fetch_stream(std::integral_constant<int, 1>{}, nm, boff);
        }
    };
    fetch_tram(s_begin, 0);
    cp_async_commit();
    if (tma) fetch_tma(s_begin, 0);                     // (after the block-wide barrier above: the mbarriers are initialised and visible)

    for (int s0 = s_begin; s0 < s_end; s0 += M, cx.out_b += out_step) {
        const int mb = min(M, s_end - s0);
        // threads sharing a column work on different samples of the same batch: nobody starts fetching batch b + 1's
        // TRAM reads before everybody's writes of batch b - 1 are out
        if (TRAM && P > 1) __syncthreads();
        if (tma) {
            __syncthreads();                            // everybody is done reading the buffer the next copy overwrites
            fetch_tma(s0 + M, cx.boff ^ buf_bytes);
            mbar_wait(bar_s + (cx.boff ? 8u : 0u), (uint32_t)(((s0 - s_begin) / M) >> 1) & 1u);   // this buffer's k-th use completes phase k
        } else {
        fetch_batch(s0 + M, cx.boff ^ buf_bytes);
        fetch_tram(s0 + M, cx.boff ^ buf_bytes);
        cp_async_commit();
        cp_async_wait<1>();
        }
        // a column whose delay is too short for prefetching is run by its first thread alone, which then reads stage rows
        // the other threads of the column fetched: their copies must have landed and be visible
        if (TRAM && P > 1 && !cx.tram_fast) __syncthreads();
        bool owner = true;                              // this thread computes the batch's last sample (final state, if it is the call's last)
        if (!TRAM) {
            if (p.ccr_live) sl_exec<K, false, true, TRAM>(p, cx, 0, mb);
            else sl_exec<K, false, false, TRAM>(p, cx, 0, mb);
        } else {
            // prefetched TRAM reads: this thread's share of the batch in one go; a delay shorter than two batches: the
            // same code one sample at a time, by the column's first thread alone (one call site: the bulk code is
            // instantiated once)
            const int lo = cx.tram_fast ? my_lo(mb) : (part == 0 ? 0 : mb);
            const int hi = cx.tram_fast ? my_hi(mb) : (part == 0 ? mb : mb);
            const int step = cx.tram_fast ? max(1, hi - lo) : 1;
            owner = (lo < hi) && (hi == mb);
            for (int m0 = lo; m0 < hi; m0 += step) {
                if (cx.split) {                         // TRAM pointers at sample m0 (every pointer moves tr_ops per sample period)
#pragma unroll
                    for (int k = 0; k < K; ++k) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (p.tr_ops[j]) cx.tp[j][k] = ring_add(tp0[j][k], m0 * p.tr_ops[j], cx.rsize[j >> 1]);
                    }
                }
                if (p.ccr_live) sl_exec<K, false, true, TRAM>(p, cx, m0, m0 + step);
                else sl_exec<K, false, false, TRAM>(p, cx, m0, m0 + step);
            }
            if (cx.split) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (p.tr_ops[j]) tp0[j][k] = ring_add(tp0[j][k], mb * p.tr_ops[j], cx.rsize[j >> 1]);
                }
            }
        }
        if (s0 + mb == p.n_samples && blk == p.n_blk - 1 && valid && owner) {
            // This thread owns the call's last sample: leave the final state behind (cold path).
            if (p.pdl_late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");   // state writes follow
            sl_exec<K, true, true, TRAM>(p, cx, mb - 1, mb);
            for (int i = 0; i < p.n_wb; ++i) {
                const uint2 e = p.wb_list[i];
                const unsigned char* src = col + (e.x & SL_OFF_MASK) + ((e.x & SL_BUF) ? cx.boff : 0u) + (uint32_t)(mb - 1) * sl_stride(e.x);
                vstore<K>(p.gpr + (size_t)e.y * N + inst0, vload<K>(reinterpret_cast<const float*>(src)));
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (p.acc_writer) p.acc[inst0 + k] = (double)cx.acc_last[k];
                if (TRAM && p.has_tram) {
                    p.ptrs[inst0 + k] = cx.split ? tp0[0][k] : cx.tp[0][k]; p.ptrs[N + inst0 + k] = cx.split ? tp0[1][k] : cx.tp[1][k];
                    p.ptrs[2 * N + inst0 + k] = cx.split ? tp0[2][k] : cx.tp[2][k]; p.ptrs[3 * N + inst0 + k] = cx.split ? tp0[3][k] : cx.tp[3][k];
                }
                p.counts[inst0 + k] += (unsigned long long)p.n_samples * (unsigned long long)p.n_instrs * (unsigned long long)p.n_blk;
            }
        }
        cx.boff ^= buf_bytes;
    }
    if (cx.flags) atomicOr(p.rt_flags, cx.flags);
}

}  // namespace fxk
