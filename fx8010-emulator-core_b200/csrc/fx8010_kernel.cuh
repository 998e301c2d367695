// fx8010_kernel.cuh — the batched FX8010 interpreter kernel (sm_100a).
//
// One thread runs K emulated DSP instances ("contexts", K = 1, 2 or 4 ADJACENT instances) over one
// time segment.  The decoded program sits in __constant__ memory; the program counter is
// warp-uniform, so fetch, decode and the opcode switch run once per thread for K contexts and
// every lane executes the same DSP instruction on its own instances.  Data-dependent SKIP never
// diverges: a per-context skip counter turns into a write predicate.  The GPR file is a
// structure-of-arrays tile in shared memory (gpr[reg][thread][K], one 4/8/16-byte access per
// operand, bank-conflict free); accumulator, LFSR, TRAM pointers and counters live in hardware
// registers.  Input samples stream HBM -> shared through two cp.async buffers of `chunk` sample periods each (while
// one is consumed the next is in flight; one 16-byte LDGSTS per thread and sample at K = 4); outputs go back with
// coalesced 16-byte stores straight from registers.
//
// Semantics follow the reference's FX8010::process (reference source/FX8010.cpp:1023-1249) op by
// op; every float/double operation uses an explicitly rounded intrinsic so nothing can be
// contracted into an FMA.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "fx8010_gpu.h"

namespace fxk {

// ---- micro-ops: the reference opcodes with IDELAY/XDELAY resolved by the type of R -----------
enum Uop : uint32_t {
    U_MACS = 0, U_MACSN, U_MACW, U_MACWN, U_MACINTW, U_ACC3, U_MACMV, U_ANDXOR, U_TSTNEG, U_LIMIT,
    U_LIMITN, U_LOG, U_EXP, U_INTERP, U_SKIP, U_IREAD, U_IWRITE, U_XREAD, U_XWRITE, U_NOP, U_END
};

// flag bits (word0 bits 8..23)
constexpr uint32_t F_PRE_A = 1u << 8;    // A <- in[in_ch]   (source/FX8010.cpp:1055)
constexpr uint32_t F_PRE_X = 1u << 9;    // X <- in[in_ch]   (:1057, uses A's IOIndex)
constexpr uint32_t F_PRE_Y = 1u << 10;   // Y <- in[in_ch]   (:1059)
constexpr uint32_t F_NOISE = 1u << 11;   // noise register <- whitenoise()  (:1063-1071)
constexpr uint32_t F_OUT = 1u << 12;     // latch[out_ch] <- R after the instruction (:1229-1233)
constexpr uint32_t F_CCR = 1u << 13;     // CCR value is observable: materialise it (:211-232)
constexpr uint32_t F_TAB_SMEM = 1u << 14; // LOG/EXP: literal selector, table staged in shared memory
constexpr uint32_t F_TAB_IMM = 1u << 15;  // LOG/EXP: literal selector, table in global memory
constexpr uint32_t F_OUT_DIRECT = 1u << 16; // SKIP-free programs: the last writer of its channel in program order
                                            // stores R straight to the output block (no latch round trip)
constexpr uint32_t F_PRED = 1u << 17;    // (general interpreter) some SKIP can reach this instruction: it runs under the per-context skip predicate
                                         // (load-time reach analysis; instructions no SKIP can reach run unpredicated and uncounted — their
                                         // number per pass is added to the executed-instruction counters at once)
constexpr uint32_t F_ACC = 1u << 18;     // (general interpreter) the accumulator value this instruction leaves can be observed (a MACMV reads it
                                         // before another instruction overwrites it); the batch-final value is always kept
constexpr uint32_t F_PRE_ANY = F_PRE_A | F_PRE_X | F_PRE_Y;

// 32-byte decoded instruction (two uint4 words).  All operand locations are ready-made BYTE offsets
// into the thread's shared-memory column (register r of thread t lives at column(t) + r * RS * 4 with
// RS = blockDim.x * K; the output latches and the input stages follow the registers in the same
// column), so an operand fetch is one add and one LDS:
//   A: { uop[0:8) | flags[8:24) | output channel[24:32),  R offset,  A offset,  X offset }
//   B: { Y offset,  noise-register offset[0:24) | table slot or id[24:32),
//        input-stage offset of the preload channel,  latch offset of the output channel }
// The encoding depends on the launch geometry (RS) and is redone by the host when that changes.
constexpr int MAX_INSTR = FX8010_MAX_INSTRUCTIONS;
constexpr int SLOT_WORDS = 2 * (MAX_INSTR + 1);           // a maximum-length program (plus its pad instruction)
constexpr int ARENA_WORDS = 2 * SLOT_WORDS;              // 64 064 B of the 64 KiB constant bank: two maximum-length programs or dozens of short ones
// Decoded programs of all live handles of a device share this arena (one per kernel translation unit, see
// fx8010_families.h); the host allocates ranges first-fit and evicts the least recently launched program when it
// is full (fx8010_gpu.cu::arena_acquire), so the number of live handles is not bounded by constant memory.
static __constant__ uint4 c_prog[ARENA_WORDS];

constexpr int MAX_CHUNK = 64;             // input stage: two buffers of `chunk` samples per channel; while one is consumed the
                                         // other is in flight (a recurrence needs ~30 sample rows in flight per thread to
                                         // cover HBM latency at full bandwidth)
constexpr int MAX_SMEM_TABLES = 2;       // LOG/EXP tables replicated into shared memory
#ifndef FXK_TAB_REPL
#define FXK_TAB_REPL 8
#endif
constexpr int TAB_REPL = FXK_TAB_REPL;   // replicas per table entry; lane l reads replica l % TAB_REPL.  With 8 the eight lanes of a quarter warp (one
                                         // phase of a 128-bit shared load) hit eight different 16-byte bank groups whatever their indices: the gather is
                                         // conflict free (4 wavefronts instead of ~7 with two replicas; the shared-memory pipe is what bounds cfg2)
constexpr int TAB_SMEM_BYTES = FX8010_TABLE_ENTRIES * TAB_REPL * 16;   // 8 KiB per table

struct __align__(16) TableEntry { double y1, slope; }; // T[i], (T[i+1]-T[i])/(x2-x1) — host-computed in IEEE double

struct Params {
    // per-instance state, all [..][N]
    float* gpr;                 // [n_regs][N]
    double* acc;                // [N]
    uint32_t* lfsr;             // [2][N]
    float* latch;               // [C][N]
    int32_t* ptrs;              // [4][N]  iw, ir, xw, xr
    float* itram;               // [itram_size][N]
    float* xtram;               // [xtram_size][N]
    unsigned long long* counts; // [N]
    unsigned int* rt_flags;     // 1 word
    const uint32_t* reg_map;    // [n_regs] shared-memory row -> register index in the state arrays (only the
                                // registers the program refers to get a row)
    const uint32_t* load_rows;  // rows whose initial value the program can see (a stateless program never reads the others)
    const uint32_t* wb_regs;    // rows the program may write (write-back list)
    const uint32_t* latch_ch;   // output channels still served from the latch at the end of a sample period
    const TableEntry* tabs;     // [2][32][64]  LOG then EXP
    // I/O
    const float* in;            // element (c, s, i) at in[c * in_cstride + s * N + i]
    float* out;                 // element (c, s, i) at out[c * out_cstride + s * N + i]
    size_t in_cstride, out_cstride;
    int n_samples;              // S of this call
    int seg_len;                // samples per time segment (== S when serial)
    int n_seg;
    // geometry
    int N, C, n_regs, n_instrs, n_wb, prog_off;   // n_regs = shared-memory rows; prog_off = first word of the program in c_prog
    int n_exec;                 // encoded instructions (END/NOP are dropped for SKIP-free programs)
    int n_latch_ch;             // entries of latch_ch
    int n_unpred;               // encoded instructions without F_PRED that count as executed (END/NOP included): added to the counters per first pass
    int n_load;                 // entries of load_rows
    int load_latch, load_acc;   // the latches / the accumulator can be observed before the program rewrites them
    int chunk;                  // samples per input-stage buffer (power of two <= MAX_CHUNK)
    fx8010_trace_entry* trace;  // debug: [n_samples][n_exec] records for instance trace_inst (only the <1,true,true,0> variant looks)
    int trace_inst;
    int pdl_late_wait;          // programmatic dependent launch: 1 = this launch reads nothing the previous launch on
                                // the stream writes until its own state write-back (stateless program, disjoint
                                // buffers), so it only waits for that launch right before writing state
    int itram_size, xtram_size;
    int n_smem_tabs;            // tables staged in shared memory
    int smem_tab_id[MAX_SMEM_TABLES];   // op*32 + selector
};

// ---- K-wide context vectors -------------------------------------------------------------------
template <int K> struct VT;
template <> struct VT<1> { using T = float; };
template <> struct VT<2> { using T = float2; };
template <> struct VT<4> { using T = float4; };

template <int K> struct Vec {
    float v[K];
    __device__ __forceinline__ float& operator[](int i) { return v[i]; }
    __device__ __forceinline__ const float& operator[](int i) const { return v[i]; }
};
template <int K> __device__ __forceinline__ Vec<K> vload(const float* p) {
    Vec<K> r;
    const typename VT<K>::T t = *reinterpret_cast<const typename VT<K>::T*>(p);
    __builtin_memcpy(&r, &t, sizeof(t));
    return r;
}
template <int K> __device__ __forceinline__ void vstore(float* p, const Vec<K>& r) {
    typename VT<K>::T t;
    __builtin_memcpy(&t, &r, sizeof(t));
    *reinterpret_cast<typename VT<K>::T*>(p) = t;
}

// Explicit shared-space accessors on 32-bit shared addresses: no generic->shared window arithmetic in the
// inner loops (ptxas otherwise rebuilds the window base from SR_CgaCtaId next to many accesses).
template <int K> __device__ __forceinline__ Vec<K> lds(uint32_t a) {
    Vec<K> r;
    if (K == 4) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[K > 1 ? 1 : 0]), "=f"(r.v[K > 2 ? 2 : 0]), "=f"(r.v[K > 3 ? 3 : 0]) : "r"(a));
    else if (K == 2) asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(r.v[0]), "=f"(r.v[K > 1 ? 1 : 0]) : "r"(a));
    else asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r.v[0]) : "r"(a));
    return r;
}
template <int K> __device__ __forceinline__ void sts(uint32_t a, const Vec<K>& r) {
    if (K == 4) asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(r.v[0]), "f"(r.v[K > 1 ? 1 : 0]), "f"(r.v[K > 2 ? 2 : 0]), "f"(r.v[K > 3 ? 3 : 0]) : "memory");
    else if (K == 2) asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(r.v[0]), "f"(r.v[K > 1 ? 1 : 0]) : "memory");
    else asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(r.v[0]) : "memory");
}
__device__ __forceinline__ void lds_f64x2(uint32_t a, double& x, double& y) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}

// ---- exact scalar semantics ----------------------------------------------------------------

// static_cast<int32_t>(float) on x86-64 (cvttss2si): NaN / out of range -> 0x80000000
__device__ __forceinline__ int32_t cvt_x86(float f) {
    const int32_t i = __float2int_rz(f);        // saturating, NaN -> 0
    return (f < 2147483648.0f) ? i : INT32_MIN; // fixes +overflow and NaN; -overflow saturates to INT_MIN already
}
// saturate(): source/FX8010.cpp:275-279 (NaN passes through)
__device__ __forceinline__ float sat1(float v) {
    const float t = (v <= -1.0f) ? -1.0f : v;      // same result as the reference's nested ternary for every input
    return (t >= 1.0f) ? 1.0f : t;                 // (v >= 1 -> t == v -> 1; NaN -> NaN), two FMNMX.NAN instead of three ops
}
// setCCR(): source/FX8010.cpp:211-232
__device__ __forceinline__ float ccr_of(float r) {
    const float ar = fabsf(r);
    const float c = (ar < 1.0f) ? ((r < 0.0f) ? 6.0f : 2.0f) : ((ar == 1.0f) ? ((r < 0.0f) ? 20.0f : 16.0f) : 0.0f);
    return (r == 0.0f) ? 8.0f : c;
}
// wrapAround() value path: source/FX8010.cpp:299-328
__device__ __forceinline__ float wrap1(float v) {
    return (v >= 1.0f) ? __fadd_rn(v, -2.0f) : ((v < -1.0f) ? __fadd_rn(v, 2.0f) : v);
}
// logicOps(): source/FX8010.cpp:330-360 (first matching rule wins; written last-to-first)
__device__ __forceinline__ int32_t logic_ops(float fa, float fx, float fy) {
    const int32_t A = cvt_x86(fa), X = cvt_x86(fx), Y = cvt_x86(fy);
    int32_t r = (A & X) ^ Y;
    if (Y == 0xFFFFFF) r = ~A & X;
    if (Y == ~X) r = A | Y;
    if (X == 0xFFFFFFF && Y == 0xFFFFFF) r = ~A;
    if (X == 0xFFFFFF) r = A ^ Y;
    if (Y == 0) r = A & X;
    return r;
}

// INTERP, source/FX8010.cpp:1180-1187: (float)((1.0 - (double)X) * (double)A + (double)(X * Y)) for K contexts; omx = 1.0 - (double)X.
// (Widening A and X * Y with integer-pipe bit arithmetic instead of F2F.F64.F32 — 16 lanes per clock and SM on the conversion
//  unit, profiles/pipe_peaks.json — was measured and is SLOWER: cfg4 139 -> 170 us at 65 536 instances, 86 -> 98 us at 8 192;
//  the five dependent integer instructions are as long a chain as the conversion's 18 cycles and cost issue slots.)
template <int K> __device__ __forceinline__ void interp_core(const double (&omx)[K], const float (&a)[K], const float (&x)[K], const float (&y)[K], float (&out)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = __double2float_rn(__dadd_rn(__dmul_rn(omx[k], (double)a[k]), (double)__fmul_rn(x[k], y[k])));
}
template <int K> __device__ __forceinline__ void one_minus(const float (&x)[K], double (&omx)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) omx[k] = __dsub_rn(1.0, (double)x[k]);
}

// linearInterpolate() index: static_cast<int>((x - x_min) / step), source/FX8010.cpp:285-286.
// (x + 1.0) * 31.5 in double gives the same integer as (x + 1.0) / (2.0/63) for every binary32 x in
// [-1, 1] (exhaustively checked in C, sampled in tests/test_oracle.py::test_table_index_formula, swept
// on the GPU by test_table_sweep_all_selectors); there the result is already inside 0..63.
// The conversion unit (F2I.F64 / I2F.F64, about 11 issue cycles per warp instruction on this chip) is kept out of it:
// for 0 <= q < 2^31, q + 2^52 rounded toward -infinity is exactly 2^52 + floor(q), whose low mantissa word is the
// index and from which (double)index = t - 2^52 follows exactly — two FP64-pipe adds instead of two conversions.
__device__ __forceinline__ int table_index_inrange(double xd, double& di) {
    const double q = __dmul_rn(__dadd_rn(xd, 1.0), 31.5);
    const double t = __dadd_rd(q, 4503599627370496.0);
    di = __dsub_rn(t, 4503599627370496.0);
    return __double2loint(t);
}
// Outside [-1, 1] (rule U6) the reference's index is clamped; cvttsd2si overflow / NaN -> INT_MIN -> 0.
__device__ __forceinline__ int table_index_wild(float a) {   // (inlined: a real call makes ptxas spill the callers' long-lived state around it)
    const double q = __dmul_rn(__dadd_rn((double)a, 1.0), 31.5);
    int i = __double2int_rz(q);
    i = min(max(i, 0), FX8010_TABLE_ENTRIES - 1);
    return (q < 2147483648.0) ? i : 0;
}
// y = (y2 - y1) / (x2 - x1) * (x - x1) + y1 with the quotient precomputed on the host (:287-293)
__device__ __forceinline__ float table_finish(double xd, double di, double y1, double slope) {   // di = (double)index
    const double step = 2.0 / 63.0;
    const double x1 = __dadd_rn(-1.0, __dmul_rn(di, step));
    return __double2float_rn(__dadd_rn(__dmul_rn(slope, __dsub_rn(xd, x1)), y1));
}

template <int BYTES> __device__ __forceinline__ void cp_async(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
    else if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// An opaque move: the front end can neither rematerialise the value nor reason about its range.
__device__ __forceinline__ uint32_t pin32(uint32_t v) { uint32_t o; asm volatile("mov.b32 %0, %1;" : "=r"(o) : "r"(v)); return o; }
__device__ __forceinline__ uint64_t pin64(uint64_t v) { uint64_t o; asm volatile("mov.b64 %0, %1;" : "=l"(o) : "l"(v)); return o; }

// Loop invariants that come from the kernel parameter bank or from __constant__ memory: ptxas likes to
// re-read them inside the sample loop ("rematerialisation"), and a constant-bank load in front of a
// dependent instruction costs a lone warp tens of cycles of its per-sample chain (measured: half of a
// one-instruction recurrence's time).  They are therefore bounced through shared memory once and read
// back with VOLATILE loads before the loop: a volatile access cannot be repeated, so the value has to
// stay in a register.
constexpr int INV_WORDS = 8;             // the scalars below
__device__ __forceinline__ uint32_t inv_read(const uint32_t* s_inv, int i) { return *reinterpret_cast<const volatile uint32_t*>(s_inv + i); }

// Shared-memory layout of one block (host and device agree through these formulas):
//   [ tables ][ registers n_regs ][ latches C ][ input stage C x 2 x chunk ]   (the last three per column)
__host__ __device__ inline size_t smem_bytes(int n_regs, int C, int B, int K, int n_smem_tabs, int chunk) {
    return (size_t)n_smem_tabs * TAB_SMEM_BYTES +
           (size_t)B * K * sizeof(float) * ((size_t)n_regs + C + 2 * (size_t)C * chunk);
}
__host__ __device__ inline uint32_t reg_offset(int r, int RS) { return (uint32_t)r * RS * 4u; }
__host__ __device__ inline uint32_t latch_offset(int n_regs, int c, int RS) { return (uint32_t)(n_regs + c) * RS * 4u; }
__host__ __device__ inline uint32_t stage_offset(int n_regs, int C, int c, int RS, int chunk) { return (uint32_t)(n_regs + C + c * 2 * chunk) * RS * 4u; }

// ---- the kernel ------------------------------------------------------------------------------
//
// grid = (ceil(N / (K * B)), n_seg), block = B threads.
//   SKIP: the program contains SKIP (per-context predicate, extra passes when END is skipped)
//   EXT : the program uses TRAM, the noise LFSR or MACMV (their state stays out of the registers
//         of simpler programs)
template <int K, bool SKIP, bool EXT>
__global__ void __launch_bounds__(128, 3) fx_interp_kernel(const Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int B = blockDim.x;
    const int tid = threadIdx.x;
    const int N = p.N, C = p.C;
    const int tslot_raw = blockIdx.x * B + tid;
    const bool valid = tslot_raw * K < N;
    const int inst0 = valid ? tslot_raw * K : N - K;       // N % K == 0 (host guarantees it)
    const int seg = blockIdx.y;
    const bool last_seg = (seg == p.n_seg - 1);
    const int s_begin = seg * p.seg_len;
    const int s_end = min(p.n_samples, s_begin + p.seg_len);
    const uint4* const prog = c_prog + p.prog_off;
    const int RS = B * K;                                  // register stride (floats)

    // Programmatic dependent launch: let the next launch on the stream start filling SMs as this one
    // drains; unless told otherwise, wait here for the previous launch (it may still be writing the
    // state arrays or the buffers this one reads).
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (!p.pdl_late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");

    // shared-memory carve-up
    TableEntry* const s_tab = reinterpret_cast<TableEntry*>(smem_raw);                 // [n_smem_tabs][64][TAB_REPL]
    unsigned char* const col = smem_raw + (size_t)p.n_smem_tabs * TAB_SMEM_BYTES + (size_t)tid * K * 4;   // this thread's column
    auto at = [&](uint32_t byte_off) { return reinterpret_cast<float*>(col + byte_off); };
    const uint32_t latch0 = latch_offset(p.n_regs, 0, RS);
    const uint32_t stage0 = stage_offset(p.n_regs, C, 0, RS, p.chunk);
    const uint32_t row_bytes = (uint32_t)RS * 4u;

    // input stage: chunk starting at sample s0 -> buffer `buf` of every channel; one cp.async group per chunk
    const bool has_in = (p.in != nullptr);
    const int chunk = p.chunk;
    const uint32_t buf_bytes = (uint32_t)chunk * row_bytes;
    auto fetch_chunk = [&](int s0, uint32_t boff) {
        if (has_in && s0 < s_end) {
            const int n = min(chunk, s_end - s0);
            const float* g = p.in + (size_t)s0 * N + inst0;
            for (int c = 0; c < C; ++c, g += p.in_cstride) {
                unsigned char* d = reinterpret_cast<unsigned char*>(at(stage0 + (uint32_t)c * 2u * buf_bytes + boff));
                const float* gs = g;
#pragma unroll 4
                for (int m = 0; m < n; ++m, d += row_bytes, gs += N) cp_async<4 * K>(d, gs);
            }
        }
        cp_async_commit();
    };
    fetch_chunk(s_begin, 0);

    // literal-selector LOG/EXP tables -> shared; entry e of replica q lives at slot e * TAB_REPL + q and lane l
    // reads replica l % TAB_REPL, which spreads the 128-bit gathers over more bank groups.
    for (int t = 0; t < p.n_smem_tabs; ++t) {
        const TableEntry* src = p.tabs + (size_t)p.smem_tab_id[t] * FX8010_TABLE_ENTRIES;
#pragma unroll 4
        for (int i = tid; i < FX8010_TABLE_ENTRIES * TAB_REPL; i += B)      // consecutive lanes -> consecutive slots (conflict free)
            s_tab[t * FX8010_TABLE_ENTRIES * TAB_REPL + i] = src[i / TAB_REPL];
    }

    {   // register rows: batches of four independent loads keep the startup off the L2 latency chain
        int j = 0;
        for (; j + 4 <= p.n_load; j += 4) {
            Vec<K> t[4];
            uint32_t row[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { row[q] = p.load_rows[j + q]; t[q] = vload<K>(p.gpr + (size_t)p.reg_map[row[q]] * N + inst0); }
#pragma unroll
            for (int q = 0; q < 4; ++q) vstore<K>(at(reg_offset(row[q], RS)), t[q]);
        }
        for (; j < p.n_load; ++j) { const uint32_t row = p.load_rows[j]; vstore<K>(at(reg_offset(row, RS)), vload<K>(p.gpr + (size_t)p.reg_map[row] * N + inst0)); }
    }
    if (p.load_latch)
        for (int c = 0; c < C; ++c) vstore<K>(at(latch0 + (uint32_t)c * RS * 4u), vload<K>(p.latch + (size_t)c * N + inst0));

    float acc_f[K];
    double acc_d[K];
    bool acc_is_f[K];                 // the accumulator currently holds the float acc_f (else acc_d)
    int skip[K];
    unsigned int count[K];
    uint32_t g1[K], g2[K];
    int32_t iw[K], ir[K], xw[K], xr[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        acc_d[k] = p.load_acc ? p.acc[inst0 + k] : 0.0; acc_f[k] = 0.0f; acc_is_f[k] = false;
        skip[k] = 0; count[k] = 0;
        if (EXT) {
            g1[k] = p.lfsr[inst0 + k]; g2[k] = p.lfsr[N + inst0 + k];
            iw[k] = p.ptrs[inst0 + k]; ir[k] = p.ptrs[N + inst0 + k];
            xw[k] = p.ptrs[2 * N + inst0 + k]; xr[k] = p.ptrs[3 * N + inst0 + k];
        }
    }
    unsigned int flags = 0;
    __shared__ uint32_t s_inv[INV_WORDS];
    if (tid == 0) {
        s_inv[0] = (uint32_t)p.n_exec; s_inv[1] = (uint32_t)p.n_latch_ch; s_inv[2] = (uint32_t)(p.n_samples - 1);
        s_inv[3] = (uint32_t)(p.out_cstride & 0xffffffffu); s_inv[4] = (uint32_t)(p.out_cstride >> 32);
        s_inv[5] = (uint32_t)p.N; s_inv[6] = (uint32_t)p.n_unpred;
    }
    __syncthreads();                  // s_tab / s_inv visible (the only block-wide dependency)
    const int n_exec = (int)inv_read(s_inv, 0), n_latch_ch = (int)inv_read(s_inv, 1), last_s = (int)inv_read(s_inv, 2);
    const size_t out_cstride = (size_t)inv_read(s_inv, 3) | ((size_t)inv_read(s_inv, 4) << 32);
    const int Nv = (int)inv_read(s_inv, 5);        // N for use inside the sample loop
    const unsigned int n_unpred = inv_read(s_inv, 6);

    const bool tracing = (K == 1 && SKIP && EXT) && p.trace != nullptr && valid && inst0 == p.trace_inst;
    const int lane_rep = tid & (TAB_REPL - 1);
    float* out_s = p.out + (size_t)s_begin * N + inst0;   // this thread's slot in the current output row
    uint32_t boff = 0;                                     // byte offset of the stage buffer being consumed
    for (int c0 = s_begin; c0 < s_end; c0 += chunk, boff ^= buf_bytes) {
      fetch_chunk(c0 + chunk, boff ^ buf_bytes);
      cp_async_wait<1>();                                          // the chunk starting at c0 has landed
      const int c_end = min(s_end, c0 + chunk);
      uint32_t stage_s = boff;                                     // byte offset of the current sample's stage row
      for (int sidx = c0; sidx < c_end; ++sidx, out_s += Nv, stage_s += row_bytes) {
        {
            // ---- one sample period: FX8010::process, source/FX8010.cpp:1023-1249 ----
            // The final CCR / latch must be in shared memory when the batch ends (state write-back).
            const bool last_sample = (sidx == last_s);
            const uint32_t force_flags = last_sample ? (F_CCR | F_ACC) : 0u;
            bool saw_end[K];
#pragma unroll
            for (int k = 0; k < K; ++k) { skip[k] = 0; saw_end[k] = false; }
            int pass = 0;
            do {
                int trace_pc = 0;
                // PRED = this instruction runs under the skip predicate (a SKIP can reach it, or this is an extra pass in which
                // finished contexts idle); instructions no SKIP can reach take the unpredicated copy: vector stores, no
                // per-context bookkeeping.
                auto exec_instr = [&](auto pred_tag, const uint4 wA, const uint4 wB) {
                    constexpr bool PRED = decltype(pred_tag)::value;
                    const uint32_t w0 = wA.x | force_flags;           // the call's last sample materialises CCR and accumulator whatever the liveness says
                    const uint32_t uop = w0 & 0xffu;                  // (a compare tree: measured faster than the LDC + BRX jump table, 141 vs 155 ms on cfg5)
                    bool act[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        if (PRED) {                                   // :1037 / :1235-1241
                            act[k] = (skip[k] == 0);
                            skip[k] = (skip[k] > 0) ? skip[k] - 1 : 0; // a negative count skips exactly one
                        } else act[k] = true;
                    }
                    float* const pr = at(wA.y);
                    float* const pa = at(wA.z);
                    float* const px = at(wA.w);
                    float* const py = at(wB.x);
                    if (w0 & (F_PRE_ANY | F_NOISE)) {                 // the less common pre-work behind one test
                    if (w0 & F_PRE_ANY) {                             // :1053-1061
                        Vec<K> v;
                        if (has_in) v = vload<K>(at(wB.z + stage_s));
                        else {
#pragma unroll
                            for (int k = 0; k < K; ++k) v[k] = 0.0f;
                        }
                        if (!PRED) {
                            if (w0 & F_PRE_A) vstore<K>(pa, v);
                            if (w0 & F_PRE_X) vstore<K>(px, v);
                            if (w0 & F_PRE_Y) vstore<K>(py, v);
                        } else {
#pragma unroll
                            for (int k = 0; k < K; ++k)
                                if (act[k]) {
                                    if (w0 & F_PRE_A) pa[k] = v[k];
                                    if (w0 & F_PRE_X) px[k] = v[k];
                                    if (w0 & F_PRE_Y) py[k] = v[k];
                                }
                        }
                    }
                    if (EXT && (w0 & F_NOISE)) {                      // :1063-1071, whitenoise :993-1000
                        float* const pn = at(wB.y & 0xffffffu);
#pragma unroll
                        for (int k = 0; k < K; ++k)
                            if (act[k]) {
                                g1[k] ^= g2[k];
                                pn[k] = __fmul_rn(__int2float_rn((int32_t)g2[k]), 4.656612873077392578125e-10f);
                                g2[k] += g1[k];
                            }
                    }
                    }
                    // All three operands are fetched up front (every operand offset names a valid row): the loads are
                    // in flight while the opcode is dispatched.
                    const Vec<K> a = vload<K>(pa), x = vload<K>(px), y = vload<K>(py);
                    Vec<K> r;
                    bool writes_r = true;
#define FX_EACH _Pragma("unroll") for (int k = 0; k < K; ++k)
                    const bool want_acc = (w0 & F_ACC) != 0u;         // nobody can observe the accumulator otherwise (it is overwritten first)
#define FX_ACC(val) { if (want_acc && act[k]) { acc_f[k] = (val); acc_is_f[k] = true; } }
                    switch (uop) {
                    case U_MACS: {   // :1077-1085 (MACINTS :1095-1103 is identical)
                        FX_EACH { const float t = __fadd_rn(a[k], __fmul_rn(x[k], y[k])); FX_ACC(t); r[k] = sat1(t); } break; }
                    case U_MACSN: {  // :1086-1094
                        FX_EACH { const float t = __fsub_rn(a[k], __fmul_rn(x[k], y[k])); FX_ACC(t); r[k] = sat1(t); } break; }
                    case U_ACC3: {   // :1104-1112
                        FX_EACH { const float t = __fadd_rn(__fadd_rn(a[k], x[k]), y[k]); FX_ACC(t); r[k] = sat1(t); } break; }
                    case U_MACW: {   // :1126-1131
                        FX_EACH { r[k] = __fadd_rn(a[k], wrap1(__fmul_rn(x[k], y[k]))); FX_ACC(r[k]); } break; }
                    case U_MACWN: {  // :1132-1137
                        FX_EACH { r[k] = __fsub_rn(a[k], wrap1(__fmul_rn(x[k], y[k]))); FX_ACC(r[k]); } break; }
                    case U_MACINTW: { // :1138-1143
                        FX_EACH { r[k] = wrap1(__fadd_rn(a[k], __fmul_rn(x[k], y[k]))); FX_ACC(r[k]); } break; }
                    case U_MACMV: {  // :1144-1149
                        FX_EACH {
                            if (EXT && act[k]) {
                                const double base = acc_is_f[k] ? (double)acc_f[k] : acc_d[k];
                                acc_d[k] = __dadd_rn(base, (double)__fmul_rn(x[k], y[k]));
                                acc_is_f[k] = false;
                            }
                            r[k] = a[k];
                        } break; }
                    case U_ANDXOR: { // :1150-1154 (accumulator untouched)
                        FX_EACH { r[k] = __int2float_rn(logic_ops(a[k], x[k], y[k])); } break; }
                    case U_TSTNEG: { // :1155-1162
                        FX_EACH {
                            const int32_t q = cvt_x86(__fmul_rn(x[k], 2147483648.0f));
                            r[k] = (a[k] >= y[k]) ? x[k] : __fmul_rn(__int2float_rn(~q), 4.656612873077392578125e-10f);
                            FX_ACC(r[k]);
                        } break; }
                    case U_LIMIT: {  // :1163-1168
                        FX_EACH { r[k] = (a[k] >= y[k]) ? x[k] : y[k]; FX_ACC(r[k]); } break; }
                    case U_LIMITN: { // :1169-1174
                        FX_EACH { r[k] = (a[k] < y[k]) ? x[k] : y[k]; FX_ACC(r[k]); } break; }
                    case U_LOG:                                       // :1113-1119
                    case U_EXP: {                          // :1120-1125, linearInterpolate :283-296
                        int idx[K];
                        double di[K];
                        bool wild = false;
                        FX_EACH { wild |= !(fabsf(a[k]) <= 1.0f); }
                        if (!wild) { FX_EACH { idx[k] = table_index_inrange((double)a[k], di[k]); } }
                        else {                                        // rule U6: clamp and flag (rare, may diverge)
                            FX_EACH {
                                idx[k] = table_index_wild(a[k]); di[k] = (double)idx[k];
                                if (!(fabsf(a[k]) <= 1.0f) && act[k]) flags |= FX8010_RT_TABLE_RANGE;
                            }
                        }
                        if (w0 & F_TAB_SMEM) {
                            const TableEntry* const tb = s_tab + (size_t)(wB.y >> 24) * (FX8010_TABLE_ENTRIES * TAB_REPL) + lane_rep;
                            FX_EACH {
                                const TableEntry e = tb[idx[k] * TAB_REPL];
                                r[k] = table_finish((double)a[k], di[k], e.y1, e.slope); FX_ACC(r[k]);
                            }
                        } else {
                            FX_EACH {
                                int tsel;
                                if (w0 & F_TAB_IMM) tsel = (int)(wB.y >> 24);
                                else {
                                    int32_t sel = cvt_x86(x[k]);
                                    if (sel < 0 || sel > FX8010_TABLE_COUNT - 1) {
                                        if (act[k]) flags |= FX8010_RT_TABLE_RANGE;
                                        sel = sel < 0 ? 0 : FX8010_TABLE_COUNT - 1;
                                    }
                                    tsel = (uop == U_EXP ? FX8010_TABLE_COUNT : 0) + sel;
                                }
                                const double2 e = __ldg(reinterpret_cast<const double2*>(p.tabs + tsel * FX8010_TABLE_ENTRIES + idx[k]));
                                r[k] = table_finish((double)a[k], di[k], e.x, e.y); FX_ACC(r[k]);
                            }
                        }
                        break; }
                    case U_INTERP: { // :1180-1187
                        double omx[K];
                        float t[K];
                        one_minus<K>(x.v, omx);
                        interp_core<K>(omx, a.v, x.v, y.v, t);
                        FX_EACH { FX_ACC(t[k]); r[k] = sat1(t[k]); } break; }
                    case U_SKIP:                                      // :1175-1179
                        if (SKIP) { const Vec<K> c = vload<K>(at(0));
                            FX_EACH { if (act[k] && __int2float_rn(cvt_x86(x[k])) == c[k]) skip[k] = cvt_x86(y[k]); } }
                        writes_r = false; break;
                    case U_IREAD: case U_XREAD:                       // :1190-1193 / :1202-1205, readSmallDelay :934-956
                        if (EXT) { 
                            const bool isx = (uop == U_XREAD);
                            const float* const ring = isx ? p.xtram : p.itram;
                            const int size = isx ? p.xtram_size : p.itram_size;
                            int ridx[K];
                            bool same = true;                         // all K contexts active and at the same ring position
                            FX_EACH {
                                int32_t& rp = isx ? xr[k] : ir[k];
                                const int pos = min(max(cvt_x86(y[k]), 0), size - 1);
                                int idx = rp - pos;
                                idx += (idx < 0) ? size : 0;          // rule U1: mathematical modulo
                                ridx[k] = idx;
                                if (act[k]) rp = (rp + 1 == size) ? 0 : rp + 1;
                                same = same && act[k] && (idx == ridx[0]);
                            }
                            if (K > 1 && same) vstore<K>(pa, vload<K>(ring + (size_t)ridx[0] * N + inst0));   // one coalesced 8/16-byte access
                            else { FX_EACH { if (act[k]) pa[k] = ring[(size_t)ridx[k] * N + inst0 + k]; } }
                        }
                        writes_r = false; break;
                    case U_IWRITE: case U_XWRITE:                     // :1195-1198 / :1207-1210, writeSmallDelay :909-917
                        if (EXT) { 
                            const bool isx = (uop == U_XWRITE);
                            float* const ring = isx ? p.xtram : p.itram;
                            const int size = isx ? p.xtram_size : p.itram_size;
                            int widx[K];
                            bool same = true;
                            FX_EACH {
                                int32_t& wp = isx ? xw[k] : iw[k];
                                const int pos = min(max(cvt_x86(y[k]), 0), size - 1);
                                widx[k] = wp + pos;                   // the reference does not wrap wp + pos: slots at or beyond
                                if (act[k]) wp = (wp + 1 == size) ? 0 : wp + 1;   // the ring are never read back -> dropped
                                same = same && act[k] && (widx[k] == widx[0]);
                            }
                            if (K > 1 && same) { if (widx[0] < size && valid) vstore<K>(ring + (size_t)widx[0] * N + inst0, a); }
                            else { FX_EACH { if (act[k] && widx[k] < size && valid) ring[(size_t)widx[k] * N + inst0 + k] = a[k]; } }
                        }
                        writes_r = false; break;
                    case U_END:                                       // :1212-1215
                        FX_EACH { if (act[k]) saw_end[k] = true; }
                        writes_r = false; break;
                    default: writes_r = false; break;                 // U_NOP: IDELAY/XDELAY whose R is neither read nor write
                    }
#undef FX_ACC
                    if (writes_r) {
                        if (!PRED) vstore<K>(pr, r);
                        else { FX_EACH { if (act[k]) pr[k] = r[k]; } }
                        if (w0 & F_CCR) {                             // setCCR :211-232 (after the R store: R may be ccr)
                            if (!PRED) { Vec<K> c; FX_EACH { c[k] = ccr_of(r[k]); } vstore<K>(at(0), c); }
                            else { FX_EACH { if (act[k]) at(0)[k] = ccr_of(r[k]); } }
                        }
                    }
                    if (PRED) { FX_EACH { count[k] += act[k] ? 1u : 0u; } }   // :1222 (unpredicated instructions: p.n_unpred per first pass)
                    if (w0 & F_OUT) {                                 // :1229-1233 (after EVERY executed instruction)
                        if (!SKIP && (w0 & F_OUT_DIRECT)) {
                            // R was just written by this instruction and nothing later in the sample period
                            // touches the channel: the value is the period's output (:1248)
                            if (valid) vstore<K>(out_s + (size_t)(w0 >> 24) * out_cstride, r);
                            if (last_sample) vstore<K>(at(wB.w), r);
                        } else if (!PRED) vstore<K>(at(wB.w), vload<K>(pr));
                        else { float* const pl = at(wB.w); FX_EACH { if (act[k]) pl[k] = pr[k]; } }
                    }
                    if (K == 1 && SKIP && EXT) {          // debug trace (fx8010_gpu_trace), compiled into one variant only
                        if (tracing) {
                            fx8010_trace_entry e;
                            e.index = trace_pc; e.executed = act[0] ? 1 : 0;
                            e.r = pr[0]; e.a = pa[0]; e.x = px[0]; e.y = py[0]; e.ccr = at(0)[0];
                            e.opcode = (w0 & (F_TAB_SMEM | F_TAB_IMM)) ? (uop == U_LOG ? FX_LOG : FX_EXP) : (int32_t)(wB.y >> 24);
                            e.acc = acc_is_f[0] ? (double)acc_f[0] : acc_d[0];
                            p.trace[(size_t)sidx * p.n_exec + trace_pc] = e;
                        }
                        ++trace_pc;
                    }
                };
                {
                    uint4 nA = prog[0], nB = prog[1];
                    const bool all_pred = SKIP && pass > 0;           // an extra pass: contexts that saw END idle through it on their skip counters
                    for (int pc = 0; pc < n_exec; ++pc) {
                        const uint4 wA = nA, wB = nB;
                        nA = prog[2 * pc + 2]; nB = prog[2 * pc + 3]; // the slot is padded with one extra instruction
                        if (SKIP && ((wA.x & F_PRED) || all_pred)) exec_instr(std::true_type{}, wA, wB);
                        else exec_instr(std::false_type{}, wA, wB);
                    }
                    if (SKIP && pass == 0) { FX_EACH { count[k] += n_unpred; } }
                }
                ++pass;
                if (!SKIP) break;
                // END skipped by some context: run the program again (source/FX8010.cpp:1243); contexts
                // that saw END idle through the extra pass via an "infinite" skip count.
                bool again = false;
                FX_EACH {
                    const bool more = !saw_end[k] && pass < FX8010_MAX_PASSES;
                    if (!saw_end[k] && pass >= FX8010_MAX_PASSES) flags |= FX8010_RT_END_SKIPPED_CAP;   // rule U9
                    again |= more;
                    if (!more) skip[k] = 0x7fffffff;
                }
                if (!__any_sync(0xffffffffu, again)) break;
            } while (true);
            if (valid)                                                // :1248 — coalesced, lane = K adjacent instances
                for (int j = 0; j < n_latch_ch; ++j) {
                    const uint32_t c = p.latch_ch[j];
                    vstore<K>(out_s + (size_t)c * out_cstride, vload<K>(at(latch0 + c * RS * 4u)));
                }
        }
      }
    }
#undef FX_EACH

    // ---- write the state back (the last time segment carries the final state) ----
    if (p.pdl_late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (flags) atomicOr(p.rt_flags, flags);
    if (!valid || !last_seg) return;
    for (int i = 0; i < p.n_wb; ++i) { const uint32_t r = p.wb_regs[i]; vstore<K>(p.gpr + (size_t)p.reg_map[r] * N + inst0, vload<K>(at(reg_offset(r, RS)))); }
    for (int c = 0; c < C; ++c) vstore<K>(p.latch + (size_t)c * N + inst0, vload<K>(at(latch0 + (uint32_t)c * RS * 4u)));
#pragma unroll
    for (int k = 0; k < K; ++k) {
        p.acc[inst0 + k] = acc_is_f[k] ? (double)acc_f[k] : acc_d[k];
        if (EXT) {
            p.lfsr[inst0 + k] = g1[k]; p.lfsr[N + inst0 + k] = g2[k];
            p.ptrs[inst0 + k] = iw[k]; p.ptrs[N + inst0 + k] = ir[k];
            p.ptrs[2 * N + inst0 + k] = xw[k]; p.ptrs[3 * N + inst0 + k] = xr[k];
        }
        const unsigned long long c = SKIP ? (unsigned long long)count[k]
                                          : (unsigned long long)p.n_samples * (unsigned long long)p.n_instrs;
        p.counts[inst0 + k] += c;
    }
}

// Fills register `reg` of every instance with one value (broadcast setRegisterValue).
static __global__ void fx_fill_kernel(float* dst, float v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}

}  // namespace fxk
