// fx8010_stateless.cuh — the sample-parallel kernel for STATELESS programs (sm_100a).
//
// A program is stateless when no sample period reads anything an earlier period wrote (no SKIP, TRAM,
// noise or MACMV; every register it reads is either never written or written earlier in the same
// period — decided at load time, fx8010_gpu.cu::analyse).  Its sample periods are independent, so the
// interpreter may run them in any order.  This kernel runs them INSTRUCTION-MAJOR over mini-batches of
// M samples: each DSP instruction is fetched from __constant__ memory and decoded once, then executed
// for the M samples of the batch (and K adjacent instances each), which amortises the interpretive
// overhead M-fold and gives the warp M x K independent dependency chains.  Registers the program
// writes and reads back get one shared-memory row PER SAMPLE of the batch; registers it only reads
// keep a single row; registers it only writes (outputs, an unobserved CCR) are stored once, on the
// batch-final sample, for the state write-back.  INPUT registers are not copied at all: the cp.async
// input stage rows ARE their rows (every read of an INPUT register is preceded by its preload in the
// same instruction, reference source/FX8010.cpp:1053-1061).  The time axis is also cut into segments
// across blockIdx.y; only the thread that owns the last sample writes state back.
//
// Arithmetic is the same code as the generic kernel (fx8010_kernel.cuh): bit-exact with the reference.
#pragma once

#include "fx8010_kernel.cuh"

namespace fxk {

// operand word: where a register lives in the thread's column for sample m of the batch
constexpr uint32_t SL_OFF_MASK = 0xfffffu;   // [0:20)  byte offset
constexpr int SL_STRIDE_SHIFT = 20;          // [20:31) bytes between consecutive samples of the batch, in 16-byte units
constexpr uint32_t SL_BUF = 1u << 31;        // input-stage row: add the offset of the current stage buffer
constexpr uint32_t F_ST_LAST = 1u << 17;     // R is never read by the program: store it on the batch-final sample only
constexpr int SL_MAX_M = 8;
constexpr int SL_CARRY_SHIFT = 18;           // w0 bits 18..20: operand A / X / Y is this instruction's OWN result of the previous sample
                                             // (a self recurrence): row (m - 1) mod M on the first sample of a batch, then forwarded
                                             // in a hardware register — the recurrence never waits for shared memory

// instruction: A = { uop | flags | out channel << 24, R word, A word, X word },  B = { Y word, table slot/id << 24, CCR word, 0 }

struct SLParams {
    float* gpr;                 // [n_regs][N]
    double* acc;                // [N]
    float* latch;               // [C][N]
    unsigned long long* counts; // [N]
    unsigned int* rt_flags;
    const TableEntry* tabs;
    const uint2* load_list;     // (operand word, register index): read-only rows fetched at start
    const uint2* wb_list;       // (operand word, register index): rows written back by the owner of the last sample
    const float* in;
    float* out;
    size_t in_cstride, out_cstride;
    int n_samples, seg_len, n_seg;
    int N, C, n_instrs, n_exec, slot, n_load, n_wb;
    int M;                      // samples per batch (power of two <= SL_MAX_M)
    uint32_t stage0;            // byte offset of the input stage rows: [C][2 buffers][M]
    int n_smem_tabs;
    int smem_tab_id[MAX_SMEM_TABLES];
    int acc_writer;             // some instruction sets the accumulator (else it keeps its value)
    int pdl_late_wait;
};

// Per-thread context of the stateless kernel, handed to the batch executor.
template <int K> struct SLCtx {
    uint32_t col_s;                 // this thread's shared-memory column (32-bit shared address)
    uint32_t tab_s;                 // staged tables + this lane's replica (32-bit shared address)
    const uint4* prog;
    float* out_b;                   // output row of the batch's first sample, this thread's instances
    int N, inst0, n_exec;
    size_t out_cstride;
    bool valid;
    uint32_t boff;                  // byte offset of the current input-stage buffer
    unsigned int flags;
    Vec<K> acc_last;
};

// Runs the whole program for samples [m_lo, m_hi) of the current batch, instruction-major.
//   FINAL = false: the bulk path — stores only what a later instruction or the caller can see
//                  (live registers, a CCR somebody reads, the output block).
//   FINAL = true : re-run of the call's last sample by its owner, storing EVERYTHING the state
//                  write-back needs (all result registers, CCR, output latch, accumulator) and nothing
//                  to the output block.  Stateless programs are idempotent per sample, so the re-run
//                  reproduces the same values.
// All operand addresses are running 32-bit shared addresses (one add per operand and sample).
template <int K, bool FINAL>
__device__ __forceinline__ void sl_exec(const SLParams& p, SLCtx<K>& cx, const int m_lo, const int m_hi) {
    uint4 nA = cx.prog[0], nB = cx.prog[1];
    const int n_exec = cx.n_exec;
    const int n_m = m_hi - m_lo;
    for (int pc = 0; pc < n_exec; ++pc) {
        const uint4 wA = nA, wB = nB;
        nA = cx.prog[2 * pc + 2]; nB = cx.prog[2 * pc + 3];
        const uint32_t w0 = wA.x;
        const uint32_t uop = w0 & 0xffu;
        // decode once per batch: address of sample m_lo and per-sample stride of every operand, store modes
#define SL_STRIDE(w) ((((w) >> SL_STRIDE_SHIFT) & 0x7ffu) << 4)
#define SL_ADDR(w) (cx.col_s + ((w) & SL_OFF_MASK) + (((w) & SL_BUF) ? cx.boff : 0u) + (uint32_t)m_lo * SL_STRIDE(w))
        // a carried operand: the row of the previous sample (row M - 1 before sample 0), and the pointer stays there
#define SL_ADDR_C(w) (cx.col_s + ((w) & SL_OFF_MASK) + (uint32_t)((m_lo == 0 ? p.M : m_lo) - 1) * SL_STRIDE(w))
        const bool ca = (w0 >> SL_CARRY_SHIFT) & 1u, cxx = (w0 >> SL_CARRY_SHIFT) & 2u, cy = (w0 >> SL_CARRY_SHIFT) & 4u;
        const uint32_t sr = SL_STRIDE(wA.y), sa = ca ? 0u : SL_STRIDE(wA.z), sx = cxx ? 0u : SL_STRIDE(wA.w), sy = cy ? 0u : SL_STRIDE(wB.x), sccr = SL_STRIDE(wB.z);
        uint32_t qr = SL_ADDR(wA.y), qa = ca ? SL_ADDR_C(wA.z) : SL_ADDR(wA.z), qx = cxx ? SL_ADDR_C(wA.w) : SL_ADDR(wA.w),
                 qy = cy ? SL_ADDR_C(wB.x) : SL_ADDR(wB.x), qccr = SL_ADDR(wB.z);
        Vec<K> rp;                                // the previous sample's result (carried operands)
        _Pragma("unroll") for (int k = 0; k < K; ++k) rp[k] = 0.0f;
        const bool st_r = FINAL || !(w0 & F_ST_LAST);
        const bool st_c = FINAL || (w0 & F_CCR);
        const bool st_o = (w0 & F_OUT_DIRECT) && cx.valid;
        float* qo = cx.out_b + (size_t)(w0 >> 24) * cx.out_cstride + (size_t)m_lo * cx.N;
        const int N = cx.N;
#define SL_EACH _Pragma("unroll") for (int k = 0; k < K; ++k)
#define SL_FOR_M _Pragma("unroll 1") for (int m = 0; m < n_m; ++m, qr += sr, qa += sa, qx += sx, qy += sy, qccr += sccr, qo += N)
        // R store (:1079-1082 etc.), setCCR (:211-232), output (:1229-1233, :1248) for the current sample
#define SL_WRITE(SETS_ACC)                                                                                       \
        {                                                                                                        \
            if (st_r) sts<K>(qr, r);                                                                             \
            rp = r;                                                                                              \
            if (st_c) { Vec<K> c; SL_EACH { c[k] = ccr_of(r[k]); } sts<K>(qccr, c); }                             \
            if (!FINAL) { if (st_o) vstore<K>(qo, r); }                                                          \
            else {                                                                                               \
                if (st_o) vstore<K>(p.latch + (size_t)(w0 >> 24) * cx.N + cx.inst0, r);                           \
                if (SETS_ACC) cx.acc_last = accv;                                                                \
            }                                                                                                    \
        }
#define SL_LOAD3 Vec<K> a = lds<K>(qa), x = lds<K>(qx), y = lds<K>(qy); Vec<K> r, accv;                            \
        if (m > 0) { SL_EACH { a[k] = ca ? rp[k] : a[k]; x[k] = cxx ? rp[k] : x[k]; y[k] = cy ? rp[k] : y[k]; } }
        switch (uop) {
        case U_MACS: SL_FOR_M { SL_LOAD3
            SL_EACH { accv[k] = __fadd_rn(a[k], __fmul_rn(x[k], y[k])); r[k] = sat1(accv[k]); } SL_WRITE(true) } break;
        case U_MACSN: SL_FOR_M { SL_LOAD3
            SL_EACH { accv[k] = __fsub_rn(a[k], __fmul_rn(x[k], y[k])); r[k] = sat1(accv[k]); } SL_WRITE(true) } break;
        case U_ACC3: SL_FOR_M { SL_LOAD3
            SL_EACH { accv[k] = __fadd_rn(__fadd_rn(a[k], x[k]), y[k]); r[k] = sat1(accv[k]); } SL_WRITE(true) } break;
        case U_MACW: SL_FOR_M { SL_LOAD3
            SL_EACH { r[k] = __fadd_rn(a[k], wrap1(__fmul_rn(x[k], y[k]))); accv[k] = r[k]; } SL_WRITE(true) } break;
        case U_MACWN: SL_FOR_M { SL_LOAD3
            SL_EACH { r[k] = __fsub_rn(a[k], wrap1(__fmul_rn(x[k], y[k]))); accv[k] = r[k]; } SL_WRITE(true) } break;
        case U_MACINTW: SL_FOR_M { SL_LOAD3
            SL_EACH { r[k] = wrap1(__fadd_rn(a[k], __fmul_rn(x[k], y[k]))); accv[k] = r[k]; } SL_WRITE(true) } break;
        case U_ANDXOR: SL_FOR_M { SL_LOAD3
            SL_EACH { r[k] = __int2float_rn(logic_ops(a[k], x[k], y[k])); accv[k] = 0.0f; } SL_WRITE(false) } break;
        case U_TSTNEG: SL_FOR_M { SL_LOAD3
            SL_EACH {
                const int32_t q = cvt_x86(__fmul_rn(x[k], 2147483648.0f));
                r[k] = (a[k] >= y[k]) ? x[k] : __fmul_rn(__int2float_rn(~q), 4.656612873077392578125e-10f); accv[k] = r[k];
            } SL_WRITE(true) } break;
        case U_LIMIT: SL_FOR_M { SL_LOAD3
            SL_EACH { r[k] = (a[k] >= y[k]) ? x[k] : y[k]; accv[k] = r[k]; } SL_WRITE(true) } break;
        case U_LIMITN: SL_FOR_M { SL_LOAD3
            SL_EACH { r[k] = (a[k] < y[k]) ? x[k] : y[k]; accv[k] = r[k]; } SL_WRITE(true) } break;
        case U_INTERP: SL_FOR_M { SL_LOAD3
            SL_EACH {
                const double d = __dadd_rn(__dmul_rn(__dsub_rn(1.0, (double)x[k]), (double)a[k]), (double)__fmul_rn(x[k], y[k]));
                accv[k] = __double2float_rn(d); r[k] = sat1(accv[k]);
            } SL_WRITE(true) } break;
        case U_LOG:
        case U_EXP: {
            const uint32_t tb_s = cx.tab_s + (wB.y >> 24) * (uint32_t)TAB_SMEM_BYTES;
            SL_FOR_M {
                Vec<K> a = lds<K>(qa);
                if (m > 0) { SL_EACH { a[k] = ca ? rp[k] : a[k]; } }
                Vec<K> r, accv;
                int idx[K];
                double xd[K];
                bool wild = false;
                SL_EACH { wild |= !(fabsf(a[k]) <= 1.0f); xd[k] = (double)a[k]; }
                if (!wild) { SL_EACH { idx[k] = table_index_inrange(xd[k]); } }
                else { SL_EACH { idx[k] = table_index_wild(a[k]); if (!(fabsf(a[k]) <= 1.0f)) cx.flags |= FX8010_RT_TABLE_RANGE; } }   // rule U6
                if (w0 & F_TAB_SMEM) {
                    SL_EACH { double y1, slope; lds_f64x2(tb_s + (uint32_t)idx[k] * (TAB_REPL * 16u), y1, slope); r[k] = table_finish(xd[k], idx[k], y1, slope); }
                } else {
                    Vec<K> x;
                    if (!(w0 & F_TAB_IMM)) x = lds<K>(qx);
                    SL_EACH {
                        int tsel;
                        if (w0 & F_TAB_IMM) tsel = (int)(wB.y >> 24);
                        else {
                            int32_t sel = cvt_x86(x[k]);
                            if (sel < 0 || sel > FX8010_TABLE_COUNT - 1) { cx.flags |= FX8010_RT_TABLE_RANGE; sel = sel < 0 ? 0 : FX8010_TABLE_COUNT - 1; }
                            tsel = (uop == U_EXP ? FX8010_TABLE_COUNT : 0) + sel;
                        }
                        const double2 e = __ldg(reinterpret_cast<const double2*>(p.tabs + tsel * FX8010_TABLE_ENTRIES + idx[k]));
                        r[k] = table_finish(xd[k], idx[k], e.x, e.y);
                    }
                }
                SL_EACH { accv[k] = r[k]; }
                SL_WRITE(true)
            }
            break; }
        default: break;      // nothing else can appear in a stateless program's encoded stream
        }
#undef SL_EACH
#undef SL_FOR_M
#undef SL_WRITE
#undef SL_LOAD3
#undef SL_ADDR
#undef SL_ADDR_C
    }
}

template <int K>
__global__ void __launch_bounds__(128) fx_stateless_kernel(const SLParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int B = blockDim.x;
    const int tid = threadIdx.x;
    const int N = p.N, C = p.C, M = p.M;
    const int tslot_raw = blockIdx.x * B + tid;
    const bool valid = tslot_raw * K < N;
    const int inst0 = valid ? tslot_raw * K : N - K;
    const int seg = blockIdx.y;
    const int s_begin = seg * p.seg_len;
    const int s_end = min(p.n_samples, s_begin + p.seg_len);
    const uint32_t row_bytes = (uint32_t)B * K * 4u;
    const uint32_t buf_bytes = (uint32_t)M * row_bytes;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (!p.pdl_late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");

    TableEntry* const s_tab = reinterpret_cast<TableEntry*>(smem_raw);
    unsigned char* const col = smem_raw + (size_t)p.n_smem_tabs * TAB_SMEM_BYTES + (size_t)tid * K * 4;
    auto at = [&](uint32_t byte_off) { return reinterpret_cast<float*>(col + byte_off); };

    // input stage: batch starting at sample s0 -> buffer at byte offset boff (channel 0 unrolled)
    const bool has_in = (p.in != nullptr);
    auto fetch_batch = [&](int s0, uint32_t boff) {
        if (has_in && s0 < s_end) {
            const int n = s_end - s0;                    // samples left (>= 1); the batch takes min(M, n)
            const float* g = p.in + (size_t)s0 * N + inst0;
            float* d = at(p.stage0 + boff);
#pragma unroll
            for (int m = 0; m < SL_MAX_M; ++m)
                if (m < M && m < n) cp_async<4 * K>(reinterpret_cast<unsigned char*>(d) + (uint32_t)m * row_bytes, g + (size_t)m * N);
            for (int c = 1; c < C; ++c) {
                g += p.in_cstride;
                const uint32_t base = p.stage0 + (uint32_t)c * 2u * buf_bytes + boff;
                for (int m = 0; m < M && m < n; ++m) cp_async<4 * K>(at(base + (uint32_t)m * row_bytes), g + (size_t)m * N);
            }
        }
        cp_async_commit();
    };
    fetch_batch(s_begin, 0);

    for (int t = 0; t < p.n_smem_tabs; ++t) {
        const TableEntry* src = p.tabs + (size_t)p.smem_tab_id[t] * FX8010_TABLE_ENTRIES;
#pragma unroll 4
        for (int i = tid; i < FX8010_TABLE_ENTRIES * TAB_REPL; i += B)
            s_tab[t * FX8010_TABLE_ENTRIES * TAB_REPL + i] = src[i / TAB_REPL];
    }
    {   // read-only rows (controls, literals): the only state a stateless program can see
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
    cx.prog = c_prog[p.slot];
    cx.N = N; cx.inst0 = inst0; cx.valid = valid; cx.boff = 0; cx.flags = 0;
    cx.n_exec = p.n_exec; cx.out_cstride = p.out_cstride;
#pragma unroll
    for (int k = 0; k < K; ++k) cx.acc_last[k] = 0.0f;
    cx.out_b = p.out + (size_t)s_begin * N + inst0;
    const size_t out_step = (size_t)M * N;

    for (int s0 = s_begin; s0 < s_end; s0 += M, cx.out_b += out_step) {
        const int mb = min(M, s_end - s0);
        fetch_batch(s0 + M, cx.boff ^ buf_bytes);
        cp_async_wait<1>();
        sl_exec<K, false>(p, cx, 0, mb);
        if (s0 + mb == p.n_samples && valid) {
            // This thread owns the call's last sample: leave the final state behind (cold path).
            if (p.pdl_late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");   // state writes follow
            sl_exec<K, true>(p, cx, mb - 1, mb);
            for (int i = 0; i < p.n_wb; ++i) {
                const uint2 e = p.wb_list[i];
                const unsigned char* src = col + (e.x & SL_OFF_MASK) + ((e.x & SL_BUF) ? cx.boff : 0u) + (uint32_t)(mb - 1) * SL_STRIDE(e.x);
                vstore<K>(p.gpr + (size_t)e.y * N + inst0, vload<K>(reinterpret_cast<const float*>(src)));
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (p.acc_writer) p.acc[inst0 + k] = (double)cx.acc_last[k];
                p.counts[inst0 + k] += (unsigned long long)p.n_samples * (unsigned long long)p.n_instrs;
            }
        }
        cx.boff ^= buf_bytes;
    }
    if (cx.flags) atomicOr(p.rt_flags, cx.flags);
#undef SL_STRIDE
}

}  // namespace fxk
