// fx8010_short.cuh — the kernel for SHORT SKIP-free programs (at most SH_MAX_NI executed instructions): one-pole
// filters, feedback delay lines, gain stages — the recurrences that cannot be cut along time (sm_100a).
//
// Such a launch is bound by ONE warp's per-sample dependency chain (1 024 samples x the chain, however many SMs
// idle), and in the general interpreter that chain is mostly control flow: ~20 branches per sample at ~25 cycles
// each for a lone warp.  Here everything that does not depend on the sample is decided once per thread, before the
// sample loop:
//   * the program's NI instructions are read once and unrolled (NI is a template parameter: no length tests);
//   * every operand becomes a ready 32-bit shared-memory address plus a 0/1 factor telling whether it follows the
//     input stage (INPUT operands read the cp.async stage rows directly — reference source/FX8010.cpp:1053-1061
//     always preloads them first, so the copy into the register row is only needed for the final state);
//   * output stores are predicated (no branch); LOG/EXP table placement is folded into the micro-op;
//   * the call's last sample is peeled into a second, cold copy of the body (LAST = true) that additionally leaves
//     behind what the state write-back needs (INPUT register rows, CCR, output latches, accumulator).
// What is left per DSP instruction and sample: three address IMADs, three LDS, ONE indirect branch (the opcode
// switch compiles to a jump table), the arithmetic, one STS and a predicated STG.
//
// Eligibility is decided at load time (fx8010_gpu.cu::analyse, `short_ok`): no SKIP, no noise, no MACMV, nobody
// reads `ccr`, every output channel has a writer, every INPUT operand is preloaded by its own instruction.
// Arithmetic helpers are the generic kernel's (fx8010_kernel.cuh): bit-exact with the reference.
#pragma once

#include <type_traits>

#include "fx8010_kernel.cuh"

namespace fxk {

constexpr int SH_MAX_NI = 4;
// micro-ops private to this kernel: LOG/EXP with the table placement folded in
constexpr uint32_t SH_TAB_SMEM = U_END + 1;   // literal selector, table staged in shared memory
constexpr uint32_t SH_TAB_IMM = U_END + 2;    // literal selector, table in global memory


// Predicated global store of K floats (a branch here would sit on the per-sample chain).
template <int K> __device__ __forceinline__ void stg_if(uint32_t pred, float* p, const Vec<K>& r) {
    if (K == 4) asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.global.v4.f32 [%1], {%2,%3,%4,%5}; }" ::"r"(pred), "l"(p), "f"(r.v[0]), "f"(r.v[K > 1 ? 1 : 0]), "f"(r.v[K > 2 ? 2 : 0]), "f"(r.v[K > 3 ? 3 : 0]) : "memory");
    else if (K == 2) asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.global.v2.f32 [%1], {%2,%3}; }" ::"r"(pred), "l"(p), "f"(r.v[0]), "f"(r.v[K > 1 ? 1 : 0]) : "memory");
    else asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.global.f32 [%1], %2; }" ::"r"(pred), "l"(p), "f"(r.v[0]) : "memory");
}

// grid = (ceil(N / (K * B)), n_seg), block = B threads; same shared-memory layout and encoding as fx_interp_kernel.
template <int K, bool EXT, int NI>
__global__ void __launch_bounds__(128, 3) fx_short_kernel(const Params p) {
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
    const int RS = B * K;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (!p.pdl_late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");

    TableEntry* const s_tab = reinterpret_cast<TableEntry*>(smem_raw);
    unsigned char* const col = smem_raw + (size_t)p.n_smem_tabs * TAB_SMEM_BYTES + (size_t)tid * K * 4;
    auto at = [&](uint32_t byte_off) { return reinterpret_cast<float*>(col + byte_off); };
    const uint32_t col_s = (uint32_t)__cvta_generic_to_shared(col);
    const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(smem_raw) + (uint32_t)(tid & (TAB_REPL - 1)) * 16u;
    const uint32_t latch0 = latch_offset(p.n_regs, 0, RS);
    const uint32_t stage0 = stage_offset(p.n_regs, C, 0, RS, p.chunk);
    const uint32_t row_bytes = (uint32_t)RS * 4u;

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
    if (!has_in) {                                         // no input block: INPUT operands read zeros (the stage rows stand in for them)
        Vec<K> z;
#pragma unroll
        for (int k = 0; k < K; ++k) z[k] = 0.0f;
        for (int j = 0; j < 2 * C * chunk; ++j) vstore<K>(at(stage0 + (uint32_t)j * row_bytes), z);
    }
    for (int t = 0; t < p.n_smem_tabs; ++t) {
        const TableEntry* src = p.tabs + (size_t)p.smem_tab_id[t] * FX8010_TABLE_ENTRIES;
#pragma unroll 4
        for (int i = tid; i < FX8010_TABLE_ENTRIES * TAB_REPL; i += B)
            s_tab[t * FX8010_TABLE_ENTRIES * TAB_REPL + i] = src[i / TAB_REPL];
    }
    {
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
    int32_t iw[K], ir[K], xw[K], xr[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (EXT) {
            iw[k] = p.ptrs[inst0 + k]; ir[k] = p.ptrs[N + inst0 + k];
            xw[k] = p.ptrs[2 * N + inst0 + k]; xr[k] = p.ptrs[3 * N + inst0 + k];
        }
    }
    __syncthreads();                  // tables visible; also orders the start-up stores before the asm accesses below

    // ---- decode once: everything the sample loop needs, pinned in registers ----
    uint32_t s_uop[NI], s_w0[NI], s_qr[NI], s_qa[NI], s_qx[NI], s_qy[NI], s_fa[NI], s_fx[NI], s_fy[NI], s_aux[NI], s_out[NI];
    uint64_t s_ooff[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const uint4 wA = prog[2 * i], wB = prog[2 * i + 1];
        const uint32_t w0 = wA.x;
        uint32_t uop = w0 & 0xffu, aux = 0;
        if (w0 & F_TAB_SMEM) { uop = SH_TAB_SMEM; aux = tab_s + (wB.y >> 24) * (uint32_t)TAB_SMEM_BYTES; }
        else if (w0 & F_TAB_IMM) { uop = SH_TAB_IMM; aux = (wB.y >> 24) * (uint32_t)FX8010_TABLE_ENTRIES; }
        s_uop[i] = pin32(uop); s_w0[i] = pin32(w0); s_aux[i] = pin32(aux);
        s_qr[i] = pin32(col_s + wA.y);
        s_fa[i] = pin32((w0 & F_PRE_A) ? 1u : 0u); s_qa[i] = pin32(col_s + ((w0 & F_PRE_A) ? wB.z : wA.z));
        s_fx[i] = pin32((w0 & F_PRE_X) ? 1u : 0u); s_qx[i] = pin32(col_s + ((w0 & F_PRE_X) ? wB.z : wA.w));
        s_fy[i] = pin32((w0 & F_PRE_Y) ? 1u : 0u); s_qy[i] = pin32(col_s + ((w0 & F_PRE_Y) ? wB.z : wB.x));
        s_out[i] = pin32((valid && (w0 & F_OUT_DIRECT)) ? 1u : 0u);
        s_ooff[i] = pin64((uint64_t)(w0 >> 24) * (uint64_t)p.out_cstride);
    }
    const uint64_t Nl = pin64((uint64_t)N);
    float* const iring = reinterpret_cast<float*>(pin64((uint64_t)(EXT ? p.itram + inst0 : nullptr)));
    float* const xring = reinterpret_cast<float*>(pin64((uint64_t)(EXT ? p.xtram + inst0 : nullptr)));
    const int isize = (int)pin32((uint32_t)p.itram_size), xsize = (int)pin32((uint32_t)p.xtram_size);
    const TableEntry* const gtabs = reinterpret_cast<const TableEntry*>(pin64((uint64_t)p.tabs));
    const uint32_t valid_u = pin32(valid ? 1u : 0u);

    unsigned int flags = 0;
    Vec<K> acc_last;
    bool acc_set = false;
#pragma unroll
    for (int k = 0; k < K; ++k) acc_last[k] = 0.0f;

#define SH_EACH _Pragma("unroll") for (int k = 0; k < K; ++k)
    // ---- one sample period: FX8010::process, source/FX8010.cpp:1023-1249, for a SKIP-free program ----
    auto sample = [&](auto last_tag, const uint32_t stage_s, float* const out_s) {
        constexpr bool LAST = decltype(last_tag)::value;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const uint32_t qa = s_qa[i] + s_fa[i] * stage_s, qx = s_qx[i] + s_fx[i] * stage_s, qy = s_qy[i] + s_fy[i] * stage_s;
            if (LAST && (s_w0[i] & F_PRE_ANY)) {              // :1053-1061 — the preload's copy into the register row, for the final state
                const uint4 wA = prog[2 * i], wB = prog[2 * i + 1];
                const Vec<K> v = lds<K>(col_s + wB.z + stage_s);
                if (s_w0[i] & F_PRE_A) sts<K>(col_s + wA.z, v);
                if (s_w0[i] & F_PRE_X) sts<K>(col_s + wA.w, v);
                if (s_w0[i] & F_PRE_Y) sts<K>(col_s + wB.x, v);
            }
            const Vec<K> a = lds<K>(qa), x = lds<K>(qx), y = lds<K>(qy);
            Vec<K> r, accv;
            // R store (:1079-1082 etc.) and the period's output (:1229-1233, :1248); on the call's last sample also
            // setCCR (:211-232), the output latch and the accumulator
#define SH_WRITE(SETS_ACC)                                                                                   \
            {                                                                                                \
                sts<K>(s_qr[i], r);                                                                          \
                stg_if<K>(s_out[i], out_s + s_ooff[i], r);                                                   \
                if (LAST) {                                                                                  \
                    Vec<K> c; SH_EACH { c[k] = ccr_of(r[k]); } sts<K>(col_s, c);                             \
                    if (s_w0[i] & F_OUT_DIRECT) sts<K>(col_s + prog[2 * i + 1].w, r);                        \
                    if (SETS_ACC) { acc_last = accv; acc_set = true; }                                       \
                }                                                                                            \
            }
            switch (s_uop[i]) {
            case U_MACS: {                                    // :1077-1085 (MACINTS :1095-1103 is identical)
                SH_EACH { accv[k] = __fadd_rn(a[k], __fmul_rn(x[k], y[k])); r[k] = sat1(accv[k]); } SH_WRITE(true) } break;
            case U_MACSN: {                                   // :1086-1094
                SH_EACH { accv[k] = __fsub_rn(a[k], __fmul_rn(x[k], y[k])); r[k] = sat1(accv[k]); } SH_WRITE(true) } break;
            case U_ACC3: {                                    // :1104-1112
                SH_EACH { accv[k] = __fadd_rn(__fadd_rn(a[k], x[k]), y[k]); r[k] = sat1(accv[k]); } SH_WRITE(true) } break;
            case U_MACW: {                                    // :1126-1131
                SH_EACH { r[k] = __fadd_rn(a[k], wrap1(__fmul_rn(x[k], y[k]))); accv[k] = r[k]; } SH_WRITE(true) } break;
            case U_MACWN: {                                   // :1132-1137
                SH_EACH { r[k] = __fsub_rn(a[k], wrap1(__fmul_rn(x[k], y[k]))); accv[k] = r[k]; } SH_WRITE(true) } break;
            case U_MACINTW: {                                 // :1138-1143
                SH_EACH { r[k] = wrap1(__fadd_rn(a[k], __fmul_rn(x[k], y[k]))); accv[k] = r[k]; } SH_WRITE(true) } break;
            case U_ANDXOR: {                                  // :1150-1154 (accumulator untouched)
                SH_EACH { r[k] = __int2float_rn(logic_ops(a[k], x[k], y[k])); } SH_WRITE(false) } break;
            case U_TSTNEG: {                                  // :1155-1162
                SH_EACH {
                    const int32_t q = cvt_x86(__fmul_rn(x[k], 2147483648.0f));
                    r[k] = (a[k] >= y[k]) ? x[k] : __fmul_rn(__int2float_rn(~q), 4.656612873077392578125e-10f);
                    accv[k] = r[k];
                } SH_WRITE(true) } break;
            case U_LIMIT: {                                   // :1163-1168
                SH_EACH { r[k] = (a[k] >= y[k]) ? x[k] : y[k]; accv[k] = r[k]; } SH_WRITE(true) } break;
            case U_LIMITN: {                                  // :1169-1174
                SH_EACH { r[k] = (a[k] < y[k]) ? x[k] : y[k]; accv[k] = r[k]; } SH_WRITE(true) } break;
            case U_INTERP: {                                  // :1180-1187
                double omx[K];
                one_minus<K>(x.v, omx);
                interp_core<K>(omx, a.v, x.v, y.v, accv.v);
                SH_EACH { r[k] = sat1(accv[k]); } SH_WRITE(true) } break;
            case U_LOG: case U_EXP: case SH_TAB_SMEM: case SH_TAB_IMM: {   // :1113-1125, linearInterpolate :283-296
                int idx[K];
                double di[K];
                bool wild = false;
                SH_EACH { wild |= !(fabsf(a[k]) <= 1.0f); }
                if (!wild) { SH_EACH { idx[k] = table_index_inrange((double)a[k], di[k]); } }
                else {                                        // rule U6: clamp and flag (rare)
                    SH_EACH { idx[k] = table_index_wild(a[k]); di[k] = (double)idx[k]; if (!(fabsf(a[k]) <= 1.0f)) flags |= FX8010_RT_TABLE_RANGE; }
                }
                if (s_uop[i] == SH_TAB_SMEM) {
                    SH_EACH {
                        double y1, sl;
                        lds_f64x2(s_aux[i] + (uint32_t)idx[k] * (uint32_t)(TAB_REPL * 16), y1, sl);
                        r[k] = table_finish((double)a[k], di[k], y1, sl); accv[k] = r[k];
                    }
                } else {
                    SH_EACH {
                        uint32_t tsel;
                        if (s_uop[i] == SH_TAB_IMM) tsel = s_aux[i];
                        else {
                            int32_t sel = cvt_x86(x[k]);
                            if (sel < 0 || sel > FX8010_TABLE_COUNT - 1) { flags |= FX8010_RT_TABLE_RANGE; sel = sel < 0 ? 0 : FX8010_TABLE_COUNT - 1; }
                            tsel = (uint32_t)((s_uop[i] == U_EXP ? FX8010_TABLE_COUNT : 0) + sel) * (uint32_t)FX8010_TABLE_ENTRIES;
                        }
                        const double2 e = __ldg(reinterpret_cast<const double2*>(gtabs + tsel + idx[k]));
                        r[k] = table_finish((double)a[k], di[k], e.x, e.y); accv[k] = r[k];
                    }
                }
                SH_WRITE(true) } break;
            case U_IREAD: case U_XREAD:                       // :1190-1193 / :1202-1205, readSmallDelay :934-956
                if (EXT) {
                    const bool isx = (s_uop[i] == U_XREAD);
                    const float* const ring = isx ? xring : iring;
                    const int size = isx ? xsize : isize;
                    int ridx[K];
                    bool same = true;
                    SH_EACH {
                        int32_t& rp = isx ? xr[k] : ir[k];
                        const int pos = min(max(cvt_x86(y[k]), 0), size - 1);
                        int idx = rp - pos;
                        idx += (idx < 0) ? size : 0;          // rule U1: mathematical modulo
                        ridx[k] = idx;
                        rp = (rp + 1 == size) ? 0 : rp + 1;
                        same = same && (idx == ridx[0]);
                    }
                    Vec<K> v;
                    if (K > 1 && same) v = vload<K>(ring + (uint64_t)ridx[0] * Nl);
                    else { SH_EACH { v[k] = ring[(uint64_t)ridx[k] * Nl + k]; } }
                    sts<K>(s_qa[i], v);
                }
                break;
            case U_IWRITE: case U_XWRITE:                     // :1195-1198 / :1207-1210, writeSmallDelay :909-917
                if (EXT) {
                    const bool isx = (s_uop[i] == U_XWRITE);
                    float* const ring = isx ? xring : iring;
                    const int size = isx ? xsize : isize;
                    int widx[K];
                    bool same = true;
                    SH_EACH {
                        int32_t& wp = isx ? xw[k] : iw[k];
                        const int pos = min(max(cvt_x86(y[k]), 0), size - 1);
                        widx[k] = wp + pos;                   // not wrapped in the reference: slots at or beyond the ring are never read back
                        wp = (wp + 1 == size) ? 0 : wp + 1;
                        same = same && (widx[k] == widx[0]);
                    }
                    if (K > 1 && same) stg_if<K>((widx[0] < size) ? valid_u : 0u, ring + (uint64_t)widx[0] * Nl, a);
                    else { SH_EACH { if (widx[k] < size && valid_u) ring[(uint64_t)widx[k] * Nl + k] = a[k]; } }
                }
                break;
            default: break;                                   // cannot happen (END/NOP are not encoded for SKIP-free programs)
            }
#undef SH_WRITE
        }
    };

    float* out_s = p.out + (size_t)s_begin * N + inst0;
    const int last_s = p.n_samples - 1;
    uint32_t boff = 0;
    for (int c0 = s_begin; c0 < s_end; c0 += chunk, boff ^= buf_bytes) {
        fetch_chunk(c0 + chunk, boff ^ buf_bytes);
        cp_async_wait<1>();                                   // the chunk starting at c0 has landed
        const int c_end = min(s_end, c0 + chunk);
        const bool has_last = (c_end - 1 == last_s);
        const int n_bulk = c_end - c0 - (has_last ? 1 : 0);
        uint32_t stage_s = boff;
#pragma unroll 1
        for (int m = 0; m < n_bulk; ++m, out_s += Nl, stage_s += row_bytes) sample(std::false_type{}, stage_s, out_s);
        if (has_last) { sample(std::true_type{}, stage_s, out_s); out_s += Nl; }
    }
#undef SH_EACH

    // ---- write the state back (the last time segment carries the final state) ----
    if (p.pdl_late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (flags) atomicOr(p.rt_flags, flags);
    if (!valid || !last_seg) return;
    for (int i = 0; i < p.n_wb; ++i) { const uint32_t r = p.wb_regs[i]; vstore<K>(p.gpr + (size_t)p.reg_map[r] * N + inst0, vload<K>(at(reg_offset(r, RS)))); }
    for (int c = 0; c < C; ++c) vstore<K>(p.latch + (size_t)c * N + inst0, vload<K>(at(latch0 + (uint32_t)c * RS * 4u)));
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (acc_set) p.acc[inst0 + k] = (double)acc_last[k];
        if (EXT) {
            p.ptrs[inst0 + k] = iw[k]; p.ptrs[N + inst0 + k] = ir[k];
            p.ptrs[2 * N + inst0 + k] = xw[k]; p.ptrs[3 * N + inst0 + k] = xr[k];
        }
        p.counts[inst0 + k] += (unsigned long long)p.n_samples * (unsigned long long)p.n_instrs;
    }
}

}  // namespace fxk
