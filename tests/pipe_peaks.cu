// pipe_peaks.cu — issue-rate and latency microbenchmarks of the SM pipes the FX8010 interpreter leans on
// ("replace the nominal peaks with microbenchmarked values").
//
//   throughput: one block of 1 024 threads (32 warps) per SM, each thread runs 8 independent chains of one instruction
//               for ITER iterations; lane-ops / clock / SM = the block's lane-ops / its clock64() window (independent of
//               the MHz the chip happens to run at), averaged over the blocks that had an SM to themselves (%smid recorded);
//   latency   : one warp per SM, ONE dependent chain, cycles per instruction.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/pipe_peaks tests/pipe_peaks.cu
// Run (on the GPU box): tests/pipe_peaks > profiles/pipe_peaks.json
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e__)); return 1; } } while (0)

constexpr int ITER = 4096;
constexpr int CHAINS = 8;

enum Op { OP_FADD = 0, OP_FMUL, OP_FFMA, OP_FMNMX, OP_FSEL, OP_IMAD, OP_IADD3, OP_LOP3, OP_DADD, OP_DMUL, OP_DFMA, OP_F2F_64_32, OP_F2F_32_64,
          OP_F2I_32, OP_I2F_32, OP_F2I_64, OP_I2F_64, OP_LDS128, OP_LDS32, OP_COUNT };
static const char* OP_NAME[OP_COUNT] = {"fadd_f32", "fmul_f32", "ffma_f32", "fmnmx_f32", "fsetp_sel_f32", "imad_s32", "iadd3_s32", "prmt_b32", "dadd_f64", "dmul_f64",
                                         "dfma_f64", "f2f_f64_f32", "f2f_f32_f64", "f2i_s32_f32", "i2f_f32_s32", "f2i_s32_f64", "i2f_f64_s32", "lds_128", "lds_32"};

// one instruction of kind OP on chain registers (f: float, d: double, i: int); the asm is volatile so nothing is merged or dropped
template <int OP> __device__ __forceinline__ void step(float& f, double& d, int& i, float fc, double dc, int ic, uint32_t saddr) {
    if (OP == OP_FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f) : "f"(fc));
    else if (OP == OP_FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f) : "f"(fc));
    else if (OP == OP_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f) : "f"(fc));
    else if (OP == OP_FMNMX) asm volatile("min.NaN.f32 %0, %0, %1;" : "+f"(f) : "f"(fc));
    else if (OP == OP_FSEL) asm volatile("{ .reg .pred p; setp.ge.f32 p, %0, %1; selp.f32 %0, %1, %0, p; }" : "+f"(f) : "f"(fc));   // FSETP + FSEL pair
    else if (OP == OP_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(i) : "r"(ic));
    else if (OP == OP_IADD3) asm volatile("add.s32 %0, %0, %1;" : "+r"(i) : "r"(ic));
    else if (OP == OP_LOP3) asm volatile("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(i) : "r"(ic));   // ALU pipe (LOP3 chains get folded by ptxas)
    else if (OP == OP_DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d) : "d"(dc));
    else if (OP == OP_DMUL) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d) : "d"(dc));
    else if (OP == OP_DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d) : "d"(dc));
    else if (OP == OP_F2F_64_32) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(f)); asm volatile("mov.b64 {%0, _}, %1;" : "=f"(f) : "d"(d)); }   // (the mov keeps the chain; it is a plain register move)
    else if (OP == OP_F2F_32_64) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(d)); asm volatile("mov.b64 %0, {%1, %1};" : "=d"(d) : "f"(f)); }
    else if (OP == OP_F2I_32) { asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(i) : "f"(f)); asm volatile("mov.b32 %0, %1;" : "=f"(f) : "r"(i)); }
    else if (OP == OP_I2F_32) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(i)); asm volatile("mov.b32 %0, %1;" : "=r"(i) : "f"(f)); }
    else if (OP == OP_F2I_64) { asm volatile("cvt.rzi.s32.f64 %0, %1;" : "=r"(i) : "d"(d)); asm volatile("mov.b64 %0, {%1, %1};" : "=d"(d) : "r"(i)); }
    else if (OP == OP_I2F_64) { asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d) : "r"(i)); asm volatile("mov.b64 {%0, _}, %1;" : "=r"(i) : "d"(d)); }
    else if (OP == OP_LDS128) { float a, b, c, e; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(e) : "r"(saddr + ((uint32_t)i & 0x3f0u))); i += __float_as_int(a) & 16; }
    else if (OP == OP_LDS32) { float a; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(saddr + ((uint32_t)i & 0x3fcu))); i += __float_as_int(a) & 4; }
}

template <int OP, int NCH>
__global__ void __launch_bounds__(1024) k_pipe(unsigned long long* cycles, float* sink, int iters) {
    __shared__ __align__(16) float s_buf[4096 + 1024];
    for (int j = threadIdx.x; j < (int)(sizeof(s_buf) / 4); j += blockDim.x) s_buf[j] = 0.0f;
    __syncthreads();
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(s_buf) + (threadIdx.x * 16u & 0x3fffu);   // consecutive lanes, consecutive 16-byte slots
    float f[NCH]; double d[NCH]; int i[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { f[c] = 0.25f + 0.001f * (float)(threadIdx.x + c); d[c] = 0.5 + 0.001 * (double)(threadIdx.x + c); i[c] = (int)threadIdx.x + c; }
    const float fc = 0.999f + 1e-9f * (float)blockIdx.x; const double dc = 0.999 + 1e-12 * (double)blockIdx.x; const int ic = 3 + (int)(blockIdx.x & 1);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) step<OP>(f[c], d[c], i[c], fc, dc, ic, saddr);
        }
    }
    const long long t1 = clock64();
    float acc = 0.0f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) acc += f[c] + (float)d[c] + (float)i[c];
    if (acc == 123.456f) sink[0] = acc;
    if (threadIdx.x == 0) {
        unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        cycles[3 * blockIdx.x] = (unsigned long long)t0; cycles[3 * blockIdx.x + 1] = (unsigned long long)t1; cycles[3 * blockIdx.x + 2] = smid;
    }
}

typedef void (*Kfn)(unsigned long long*, float*, int);
template <int NCH> static Kfn pick(int op) {
    switch (op) {
#define C(o) case o: return k_pipe<o, NCH>;
    C(OP_FADD) C(OP_FMUL) C(OP_FFMA) C(OP_FMNMX) C(OP_FSEL) C(OP_IMAD) C(OP_IADD3) C(OP_LOP3) C(OP_DADD) C(OP_DMUL) C(OP_DFMA) C(OP_F2F_64_32) C(OP_F2F_32_64)
    C(OP_F2I_32) C(OP_I2F_32) C(OP_F2I_64) C(OP_I2F_64) C(OP_LDS128) C(OP_LDS32)
#undef C
    default: return nullptr;
    }
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    unsigned long long* d_cyc; float* d_sink;
    CK(cudaMalloc(&d_cyc, sizeof(unsigned long long) * 3 * sms * 8));
    CK(cudaMalloc(&d_sink, 64));
    std::vector<unsigned long long> h(3 * sms * 8);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_rate_khz_attr\": %d, \"method\": \"per-SM resident set of 32 warps x %d independent chains, clock64 window, lane-ops per clock per SM; latency: 1 warp, 1 chain\",\n \"pipes\": {\n", prop.name, sms, clk_khz, CHAINS);
    for (int op = 0; op < OP_COUNT; ++op) {
        // throughput: one block of 1 024 threads per SM
        const int blocks = sms;
        Kfn fn = pick<CHAINS>(op);
        for (int rep = 0; rep < 2; ++rep) { fn<<<blocks, 1024>>>(d_cyc, d_sink, ITER); }
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); fn<<<blocks, 1024>>>(d_cyc, d_sink, ITER); cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        CK(cudaMemcpy(h.data(), d_cyc, sizeof(unsigned long long) * 3 * blocks, cudaMemcpyDeviceToHost));
        std::vector<int> per_sm(1024, 0);
        for (int b = 0; b < blocks; ++b) per_sm[h[3 * b + 2] & 1023]++;
        double cyc_sum = 0; unsigned long long cmax = 0; int alone = 0;
        for (int b = 0; b < blocks; ++b) {
            const unsigned long long c = h[3 * b + 1] - h[3 * b];
            cmax = c > cmax ? c : cmax;
            if (per_sm[h[3 * b + 2] & 1023] == 1) { cyc_sum += (double)c; ++alone; }
        }
        const double cyc_avg = alone ? cyc_sum / alone : (double)cmax;
        const double lane_ops_per_sm = 1024.0 * CHAINS * 4.0 * ITER;                 // resident threads x chains x unroll x iterations
        const double thr = lane_ops_per_sm / cyc_avg;
        // latency: one warp per SM, one chain
        Kfn fl = pick<1>(op);
        fl<<<sms, 32>>>(d_cyc, d_sink, ITER); CK(cudaDeviceSynchronize());
        fl<<<sms, 32>>>(d_cyc, d_sink, ITER); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), d_cyc, sizeof(unsigned long long) * 3 * sms, cudaMemcpyDeviceToHost));
        double lat = 0;
        for (int b = 0; b < sms; ++b) lat += (double)(h[3 * b + 1] - h[3 * b]);
        lat /= (double)sms * 4.0 * ITER;
        printf("  \"%s\": {\"lane_ops_per_clk_per_sm\": %.2f, \"warp_instr_per_clk_per_sm\": %.3f, \"dependent_latency_cycles\": %.1f, \"kernel_ms\": %.3f, \"sm_mhz_effective\": %.0f, \"blocks_alone_on_their_sm\": %d}%s\n",
               OP_NAME[op], thr, thr / 32.0, lat, ms, (double)cmax / (ms * 1e3), alone, op + 1 < OP_COUNT ? "," : "");
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    printf(" },\n \"notes\": \"conversion rows (f2f/f2i/i2f) include one register move per conversion to keep a dependent chain; lds rows add one IADD+LOP per load\"\n}\n");
    return 0;
}
