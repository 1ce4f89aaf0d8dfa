// Pipe-throughput microbenchmark for the epilogue instruction mix (sm_100a).
// Each kernel runs ITER x 8 independent chains per thread; reports warp-instructions / cycle / SM.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITER 4096
#define NCH 8

template <int MODE>
__global__ void __launch_bounds__(512) k(uint32_t *out, uint32_t seed, long long *cycles)
{
    uint32_t a[NCH], b[NCH], c[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { a[i] = seed * (threadIdx.x + i + 1); b[i] = seed ^ (i * 77 + threadIdx.x); c[i] = seed + i * threadIdx.x; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (MODE == 0) { a[i] = __vminu2(a[i], c[i]); b[i] = __vmaxu2(b[i], c[i]); }                       // 2 VIMNMX.U16x2
            if (MODE == 1) { __half2 x = *(__half2 *)&a[i], y = *(__half2 *)&b[i]; __half2 z = *(__half2 *)&c[i]; x = __hmin2(x, z); y = __hmax2(y, z); a[i] = *(uint32_t *)&x; b[i] = *(uint32_t *)&y; }  // 2 HMNMX2
            if (MODE == 2) { a[i] = a[i] * 0xFFFFFF80u + b[i]; b[i] = b[i] * 0x10001u + a[i]; }               // 2 IMAD
            if (MODE == 3) { a[i] = __vminu2(a[i], c[i]); __half2 x = *(__half2 *)&c[i], y = *(__half2 *)&b[i]; y = __hmax2(y, x); b[i] = *(uint32_t *)&y; }   // VIMNMX + HMNMX2
            if (MODE == 4) { a[i] = __vminu2(a[i], c[i]); b[i] = b[i] * 0x10001u + c[i]; }                     // VIMNMX + IMAD
            if (MODE == 5) { a[i] = __byte_perm(a[i], b[i], 0x5410); b[i] = __byte_perm(b[i], a[i], 0x1032); } // 2 PRMT
            if (MODE == 6) { a[i] = __float_as_uint(__uint_as_float(a[i]) + 8388608.0f); b[i] = __float_as_uint(__uint_as_float(b[i]) + 8388608.0f); }  // 2 FADD
            if (MODE == 7) { a[i] = min((int)a[i], (int)c[i]); b[i] = max((int)b[i], (int)c[i]); }             // 2 VIMNMX (s32)
            if (MODE == 9) { a[i] = __popc(a[i] ^ c[i]) + a[i]; b[i] = __popc(b[i] ^ c[i]) + b[i]; }           // 2 (POPC + dependent IADD): the POPC issue rate is what is read off
            if (MODE == 10) { a[i] = __dp2a_lo((int)a[i], (int)c[i], (int)a[i]); b[i] = __dp2a_hi((int)b[i], (int)c[i], (int)b[i]); }   // 2 IDP.2A
            if (MODE == 8) { __half2 x = *(__half2 *)&a[i], y = *(__half2 *)&b[i]; __half2 z = *(__half2 *)&c[i]; x = __hmin2(x, z); a[i] = *(uint32_t *)&x; b[i] = b[i] * 0x10001u + c[i]; }  // HMNMX2 + IMAD
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += a[i] ^ b[i] ^ c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static double g_last = 0;   // warp-instructions / cycle / SM of the last run()

template <int MODE>
void run(const char *name)
{
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    k<MODE><<<148, 512>>>(out, 12345u, cyc);
    k<MODE><<<148, 512>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double winst = 16.0 * ITER * NCH * 2;   // warp-instructions per SM (16 warps)
    printf("%-22s %8.0f cycles  %.2f warp-inst/cycle/SM (%.2f per SMSP)\n", name, c, winst / c, winst / c / 4);
    g_last = winst / c;
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("VIMNMX.U16x2 x2"); run<7>("VIMNMX.S32 x2"); run<1>("HMNMX2 x2"); run<2>("IMAD x2"); run<5>("PRMT x2"); run<6>("FADD x2");
    run<3>("VIMNMX16 + HMNMX2"); run<4>("VIMNMX16 + IMAD"); run<8>("HMNMX2 + IMAD");
    run<10>("IDP.2A x2");
    // POPC: the loop body issues one POPC and one IADD per chain step; the pair count per cycle is the POPC rate when POPC is the slower pipe
    run<9>("(POPC + IADD) x2");
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double popc_warp_per_clk_sm = g_last;                 // run() counts two instructions per chain step = the two POPCs
    printf("{\"popc_warp_inst_per_clk_per_sm\": %.3f, \"popc_lanes_per_clk_per_sm\": %.1f, \"sms\": %d, \"sm_clock_mhz\": %.0f, \"popc32_per_s\": %.4e}\n",
           popc_warp_per_clk_sm, popc_warp_per_clk_sm * 32, sms, clk_khz / 1e3, popc_warp_per_clk_sm * 32 * sms * (clk_khz * 1e3));
    return 0;
}
