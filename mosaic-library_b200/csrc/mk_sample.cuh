// mk_sample.cuh -- OpenCV-exact inverse-map coordinate pipeline and fixed-point samplers.
//
// Restates (never copies) what cv::warpPerspective + cv::remap do for uint8 images with
// BORDER_CONSTANT(0), the call of src/mosaicking/mosaic.py:114:
//   * destination walked in blocks 64 px wide; X0/Y0/W0 evaluated in double at the block's
//     first column, advanced by M*x1 inside the block; no FMA contraction anywhere
//   * W = W ? 32/W : 0; coordinates rounded half-to-even to 1/32 px (cvt.rni saturates like
//     the INT_MIN/INT_MAX clamp for every finite value)
//   * INTER_LINEAR: integer weights (32-ax)(32-ay).. , (sum + 512) >> 10
//   * INTER_CUBIC : 32x32 table of 4x4 int16 weights (sum 32768), (sum + 2^14) >> 15, saturate
//   * a tap outside the source contributes 0 (per tap)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mk {

constexpr int BLOCK_W = 64;  // OpenCV's warpPerspective block width for dst width >= 64

struct InvMap {
    double m[9];  // inverse homography (canvas -> frame), cv::invert bits
    double wr;    // 32 / m[8] when affine
    int affine;   // m[6] == 0 && m[7] == 0
};

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }

// X0/Y0/W0 of a block row: block first column x0, destination row y.
__device__ __forceinline__ void row_base(const InvMap &M, int x0, int y, double &X0, double &Y0, double &W0)
{
    const double dx = (double)x0, dy = (double)y;
    X0 = dadd(dadd(dmul(M.m[0], dx), dmul(M.m[1], dy)), M.m[2]);
    Y0 = dadd(dadd(dmul(M.m[3], dx), dmul(M.m[4], dy)), M.m[5]);
    W0 = dadd(dadd(dmul(M.m[6], dx), dmul(M.m[7], dy)), M.m[8]);
}

// mx/my/mw = M[0]*x1, M[3]*x1, M[6]*x1 (x1 = column inside the block).
__device__ __forceinline__ void quantize(const InvMap &M, double X0, double Y0, double W0, double mx, double my,
                                         double mw, int &X, int &Y)
{
    double W;
    if (M.affine) {
        W = M.wr;  // W0 + (+-0) == m[8] exactly, so 32/W is a per-frame constant
    } else {
        W = dadd(W0, mw);
        W = (W != 0.0) ? __ddiv_rn(32.0, W) : 0.0;
    }
    X = __double2int_rn(dmul(dadd(X0, mx), W));
    Y = __double2int_rn(dmul(dadd(Y0, my), W));
}

// Same as quantize() for |value| < 2^31 (guaranteed by the caller: the tile corners bound every pixel of the tile and
// were checked to be far from saturation): adding 1.5 * 2^52 rounds to the nearest-even integer exactly like
// cvt.rni.s32.f64, but runs on the FP64 pipe instead of the 16-lane conversion unit.
__device__ __forceinline__ int round_magic(double v) { return __double2loint(__dadd_rn(v, 6755399441055744.0)); }

template <bool AFFINE>
__device__ __forceinline__ void quantize_fast_t(const InvMap &M, double wr, double X0, double Y0, double W0, double mx, double my,
                                                double mw, int &X, int &Y)
{
    double W = wr;
    if constexpr (!AFFINE) {
        W = dadd(W0, mw);
        W = (W != 0.0) ? __ddiv_rn(32.0, W) : 0.0;
    }
    X = round_magic(dmul(dadd(X0, mx), W));
    Y = round_magic(dmul(dadd(Y0, my), W));
}

__device__ __forceinline__ uint32_t ldg32(const uint32_t *p) { return __ldg(p); }

// ------------------------------------------------------------------------------------------
// INTER_LINEAR, 3 channels.  Returns B | G<<8 | R<<16.
// Fast path: the 2x2 footprint and the three aligned words around it lie inside the row.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tap3(const uint8_t *__restrict__ src, size_t pitch, int h, int w, int y, int x)
{
    if ((unsigned)x >= (unsigned)w || (unsigned)y >= (unsigned)h) return 0u;
    const uint8_t *p = src + (size_t)y * pitch + (size_t)x * 3;
    return (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
}

__device__ __forceinline__ uint32_t lerp3(uint32_t p00, uint32_t p01, uint32_t p10, uint32_t p11, int ax, int ay)
{
    const int w00 = (32 - ax) * (32 - ay), w01 = ax * (32 - ay), w10 = (32 - ax) * ay, w11 = ax * ay;
    uint32_t out = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int s = 8 * c;
        const int v = (int)((p00 >> s) & 255u) * w00 + (int)((p01 >> s) & 255u) * w01 +
                      (int)((p10 >> s) & 255u) * w10 + (int)((p11 >> s) & 255u) * w11 + 512;
        out |= (uint32_t)(v >> 10) << s;
    }
    return out;
}

__device__ __forceinline__ uint32_t sample_linear_c3(const uint8_t *__restrict__ src, size_t pitch, int h, int w,
                                                     int X, int Y)
{
    const int sx = X >> 5, sy = Y >> 5, ax = X & 31, ay = Y & 31;
    if (sx < -1 || sx >= w || sy < -1 || sy >= h) return 0u;
    uint32_t p00, p01, p10, p11;
    if (sx >= 0 && sx <= w - 4 && sy >= 0 && sy <= h - 2) {
        const uint8_t *a = src + (size_t)sy * pitch + (size_t)sx * 3;
        const uintptr_t ai = reinterpret_cast<uintptr_t>(a);
        const uint32_t *q0 = reinterpret_cast<const uint32_t *>(ai & ~(uintptr_t)3);
        const uint32_t *q1 = reinterpret_cast<const uint32_t *>((ai & ~(uintptr_t)3) + pitch);  // pitch % 4 == 0
        const unsigned sh = (unsigned)(ai & 3) * 8;
        const uint32_t a0 = ldg32(q0), a1 = ldg32(q0 + 1), a2 = ldg32(q0 + 2);
        const uint32_t b0 = ldg32(q1), b1 = ldg32(q1 + 1), b2 = ldg32(q1 + 2);
        const uint32_t alo = __funnelshift_r(a0, a1, sh), ahi = __funnelshift_r(a1, a2, sh);
        const uint32_t blo = __funnelshift_r(b0, b1, sh), bhi = __funnelshift_r(b1, b2, sh);
        p00 = alo & 0xFFFFFFu;
        p01 = (alo >> 24) | ((ahi & 0xFFFFu) << 8);
        p10 = blo & 0xFFFFFFu;
        p11 = (blo >> 24) | ((bhi & 0xFFFFu) << 8);
    } else {
        p00 = tap3(src, pitch, h, w, sy, sx);
        p01 = tap3(src, pitch, h, w, sy, sx + 1);
        p10 = tap3(src, pitch, h, w, sy + 1, sx);
        p11 = tap3(src, pitch, h, w, sy + 1, sx + 1);
    }
    return lerp3(p00, p01, p10, p11, ax, ay);
}

// INTER_LINEAR, 1 channel (keypoint-hull ROI mask, mosaic.py:118).
__device__ __forceinline__ uint32_t tap1(const uint8_t *__restrict__ src, size_t pitch, int h, int w, int y, int x)
{
    if ((unsigned)x >= (unsigned)w || (unsigned)y >= (unsigned)h) return 0u;
    return (uint32_t)__ldg(src + (size_t)y * pitch + x);
}

__device__ __forceinline__ uint32_t sample_linear_c1(const uint8_t *__restrict__ src, size_t pitch, int h, int w,
                                                     int X, int Y)
{
    const int sx = X >> 5, sy = Y >> 5, ax = X & 31, ay = Y & 31;
    if (sx < -1 || sx >= w || sy < -1 || sy >= h) return 0u;
    const int w00 = (32 - ax) * (32 - ay), w01 = ax * (32 - ay), w10 = (32 - ax) * ay, w11 = ax * ay;
    const int v = (int)tap1(src, pitch, h, w, sy, sx) * w00 + (int)tap1(src, pitch, h, w, sy, sx + 1) * w01 +
                  (int)tap1(src, pitch, h, w, sy + 1, sx) * w10 + (int)tap1(src, pitch, h, w, sy + 1, sx + 1) * w11 + 512;
    return (uint32_t)(v >> 10);
}

// ------------------------------------------------------------------------------------------
// INTER_CUBIC.  tab = 1024 x 16 int16 weights ([ay*32+ax][ky*4+kx]) in global memory.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int clamp_u8(int v) { return min(255, max(0, v)); }

// signed 16-bit weight pairs x unsigned bytes: d = a.lo * b.byte[0|2] + a.hi * b.byte[1|3] + c   (SASS IDP.2A)
__device__ __forceinline__ int dp2a_lo_su(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ uint32_t sample_cubic_c3(const uint8_t *__restrict__ src, size_t pitch, int h, int w,
                                                    int X, int Y, const short *__restrict__ tab)
{
    const int sx = X >> 5, sy = Y >> 5, ax = X & 31, ay = Y & 31;
    if (sx < -2 || sx > w || sy < -2 || sy > h) return 0u;
    const short *wt = tab + (ay * 32 + ax) * 16;
    int acc0 = 0, acc1 = 0, acc2 = 0;
    if (sx >= 1 && sx <= w - 5 && sy >= 1 && sy <= h - 3) {
        // interior: 16 int16 weights = two 16-byte loads (word k of row ky = the pair w[2k], w[2k+1]); every row of the 4x4
        // footprint is 12 bytes = four aligned words funnel-shifted, channels gathered by PRMT, two IDP.2A per channel and row
        const uint4 *wq = reinterpret_cast<const uint4 *>(wt);
        const uint4 wa = __ldg(wq), wb = __ldg(wq + 1);
        const uint32_t wp[8] = { wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w };
        const uint8_t *a = src + (size_t)(sy - 1) * pitch + (size_t)(sx - 1) * 3;
        const uintptr_t ai = reinterpret_cast<uintptr_t>(a);
        const unsigned sh = (unsigned)(ai & 3) * 8;
        const uint8_t *base = reinterpret_cast<const uint8_t *>(ai & ~(uintptr_t)3);
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(base + (size_t)ky * pitch);
            const uint32_t r0 = ldg32(q), r1 = ldg32(q + 1), r2 = ldg32(q + 2), r3 = ldg32(q + 3);
            const uint32_t d0 = __funnelshift_r(r0, r1, sh), d1 = __funnelshift_r(r1, r2, sh),
                           d2 = __funnelshift_r(r2, r3, sh);  // 12 bytes: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
            const uint32_t vb = __byte_perm(__byte_perm(d0, d1, 0x0630), d2, 0x5210);   // B0 B1 B2 B3
            const uint32_t vg = __byte_perm(__byte_perm(d0, d1, 0x0741), d2, 0x6210);   // G0 G1 G2 G3
            const uint32_t vr = __byte_perm(__byte_perm(d0, d1, 0x0052), d2, 0x7410);   // R0 R1 R2 R3
            const uint32_t w01 = wp[2 * ky], w23 = wp[2 * ky + 1];
            acc0 = dp2a_hi_su(w23, vb, dp2a_lo_su(w01, vb, acc0));
            acc1 = dp2a_hi_su(w23, vg, dp2a_lo_su(w01, vg, acc1));
            acc2 = dp2a_hi_su(w23, vr, dp2a_lo_su(w01, vr, acc2));
        }
    } else {
#pragma unroll
        for (int ky = 0; ky < 4; ++ky)
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                const uint32_t p = tap3(src, pitch, h, w, sy - 1 + ky, sx - 1 + kx);
                const int wv = wt[ky * 4 + kx];
                acc0 += (int)(p & 255u) * wv;
                acc1 += (int)((p >> 8) & 255u) * wv;
                acc2 += (int)((p >> 16) & 255u) * wv;
            }
    }
    return (uint32_t)clamp_u8((acc0 + (1 << 14)) >> 15) | ((uint32_t)clamp_u8((acc1 + (1 << 14)) >> 15) << 8) |
           ((uint32_t)clamp_u8((acc2 + (1 << 14)) >> 15) << 16);
}

__device__ __forceinline__ uint32_t sample_cubic_c1(const uint8_t *__restrict__ src, size_t pitch, int h, int w,
                                                    int X, int Y, const short *__restrict__ tab)
{
    const int sx = X >> 5, sy = Y >> 5, ax = X & 31, ay = Y & 31;
    if (sx < -2 || sx > w || sy < -2 || sy > h) return 0u;
    const short *wt = tab + (ay * 32 + ax) * 16;
    int acc = 0;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky)
#pragma unroll
        for (int kx = 0; kx < 4; ++kx)
            acc += (int)tap1(src, pitch, h, w, sy - 1 + ky, sx - 1 + kx) * (int)wt[ky * 4 + kx];
    return (uint32_t)clamp_u8((acc + (1 << 14)) >> 15);
}

// ------------------------------------------------------------------------------------------
// Shared-memory samplers.  The CTA has staged the source rectangle rows [sy0, sy0+R) x bytes
// [bx0, bx0+rowbytes) (bx0 % 16 == 0) with zero fill outside the image, so border taps need no
// predicates.  A tap set that does not lie inside the staged rectangle (never expected: the
// rectangle is the quantised image of the tile corners plus a safety ring) returns false and the
// caller falls back to the global-memory sampler, so exactness never depends on the estimate.
// ------------------------------------------------------------------------------------------
struct SrcTile {
    const uint8_t *s;   // shared memory
    int sy0, bx0, R, rowbytes;
};

// Bilinear weights as two int16 pairs: lut[ay * 32 + ax] = ((w00 | w01 << 16), (w10 | w11 << 16)), weights (32-ax)(32-ay) ...
// <= 1024.  Per channel the 2x2 quad sits in one register as bytes (v00, v01, v10, v11), so
//   sum(v * w) + 512 = dp2a.hi(w_bottom, V, dp2a.lo(w_top, V, 512))            (exact integer arithmetic, 2 IDP.2A)
__device__ __forceinline__ uint32_t lerp3_dp2a(uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint2 wt)
{
    // alo = [B0 G0 R0 B1], ahi = [G1 R1 . .] (row sy), blo / bhi the same for row sy + 1
    const uint32_t vb = __byte_perm(alo, blo, 0x7430);          // B0 B1 B0' B1'
    const uint32_t ra = __byte_perm(alo, ahi, 0x5241);          // G0 G1 R0 R1
    const uint32_t rb = __byte_perm(blo, bhi, 0x5241);          // G0' G1' R0' R1'
    const uint32_t vg = __byte_perm(ra, rb, 0x5410);            // G0 G1 G0' G1'
    const uint32_t vr = __byte_perm(ra, rb, 0x7632);            // R0 R1 R0' R1'
    const uint32_t b = (uint32_t)dp2a_hi_su(wt.y, vb, dp2a_lo_su(wt.x, vb, 512)) >> 10;
    const uint32_t g = (uint32_t)dp2a_hi_su(wt.y, vg, dp2a_lo_su(wt.x, vg, 512)) >> 10;
    const uint32_t r = (uint32_t)dp2a_hi_su(wt.y, vr, dp2a_lo_su(wt.x, vr, 512)) >> 10;
    return b | (g << 8) | (r << 16);
}

__device__ __forceinline__ bool sample_linear_c3_smem_lut(const SrcTile &T, int X, int Y, const uint2 *__restrict__ lut,
                                                          uint32_t &out)
{
    const int r = (Y >> 5) - T.sy0, cb = (X >> 5) * 3 - T.bx0;
    if ((unsigned)r >= (unsigned)(T.R - 1) || cb < 0 || cb + 12 > T.rowbytes) return false;
    const uint2 wt = __ldg(lut + ((Y & 31) * 32 + (X & 31)));
    const uint32_t *q0 = reinterpret_cast<const uint32_t *>(T.s + r * T.rowbytes + (cb & ~3));
    const uint32_t *q1 = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(q0) + T.rowbytes);
    const unsigned sh = (unsigned)(cb & 3) * 8;
    const uint32_t a0 = q0[0], a1 = q0[1], a2 = q0[2], b0 = q1[0], b1 = q1[1], b2 = q1[2];
    out = lerp3_dp2a(__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh), __funnelshift_r(b0, b1, sh),
                     __funnelshift_r(b1, b2, sh), wt);
    return true;
}

// Same without the containment test: valid when the staged rectangle is the UNCLAMPED corner hull + safety ring, which
// provably contains the taps of every pixel of the haloed tile (W keeps its sign => the image of the tile is the convex
// hull of the corner images; rounding to 1/32 px is monotone; the ring absorbs the 1-ulp evaluation-order differences).
__device__ __forceinline__ uint32_t sample_linear_c3_smem_fast(const SrcTile &T, int X, int Y, const uint2 *__restrict__ lut)
{
    const int r = (Y >> 5) - T.sy0, cb = (X >> 5) * 3 - T.bx0;
    const uint2 wt = __ldg(lut + ((Y & 31) * 32 + (X & 31)));
    const uint32_t *q0 = reinterpret_cast<const uint32_t *>(T.s + r * T.rowbytes + (cb & ~3));
    const uint32_t *q1 = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(q0) + T.rowbytes);
    const unsigned sh = (unsigned)(cb & 3) * 8;
    const uint32_t a0 = q0[0], a1 = q0[1], a2 = q0[2], b0 = q1[0], b1 = q1[1], b2 = q1[2];
    return lerp3_dp2a(__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh), __funnelshift_r(b0, b1, sh),
                      __funnelshift_r(b1, b2, sh), wt);
}

__device__ __forceinline__ bool sample_cubic_c3_smem(const SrcTile &T, int X, int Y, const short *__restrict__ tab,
                                                     uint32_t &out)
{
    const int ax = X & 31, ay = Y & 31;
    const int r = (Y >> 5) - 1 - T.sy0, cb = ((X >> 5) - 1) * 3 - T.bx0;
    if (r < 0 || r + 3 >= T.R || cb < 0 || cb + 16 > T.rowbytes) return false;
    // 16 int16 weights = two 16-byte loads; word k of row ky holds the weight pair (w[2k], w[2k+1])
    const uint4 *wq = reinterpret_cast<const uint4 *>(tab + (ay * 32 + ax) * 16);
    const uint4 wa = __ldg(wq), wb = __ldg(wq + 1);
    const uint32_t wp[8] = { wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w };
    const unsigned sh = (unsigned)(cb & 3) * 8;
    const uint8_t *base = T.s + r * T.rowbytes + (cb & ~3);
    int acc0 = 1 << 14, acc1 = 1 << 14, acc2 = 1 << 14;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(base + ky * T.rowbytes);
        const uint32_t r0 = q[0], r1 = q[1], r2 = q[2], r3 = q[3];
        // 12 bytes: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
        const uint32_t d0 = __funnelshift_r(r0, r1, sh), d1 = __funnelshift_r(r1, r2, sh), d2 = __funnelshift_r(r2, r3, sh);
        const uint32_t vb = __byte_perm(__byte_perm(d0, d1, 0x0630), d2, 0x5210);   // B0 B1 B2 B3
        const uint32_t vg = __byte_perm(__byte_perm(d0, d1, 0x0741), d2, 0x6210);   // G0 G1 G2 G3
        const uint32_t vr = __byte_perm(__byte_perm(d0, d1, 0x0052), d2, 0x7410);   // R0 R1 R2 R3
        const uint32_t w01 = wp[2 * ky], w23 = wp[2 * ky + 1];
        acc0 = dp2a_hi_su(w23, vb, dp2a_lo_su(w01, vb, acc0));
        acc1 = dp2a_hi_su(w23, vg, dp2a_lo_su(w01, vg, acc1));
        acc2 = dp2a_hi_su(w23, vr, dp2a_lo_su(w01, vr, acc2));
    }
    out = (uint32_t)clamp_u8(acc0 >> 15) | ((uint32_t)clamp_u8(acc1 >> 15) << 8) | ((uint32_t)clamp_u8(acc2 >> 15) << 16);
    return true;
}

template <int INTERP>
__device__ __forceinline__ uint32_t sample_c3_tile(bool use_smem, const SrcTile &T, const uint8_t *__restrict__ src,
                                                   size_t pitch, int h, int w, int X, int Y, const short *__restrict__ tab,
                                                   const uint2 *__restrict__ lut)
{
    uint32_t out;
    if (use_smem) {
        if constexpr (INTERP == 1) { if (sample_linear_c3_smem_lut(T, X, Y, lut, out)) return out; }
        else { if (sample_cubic_c3_smem(T, X, Y, tab, out)) return out; }
        // outside the staged rectangle: if it is also outside the tap reach of the image the value is 0
    }
    if constexpr (INTERP == 1) return sample_linear_c3(src, pitch, h, w, X, Y);
    else return sample_cubic_c3(src, pitch, h, w, X, Y, tab);
}

template <int INTERP, int CN>
__device__ __forceinline__ uint32_t sample(const uint8_t *__restrict__ src, size_t pitch, int h, int w, int X, int Y,
                                           const short *__restrict__ tab)
{
    if constexpr (INTERP == 1) {
        if constexpr (CN == 3) return sample_linear_c3(src, pitch, h, w, X, Y);
        else return sample_linear_c1(src, pitch, h, w, X, Y);
    } else {
        if constexpr (CN == 3) return sample_cubic_c3(src, pitch, h, w, X, Y, tab);
        else return sample_cubic_c1(src, pitch, h, w, X, Y, tab);
    }
}

}  // namespace mk
