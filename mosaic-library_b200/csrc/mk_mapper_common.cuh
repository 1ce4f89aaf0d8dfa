// mk_mapper_common.cuh -- types and device helpers shared by the compositing kernels (mk_mapper.cu: tile-per-CTA kernel,
// mk_strip.cu: persistent strip kernel).
#pragma once

#include "mk_common.cuh"
#include "mk_sample.cuh"

namespace mk {

constexpr int TW = 64;   // tile width  (== BLOCK_W)
constexpr int TH = 32;   // tile height
constexpr int HALO = 2;  // 5x5 erosion radius
constexpr int EH = TH + 2 * HALO;
constexpr int NTHREADS = 256;       // K3 warp kernel
constexpr int MTHREADS = 256;       // K4 (512 threads per tile measured 30-50 % slower: serial per-tile sections, lower occupancy)
constexpr int ROWB = TW * 3;       // canvas bytes per tile row
constexpr int ROWV = ROWB / 16;    // 16-byte vectors per tile row

// one frame to composite
struct FrameDesc {
    const uint8_t *src, *roi;
    unsigned long long spitch, rpitch;
    int h, w;
    int src_aligned;     // src 16-byte aligned with pitch % 16 == 0 (needed for the bulk-copy staging)
    int pad_;
    InvMap M;
};

struct MapperParams {
    uint8_t *canvas, *cmask;
    float *wacc;
    const short *cubic;
    const uint2 *lut;
    unsigned long long cpitch, mpitch, wpitch;
    int Hc, Wc;
    int tx0, ty0;
    float alpha, beta, ramp;
    uint32_t magic;      // 0x4B000000, see blend_pair_ch
    int src_smem_bytes;  // dynamic shared memory available for the staged source rectangle
    FrameDesc f0;        // single-frame launch: the frame, passed by value
    // batched launch (mk_mapper_update_batch): CTA b owns canvas tile tile_xy[b] and applies, in order, the frames
    // tile_frame[tile_ptr[b] .. tile_ptr[b+1])  -- the "gather" formulation: the canvas tile is read and written once
    const FrameDesc *frames;
    const int *tile_ptr, *tile_frame, *tile_xy;
};

// feather: canvas = sat_u8(rint((canvas*Wacc + warped*g) / (Wacc + g)))  
__device__ __forceinline__ uint32_t feather_u8(uint32_t c, uint32_t w, float acc, float g, float s)
{
    const float t = __fdiv_rn(__fadd_rn(__fmul_rn((float)c, acc), __fmul_rn((float)w, g)), s);
    return (uint32_t)min(255, max(0, __float2int_rn(t)));
}

// ---- async-proxy helpers (mbarrier + row-wise bulk copies; SASS: UBLKCP / SYNCS) -----------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, bytes % 16 == 0, both addresses 16-byte aligned; completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read()
{
    asm volatile("cp.async.bulk.commit_group;\ncp.async.bulk.wait_group.read 0;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ---- packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2): two IEEE-rn lanes per instruction -------------------------
__device__ __forceinline__ unsigned long long pack2(uint32_t lo, uint32_t hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, uint32_t &lo, uint32_t &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// cv2.addWeighted (mosaic.py:133) for TWO pixels at once: per channel rint(fmaf(c, a32, w * b32)), the two pixels in the
// two lanes of the f32x2 instructions.  Bytes become floats by OR-ing them into 2^23 and subtracting 2^23, and the result
// is rounded to nearest-even by adding 1.5 * 2^23 (exact for 0..255): no I2F / F2I on the 16-lane conversion unit.
// `magic` = 0x4B000000 (2^23) comes from the kernel parameters (MapperParams::magic), i.e. from the constant bank: with a
// literal the compiler puts it in PRMT's immediate slot and spends one IMAD.MOV per PRMT to materialise the selector.
template <int CH>
__device__ __forceinline__ void blend_pair_ch(uint32_t ca, uint32_t cb, uint32_t wa, uint32_t wb, unsigned long long a2,
                                              unsigned long long b2, uint32_t magic, uint32_t &ra, uint32_t &rb)
{
    const unsigned long long NEG = pack2(0xCB000000u, 0xCB000000u);      // -2^23
    const unsigned long long MAGIC = pack2(0x4B400000u, 0x4B400000u);    // 1.5 * 2^23
    unsigned long long cf = pack2(__byte_perm(ca, magic, 0x7650u | CH), __byte_perm(cb, magic, 0x7650u | CH));
    unsigned long long wf = pack2(__byte_perm(wa, magic, 0x7650u | CH), __byte_perm(wb, magic, 0x7650u | CH));
    cf = fadd2(cf, NEG);
    wf = fadd2(wf, NEG);
    const unsigned long long r = fadd2(ffma2(cf, a2, fmul2(wf, b2)), MAGIC);
    unpack2(r, ra, rb);
}
__device__ __forceinline__ void blend_pair(uint32_t ca, uint32_t cb, uint32_t wa, uint32_t wb, unsigned long long a2,
                                           unsigned long long b2, uint32_t magic, uint32_t &oa, uint32_t &ob)
{
    uint32_t a0, b0, a1, b1, a2r, b2r;
    blend_pair_ch<0>(ca, cb, wa, wb, a2, b2, magic, a0, b0);
    blend_pair_ch<1>(ca, cb, wa, wb, a2, b2, magic, a1, b1);
    blend_pair_ch<2>(ca, cb, wa, wb, a2, b2, magic, a2r, b2r);
    // low byte of every result (0..255: t <= 255.00003 rounds to at most 255) -> B | G << 8 | R << 16
    oa = __byte_perm(__byte_perm(a0, a1, 0x0040), a2r, 0x3410) & 0xFFFFFFu;
    ob = __byte_perm(__byte_perm(b0, b1, 0x0040), b2r, 0x3410) & 0xFFFFFFu;
}

struct DeviceTables {
    const short *cubic;
    const uint2 *lut;
};
// read-only interpolation tables, one copy per device, created on first use (mk_mapper.cu)
int device_tables(DeviceTables &out);

// mk_strip.cu: persistent strip kernel for the common case (alpha blend, no ROI, 16-byte aligned source, moderate
// minification).  Returns MK_OK and sets *done = true when it launched; *done = false means "not applicable, use the tile kernel".
int launch_strip_update(const MapperParams &P, const double *H_host, int interp, const int32_t *rect, cudaStream_t st, bool *done);

// mk_strip.cu: staged form of the plain warp (3 channels, 16-byte aligned source); *done = false -> use warp_kernel
int launch_warp_staged(const uint8_t *src, int h, int w, size_t spitch, const InvMap &M, uint8_t *dst, int Hd, int Wd, size_t dpitch, int interp,
                       const short *cubic, const uint2 *lut, cudaStream_t st, bool *done);

}  // namespace mk
