// mk_strip.cu -- K4s: persistent strip form of the fused warp + erode + blend + mask pass (one Mapper.update,
// src/mosaicking/mosaic.py:159-172 -> _update_cpu :111-138), the kernel mk_mapper_update launches whenever it applies
// (alpha blend, no keypoint ROI, 16-byte aligned source, moderate minification).  Same per-pixel arithmetic as the
// tile-per-CTA kernel in mk_mapper.cu (which stays the general fallback), different work decomposition:
//
//   * the visited canvas rectangle is cut into column strips 64 px wide = OpenCV's warpPerspective coordinate blocks;
//     the host clips the forward-mapped frame quad against every strip, so only rows the frame can reach are work;
//   * the rows of all strips form one flat list of 16-row chunks that is dealt out evenly to a persistent grid
//     (<= 4 CTAs per SM): every CTA walks DOWN its part of a strip, so the 2-row erosion halo above a chunk is the
//     tail of the previous chunk (kept in 32-row rings of warped pixels and validity bit-rows) instead of being
//     re-warped, and the per-tile set-up is amortised;
//   * every transfer is a 2-D TMA box (cp.async.bulk.tensor, SASS UTMALDG / UTMASTG): the source rectangle of chunk
//     j+2 and the canvas + mask tile of chunk j+1 are in flight while chunk j is computed (two mbarrier-tracked
//     buffers each); out-of-frame source bytes and out-of-canvas tile bytes are the TMA's zero fill / clipping, which
//     is exactly BORDER_CONSTANT(0) and the canvas edge handling;
//   * phase 1 handles two rows per warp iteration (two independent dependency chains per thread);
//   * phase 2 works on the packed BGR words: byte-mask selects, and for alpha = 0.5 (the Mapper default), 0 and 1 the
//     blend itself is word arithmetic -- rint(fmaf(c, .5, w * .5)) == round-half-even((c + w) / 2) exactly.
#include "mk_mapper_common.cuh"

#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <type_traits>

namespace mk {

constexpr int CH = 16;               // canvas rows finalised per chunk
constexpr int BR = CH + 2 * HALO;    // rows of the first phase-1 block of a segment
constexpr int RING = 64;             // ring of warped rows / bit rows, indexed by a running row counter of the CTA (three chunks in flight)
constexpr int STHREADS = 256;
constexpr int MAX_STRIPS = 200;
#ifndef MK_MAX_CHUNKS
#define MK_MAX_CHUNKS 64
#endif
constexpr int MAX_CHUNKS = MK_MAX_CHUNKS;      // chunks one CTA may own
#ifndef MK_NSRC
#define MK_NSRC 3
#endif
constexpr int NSRC = MK_NSRC;             // source boxes in flight per CTA (three: the box of chunk j+3 is requested while chunk j is computed --
                                    // enough lead for a frame that lives in another GPU's memory and arrives over NVLink)
constexpr int NCV = 4;              // canvas / mask tile buffers: tile i+1 is loading, i-1 is blended, i-2 is being stored, i-3 has left
constexpr int NROWB = 4;            // coordinate-base buffers (bases of chunk i+2 are written while chunk i-1 may still be read)
#ifndef MK_STRIP_OCC
#define MK_STRIP_OCC 3
#endif
#ifndef MK_STRIP_ILP
#define MK_STRIP_ILP 4
#endif
constexpr int STRIP_OCC = MK_STRIP_OCC;   // CTAs per SM the register budget is set for (four rows per warp in flight need ~76 registers)
#ifndef MK_P2_FAST
#define MK_P2_FAST 0                // build knob (A/B measured, off): phase 2 short path for groups of four pixels inside an existing overlap
                                    // (alpha = 0.5, mask already 255: no byte masks, no mask store; ~20 % fewer phase-2 instructions) -- 15.9 -> 15.8 us
                                    // per 1080p launch in a train, 49.4 -> 49.4 at 4K: the blend arithmetic is not what the pipeline waits for
#endif
#ifndef MK_STAGGER_NS
#define MK_STAGGER_NS 0
#endif
#ifndef MK_STAGGER_SMS
#define MK_STAGGER_SMS 148
#endif
#ifndef MK_SVC_WARP
#define MK_SVC_WARP 0
#endif
// Build knob (A/B measured, off): a ninth "service" warp erodes the validity rows and requests the source boxes instead of warp 1.
// In the ncu source view the other warps poll the "bit rows ready" barrier 13 times per chunk (12 % of all issued instructions), which
// looked like a critical chain on warp 1 -- but neither this warp, nor MK_P1_BALANCE, nor MK_WAIT_BACKOFF changes the launch time
// (16.0-16.4 us back to back at 1080p, 50-51.5 us at 4K): the polls only fill issue slots nobody else wants.  What does move the
// time is the number of USEFUL instructions (MK_EXP_SKIP_* below: phase-1 arithmetic -30 %, phase-2 arithmetic -15 %, side halo -3 %).
constexpr int SVC = MK_SVC_WARP;
#ifndef MK_P1_BALANCE
#define MK_P1_BALANCE 0
#endif
// Phase-1 rows of a chunk per warp (warp & 1 = which 32-column half, the four warps of a half split its rows): first row | rows << 8
// for a 16-row chunk, the same << 16 for the 20-row first chunk of a segment.  The warps with a side job (4-6: side-halo columns,
// ~150 instructions per chunk; 2-3: coordinate bases; 7: TMA) take fewer rows than warps 0 and 1, so that all eight finish a chunk
// at about the same time -- the pipeline runs at the pace of its slowest warp, and the others spend their slack polling.
#define MK_SH(a16, n16, a20, n20) ((uint32_t)(a16) | (uint32_t)(n16) << 8 | (uint32_t)(a20) << 16 | (uint32_t)(n20) << 24)
__constant__ uint32_t c_p1_share[8] = {
    MK_SH(0, 5, 0, 6),  MK_SH(0, 5, 0, 6),      // warps 0, 1
    MK_SH(5, 5, 6, 6),  MK_SH(5, 4, 6, 5),      // warps 2, 3
    MK_SH(10, 3, 12, 4), MK_SH(9, 3, 11, 4),    // warps 4, 5
    MK_SH(13, 3, 16, 4), MK_SH(12, 4, 15, 5),   // warps 6, 7
};
constexpr int NTHR = STHREADS + 32 * SVC;  // threads per CTA
constexpr int ERODE_WARP = SVC ? 8 : 1;

// shared-memory layout of the strip kernel: ONE dynamic block, every array at a constant offset from its 128-byte aligned base, so
// every shared address is `base register + immediate` (static __shared__ arrays cost an S2R / LEA pair per address on sm_100)
constexpr int OFF_CANVAS = 0;                                   // [NCV][CH][ROWB]   TMA destination / source
constexpr int OFF_MASK = OFF_CANVAS + NCV * CH * ROWB;          // [NCV][CH][TW]
constexpr int OFF_WARP = OFF_MASK + NCV * CH * TW;              // [RING][TW] uint32 warped pixels
constexpr int OFF_ROW = OFF_WARP + RING * TW * 4;               // [NROWB][3][BR] x 32 bytes: X0, Y0, W0 of the left / centre / right block
constexpr int OFF_M0 = OFF_ROW + NROWB * 3 * BR * 32;           // [RING] 64-bit validity rows
constexpr int OFF_H = OFF_M0 + RING * 8;                        // [RING] horizontally eroded rows
constexpr int OFF_V = OFF_H + RING * 8;                         // [2][CH] fully eroded rows of a chunk
constexpr int OFF_SIDE = OFF_V + 2 * CH * 8;                    // [RING][4] validity of the 2 + 2 side-halo columns
constexpr int OFF_BAR = OFF_SIDE + RING * 4;                    // mbarriers, see BAR_*
constexpr int OFF_ANY = OFF_BAR + 128;                          // [4] "chunk has a valid pixel"
constexpr int OFF_BLK = OFF_ANY + 16;                           // [MAX_CHUNKS] int4: tileX, first finalised row, first phase-1 row, phase-1 rows
constexpr int OFF_RECT = OFF_BLK + MAX_CHUNKS * 16;             // [MAX_CHUNKS] int4: source box row, byte, fits flag, ring offset
constexpr int OFF_SRC = (OFF_RECT + MAX_CHUNKS * 16 + 127) & ~127;   // [NSRC] source boxes
constexpr int BAR_SRC = OFF_BAR;                 // [NSRC] source box landed (tx bytes)
constexpr int BAR_CV = BAR_SRC + 8 * NSRC;       // [NCV]  canvas + mask tile landed
constexpr int BAR_V = BAR_CV + 8 * NCV;          // [2]    eroded bit rows of chunk i ready (warp 1 -> everyone)
constexpr int BAR_RR = BAR_V + 16;               // [2]    all eight warps have written their phase-1 rows of chunk i
constexpr int BAR_P2 = BAR_RR + 16;              // [2]    all eight warps have blended chunk i
// Two barriers, used alternately, wherever several warps arrive: a warp may be one chunk ahead of the slowest one, and an mbarrier
// cannot tell which phase an arrival is meant for -- on a single barrier the early arrival would complete the previous phase.
static_assert(BAR_P2 + 16 <= OFF_BAR + 128, "barrier block");
static_assert((OFF_M0 & 7) == 0 && (OFF_BLK & 15) == 0 && (OFF_ROW & 31) == 0, "alignment");

struct StripParams {
    const short *cubic;
    const uint2 *lut;
    int Hc, Wc;
    int tx0;                // first strip (units of 64 px)
    int nstrips, nchunks, chunks_per_cta;
    int boxW, boxR;         // source box: bytes per row (multiple of 16), rows
    int esz_shift;          // source tensor map element: 0 = uint8, 2 = uint32 (rows wider than 256 bytes)
    float alpha, beta;
    uint32_t magic;
    int blend_mode;         // 0 general alpha, 1 alpha == 0.5, 2 alpha == 1 (keep canvas), 3 alpha == 0 (take frame)
    FrameDesc f;
    int2 rows[MAX_STRIPS + 1]; // .x: first row of every strip, .y: chunks in the strips before it (kernel parameters are copied per launch: keep them small)
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, unsigned long long *bar)
{
    tma_load_2d(dst, map, c0, c1, smem_u32(bar));
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, uint32_t src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
// mbarrier wait (every lane polls; the hardware suspends the warp for up to the hint between polls) that traps instead of hanging
// the GPU if a transfer never completes (bad tensor map, lost transaction)
__device__ __forceinline__ void mbar_wait_addr(uint32_t a, uint32_t parity, int what = 0)
{
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 20); ++spin) {
        uint32_t ok;
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\nselp.u32 %0, 1, 0, P1;\n}\n"
                     : "=r"(ok) : "r"(a), "r"(parity), "r"(10000u) : "memory");
        if (ok) return;
    }
#ifdef MK_STRIP_DEBUG
    if ((threadIdx.x & 31) == 0) printf("strip kernel: wait timeout block %d warp %d what %d parity %u\n", blockIdx.x, threadIdx.x >> 5, what, parity);
    return;
#endif
    __trap();
}
// The same for waits that are expected to be long (a warp that is ahead of the pipeline): the hardware wakes a suspended try_wait on
// every barrier event of the CTA (13 polls of 7 instructions per wait in the ncu source view), so sleep blind between the polls.
#ifndef MK_WAIT_BACKOFF
#define MK_WAIT_BACKOFF 0          // build knob (A/B measured, off): blind sleep in ns between polls of the long waits
#endif
__device__ __forceinline__ void mbar_wait_slack(uint32_t a, uint32_t parity, int what = 0)
{
#if MK_WAIT_BACKOFF > 0
    uint32_t ok;
    asm volatile("{\n.reg .pred P1;\nmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 16); ++spin) {
        __nanosleep(MK_WAIT_BACKOFF);
        asm volatile("{\n.reg .pred P1;\nmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
#else
    mbar_wait_addr(a, parity, what);
#endif
}
__device__ __forceinline__ void mbar_wait_or_trap(unsigned long long *bar, uint32_t parity, int what = 0) { mbar_wait_addr(smem_u32(bar), parity, what); }
__device__ __forceinline__ void mbar_init_addr(uint32_t a, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive_addr(uint32_t a) { asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];\n" ::"r"(a) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx_addr(uint32_t a, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned long long lds64(uint32_t a)
{
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];\n" : "=l"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int4 lds128i(uint32_t a)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;\n" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;\n" ::"r"(a), "r"(v) : "memory"); }

// INTER_LINEAR from the staged box, no containment test (see sample_linear_c3_smem_fast).  `boxrel` = shared address of the box
// minus bx0, the byte column its first byte stands for (both multiples of 16), so the first tap lives at
// t = boxrel + r * pitch + 3 * sx: three aligned words from t & ~3 per row, funnel-shifted by 8 * (t & 3) (the shift wraps mod 32).
__device__ __forceinline__ uint32_t sample_linear_box(uint32_t boxrel, int pitch, int sy0, int X, int Y)
{
    const int r = (Y >> 5) - sy0;
    // the four integer weights as two int16 pairs, (w00 | w01 << 16, w10 | w11 << 16) = ((32-ay) * u, ay * u) with
    // u = (32-ax) | ax << 16: three IMADs on the FMA pipe instead of a 32-way divergent table read (8.5 L1 wavefronts per warp)
    const uint32_t ax = (uint32_t)X & 31u, ay = (uint32_t)Y & 31u;
    const uint32_t u = ax * 65535u + 32u;
    const uint32_t wbot = ay * u, wtop = (u << 5) - wbot;
    const uint2 wt = make_uint2(wtop, wbot);
    const uint32_t t = (uint32_t)(r * pitch) + ((uint32_t)(X >> 5) * 3u + boxrel);
    const uint32_t a = t & ~3u, b = a + (uint32_t)pitch;
    const unsigned sh = t << 3;
    const uint32_t a0 = lds32(a), a1 = lds32(a + 4), a2 = lds32(a + 8), b0 = lds32(b), b1 = lds32(b + 4), b2 = lds32(b + 8);
    return lerp3_dp2a(__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh), __funnelshift_r(b0, b1, sh),
                      __funnelshift_r(b1, b2, sh), wt);
}

// The kernel is a software pipeline over the chunks of a CTA with NO CTA-wide barrier in the loop.  Iteration i of every warp:
//     phase 1 of chunk i (warp its rows into the ring)  ->  arrive on "rows ready"  ->  its side job  ->  phase 2 of chunk i-1 (blend)
// so the erosion of chunk i (warp 1, needs everybody's rows) runs while the other warps blend chunk i-1 and warp chunk i+1, and the only
// waits a warp meets are on things produced an iteration earlier.  Side jobs: warp 1 erodes chunk i and requests source box i+3,
// warps 2-3 the coordinate bases of chunk i+2, warps 4-6 the side-halo columns of chunk i+1, warp 7 stores tile i-2 and loads tile i+1.
template <int INTERP>
__global__ void __launch_bounds__(NTHR, STRIP_OCC)
mapper_strip_kernel(const __grid_constant__ CUtensorMap map_src, const __grid_constant__ CUtensorMap map_cv,
                    const __grid_constant__ CUtensorMap map_mk, const __grid_constant__ StripParams p)
{
    extern __shared__ __align__(128) uint8_t s_dyn[];
    const uint32_t sb = (smem_u32(s_dyn) + 127u) & ~127u;          // shared address of the layout's base
    uint8_t *const sp = s_dyn + (sb - smem_u32(s_dyn));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // A launch costs about 6 us before and after its work (measured on 64 x 32 px frames): let the next launch of a train be
    // scheduled now (programmatic dependent launch: its CTAs move in as ours leave and run their prologue; they touch global memory
    // only after griddepcontrol.wait below, i.e. after this grid has completed), and start fetching the three tensor maps.
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_src) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_cv) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_mk) : "memory");
    }
    const FrameDesc &F = p.f;
    const int boxbytes = p.boxW * p.boxR;                 // bytes one TMA box delivers
    const int boxstride = (boxbytes + 127) & ~127;        // TMA destinations are 128-byte aligned
    const uint32_t src_base = sb + OFF_SRC;

    // ---- chunks [c0, c0 + n) of the flat list belong to this CTA: strip (binary search in the host's prefix sums) and rows ----
    const int c0 = (int)blockIdx.x * p.chunks_per_cta, n = min(p.nchunks, c0 + p.chunks_per_cta) - c0;
    if (n <= 0) return;
#if MK_STAGGER_NS > 0
    // build knob: the CTAs of a launch start together and would walk through their phases in lockstep (all three CTAs of an SM
    // in phase 1, then all three waiting): delay the second and third CTA of every SM by a fraction of an iteration
    if (blockIdx.x >= MK_STAGGER_SMS) __nanosleep((blockIdx.x / MK_STAGGER_SMS) * MK_STAGGER_NS);
#endif
    if (tid < n) {
        const int c = c0 + tid;
        int lo = 0, hi = p.nstrips;      // rows[lo].y <= c < rows[hi].y
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (p.rows[mid].y <= c) lo = mid; else hi = mid;
        }
        const int q = c - p.rows[lo].y, fy = p.rows[lo].x + q * CH;
        const bool seg_start = (tid == 0) || (q == 0);     // no rows of the previous chunk in the rings: warp the 2-row halo above too
        reinterpret_cast<int4 *>(sp + OFF_BLK)[tid] = make_int4((p.tx0 + lo) * TW, fy, seg_start ? fy - HALO : fy + HALO, seg_start ? BR : CH);
    }
    if (tid == 224) {
        for (int i = 0; i < NSRC; ++i) mbar_init_addr(sb + BAR_SRC + 8 * i, 1);
        for (int i = 0; i < NCV; ++i) mbar_init_addr(sb + BAR_CV + 8 * i, 1);
        mbar_init_addr(sb + BAR_V, 32); mbar_init_addr(sb + BAR_V + 8, 32);
        mbar_init_addr(sb + BAR_RR, STHREADS); mbar_init_addr(sb + BAR_RR + 8, STHREADS);
        mbar_init_addr(sb + BAR_P2, STHREADS); mbar_init_addr(sb + BAR_P2 + 8, STHREADS);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;\n" ::: "memory");      // everything the previous grid of the stream wrote (the canvas!) is visible from here on

    // ---- source boxes of ALL chunks of this CTA, four lanes per chunk: the quantised images of the four corners of the haloed
    //      block (columns tileX-2 .. tileX+65) bound every tap of the block as long as W keeps its sign; and the ring offset of the
    //      chunk (ring slot of canvas row y = (y + offset) & 63: a running count of the rows this CTA warps, so consecutive chunks of
    //      a strip share their halo rows and chunks of different strips never alias) ----
    if (tid == 224) {              // canvas + mask tile of chunk 0
        const int4 blk = lds128i(sb + OFF_BLK);
        mbar_expect_tx_addr(sb + BAR_CV, (uint32_t)(CH * ROWB + CH * TW));
        tma_load_2d(sb + OFF_CANVAS, &map_cv, blk.x * 3, blk.y, sb + BAR_CV);
        tma_load_2d(sb + OFF_MASK, &map_mk, blk.x, blk.y, sb + BAR_CV);
    }
    if (warp * 8 < n) {
        const int jj = min(tid >> 2, n - 1), cnr = tid & 3;
        const int4 blk = lds128i(sb + OFF_BLK + 16 * jj);
        const int tileX = blk.x, first = blk.z, nr = blk.w;
        int X, Y, sg;
        {
            const int x0 = (cnr & 1) ? tileX + BLOCK_W : tileX - BLOCK_W, x1 = (cnr & 1) ? HALO - 1 : BLOCK_W - HALO;
            const int y = (cnr & 2) ? min(first + nr - 1, p.Hc - 1) : max(first, 0);
            double X0, Y0, W0;
            row_base(F.M, x0, y, X0, Y0, W0);
            const double d1 = (double)x1;
            quantize(F.M, X0, Y0, W0, dmul(F.M.m[0], d1), dmul(F.M.m[3], d1), dmul(F.M.m[6], d1), X, Y);
            const double W = F.M.affine ? F.M.m[8] : dadd(W0, dmul(F.M.m[6], d1));
            const bool sane = abs(X) < (1 << 30) && abs(Y) < (1 << 30);
            sg = !sane ? 0 : ((W > 0.0) ? 1 : ((W < 0.0) ? -1 : 0));
        }
        int xlo = X, xhi = X, ylo = Y, yhi = Y, smin = sg, smax = sg;
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
            xlo = min(xlo, __shfl_xor_sync(0xffffffffu, xlo, o)); xhi = max(xhi, __shfl_xor_sync(0xffffffffu, xhi, o));
            ylo = min(ylo, __shfl_xor_sync(0xffffffffu, ylo, o)); yhi = max(yhi, __shfl_xor_sync(0xffffffffu, yhi, o));
            smin = min(smin, __shfl_xor_sync(0xffffffffu, smin, o)); smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, o));
        }
        if (cnr == 0 && (tid >> 2) < n) {
            constexpr int LO = (INTERP == 1) ? 1 : 2, HI = (INTERP == 1) ? 2 : 3;   // tap reach + 1 px safety ring
            const int sx_lo = (xlo >> 5) - LO, sx_hi = (xhi >> 5) + HI, sy_lo = (ylo >> 5) - LO, sy_hi = (yhi >> 5) + HI;
            const int bx0 = (sx_lo * 3) & ~15;      // the TMA needs a 16-byte aligned global start address
            const bool convex = smin == smax && smin != 0;
            const bool fits = convex && (sx_hi * 3 + 3 - bx0) <= p.boxW && (sy_hi - sy_lo + 1) <= p.boxR;
            int rows_before = 0;
            for (int k = 0; k < jj; ++k) rows_before += (int)lds32(sb + OFF_BLK + 16 * k + 12);
            reinterpret_cast<int4 *>(sp + OFF_RECT)[jj] = make_int4(sy_lo, bx0, fits ? 1 : 0, rows_before - first);
            if (jj < NSRC) {               // the first boxes leave right away: nothing can be computed before box 0 is here
                mbar_expect_tx_addr(sb + BAR_SRC + 8 * jj, fits ? (uint32_t)boxbytes : 0u);
                if (fits) tma_load_2d(src_base + (uint32_t)(jj * boxstride), &map_src, bx0 >> p.esz_shift, sy_lo, sb + BAR_SRC + 8 * jj);
            }
        }
    }
    // per-row coordinate bases of block j for the left / centre / right 64-px blocks; 60 threads (idx)
    auto row_bases = [&](int j, int idx) {
        const int4 blk = lds128i(sb + OFF_BLK + 16 * j);
        const int b = idx / BR, r = idx - b * BR;
        if (b < 3 && r < blk.w) {
            double X0, Y0, W0;
            row_base(F.M, blk.x + (b - 1) * BLOCK_W, blk.z + r, X0, Y0, W0);
            double *d = reinterpret_cast<double *>(sp + OFF_ROW + (((j & (NROWB - 1)) * 3 + b) * BR + r) * 32);
            d[0] = X0; d[1] = Y0; d[2] = W0;
        }
    };
    if (warp == 2 || warp == 3) row_bases(0, tid - 64);
    else if ((warp == 4 || warp == 5) && n > 1) row_bases(1, tid - 128);
    __syncthreads();

    auto issue_src = [&](int j) {                 // one thread
        const int buf = j % NSRC;
        const int4 rc = lds128i(sb + OFF_RECT + 16 * j);
        mbar_expect_tx_addr(sb + BAR_SRC + 8 * buf, rc.z ? (uint32_t)boxbytes : 0u);
        if (rc.z) tma_load_2d(src_base + (uint32_t)(buf * boxstride), &map_src, rc.y >> p.esz_shift, rc.x, sb + BAR_SRC + 8 * buf);
    };
    auto load_canvas = [&](int j) {               // one thread
        const int cb = j & (NCV - 1);
        const int4 blk = lds128i(sb + OFF_BLK + 16 * j);
        mbar_expect_tx_addr(sb + BAR_CV + 8 * cb, (uint32_t)(CH * ROWB + CH * TW));
        tma_load_2d(sb + OFF_CANVAS + cb * (CH * ROWB), &map_cv, blk.x * 3, blk.y, sb + BAR_CV + 8 * cb);
        tma_load_2d(sb + OFF_MASK + cb * (CH * TW), &map_mk, blk.x, blk.y, sb + BAR_CV + 8 * cb);
    };
    // the 2 + 2 side-halo columns of block j (they live in the neighbouring coordinate blocks): thread t = 4 * row + column
    auto side_halo = [&](int j, int t) {
        const int4 blk = lds128i(sb + OFF_BLK + 16 * j);
        const int tileX = blk.x, first = blk.z, nr = blk.w;
        if (t >= nr * 4) return;
#ifdef MK_EXP_SKIP_SIDE        // timing experiment only (wrong at frame edges): how much do the side-halo columns cost?
        { const int4 rq = lds128i(sb + OFF_RECT + 16 * j); sts8(sb + OFF_SIDE + (uint32_t)((first + (t >> 2) + rq.w) & (RING - 1)) * 4u + (uint32_t)(t & 3), 255u); return; }
#endif
        const int4 rc = lds128i(sb + OFF_RECT + 16 * j);
        SrcTile T;
        T.s = sp + OFF_SRC + (j % NSRC) * boxstride; T.sy0 = rc.x; T.bx0 = rc.y; T.R = p.boxR; T.rowbytes = p.boxW;
        const int r = t >> 2, cc = t & 3, y = first + r;
        const int x = (cc < 2) ? tileX - HALO + cc : tileX + TW + (cc - 2);
        const int b = (cc < 2) ? 0 : 2, x1 = (cc < 2) ? BLOCK_W - HALO + cc : cc - 2;
        uint32_t mv = 255;
        if (y >= 0 && y < p.Hc && x >= 0 && x < p.Wc) {
            const double d1 = (double)x1;
            const double *rb = reinterpret_cast<const double *>(sp + OFF_ROW + (((j & (NROWB - 1)) * 3 + b) * BR + r) * 32);
            const double sx = dmul(F.M.m[0], d1), sy = dmul(F.M.m[3], d1), sw = dmul(F.M.m[6], d1);
            int X, Y;
            if (rc.z != 0 && INTERP == 1) {        // staged: the box holds the halo columns' taps too (its corners are columns tileX-2 / tileX+65)
                if (F.M.affine) quantize_fast_t<true>(F.M, F.M.wr, rb[0], rb[1], rb[2], sx, sy, sw, X, Y);
                else quantize_fast_t<false>(F.M, F.M.wr, rb[0], rb[1], rb[2], sx, sy, sw, X, Y);
                mv = sample_linear_box(src_base + (uint32_t)((j % NSRC) * boxstride) - (uint32_t)rc.y, p.boxW, rc.x, X, Y) ? 255u : 0u;
            } else {
                quantize(F.M, rb[0], rb[1], rb[2], sx, sy, sw, X, Y);
                mv = sample_c3_tile<INTERP>(rc.z != 0, T, F.src, F.spitch, F.h, F.w, X, Y, p.cubic, p.lut) ? 255u : 0u;
            }
        }
        sts8(sb + OFF_SIDE + (uint32_t)((y + rc.w) & (RING - 1)) * 4u + (uint32_t)cc, mv);
    };

    if (warp < 3) {                       // side halo of chunk 0 (later chunks get theirs one iteration ahead)
        mbar_wait_addr(sb + BAR_SRC, 0u, 90);
        side_halo(0, tid);
    }

    // erosion of chunk i (everybody's rows are in) and the request for source box i+NSRC: one warp
    auto erode_chunk = [&](int i) {
            mbar_wait_slack(sb + BAR_RR + 8 * (i & 1), (uint32_t)(i >> 1) & 1u, 400 + i);
            if (lane == 0 && i + NSRC < n) issue_src(i + NSRC);      // into the buffer phase 1 of this chunk has finished reading
            const int4 blk = lds128i(sb + OFF_BLK + 16 * i);
            const int fy = blk.y, first = blk.z, nr = blk.w;
            const int roff = (int)lds32(sb + OFF_RECT + 16 * i + 12);
            if (lane < nr) {
                const uint32_t ri = (uint32_t)(first + lane + roff) & (RING - 1);
                const unsigned long long m = lds64(sb + OFF_M0 + ri * 8u);
                const uint32_t sd = lds32(sb + OFF_SIDE + ri * 4u);
                const unsigned long long l2 = (sd & 0xFFu) ? 1ull : 0ull, l1 = (sd & 0xFF00u) ? 1ull : 0ull;
                const unsigned long long r1 = (sd & 0xFF0000u) ? 1ull : 0ull, r2 = (sd & 0xFF000000u) ? 1ull : 0ull;
                const unsigned long long sl1 = (m << 1) | l1, sl2 = (m << 2) | (l1 << 1) | l2;
                const unsigned long long sr1 = (m >> 1) | (r1 << 63), sr2 = (m >> 2) | (r1 << 62) | (r2 << 63);
                sts64(sb + OFF_H + ri * 8u, m & sl1 & sl2 & sr1 & sr2);
            }
            __syncwarp();
            unsigned long long ev = 0;
            if (lane < CH) {
                const int y = fy + lane + roff;
                ev = lds64(sb + OFF_H + (uint32_t)((y - 2) & (RING - 1)) * 8u) & lds64(sb + OFF_H + (uint32_t)((y - 1) & (RING - 1)) * 8u) &
                     lds64(sb + OFF_H + (uint32_t)(y & (RING - 1)) * 8u) & lds64(sb + OFF_H + (uint32_t)((y + 1) & (RING - 1)) * 8u) &
                     lds64(sb + OFF_H + (uint32_t)((y + 2) & (RING - 1)) * 8u);
                sts64(sb + OFF_V + (uint32_t)((i & 1) * CH + lane) * 8u, ev);
            }
            const int any = __any_sync(0xffffffffu, ev != 0ull);
            if (lane == 0) sts32(sb + OFF_ANY + 4 * (i & 3), (uint32_t)any);
            mbar_arrive_addr(sb + BAR_V + 8 * (i & 1));
    };
    if (SVC && warp == ERODE_WARP) {
#pragma unroll 1
        for (int i = 0; i < n; ++i) erode_chunk(i);
        return;
    }
    const int half = warp & 1, cx = half * 32 + lane;
    // lanes whose column lies outside the canvas (last strip only) sample column 0 of the strip instead and discard the result:
    // no divergence in the row loop, and every tap stays inside the staged box
    const bool xvalid_all = (p.tx0 + p.nstrips) * TW <= p.Wc;      // no strip of this launch reaches beyond the canvas
    const double dx1 = (double)cx;
    double mx = dmul(F.M.m[0], dx1), my = dmul(F.M.m[3], dx1), mw = dmul(F.M.m[6], dx1);
    const uint32_t warp_base = sb + OFF_WARP + (uint32_t)cx * 4u;
    const uint32_t m0_base = sb + OFF_M0 + (uint32_t)half * 4u;
#if MK_P1_BALANCE
    const uint32_t share = c_p1_share[warp];
#endif

#pragma unroll 1
    for (int i = 0; i <= n + 1; ++i) {
        // =========================== phase 1 of chunk i: lanes = consecutive columns, four rows in flight ===========================
        if (i < n) {
            const int4 blk = lds128i(sb + OFF_BLK + 16 * i);
            const int tileX = blk.x, first = blk.z, nr = blk.w;
            const int4 rc = lds128i(sb + OFF_RECT + 16 * i);
            const int sy0 = rc.x, bx0 = rc.y;
            const bool staged = rc.z != 0;
            const int rbase = first + rc.w;                 // ring slot of block row r = (rbase + r) & 63
            const int sbuf = i % NSRC;
            const uint32_t boxrel = src_base + (uint32_t)(sbuf * boxstride) - (uint32_t)bx0;
            // the ring slots of chunk i were last used by chunk i-4 (or later): everybody must have blended chunk i-3
            if (i >= 3) mbar_wait_addr(sb + BAR_P2 + 8 * ((i - 3) & 1), (uint32_t)((i - 3) >> 1) & 1u, 600 + i);
            mbar_wait_addr(sb + BAR_SRC + 8 * sbuf, (uint32_t)(i / NSRC) & 1u, 100 + i);

            const bool xvalid = xvalid_all || tileX + cx < p.Wc;
            if (!xvalid_all) {
                const double dxs = xvalid ? (double)cx : 0.0;
                mx = dmul(F.M.m[0], dxs); my = dmul(F.M.m[3], dxs); mw = dmul(F.M.m[6], dxs);
            }
            const int r_lo = max(0, -first), r_hi = min(nr, p.Hc - first);    // block rows inside the canvas
            if (r_lo > 0 || r_hi < nr) {      // rows outside the canvas count as set (cv2.erode's default border)
                for (int r = warp >> 1; r < nr; r += 4)
                    if ((r < r_lo || r >= r_hi) && lane == 0) sts32(m0_base + (uint32_t)((rbase + r) & (RING - 1)) * 8u, ~0u);
            }
            uint32_t rowb = sb + OFF_ROW + (uint32_t)(((i & (NROWB - 1)) * 3 + 1) * BR * 32);
            asm volatile("" : "+r"(rowb));           // keep the address in a register (otherwise it is re-derived inside the row loop)
            auto finish_row = [&](int r, uint32_t px) {
                const unsigned bal = __ballot_sync(0xffffffffu, xvalid ? (px != 0u) : true);   // columns outside the canvas count as set
                const uint32_t ri = (uint32_t)(rbase + r) & (RING - 1);
                if (lane == 0) sts32(m0_base + ri * 8u, bal);
                sts32(warp_base + ri * (TW * 4), px);
            };
            auto body = [&](auto fast_tag, auto affine_tag) {
                constexpr bool FAST = decltype(fast_tag)::value, AFF = decltype(affine_tag)::value;
                const double wr = F.M.wr;
                SrcTile T;
                T.s = sp + OFF_SRC + sbuf * boxstride; T.sy0 = sy0; T.bx0 = bx0; T.R = p.boxR; T.rowbytes = p.boxW;
                auto one = [&](int r) -> uint32_t {
#ifdef MK_EXP_SKIP_P1          // timing experiment only (wrong results): how much of the launch is phase-1 arithmetic?
                    return 0x00010101u + (uint32_t)r;
#endif
                    double X0, Y0, W0 = 0.0;
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(X0), "=d"(Y0) : "r"(rowb + (uint32_t)r * 32u));
                    if constexpr (!AFF) asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(W0) : "r"(rowb + (uint32_t)r * 32u + 16u));
                    int X, Y;
                    if constexpr (FAST) quantize_fast_t<AFF>(F.M, wr, X0, Y0, W0, mx, my, mw, X, Y);
                    else quantize(F.M, X0, Y0, W0, mx, my, mw, X, Y);
                    if constexpr (FAST && INTERP == 1) return sample_linear_box(boxrel, p.boxW, sy0, X, Y);
                    else return sample_c3_tile<INTERP>(FAST, T, F.src, F.spitch, F.h, F.w, X, Y, p.cubic, p.lut);
                };
#if MK_P1_BALANCE
                // this warp's contiguous rows of its half: fewer for the warps that carry a side job (see c_p1_share)
                const uint32_t shr = share >> (nr == BR ? 16 : 0);
                int r = max(r_lo, (int)(shr & 0xFFu));
                const int re = min(r_hi, (int)((shr & 0xFFu) + ((shr >> 8) & 0xFFu)));
#pragma unroll 1
                for (; r + 3 <= re; r += 3) {     // three independent rows in flight
                    const uint32_t pa = one(r), pb = one(r + 1), pc = one(r + 2);
                    finish_row(r, xvalid ? pa : 0u); finish_row(r + 1, xvalid ? pb : 0u); finish_row(r + 2, xvalid ? pc : 0u);
                }
                if (r + 2 <= re) {
                    const uint32_t pa = one(r), pb = one(r + 1);
                    finish_row(r, xvalid ? pa : 0u); finish_row(r + 1, xvalid ? pb : 0u);
                } else if (r < re) {
                    const uint32_t pa = one(r);
                    finish_row(r, xvalid ? pa : 0u);
                }
#else
                int r = r_lo + (warp >> 1);
#if MK_STRIP_ILP == 4
                if (r + 12 < r_hi) {              // the usual case (16 or 20 rows inside the canvas): four independent rows in flight
                    const uint32_t pa = one(r), pb = one(r + 4), pc = one(r + 8), pd = one(r + 12);
                    finish_row(r, xvalid ? pa : 0u); finish_row(r + 4, xvalid ? pb : 0u);
                    finish_row(r + 8, xvalid ? pc : 0u); finish_row(r + 12, xvalid ? pd : 0u);
                    r += 16;
                }
#else
#pragma unroll 1
                for (; r + 4 < r_hi; r += 8) {    // two independent rows in flight
                    const uint32_t pa = one(r), pb = one(r + 4);
                    finish_row(r, xvalid ? pa : 0u); finish_row(r + 4, xvalid ? pb : 0u);
                }
#endif
#pragma unroll 1
                for (; r < r_hi; r += 4) {
                    const uint32_t pa = one(r);
                    finish_row(r, xvalid ? pa : 0u);
                }
#endif
            };
            if (staged) {
                if (F.M.affine) body(std::true_type{}, std::true_type{});
                else body(std::true_type{}, std::false_type{});
            } else {
                body(std::false_type{}, std::false_type{});
            }
            // coordinate bases of chunk i+2: published by this warp's "rows ready" arrival -> warp 1's erosion -> BAR_V of chunk i,
            // which every warp has waited for before it warps chunk i+2
            if ((warp == 2 || warp == 3) && i + 2 < n) row_bases(i + 2, tid - 64);
            mbar_arrive_addr(sb + BAR_RR + 8 * (i & 1));
        }

        // =========================== side jobs (warp 1 does its job after blending) ===========================
        if (warp >= 4 && warp < 7) {
            if (i + 1 < n) {               // side halo of chunk i+1: its bases are visible once BAR_V of chunk i-1 has completed
                if (i >= 1) mbar_wait_addr(sb + BAR_V + 8 * ((i - 1) & 1), (uint32_t)((i - 1) >> 1) & 1u, 700 + i);
                mbar_wait_addr(sb + BAR_SRC + 8 * ((i + 1) % NSRC), (uint32_t)((i + 1) / NSRC) & 1u, 300 + i);
                side_halo(i + 1, tid - 128);
            }
        } else if (warp == 7) {
            if (lane == 0) {
                if (i >= 2) {              // tile i-2 has been blended by everybody: store it
                    const int j = i - 2;
                    mbar_wait_addr(sb + BAR_P2 + 8 * (j & 1), (uint32_t)(j >> 1) & 1u, 800 + i);
                    if (lds32(sb + OFF_ANY + 4 * (j & 3))) {
                        const int4 blk = lds128i(sb + OFF_BLK + 16 * j);
                        tma_store_2d(&map_cv, blk.x * 3, blk.y, sb + OFF_CANVAS + (j & (NCV - 1)) * (CH * ROWB));
                        tma_store_2d(&map_mk, blk.x, blk.y, sb + OFF_MASK + (j & (NCV - 1)) * (CH * TW));
                    }
                    bulk_commit();
                }
                if (i + 1 < n) {
                    if (i + 1 >= NCV) bulk_wait_read1();     // the store of tile i-3 (issued an iteration ago) has read this buffer
                    load_canvas(i + 1);
                }
            }
            __syncwarp();
        }

        // =========================== phase 2 of chunk i-1: thread = 4 consecutive pixels = 3 packed BGR words ===========================
        if (i >= 1 && i <= n) {
            const int j = i - 1, cbuf = j & (NCV - 1);
            const int4 blk = lds128i(sb + OFF_BLK + 16 * j);
            const int fy = blk.y;
            const int roff = (int)lds32(sb + OFF_RECT + 16 * j + 12);
            mbar_wait_slack(sb + BAR_V + 8 * (j & 1), (uint32_t)(j >> 1) & 1u, 500 + i);
            const int r = tid >> 4, g = tid & 15;
            const unsigned long long v = lds64(sb + OFF_V + (uint32_t)((j & 1) * CH + r) * 8u);
            const uint32_t vw = (g & 8) ? (uint32_t)(v >> 32) : (uint32_t)v;
            const unsigned mb = (vw >> (4 * (g & 7))) & 15u;
            mbar_wait_addr(sb + BAR_CV + 8 * cbuf, (uint32_t)(j / NCV) & 1u, 200 + i);
#ifdef MK_EXP_SKIP_P2          // timing experiment only (wrong results): how much of the launch is phase-2 arithmetic?
            if (false) {
#else
            if (mb) {
#endif
                uint32_t *cw = reinterpret_cast<uint32_t *>(sp + OFF_CANVAS + cbuf * (CH * ROWB) + r * ROWB + 12 * g);
                uint32_t *ms = reinterpret_cast<uint32_t *>(sp + OFF_MASK + cbuf * (CH * TW) + r * TW + 4 * g);
                const uint32_t c0w = cw[0], c1w = cw[1], c2w = cw[2], mk4 = *ms;
                const uint4 wv = *reinterpret_cast<const uint4 *>(sp + OFF_WARP + ((fy + r + roff) & (RING - 1)) * (TW * 4) + 16 * g);
                const uint32_t W0 = __byte_perm(wv.x, wv.y, 0x4210), W1 = __byte_perm(wv.y, wv.z, 0x5421), W2 = __byte_perm(wv.z, wv.w, 0x6542);
#if MK_P2_FAST
                // interior of an overlap (all four pixels valid, all four already covered with mask 255 -- the bulk of a frame that
                // lands on existing mosaic): the result is the blend, the mask does not change, no byte masks needed
                if (p.blend_mode == 1 && mb == 15u && mk4 == 0xFFFFFFFFu) {
                    auto avg = [](uint32_t a, uint32_t b) {
                        const uint32_t x = a ^ b, h = (a & b) + ((x & 0xFEFEFEFEu) >> 1);
                        return h + (x & h & 0x01010101u);
                    };
                    cw[0] = avg(c0w, W0); cw[1] = avg(c1w, W1); cw[2] = avg(c2w, W2);
                    fence_proxy_async();
                } else {
#endif
                // per-pixel byte masks: valid (eroded mask bit), old (canvas mask byte > 1, mosaic.py:123)
                const uint32_t Mv = ((mb * 0x00204081u) & 0x01010101u) * 0xFFu;
                const uint32_t t = mk4 & 0xFEFEFEFEu;
                const uint32_t Mo = (((((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u) >> 7) * 0xFFu;
                uint32_t B0, B1, B2;
                if (p.blend_mode == 1) {            // alpha = 0.5: round-half-even of (c + w) / 2 on four bytes at a time
                    auto avg = [](uint32_t a, uint32_t b) {
                        const uint32_t x = a ^ b, h = (a & b) + ((x & 0xFEFEFEFEu) >> 1);
                        return h + (x & h & 0x01010101u);
                    };
                    B0 = avg(c0w, W0); B1 = avg(c1w, W1); B2 = avg(c2w, W2);
                } else if (p.blend_mode == 2) {     // fmaf(c, 1, w * 0) == c
                    B0 = c0w; B1 = c1w; B2 = c2w;
                } else if (p.blend_mode == 3) {     // fmaf(c, 0, w * 1) == w
                    B0 = W0; B1 = W1; B2 = W2;
                } else {
                    B0 = c0w; B1 = c1w; B2 = c2w;
                    const uint32_t need = Mv & Mo;
                    if (need) {
                        const uint32_t px0 = c0w & 0xFFFFFFu, px1 = (c0w >> 24) | ((c1w & 0xFFFFu) << 8),
                                       px2 = (c1w >> 16) | ((c2w & 0xFFu) << 16), px3 = c2w >> 8;
                        const uint32_t ab = __float_as_uint(p.alpha), bb = __float_as_uint(p.beta);
                        const unsigned long long a2 = pack2(ab, ab), b2 = pack2(bb, bb);
                        uint32_t b0 = px0, b1 = px1, b2r = px2, b3 = px3;
                        if (need & 0x0000FFFFu) blend_pair(px0, px1, wv.x, wv.y, a2, b2, p.magic, b0, b1);
                        if (need & 0xFFFF0000u) blend_pair(px2, px3, wv.z, wv.w, a2, b2, p.magic, b2r, b3);
                        B0 = __byte_perm(b0, b1, 0x4210); B1 = __byte_perm(b1, b2r, 0x5421); B2 = __byte_perm(b2r, b3, 0x6542);
                    }
                }
                // result = valid ? (old ? blend : warped) : canvas, per pixel, on the 12 bytes of the group
                const uint32_t Ev0 = __byte_perm(Mv, Mv, 0x1000), Ev1 = __byte_perm(Mv, Mv, 0x2211), Ev2 = __byte_perm(Mv, Mv, 0x3332);
                const uint32_t Eo0 = __byte_perm(Mo, Mo, 0x1000), Eo1 = __byte_perm(Mo, Mo, 0x2211), Eo2 = __byte_perm(Mo, Mo, 0x3332);
                const uint32_t n0 = (Eo0 & B0) | (~Eo0 & W0), n1 = (Eo1 & B1) | (~Eo1 & W1), n2 = (Eo2 & B2) | (~Eo2 & W2);
                cw[0] = (Ev0 & n0) | (~Ev0 & c0w);
                cw[1] = (Ev1 & n1) | (~Ev1 & c1w);
                cw[2] = (Ev2 & n2) | (~Ev2 & c2w);
                *ms = mk4 | Mv;                                                     // mosaic.py:138
                fence_proxy_async();
#if MK_P2_FAST
                }
#endif
            }
            mbar_arrive_addr(sb + BAR_P2 + 8 * (j & 1));
        }

        if (!SVC && warp == 1 && i < n) erode_chunk(i);
    }
    if (tid == 224) bulk_wait_read0();     // the last stores must have read their tiles out of shared memory before the CTA exits
}

// ---------------------------------------------------------------------------------------------
// K3s: plain cv2.warpPerspective (3 channels) with the source footprint staged in shared memory.  One CTA per 64 x 32 output
// tile (= one OpenCV coordinate block wide): the quantised images of the tile corners give the source rectangle, ONE 2-D TMA box
// brings it in (zero fill outside the frame = BORDER_CONSTANT 0), the 8 warps sample it (two rows per iteration, the same
// unchecked sampler as the compositing kernel), and the tile leaves as whole 16-byte vectors (three per 16 pixels).
// ---------------------------------------------------------------------------------------------
struct WarpStagedParams {
    const uint8_t *src;
    uint8_t *dst;
    const short *cubic;
    const uint2 *lut;
    unsigned long long spitch, dpitch;
    int h, w, Hd, Wd;
    int boxW, boxR, esz_shift;
    InvMap M;
};

template <int INTERP>
__global__ void __launch_bounds__(256, 4) warp_staged_kernel(const __grid_constant__ CUtensorMap map_src, const __grid_constant__ WarpStagedParams p)
{
    __shared__ __align__(16) uint32_t s_px[TH][TW];
    __shared__ __align__(16) double s_rowb[TH][4];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int s_rc[4];
    extern __shared__ __align__(128) uint8_t s_dyn[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tileX = (int)blockIdx.x * TW, tileY = (int)blockIdx.y * TH;
    const uint32_t box = (smem_u32(s_dyn) + 127u) & ~127u;
    const int rows = min(TH, p.Hd - tileY);
    if (tid == 64) mbar_init(&s_bar, 1);
    if (warp == 0) {
        int X = 0, Y = 0, sg = 0;
        if (lane < 4) {
            const int x1 = (lane & 1) ? min(TW, p.Wd - tileX) - 1 : 0, y = (lane & 2) ? tileY + rows - 1 : tileY;
            double X0, Y0, W0;
            row_base(p.M, tileX, y, X0, Y0, W0);
            const double d1 = (double)x1;
            quantize(p.M, X0, Y0, W0, dmul(p.M.m[0], d1), dmul(p.M.m[3], d1), dmul(p.M.m[6], d1), X, Y);
            const double W = p.M.affine ? p.M.m[8] : dadd(W0, dmul(p.M.m[6], d1));
            const bool sane = abs(X) < (1 << 30) && abs(Y) < (1 << 30);
            sg = !sane ? 0 : ((W > 0.0) ? 1 : ((W < 0.0) ? -1 : 0));
        }
        int xlo = X, xhi = X, ylo = Y, yhi = Y, smin = sg, smax = sg;
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
            xlo = min(xlo, __shfl_xor_sync(0xffffffffu, xlo, o)); xhi = max(xhi, __shfl_xor_sync(0xffffffffu, xhi, o));
            ylo = min(ylo, __shfl_xor_sync(0xffffffffu, ylo, o)); yhi = max(yhi, __shfl_xor_sync(0xffffffffu, yhi, o));
            smin = min(smin, __shfl_xor_sync(0xffffffffu, smin, o)); smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, o));
        }
        if (lane == 0) {
            constexpr int LO = (INTERP == 1) ? 1 : 2, HI = (INTERP == 1) ? 2 : 3;
            const int sx_lo = (xlo >> 5) - LO, sx_hi = (xhi >> 5) + HI, sy_lo = (ylo >> 5) - LO, sy_hi = (yhi >> 5) + HI;
            const int bx0 = (sx_lo * 3) & ~15;
            const bool fits = smin == smax && smin != 0 && (sx_hi * 3 + 3 - bx0) <= p.boxW && (sy_hi - sy_lo + 1) <= p.boxR;
            s_rc[0] = sy_lo; s_rc[1] = bx0; s_rc[2] = fits ? 1 : 0;
        }
    } else if (warp == 1) {
        if (lane < rows) {
            double X0, Y0, W0;
            row_base(p.M, tileX, tileY + lane, X0, Y0, W0);
            s_rowb[lane][0] = X0; s_rowb[lane][1] = Y0; s_rowb[lane][2] = W0;
        }
    }
    __syncthreads();
    const int sy0 = s_rc[0], bx0 = s_rc[1];
    const bool staged = s_rc[2] != 0;
    if (tid == 0) {
        mbar_arrive_expect_tx(&s_bar, staged ? (uint32_t)(p.boxW * p.boxR) : 0u);
        if (staged) tma_load_2d(box, &map_src, bx0 >> p.esz_shift, sy0, &s_bar);
    }
    const int cx = (warp & 1) * 32 + lane;
    const double d1 = (double)cx;
    const double mx = dmul(p.M.m[0], d1), my = dmul(p.M.m[3], d1), mw = dmul(p.M.m[6], d1);
    if (lane == 0) mbar_wait_or_trap(&s_bar, 0u, 400);
    __syncwarp();
    {
        const bool xvalid = tileX + cx < p.Wd;
        auto body = [&](auto fast_tag, auto affine_tag) {
            constexpr bool FAST = decltype(fast_tag)::value, AFF = decltype(affine_tag)::value;
            const double wr = p.M.wr;
            const uint32_t boxrel = box - (uint32_t)bx0;
            SrcTile T;
            T.s = s_dyn + (box - smem_u32(s_dyn)); T.sy0 = sy0; T.bx0 = bx0; T.R = p.boxR; T.rowbytes = p.boxW;
            auto one = [&](int r) -> uint32_t {
                const double X0 = s_rowb[r][0], Y0 = s_rowb[r][1], W0 = s_rowb[r][2];
                int X, Y;
                if constexpr (FAST) quantize_fast_t<AFF>(p.M, wr, X0, Y0, W0, mx, my, mw, X, Y);
                else quantize(p.M, X0, Y0, W0, mx, my, mw, X, Y);
                if constexpr (FAST && INTERP == 1) return sample_linear_box(boxrel, p.boxW, sy0, X, Y);
                else return sample_c3_tile<INTERP>(FAST, T, p.src, p.spitch, p.h, p.w, X, Y, p.cubic, p.lut);
            };
            int r = warp >> 1;
#pragma unroll 1
            for (; r + 4 < rows; r += 8) {
                uint32_t pa = 0, pb = 0;
                if (xvalid) { pa = one(r); pb = one(r + 4); }
                s_px[r][cx] = pa; s_px[r + 4][cx] = pb;
            }
            if (r < rows) s_px[r][cx] = xvalid ? one(r) : 0u;
        };
        if (staged) {
            if (p.M.affine) body(std::true_type{}, std::true_type{});
            else body(std::true_type{}, std::false_type{});
        } else {
            body(std::false_type{}, std::false_type{});
        }
    }
    __syncthreads();
    // 16 pixels = 48 bytes = three 16-byte vectors per thread
    if (tid < TH * 4) {
        const int r = tid >> 2, seg = tid & 3, y = tileY + r, x = tileX + 16 * seg;
        if (r < rows && x < p.Wd) {
            const uint4 *q = reinterpret_cast<const uint4 *>(&s_px[r][16 * seg]);
            uint32_t o[12];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint4 v = q[g];
                o[3 * g + 0] = __byte_perm(v.x, v.y, 0x4210); o[3 * g + 1] = __byte_perm(v.y, v.z, 0x5421); o[3 * g + 2] = __byte_perm(v.z, v.w, 0x6542);
            }
            uint8_t *d = p.dst + (size_t)y * p.dpitch + (size_t)x * 3;
            const int nb = min(16, p.Wd - x) * 3;
            if (nb == 48) {
                uint4 *d4 = reinterpret_cast<uint4 *>(d);
                d4[0] = make_uint4(o[0], o[1], o[2], o[3]); d4[1] = make_uint4(o[4], o[5], o[6], o[7]); d4[2] = make_uint4(o[8], o[9], o[10], o[11]);
            } else {
                for (int b = 0; b < nb; ++b) d[b] = (uint8_t)(o[b >> 2] >> (8 * (b & 3)));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn strip_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *q = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(q);
    });
    return fn;
}

// 2-D map over a pitched byte image: dim0 bytes per row (in elements of esz bytes), dim1 rows; no swizzle, zero fill outside.
static int encode_image_map(CUtensorMap *map, const void *base, int esz, long dim0_elems, long rows, size_t pitch, int box0_elems, int box_rows)
{
    struct Entry { const void *base; int esz, box0, box1; long d0, d1; size_t pitch; CUtensorMap map; };
    static thread_local Entry cache[16];
    static thread_local int next = 0;
    for (const Entry &e : cache)
        if (e.base == base && base && e.esz == esz && e.d0 == dim0_elems && e.d1 == rows && e.pitch == pitch && e.box0 == box0_elems && e.box1 == box_rows) {
            *map = e.map;
            return MK_OK;
        }
    EncodeTiledFn fn = strip_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return MK_ERR_CUDA; }
    const cuuint64_t dims[2] = { (cuuint64_t)dim0_elems, (cuuint64_t)rows };
    const cuuint64_t strides[1] = { (cuuint64_t)pitch };
    const cuuint32_t box[2] = { (cuuint32_t)box0_elems, (cuuint32_t)box_rows };
    const cuuint32_t estr[2] = { 1, 1 };
    const CUresult r = fn(map, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(image) failed (%d)", (int)r); return MK_ERR_CUDA; }
    Entry &slot = cache[next];
    next = (next + 1) % 16;
    slot.base = base; slot.esz = esz; slot.d0 = dim0_elems; slot.d1 = rows; slot.pitch = pitch; slot.box0 = box0_elems; slot.box1 = box_rows; slot.map = *map;
    return MK_OK;
}

// [ymin, ymax] of the convex quad `pts` inside the vertical band xa <= x <= xb; false when the band misses it
static bool quad_band_rows(const double pts[4][2], double xa, double xb, double &ymin, double &ymax)
{
    ymin = 1e300; ymax = -1e300;
    auto take = [&](double y) { ymin = fmin(ymin, y); ymax = fmax(ymax, y); };
    for (int i = 0; i < 4; ++i) {
        const double *a = pts[i], *b = pts[(i + 1) & 3];
        if (a[0] >= xa && a[0] <= xb) take(a[1]);
        for (const double xc : { xa, xb }) {
            if ((a[0] - xc) * (b[0] - xc) < 0.0) {           // edge crosses the line x = xc
                const double t = (xc - a[0]) / (b[0] - a[0]);
                take(a[1] + t * (b[1] - a[1]));
            }
        }
    }
    return ymin <= ymax;
}

template <int INTERP>
static int strip_occupancy(int dyn_smem, int *occ)
{
    static std::mutex mu;
    static int configured_dev = -1, configured_smem = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    if (configured_dev != dev || dyn_smem > configured_smem) {
        const int want = std::max(dyn_smem, 48 * 1024);
        MK_CUDA_TRY(cudaFuncSetAttribute(mapper_strip_kernel<INTERP>, cudaFuncAttributeMaxDynamicSharedMemorySize, want));
        configured_dev = dev; configured_smem = want;
    }
    static int cached_dev = -1, cached_smem = -1, cached_occ = 0;       // the query costs microseconds: once per (device, shared-memory size)
    if (cached_dev != dev || cached_smem != dyn_smem) {
        MK_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cached_occ, mapper_strip_kernel<INTERP>, NTHR, (size_t)dyn_smem));
        cached_dev = dev; cached_smem = dyn_smem;
    }
    *occ = cached_occ;
    return MK_OK;
}

int launch_strip_update(const MapperParams &P, const double *H, int interp, const int32_t *rect, cudaStream_t st, bool *done)
{
    *done = false;
    static const bool disabled = [] { const char *e = getenv("MK_STRIP"); return e && e[0] == '0'; }();
    if (disabled) return MK_OK;
    const FrameDesc &F = P.f0;
    if (!F.src_aligned || F.roi) return MK_OK;
    const int tx0 = rect[0] / TW, tx1 = (rect[2] + TW - 1) / TW, ns = tx1 - tx0;
    if (ns <= 0 || ns > MAX_STRIPS) return MK_OK;
    const InvMap &M = F.M;
    // W must keep its sign over the visited rectangle (then every block maps to a convex quad and the frame quad is convex)
    {
        int sign = 0;
        for (int i = 0; i < 4; ++i) {
            const double x = (i & 1) ? rect[2] + 3.0 : rect[0] - 3.0, y = (i & 2) ? rect[3] + 3.0 : rect[1] - 3.0;
            const double W = M.m[6] * x + M.m[7] * y + M.m[8];
            const int sg = (W > 0) - (W < 0);
            if (sg == 0 || (sign && sg != sign)) return MK_OK;
            sign = sg;
        }
    }
    // forward image of the source rectangle grown by the interpolation support (same construction as footprint())
    double pts[4][2];
    {
        const double grow = (interp == MK_INTER_CUBIC) ? 3.0 : 2.0;
        const double xs[4] = { -grow, (double)F.w - 1.0 + grow, (double)F.w - 1.0 + grow, -grow };
        const double ys[4] = { -grow, -grow, (double)F.h - 1.0 + grow, (double)F.h - 1.0 + grow };
        int sign = 0;
        for (int i = 0; i < 4; ++i) {
            const double X = H[0] * xs[i] + H[1] * ys[i] + H[2], Y = H[3] * xs[i] + H[4] * ys[i] + H[5], W = H[6] * xs[i] + H[7] * ys[i] + H[8];
            const int sg = (W > 0) - (W < 0);
            if (sg == 0 || (sign && sg != sign) || !isfinite(X) || !isfinite(Y) || !isfinite(W)) return MK_OK;
            sign = sg;
            pts[i][0] = X / W; pts[i][1] = Y / W;
            if (!isfinite(pts[i][0]) || !isfinite(pts[i][1])) return MK_OK;
        }
    }
    StripParams S{};
    S.cubic = P.cubic; S.lut = P.lut; S.Hc = P.Hc; S.Wc = P.Wc; S.tx0 = tx0; S.nstrips = ns;
    S.alpha = P.alpha; S.beta = P.beta; S.magic = P.magic; S.f = F;
    S.blend_mode = (P.alpha == 0.5f && P.beta == 0.5f) ? 1 : (P.alpha == 1.0f && P.beta == 0.0f) ? 2 : (P.alpha == 0.0f && P.beta == 1.0f) ? 3 : 0;
    long nchunks = 0;
    int yhis[MAX_STRIPS];
    for (int s = 0; s < ns; ++s) {
        const double xa = (double)(tx0 + s) * TW - HALO - 3.0, xb = (double)(tx0 + s) * TW + TW + HALO + 2.0;
        double ymin, ymax;
        int ylo = 0, yhi = 0;
        if (quad_band_rows(pts, xa, xb, ymin, ymax)) {
            ylo = (int)fmax((double)rect[1], fmin((double)rect[3], floor(ymin) - 3.0));
            yhi = (int)fmax((double)rect[1], fmin((double)rect[3], ceil(ymax) + 4.0));
        }
        S.rows[s] = make_int2(ylo, (int)nchunks); yhis[s] = yhi;
        if (yhi > ylo) nchunks += (yhi - ylo + CH - 1) / CH;
    }
    S.rows[ns] = make_int2(0, (int)nchunks);
    if (nchunks <= 0) { *done = true; return MK_OK; }
    // source box: extents of the haloed 20-row block at probes over the visited region (plain double math; a block whose
    // exact rectangle does not fit samples from global memory instead, so this only has to be a good estimate)
    int boxW = 0, boxR = 0;
    {
        const int lohi = (interp == MK_INTER_CUBIC) ? 5 : 3;
        double maxc = 0, maxr = 0;
        for (int a = 0; a < 3; ++a) {
            const int s = (a == 0) ? 0 : (a == 1) ? ns / 2 : ns - 1;
            const int ylo = S.rows[s].x, yhi = yhis[s];
            if (yhi <= ylo) continue;
            for (int b = 0; b < 3; ++b) {
                const double y0 = (b == 0) ? ylo - HALO : (b == 1) ? (ylo + yhi) / 2 : std::max(ylo - HALO, yhi + HALO - BR);
                const double x0 = (double)(tx0 + s) * TW - HALO;
                double minu = 1e300, maxu = -1e300, minv = 1e300, maxv = -1e300;
                for (int c = 0; c < 4; ++c) {
                    const double x = x0 + ((c & 1) ? TW + 2 * HALO - 1 : 0), y = y0 + ((c & 2) ? BR - 1 : 0);
                    const double W = M.m[6] * x + M.m[7] * y + M.m[8];
                    const double u = (M.m[0] * x + M.m[1] * y + M.m[2]) / W, v = (M.m[3] * x + M.m[4] * y + M.m[5]) / W;
                    minu = fmin(minu, u); maxu = fmax(maxu, u); minv = fmin(minv, v); maxv = fmax(maxv, v);
                }
                if (!isfinite(minu) || !isfinite(maxu) || !isfinite(minv) || !isfinite(maxv)) return MK_OK;
                maxc = fmax(maxc, ceil(maxu - minu) + 2 + lohi); maxr = fmax(maxr, ceil(maxv - minv) + 2 + lohi);
            }
        }
        maxc += ceil(maxc / 32); maxr += ceil(maxr / 32);      // slack for blocks between the probes (perspective)
        if (maxc <= 0 || maxr <= 0 || maxc > 1000 || maxr > 256) return MK_OK;
        boxW = (((int)maxc * 3 + 3 + 15) + 15) & ~15;      // + 15: the box starts at a 16-byte boundary
        if (!(boxW & 16)) boxW += 16;                      // odd number of 16-byte units per row: the rows a warp touches fall into different banks
        boxR = (int)maxr;
    }
    int esz = 1;
    if (boxW > 256) {
        if ((F.w & 3) != 0 || boxW > 1024) return MK_OK;
        esz = 4;
    }
    const int boxbytes = boxW * boxR;
    if (OFF_SRC + NSRC * boxbytes > 160 * 1024) return MK_OK;
    S.boxW = boxW; S.boxR = boxR; S.esz_shift = esz == 4 ? 2 : 0;
    const int dyn = OFF_SRC + NSRC * ((boxbytes + 127) & ~127) + 128;
    int occ = 0;
    {
        const int rc = (interp == MK_INTER_LINEAR) ? strip_occupancy<MK_INTER_LINEAR>(dyn, &occ) : strip_occupancy<MK_INTER_CUBIC>(dyn, &occ);
        if (rc != MK_OK) return rc;
    }
    if (occ <= 0) return MK_OK;
    const long gmax = (long)stream_sm_count(st) * std::min(occ, STRIP_OCC);
    const long per = (nchunks + gmax - 1) / gmax;
    if (per > MAX_CHUNKS) return MK_OK;
    const long grid = (nchunks + per - 1) / per;
    S.nchunks = (int)nchunks; S.chunks_per_cta = (int)per;

    alignas(64) CUtensorMap map_src, map_cv, map_mk;
    int rc = encode_image_map(&map_src, F.src, esz, (long)F.w * 3 / esz, F.h, (size_t)F.spitch, boxW / esz, boxR);
    if (rc == MK_OK) rc = encode_image_map(&map_cv, P.canvas, 1, (long)P.Wc * 3, P.Hc, (size_t)P.cpitch, ROWB, CH);
    if (rc == MK_OK) rc = encode_image_map(&map_mk, P.cmask, 1, (long)P.Wc, P.Hc, (size_t)P.mpitch, TW, CH);
    if (rc != MK_OK) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(NTHR); cfg.dynamicSmemBytes = (size_t)dyn; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // see griddepcontrol in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("MK_NO_PDL") != nullptr;          // diagnosis switch: plain stream-ordered launch
    cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
    if (interp == MK_INTER_LINEAR) MK_CUDA_TRY(cudaLaunchKernelEx(&cfg, mapper_strip_kernel<MK_INTER_LINEAR>, map_src, map_cv, map_mk, S));
    else MK_CUDA_TRY(cudaLaunchKernelEx(&cfg, mapper_strip_kernel<MK_INTER_CUBIC>, map_src, map_cv, map_mk, S));
    MK_LAUNCH_CHECK("mapper_strip_kernel");
    *done = true;
    return MK_OK;
}

// Staged form of mk_warp_perspective_u8 for 3-channel images whose source is 16-byte aligned; *done = false -> use warp_kernel.
int launch_warp_staged(const uint8_t *src, int h, int w, size_t spitch, const InvMap &M, uint8_t *dst, int Hd, int Wd, size_t dpitch, int interp,
                       const short *cubic, const uint2 *lut, cudaStream_t st, bool *done)
{
    *done = false;
    static const bool disabled = [] { const char *e = getenv("MK_STRIP"); return e && e[0] == '0'; }();
    if (disabled || !aligned(src, 16) || (spitch & 15)) return MK_OK;
    const int lohi = (interp == MK_INTER_CUBIC) ? 5 : 3;
    double maxc = 0, maxr = 0;
    int sign = 0;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            const double x0 = (a == 0) ? 0 : (a == 1) ? (Wd / 2 / TW) * TW : ((Wd - 1) / TW) * TW, y0 = (b == 0) ? 0 : (b == 1) ? (Hd / 2 / TH) * TH : ((Hd - 1) / TH) * TH;
            double minu = 1e300, maxu = -1e300, minv = 1e300, maxv = -1e300;
            for (int c = 0; c < 4; ++c) {
                const double x = x0 + ((c & 1) ? TW - 1 : 0), y = y0 + ((c & 2) ? TH - 1 : 0);
                const double W = M.m[6] * x + M.m[7] * y + M.m[8];
                const int sg = (W > 0) - (W < 0);
                if (sg == 0 || (sign && sg != sign)) return MK_OK;
                sign = sg;
                const double u = (M.m[0] * x + M.m[1] * y + M.m[2]) / W, v = (M.m[3] * x + M.m[4] * y + M.m[5]) / W;
                minu = fmin(minu, u); maxu = fmax(maxu, u); minv = fmin(minv, v); maxv = fmax(maxv, v);
            }
            if (!isfinite(minu) || !isfinite(maxu) || !isfinite(minv) || !isfinite(maxv)) return MK_OK;
            maxc = fmax(maxc, ceil(maxu - minu) + 2 + lohi); maxr = fmax(maxr, ceil(maxv - minv) + 2 + lohi);
        }
    maxc += ceil(maxc / 32); maxr += ceil(maxr / 32);
    if (maxc <= 0 || maxr <= 0 || maxc > 1000 || maxr > 256) return MK_OK;
    int boxW = (((int)maxc * 3 + 3 + 15) + 15) & ~15;
    if (!(boxW & 16)) boxW += 16;          // odd number of 16-byte units per row (shared-memory banks)
    const int boxR = (int)maxr;
    int esz = 1;
    if (boxW > 256) {
        if ((w & 3) != 0 || boxW > 1024) return MK_OK;
        esz = 4;
    }
    const int dyn = ((boxW * boxR + 127) & ~127) + 128;
    if (dyn > 64 * 1024) return MK_OK;
    {
        static std::mutex mu;
        static int conf_dev[2] = { -1, -1 }, conf_smem[2] = { 0, 0 };
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(mu);
        const int k = interp == MK_INTER_LINEAR ? 0 : 1;
        if (conf_dev[k] != dev || dyn > conf_smem[k]) {
            const int want = std::max(dyn, 48 * 1024);
            if (k == 0) MK_CUDA_TRY(cudaFuncSetAttribute(warp_staged_kernel<MK_INTER_LINEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, want));
            else MK_CUDA_TRY(cudaFuncSetAttribute(warp_staged_kernel<MK_INTER_CUBIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, want));
            conf_dev[k] = dev; conf_smem[k] = want;
        }
    }
    WarpStagedParams P{};
    P.src = src; P.dst = dst; P.cubic = cubic; P.lut = lut; P.spitch = spitch; P.dpitch = dpitch; P.h = h; P.w = w; P.Hd = Hd; P.Wd = Wd;
    P.boxW = boxW; P.boxR = boxR; P.esz_shift = esz == 4 ? 2 : 0; P.M = M;
    alignas(64) CUtensorMap map_src;
    const int rc = encode_image_map(&map_src, src, esz, (long)w * 3 / esz, h, spitch, boxW / esz, boxR);
    if (rc != MK_OK) return rc;
    dim3 grid((Wd + TW - 1) / TW, (Hd + TH - 1) / TH);
    if (interp == MK_INTER_LINEAR) warp_staged_kernel<MK_INTER_LINEAR><<<grid, 256, (size_t)dyn, st>>>(map_src, P);
    else warp_staged_kernel<MK_INTER_CUBIC><<<grid, 256, (size_t)dyn, st>>>(map_src, P);
    MK_LAUNCH_CHECK("warp_staged_kernel");
    *done = true;
    return MK_OK;
}

}  // namespace mk
