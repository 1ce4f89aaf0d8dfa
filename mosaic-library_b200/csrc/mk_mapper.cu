// mk_mapper.cu -- inverse-homography warp (K3) and the fused warp + erode + blend + mask
// accumulation pass (K4) for sm_100a.
//
// K4 is one Mapper.update step (src/mosaicking/mosaic.py:159-172 -> _update_cpu :111-138) done in
// a single pass over the warped footprint.  Per canvas pixel p:
//     warped(p) = cv2.warpPerspective(...)                          mosaic.py:114
//     m0(p)     = any_c(warped(p,c) != 0)  | warped ROI value       mosaic.py:116 | :118
//     m(p)      = 5x5 erosion of m0, outside the canvas counts as set   mosaic.py:120-121
//     old(p)    = canvas_mask(p) > 1                                mosaic.py:123
//     m & !old : canvas = warped                                    mosaic.py:127,130
//     m &  old : canvas = sat_u8(rint(fmaf(canvas, a32, warped*b32)))   mosaic.py:133,135
//     canvas_mask |= m                                              mosaic.py:138
//
// Work decomposition: the canvas is cut into 64 x 32 tiles aligned to OpenCV's 64-px coordinate
// blocks; one CTA owns one tile (single writer per canvas byte).  Phase 1 warps the tile plus a
// 2-px halo (the halo only feeds the erosion), lanes = 32 consecutive columns so a ballot yields
// the validity bit-rows; the erosion is done on 64-bit rows; phase 2 re-maps threads to groups of
// 4 pixels = 3 aligned words of BGR so that canvas I/O is whole 16-byte vectors (cp.async in,
// st.global.v4 out) and only rows that changed are written back.
#include "mk_common.cuh"
#include "mk_sample.cuh"

#include <math.h>

#include <algorithm>
#include <array>
#include <mutex>
#include <type_traits>
#include <vector>

#include "mk_mapper_common.cuh"

namespace mk {


template <int INTERP, int BLEND, bool ROI, bool BATCH>
__global__ void __launch_bounds__(MTHREADS, (INTERP == 1) ? 5 : 3) mapper_update_kernel(const __grid_constant__ MapperParams p)
{
    __shared__ __align__(16) uint8_t s_canvas[TH][ROWB];
    __shared__ __align__(16) uint32_t s_warp[TH][TW];
    __shared__ __align__(16) uint8_t s_mask[TH][TW];      // canvas-mask tile, staged with the canvas rows
    __shared__ double s_row[3][EH][3];
    __shared__ unsigned long long s_m0[EH];
    __shared__ unsigned long long s_h[EH];
    __shared__ unsigned long long s_v[TH];
    __shared__ __align__(8) unsigned long long s_bar;     // source rectangle staged
    __shared__ __align__(8) unsigned long long s_bar_c;   // canvas tile staged (needed from phase 2 on only)
    __shared__ uint8_t s_side[EH][4];
    __shared__ int s_corner[4][3];
    // ROI mode keeps the (non-binary) warped ROI values: erosion = 5x5 minimum
    __shared__ uint8_t s_mval[ROI ? EH : 1][ROI ? TW + 2 * HALO + 4 : 1];
    __shared__ uint8_t s_mh[ROI ? EH : 1][ROI ? TW : 1];
    __shared__ float s_wgt[BLEND == MK_BLEND_FEATHER ? TH : 1][BLEND == MK_BLEND_FEATHER ? TW : 1];
    extern __shared__ __align__(16) uint8_t s_src[];

    __shared__ FrameDesc s_frame;
    __shared__ unsigned s_rowchg;   // bit r: canvas row r of the tile changed (write-back mask)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int tileX, tileY, nfr = 1, list0 = 0;
    if constexpr (BATCH) {
        tileX = p.tile_xy[2 * blockIdx.x] * TW; tileY = p.tile_xy[2 * blockIdx.x + 1] * TH;
        list0 = p.tile_ptr[blockIdx.x]; nfr = p.tile_ptr[blockIdx.x + 1] - list0;
    } else {
        tileX = (p.tx0 + (int)blockIdx.x) * TW; tileY = (p.ty0 + (int)blockIdx.y) * TH;
    }
    const int nvec = min(ROWV, ((p.Wc - tileX) * 3 + 15) >> 4);  // 16-byte vectors holding canvas pixels
    const int mbytes = min(TW, (p.Wc - tileX + 15) & ~15);       // mask bytes of a tile row (pitch is a multiple of 16, >= Wc)
    const FrameDesc &F = *(BATCH ? &s_frame : &p.f0);
    if (tid == 160) { mbar_init(&s_bar, 1); mbar_init(&s_bar_c, 1); s_rowchg = 0u; }
    bool canvas_loaded = false;
    uint32_t parity = 0;

#pragma unroll 1
    for (int fi = 0; fi < nfr; ++fi) {
    if constexpr (BATCH) {
        __syncthreads();            // everyone is done with the previous frame's shared state
        if (tid < (int)(sizeof(FrameDesc) / 4))
            reinterpret_cast<uint32_t *>(&s_frame)[tid] = reinterpret_cast<const uint32_t *>(p.frames + p.tile_frame[list0 + fi])[tid];
        __syncthreads();
    }
    // ---- alpha == 1 ("keep the canvas where it is already set", the reference CLI default, utils.py:124): a pixel whose
    //      mask byte is 255 can change neither its colour (fmaf(c, 1, w * 0) == c) nor its mask (255 | m == 255), so a tile
    //      whose in-canvas mask bytes are all 255 is finished for good and every later frame skips it ------------------
    if constexpr (BLEND == MK_BLEND_ALPHA) {
        if (p.alpha == 1.0f && p.beta == 0.0f) {
            bool full = true;
            if (tid < TH * (TW / 16)) {
                const int y = tileY + (tid >> 2), x0 = tileX + 16 * (tid & 3);
                if (y < p.Hc && x0 < p.Wc) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(p.cmask + (size_t)y * p.mpitch + x0);
                    if (x0 + 16 <= p.Wc) {
                        full = (v.x & v.y & v.z & v.w) == 0xFFFFFFFFu;
                    } else {
                        const uint32_t wds[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
                        for (int b = 0; b < 16; ++b)
                            if (x0 + b < p.Wc) full = full && (((wds[b >> 2] >> (8 * (b & 3))) & 255u) == 255u);
                    }
                }
            }
            if (__syncthreads_and(full)) continue;
        }
    }
    // ---- per-row coordinate bases for the left / centre / right 64-px blocks, and the source
    //      positions of the four corners of the haloed tile ---------------------------------------
    if (tid < 3 * EH) {
        const int b = tid / EH, r = tid - b * EH;
        double X0, Y0, W0;
        row_base(F.M, tileX + (b - 1) * BLOCK_W, tileY - HALO + r, X0, Y0, W0);
        s_row[b][r][0] = X0; s_row[b][r][1] = Y0; s_row[b][r][2] = W0;
        if (b == 0) s_m0[r] = ~0ull;   // rows / columns outside the canvas count as set (cv2.erode's default border)
    } else if (tid >= 128 && tid < 132) {
        const int c = tid - 128;
        const int x0 = (c & 1) ? tileX + BLOCK_W : tileX - BLOCK_W, x1 = (c & 1) ? HALO - 1 : BLOCK_W - HALO;
        const int y = (c & 2) ? tileY + TH + HALO - 1 : tileY - HALO;
        double X0, Y0, W0;
        row_base(F.M, x0, y, X0, Y0, W0);
        const double dx1 = (double)x1;
        int X, Y;
        quantize(F.M, X0, Y0, W0, dmul(F.M.m[0], dx1), dmul(F.M.m[3], dx1), dmul(F.M.m[6], dx1), X, Y);
        const double W = F.M.affine ? F.M.m[8] : dadd(W0, dmul(F.M.m[6], dx1));
        const bool sane = abs(X) < (1 << 30) && abs(Y) < (1 << 30);   // far from cvt saturation
        s_corner[c][0] = X; s_corner[c][1] = Y; s_corner[c][2] = !sane ? 0 : ((W > 0.0) ? 1 : ((W < 0.0) ? -1 : 0));
    }
    __syncthreads();

    // ---- source rectangle of this tile: the quantised corner images bound every pixel of the tile
    //      (a homography maps the rectangle to a convex quad as long as W keeps its sign) ----------
    SrcTile T;
    bool use_smem = false;
    {
        const int sg = s_corner[0][2];
        const bool convex = sg != 0 && s_corner[1][2] == sg && s_corner[2][2] == sg && s_corner[3][2] == sg;
        int sx_lo = min(min(s_corner[0][0], s_corner[1][0]), min(s_corner[2][0], s_corner[3][0])) >> 5;
        int sx_hi = max(max(s_corner[0][0], s_corner[1][0]), max(s_corner[2][0], s_corner[3][0])) >> 5;
        int sy_lo = min(min(s_corner[0][1], s_corner[1][1]), min(s_corner[2][1], s_corner[3][1])) >> 5;
        int sy_hi = max(max(s_corner[0][1], s_corner[1][1]), max(s_corner[2][1], s_corner[3][1])) >> 5;
        constexpr int LO = (INTERP == 1) ? 1 : 2, HI = (INTERP == 1) ? 2 : 3;  // tap reach + 1 px safety ring
        sx_lo -= LO; sy_lo -= LO; sx_hi += HI; sy_hi += HI;
        if (convex && (sx_hi < 0 || sy_hi < 0 || sx_lo > F.w - 1 || sy_lo > F.h - 1)) continue;  // tile misses this frame
        // the rectangle is NOT clamped to the frame: the part outside is zero-filled below, so that the fast sampler can
        // skip all containment tests (it is tile-sized by construction; the budget test below still guards the size)
        T.s = s_src; T.sy0 = sy_lo;
        T.bx0 = ((sx_lo * 3) >> 4) << 4;
        T.rowbytes = ((sx_hi * 3 + 3 - T.bx0) + 15) & ~15;
        T.R = sy_hi - sy_lo + 1;
        use_smem = convex && F.src_aligned && T.R > 1 && T.R < 4096 && T.rowbytes > 0 && T.rowbytes < 16384 &&
                   T.R * T.rowbytes <= p.src_smem_bytes;
    }

    // ---- stage the canvas tile and the source rectangle: one bulk copy per row (async proxy),
    //      completion on one mbarrier; the parts of the rectangle outside the frame are zeroed by
    //      ordinary stores (BORDER_CONSTANT 0), which only happens for tiles on the frame's edge ----
    {
        // bytes [c_lo, c_hi) of every staged row come from the frame by bulk copy
        const int row_end16 = (F.w * 3) & ~15;
        const int c_lo = use_smem ? max(0, -T.bx0) : 0;
        const int c_hi = use_smem ? max(c_lo, min(T.rowbytes, row_end16 - T.bx0)) : 0;
        const int r_lo = use_smem ? max(0, -T.sy0) : 0, r_hi = use_smem ? min(T.R, F.h - T.sy0) : 0;   // rows inside the frame
        // The source rectangle goes first and alone: phase 1 needs nothing else, and when a whole wave of CTAs starts at once
        // the HBM burst they wait for is 40 % shorter.  The canvas rows are requested when phase 1 begins (below).
        if (warp == 0) {
            if (c_hi > c_lo)
                for (int r = r_lo + lane; r < r_hi; r += 32)
                    bulk_g2s(s_src + r * T.rowbytes + c_lo, F.src + (size_t)(T.sy0 + r) * F.spitch + (T.bx0 + c_lo),
                             (uint32_t)(c_hi - c_lo), &s_bar);
        } else if (tid == 64) {
            mbar_arrive_expect_tx(&s_bar, (c_hi > c_lo && r_hi > r_lo) ? (uint32_t)(r_hi - r_lo) * (uint32_t)(c_hi - c_lo) : 0u);
        }
        if (use_smem && (c_lo > 0 || c_hi < T.rowbytes || r_lo > 0 || r_hi < T.R)) {
            const int row_end = F.w * 3, nch = T.rowbytes >> 4;
            for (int i = tid; i < T.R * nch; i += MTHREADS) {
                const int r = i / nch, c = (i - r * nch) << 4;
                if (r >= r_lo && r < r_hi && c >= c_lo && c < c_hi) continue;   // covered by the bulk copy
                const int gy = T.sy0 + r, gb = T.bx0 + c;
                uint4 z = make_uint4(0u, 0u, 0u, 0u);
                if (gy >= 0 && gy < F.h && gb + 15 >= 0 && gb < row_end) {       // chunk straddles the frame edge
                    uint8_t *zb = reinterpret_cast<uint8_t *>(&z);
                    const uint8_t *g = F.src + (size_t)gy * F.spitch;
#pragma unroll
                    for (int b = 0; b < 16; ++b)
                        if (gb + b >= 0 && gb + b < row_end) zb[b] = __ldg(g + gb + b);
                }
                *reinterpret_cast<uint4 *>(s_src + r * T.rowbytes + c) = z;
            }
        }
        if (warp == 0) mbar_wait(&s_bar, parity);   // one polling warp; the others park at the hardware barrier (no issue slots)
        parity ^= 1u;
        if (!canvas_loaded) {                       // first frame of this tile: fetch the canvas rows under phase 1
            if (warp == 1) {
                const int y = tileY + lane;
                if (y < p.Hc) {
                    bulk_g2s(&s_canvas[lane][0], p.canvas + (size_t)y * p.cpitch + (size_t)tileX * 3, (uint32_t)nvec * 16u, &s_bar_c);
                    bulk_g2s(&s_mask[lane][0], p.cmask + (size_t)y * p.mpitch + (size_t)tileX, (uint32_t)mbytes, &s_bar_c);
                }
            } else if (tid == 64) {
                mbar_arrive_expect_tx(&s_bar_c, (uint32_t)min(TH, p.Hc - tileY) * ((uint32_t)nvec * 16u + (uint32_t)mbytes));
            }
        }
        __syncthreads();
    }

    // ---- phase 1: warp tile + vertical halo; lanes = consecutive columns -------------------
    {
        const int cx = (warp & 1) * 32 + lane, x = tileX + cx;
        const bool xvalid = x < p.Wc;
        const double dx1 = (double)cx;
        const double mx = dmul(F.M.m[0], dx1), my = dmul(F.M.m[3], dx1), mw = dmul(F.M.m[6], dx1);
        // rows of the haloed tile that lie inside the canvas (the others keep their all-ones bit rows)
        const int r_lo = max(0, HALO - tileY), r_hi = min(EH, p.Hc - tileY + HALO);
        constexpr int RSTEP = MTHREADS / 64;   // rows advanced per pass: two warps cover one 64-px row
        int r0 = warp >> 1;
        if (r0 < r_lo) r0 += (r_lo - r0 + RSTEP - 1) / RSTEP * RSTEP;
        auto body = [&](auto fast_tag, auto affine_tag) {
            constexpr bool FAST = decltype(fast_tag)::value, AFF = decltype(affine_tag)::value;
            const double wr = F.M.wr;
#pragma unroll 1
            for (int r = r0; r < r_hi; r += RSTEP) {
                uint32_t px = 0, mv = 255;
                float wg = 0.f;
                if (xvalid) {
                    int X, Y;
                    if constexpr (FAST) quantize_fast_t<AFF>(F.M, wr, s_row[1][r][0], s_row[1][r][1], s_row[1][r][2], mx, my, mw, X, Y);
                    else quantize(F.M, s_row[1][r][0], s_row[1][r][1], s_row[1][r][2], mx, my, mw, X, Y);
                    if constexpr (FAST && INTERP == 1) px = sample_linear_c3_smem_fast(T, X, Y, p.lut);
                    else px = sample_c3_tile<INTERP>(FAST, T, F.src, F.spitch, F.h, F.w, X, Y, p.cubic, p.lut);
                    if constexpr (ROI) mv = sample<INTERP, 1>(F.roi, F.rpitch, F.h, F.w, X, Y, p.cubic);
                    else mv = px ? 255u : 0u;
                    if constexpr (BLEND == MK_BLEND_FEATHER) {
                        const float u = (float)X * 0.03125f, v = (float)Y * 0.03125f;
                        const float e = fminf(fminf(__fadd_rn(u, 0.5f), __fsub_rn(__fsub_rn((float)F.w, 0.5f), u)),
                                              fminf(__fadd_rn(v, 0.5f), __fsub_rn(__fsub_rn((float)F.h, 0.5f), v)));
                        wg = fminf(1.f, fmaxf(0.f, __fdiv_rn(e, p.ramp)));
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, mv != 0);
                if (lane == 0) reinterpret_cast<uint32_t *>(&s_m0[r])[warp & 1] = bal;
                if constexpr (ROI) s_mval[r][cx + HALO] = (uint8_t)mv;
                if (r >= HALO && r < HALO + TH) {
                    s_warp[r - HALO][cx] = px;
                    if constexpr (BLEND == MK_BLEND_FEATHER) s_wgt[r - HALO][cx] = wg;
                }
            }
        };
        if (use_smem) {
            if (F.M.affine) body(std::true_type{}, std::true_type{});
            else body(std::true_type{}, std::false_type{});
        } else {
            body(std::false_type{}, std::false_type{});
        }
        if constexpr (ROI) {   // rows outside the canvas: the erosion minimum must see "set" there
            for (int r = warp >> 1; r < EH; r += RSTEP)
                if (r < r_lo || r >= r_hi) s_mval[r][cx + HALO] = 255;
        }
    }
    // ---- phase 1b: the 2 + 2 side-halo columns (they live in the neighbouring blocks) -------
    if (tid < EH * 4) {
        const int r = tid >> 2, c = tid & 3, y = tileY - HALO + r;
        const int x = (c < 2) ? tileX - HALO + c : tileX + TW + (c - 2);
        const int b = (c < 2) ? 0 : 2, x1 = (c < 2) ? BLOCK_W - HALO + c : c - 2;
        uint32_t mv = 255;
        if (y >= 0 && y < p.Hc && x >= 0 && x < p.Wc) {
            const double dx1 = (double)x1;
            int X, Y;
            quantize(F.M, s_row[b][r][0], s_row[b][r][1], s_row[b][r][2], dmul(F.M.m[0], dx1), dmul(F.M.m[3], dx1),
                     dmul(F.M.m[6], dx1), X, Y);
            if constexpr (ROI) mv = sample<INTERP, 1>(F.roi, F.rpitch, F.h, F.w, X, Y, p.cubic);
            else mv = sample_c3_tile<INTERP>(use_smem, T, F.src, F.spitch, F.h, F.w, X, Y, p.cubic, p.lut) ? 255u : 0u;
        }
        s_side[r][c] = (uint8_t)mv;
        if constexpr (ROI) s_mval[r][(c < 2) ? c : TW + HALO + (c - 2)] = (uint8_t)mv;
    }
    __syncthreads();

    // ---- 5x5 erosion ----------------------------------------------------------------------
    if constexpr (!ROI) {
        if (tid < EH) {
            const unsigned long long c = s_m0[tid];
            const unsigned long long l2 = s_side[tid][0] ? 1ull : 0ull, l1 = s_side[tid][1] ? 1ull : 0ull;
            const unsigned long long r1 = s_side[tid][2] ? 1ull : 0ull, r2 = s_side[tid][3] ? 1ull : 0ull;
            const unsigned long long sl1 = (c << 1) | l1, sl2 = (c << 2) | (l1 << 1) | l2;
            const unsigned long long sr1 = (c >> 1) | (r1 << 63), sr2 = (c >> 2) | (r1 << 62) | (r2 << 63);
            s_h[tid] = c & sl1 & sl2 & sr1 & sr2;
        }
        __syncthreads();
        if (tid < TH) s_v[tid] = s_h[tid] & s_h[tid + 1] & s_h[tid + 2] & s_h[tid + 3] & s_h[tid + 4];
    } else {
        for (int i = tid; i < EH * TW; i += MTHREADS) {
            const int r = i / TW, c = i - r * TW;
            const uint8_t *q = &s_mval[r][c];
            s_mh[r][c] = (uint8_t)min(min(min((int)q[0], (int)q[1]), min((int)q[2], (int)q[3])), (int)q[4]);
        }
        __syncthreads();
        for (int i = tid; i < TH * TW; i += MTHREADS) {
            const int r = i / TW, c = i - r * TW;
            const int m = min(min(min((int)s_mh[r][c], (int)s_mh[r + 1][c]), min((int)s_mh[r + 2][c], (int)s_mh[r + 3][c])),
                              (int)s_mh[r + 4][c]);
            s_mval[r][c] = (uint8_t)m;   // the horizontal pass already consumed s_mval: reuse it for the result
        }
        __syncthreads();
        if (tid < TH) {
            unsigned long long bits = 0;
            for (int c = 0; c < TW; ++c) bits |= (unsigned long long)(s_mval[tid][c] != 0) << c;
            s_v[tid] = bits;
        }
    }
    if (!canvas_loaded) {
        if (warp == 0) mbar_wait(&s_bar_c, 0);      // used once per CTA: phase 0
        canvas_loaded = true;
    }
    __syncthreads();

    // ---- phase 2: blend; thread = 4 consecutive pixels = 3 aligned canvas words --------------
    const bool keep_old = (BLEND == MK_BLEND_ALPHA) && p.alpha == 1.0f && p.beta == 0.0f;   // fmaf(c,1,w*0) == c
    const bool take_new = (BLEND == MK_BLEND_ALPHA) && p.alpha == 0.0f && p.beta == 1.0f;   // fmaf(c,0,w*1) == w
#pragma unroll 1
    for (int item = tid; item < TH * (TW / 4); item += MTHREADS) {
        const int r = item >> 4, g = item & 15, y = tileY + r, x = tileX + 4 * g;
        if (y >= p.Hc || x >= p.Wc) continue;
        unsigned mb = (unsigned)(s_v[r] >> (4 * g)) & 15u;
        const int nvalid = min(4, p.Wc - x);
        mb &= (1u << nvalid) - 1u;
        if (!mb) continue;
        uint32_t *cw = reinterpret_cast<uint32_t *>(&s_canvas[r][12 * g]);
        const uint32_t c0 = cw[0], c1 = cw[1], c2 = cw[2];
        uint32_t px[4] = { c0 & 0xFFFFFFu, (c0 >> 24) | ((c1 & 0xFFFFu) << 8), (c1 >> 16) | ((c2 & 0xFFu) << 16), c2 >> 8 };
        uint32_t *mp = reinterpret_cast<uint32_t *>(p.cmask + (size_t)y * p.mpitch + x);
        uint32_t *ms = reinterpret_cast<uint32_t *>(&s_mask[r][4 * g]);
        uint32_t mk4 = *ms;
        const uint4 wv4 = *reinterpret_cast<const uint4 *>(&s_warp[r][4 * g]);
        const uint32_t wv[4] = { wv4.x, wv4.y, wv4.z, wv4.w };
        float4 wa;
        float *wp = nullptr;
        if constexpr (BLEND == MK_BLEND_FEATHER) {
            wp = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(p.wacc) + (size_t)y * p.wpitch) + x;
            wa = *reinterpret_cast<const float4 *>(wp);
        }
        bool touched = false;
        uint32_t bl[4] = { 0u, 0u, 0u, 0u };
        unsigned oldb = 0;
        if constexpr (BLEND == MK_BLEND_ALPHA) {
#pragma unroll
            for (int j = 0; j < 4; ++j) oldb |= (((mk4 >> (8 * j)) & 255u) > 1u ? 1u : 0u) << j;      // mosaic.py:123
            const unsigned need = (keep_old || take_new) ? 0u : (mb & oldb);
            if (need) {
                const uint32_t ab = __float_as_uint(p.alpha), bb = __float_as_uint(p.beta);
                const unsigned long long a2 = pack2(ab, ab), b2 = pack2(bb, bb);
                const uint32_t magic = p.magic;
                if (need & 3u) blend_pair(px[0], px[1], wv[0], wv[1], a2, b2, magic, bl[0], bl[1]);
                if (need & 12u) blend_pair(px[2], px[3], wv[2], wv[3], a2, b2, magic, bl[2], bl[3]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!((mb >> j) & 1u)) continue;
            if constexpr (BLEND == MK_BLEND_ALPHA) {
                const bool old = (oldb >> j) & 1u;
                if (!old || take_new) px[j] = wv[j];
                else if (!keep_old) px[j] = bl[j];
                if constexpr (ROI) mk4 |= (uint32_t)s_mval[r][4 * g + j] << (8 * j);
                else mk4 |= 255u << (8 * j);
                touched = true;
            } else {
                const float gw = s_wgt[r][4 * g + j];
                if (gw > 0.f) {
                    float *accp = (j == 0) ? &wa.x : (j == 1) ? &wa.y : (j == 2) ? &wa.z : &wa.w;
                    const float acc = *accp, s = __fadd_rn(acc, gw);
                    px[j] = feather_u8(px[j] & 255u, wv[j] & 255u, acc, gw, s) |
                            (feather_u8((px[j] >> 8) & 255u, (wv[j] >> 8) & 255u, acc, gw, s) << 8) |
                            (feather_u8((px[j] >> 16) & 255u, (wv[j] >> 16) & 255u, acc, gw, s) << 16);
                    *accp = s;
                    mk4 |= 255u << (8 * j);
                    touched = true;
                }
            }
        }
        if (touched) {
            cw[0] = px[0] | (px[1] << 24);
            cw[1] = (px[1] >> 8) | (px[2] << 16);
            cw[2] = (px[2] >> 16) | (px[3] << 8);
            *mp = mk4;
            *ms = mk4;      // later frames of a batched launch read the tile's mask from here
            if constexpr (BLEND == MK_BLEND_FEATHER) *reinterpret_cast<float4 *>(wp) = wa;
        }
    }
    if (tid < TH && s_v[tid] != 0ull) atomicOr(&s_rowchg, 1u << tid);
    }   // frames of this tile

    if (!canvas_loaded) return;   // no frame touched this tile
    __syncthreads();

    // ---- write back the rows that changed.  Plain 16-byte stores: a bulk store would make the CTA wait ~1.5 us until the
    //      copy engine has read the tile out of shared memory before it may exit; these retire with the issuing warp. --------
    const unsigned chg = s_rowchg;
    for (int v = tid; v < TH * ROWV; v += MTHREADS) {
        const int r = v / ROWV, c = v - r * ROWV, y = tileY + r;
        if (c < nvec && y < p.Hc && ((chg >> r) & 1u))
            *reinterpret_cast<uint4 *>(p.canvas + (size_t)y * p.cpitch + (size_t)tileX * 3 + c * 16) =
                *reinterpret_cast<const uint4 *>(&s_canvas[r][c * 16]);
    }
}

// ---------------------------------------------------------------------------------------------
// K3: plain warp (cv2.warpPerspective drop-in), CN = 1 or 3.  Same tiling; output rows are staged
// in shared memory and stored as 16-byte vectors (tail bytes of a partial vector individually).
// ---------------------------------------------------------------------------------------------
struct WarpParams {
    const uint8_t *src;
    uint8_t *dst;
    const short *cubic;
    unsigned long long spitch, dpitch;
    int h, w, Hd, Wd;
    InvMap M;
};

template <int INTERP, int CN>
__global__ void __launch_bounds__(NTHREADS) warp_kernel(const __grid_constant__ WarpParams p)
{
    __shared__ __align__(16) uint8_t s_out[TH][TW * CN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tileX = (int)blockIdx.x * TW, tileY = (int)blockIdx.y * TH;
    const int cx = (warp & 1) * 32 + lane, x = tileX + cx;
    const double dx1 = (double)cx;
    const double mx = dmul(p.M.m[0], dx1), my = dmul(p.M.m[3], dx1), mw = dmul(p.M.m[6], dx1);
#pragma unroll 1
    for (int k = 0; k < TH / 4; ++k) {
        const int r = (warp >> 1) + 4 * k, y = tileY + r;
        uint32_t px = 0;
        if (y < p.Hd && x < p.Wd) {
            double X0, Y0, W0;
            row_base(p.M, tileX, y, X0, Y0, W0);
            int X, Y;
            quantize(p.M, X0, Y0, W0, mx, my, mw, X, Y);
            px = sample<INTERP, CN>(p.src, p.spitch, p.h, p.w, X, Y, p.cubic);
        }
        if constexpr (CN == 3) {
            s_out[r][cx * 3] = (uint8_t)px; s_out[r][cx * 3 + 1] = (uint8_t)(px >> 8); s_out[r][cx * 3 + 2] = (uint8_t)(px >> 16);
        } else {
            s_out[r][cx] = (uint8_t)px;
        }
    }
    __syncthreads();
    constexpr int RV = TW * CN / 16;
    const int rowbytes = min(TW, p.Wd - tileX) * CN;
    for (int v = tid; v < TH * RV; v += NTHREADS) {
        const int r = v / RV, c = v - r * RV, y = tileY + r;
        if (y >= p.Hd || c * 16 >= rowbytes) continue;
        uint8_t *d = p.dst + (size_t)y * p.dpitch + (size_t)tileX * CN + c * 16;
        if (c * 16 + 16 <= rowbytes) {
            *reinterpret_cast<uint4 *>(d) = *reinterpret_cast<const uint4 *>(&s_out[r][c * 16]);
        } else {
            for (int b = 0; b < rowbytes - c * 16; ++b) d[b] = s_out[r][c * 16 + b];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
bool invert3x3(const double *S, double *D)
{
#define s(r, c) S[(r) * 3 + (c)]
    double d = s(0, 0) * (s(1, 1) * s(2, 2) - s(1, 2) * s(2, 1)) - s(0, 1) * (s(1, 0) * s(2, 2) - s(1, 2) * s(2, 0)) +
               s(0, 2) * (s(1, 0) * s(2, 1) - s(1, 1) * s(2, 0));
    if (d == 0.0) {
        for (int i = 0; i < 9; ++i) D[i] = 0.0;
        return false;
    }
    d = 1.0 / d;
    double t[9];
    t[0] = (s(1, 1) * s(2, 2) - s(1, 2) * s(2, 1)) * d;
    t[1] = (s(0, 2) * s(2, 1) - s(0, 1) * s(2, 2)) * d;
    t[2] = (s(0, 1) * s(1, 2) - s(0, 2) * s(1, 1)) * d;
    t[3] = (s(1, 2) * s(2, 0) - s(1, 0) * s(2, 2)) * d;
    t[4] = (s(0, 0) * s(2, 2) - s(0, 2) * s(2, 0)) * d;
    t[5] = (s(0, 2) * s(1, 0) - s(0, 0) * s(1, 2)) * d;
    t[6] = (s(1, 0) * s(2, 1) - s(1, 1) * s(2, 0)) * d;
    t[7] = (s(0, 1) * s(2, 0) - s(0, 0) * s(2, 1)) * d;
    t[8] = (s(0, 0) * s(1, 1) - s(0, 1) * s(1, 0)) * d;
#undef s
    for (int i = 0; i < 9; ++i) D[i] = t[i];
    return true;
}

static void make_invmap(const double *H, InvMap &M)
{
    invert3x3(H, M.m);
    M.affine = (M.m[6] == 0.0 && M.m[7] == 0.0 && M.m[8] != 0.0) ? 1 : 0;
    M.wr = M.affine ? 32.0 / M.m[8] : 0.0;
}

// INTER_CUBIC weight table, as cv::initInterTab2D(INTER_CUBIC, fixpt) builds it: fp32 A=-0.75
// cubic at k/32, outer product * 32768 rounded to int16, sum forced to 32768 by adjusting the
// largest / smallest tap among rows/cols [ksize/2, ksize/2+2) of the 4x4 block.
static void build_cubic_table(short *tab)
{
    float t1[32][4];
    for (int i = 0; i < 32; ++i) {
        const float A = -0.75f, x = (float)i * (1.f / 32);
        float *c = t1[i];
        c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
        c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
        c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
        c[3] = 1.f - c[0] - c[1] - c[2];
    }
    for (int i = 0; i < 32; ++i)
        for (int j = 0; j < 32; ++j) {
            short *it = tab + (i * 32 + j) * 16;
            int isum = 0;
            for (int k1 = 0; k1 < 4; ++k1)
                for (int k2 = 0; k2 < 4; ++k2) {
                    const float v = t1[i][k1] * t1[j][k2];
                    long r = lrintf(v * 32768.f);
                    r = r > 32767 ? 32767 : (r < -32768 ? -32768 : r);
                    it[k1 * 4 + k2] = (short)r;
                    isum += (int)r;
                }
            if (isum != 32768) {
                const int diff = isum - 32768;
                int Mk = 2 * 4 + 2, mk_ = 2 * 4 + 2;
                for (int k1 = 2; k1 < 4; ++k1)
                    for (int k2 = 2; k2 < 4; ++k2) {
                        const int k = k1 * 4 + k2;
                        if (it[k] < it[mk_]) mk_ = k;
                        else if (it[k] > it[Mk]) Mk = k;
                    }
                if (diff < 0) it[Mk] = (short)(it[Mk] - diff);
                else it[mk_] = (short)(it[mk_] - diff);
            }
        }
}

// Bilinear weight LUT for lerp3_dp2a: entry [ay*32+ax] = ((w00 | w01 << 16), (w10 | w11 << 16)) with the integer weights
// (32-ax)(32-ay), ax(32-ay), (32-ax)ay, ax*ay (<= 1024) as int16 pairs for IDP.2A.
static void build_lerp_lut(uint2 *lut)
{
    for (int ay = 0; ay < 32; ++ay)
        for (int ax = 0; ax < 32; ++ax) {
            const uint32_t w00 = (32 - ax) * (32 - ay), w01 = ax * (32 - ay), w10 = (32 - ax) * ay, w11 = ax * ay;
            lut[ay * 32 + ax] = make_uint2(w00 | (w01 << 16), w10 | (w11 << 16));
        }
}

// read-only tables, one copy per device, created on first use
int device_tables(DeviceTables &out)
{
    static std::mutex mu;
    static DeviceTables tabs[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return cuda_fail(cudaGetLastError(), "cudaGetDevice");
    std::lock_guard<std::mutex> lk(mu);
    if (!tabs[dev].cubic) {
        static short host_cubic[1024 * 16];
        static uint2 host_lut[1024];
        build_cubic_table(host_cubic);
        build_lerp_lut(host_lut);
        short *dc = nullptr;
        uint2 *dl = nullptr;
        cudaError_t e = cudaMalloc(&dc, sizeof host_cubic);
        if (e == cudaSuccess) e = cudaMalloc(&dl, sizeof host_lut);
        if (e == cudaSuccess) e = cudaMemcpy(dc, host_cubic, sizeof host_cubic, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(dl, host_lut, sizeof host_lut, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return cuda_fail(e, "interpolation table upload");
        tabs[dev].cubic = dc; tabs[dev].lut = dl;
    }
    out = tabs[dev];
    return MK_OK;
}

// Canvas rectangle that can receive a non-zero warped pixel: forward-map the source rectangle grown
// by the interpolation support, take the bounding box (+2 px slack), clip.  Conservative by design:
// the kernel evaluates every pixel exactly, the rectangle only bounds the work.
static void footprint(int Hc, int Wc, int h, int w, const double *H, int interp, int32_t *rect)
{
    const double grow = (interp == MK_INTER_CUBIC) ? 3.0 : 2.0;
    const double xs[2] = { -grow, (double)w - 1.0 + grow }, ys[2] = { -grow, (double)h - 1.0 + grow };
    double minx = 1e300, miny = 1e300, maxx = -1e300, maxy = -1e300;
    bool ok = true;
    double M[9];
    if (!invert3x3(H, M)) ok = false;
    int sign = 0;
    for (int i = 0; i < 2 && ok; ++i)
        for (int j = 0; j < 2 && ok; ++j) {
            const double X = H[0] * xs[i] + H[1] * ys[j] + H[2], Y = H[3] * xs[i] + H[4] * ys[j] + H[5],
                         W = H[6] * xs[i] + H[7] * ys[j] + H[8];
            const int sg = (W > 0) - (W < 0);
            if (sg == 0 || (sign != 0 && sg != sign) || !isfinite(X) || !isfinite(Y) || !isfinite(W)) { ok = false; break; }
            sign = sg;
            const double px = X / W, py = Y / W;
            if (!isfinite(px) || !isfinite(py)) { ok = false; break; }
            minx = fmin(minx, px); maxx = fmax(maxx, px); miny = fmin(miny, py); maxy = fmax(maxy, py);
        }
    if (!ok) { rect[0] = 0; rect[1] = 0; rect[2] = Wc; rect[3] = Hc; return; }
    const double lo_x = floor(minx) - 2.0, lo_y = floor(miny) - 2.0, hi_x = ceil(maxx) + 3.0, hi_y = ceil(maxy) + 3.0;
    rect[0] = (int)fmax(0.0, fmin((double)Wc, lo_x));
    rect[1] = (int)fmax(0.0, fmin((double)Hc, lo_y));
    rect[2] = (int)fmax(0.0, fmin((double)Wc, hi_x));
    rect[3] = (int)fmax(0.0, fmin((double)Hc, hi_y));
    if (rect[2] <= rect[0] || rect[3] <= rect[1]) { rect[0] = rect[1] = rect[2] = rect[3] = 0; }
}

constexpr int MAX_SRC_SMEM = 96 * 1024;

template <int INTERP, int BLEND, bool ROI, bool BATCH>
static cudaError_t launch_mapper_inst(dim3 grid, cudaStream_t st, const MapperParams &P)
{
    static std::atomic<int> configured{-1};  // device the attribute was last set on
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured.load(std::memory_order_acquire) != dev) {
        cudaError_t e = cudaFuncSetAttribute(mapper_update_kernel<INTERP, BLEND, ROI, BATCH>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SRC_SMEM);
        if (e != cudaSuccess) return e;
        configured.store(dev, std::memory_order_release);
    }
    mapper_update_kernel<INTERP, BLEND, ROI, BATCH><<<grid, MTHREADS, (size_t)P.src_smem_bytes, st>>>(P);
    return cudaSuccess;
}

template <int INTERP, int BLEND>
static cudaError_t launch_mapper(bool roi, bool batch, dim3 grid, cudaStream_t st, const MapperParams &P)
{
    if (batch) return launch_mapper_inst<INTERP, BLEND, false, true>(grid, st, P);
    if constexpr (BLEND == MK_BLEND_ALPHA) {
        if (roi) return launch_mapper_inst<INTERP, BLEND, true, false>(grid, st, P);
    }
    return launch_mapper_inst<INTERP, BLEND, false, false>(grid, st, P);
}

static cudaError_t launch_mapper_any(int interp, int blend, bool roi, bool batch, dim3 grid, cudaStream_t st, const MapperParams &P)
{
    if (interp == MK_INTER_LINEAR) {
        if (blend == MK_BLEND_ALPHA) return launch_mapper<MK_INTER_LINEAR, MK_BLEND_ALPHA>(roi, batch, grid, st, P);
        return launch_mapper<MK_INTER_LINEAR, MK_BLEND_FEATHER>(false, batch, grid, st, P);
    }
    if (blend == MK_BLEND_ALPHA) return launch_mapper<MK_INTER_CUBIC, MK_BLEND_ALPHA>(roi, batch, grid, st, P);
    return launch_mapper<MK_INTER_CUBIC, MK_BLEND_FEATHER>(false, batch, grid, st, P);
}

// Shared memory needed to stage the source rectangle of one haloed tile: evaluated (in plain double
// math) for tiles at the corners and the centre of the visited region; each CTA re-derives its own
// exact rectangle and falls back to global-memory sampling if it does not fit.
static int estimate_src_smem(const InvMap &M, const int32_t *rect, int interp)
{
    const int lohi = (interp == MK_INTER_CUBIC) ? 5 : 3;
    const double txs[3] = { (double)(rect[0] / TW * TW), (double)((rect[0] + rect[2]) / 2 / TW * TW), (double)((rect[2] - 1) / TW * TW) };
    const double tys[3] = { (double)(rect[1] / TH * TH), (double)((rect[1] + rect[3]) / 2 / TH * TH), (double)((rect[3] - 1) / TH * TH) };
    long need = 0;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            double minu = 1e300, maxu = -1e300, minv = 1e300, maxv = -1e300;
            bool ok = true;
            int sign = 0;
            for (int c = 0; c < 4 && ok; ++c) {
                const double x = txs[a] + ((c & 1) ? TW + HALO - 1 : -HALO), y = tys[b] + ((c & 2) ? TH + HALO - 1 : -HALO);
                const double W = M.m[6] * x + M.m[7] * y + M.m[8];
                const int sg = (W > 0) - (W < 0);
                if (sg == 0 || (sign && sg != sign)) { ok = false; break; }
                sign = sg;
                const double u = (M.m[0] * x + M.m[1] * y + M.m[2]) / W, v = (M.m[3] * x + M.m[4] * y + M.m[5]) / W;
                if (!isfinite(u) || !isfinite(v)) { ok = false; break; }
                minu = fmin(minu, u); maxu = fmax(maxu, u); minv = fmin(minv, v); maxv = fmax(maxv, v);
            }
            if (!ok) continue;
            const double rows = ceil(maxv - minv) + 2 + lohi, cols = ceil(maxu - minu) + 2 + lohi;
            if (rows > 4096 || cols > 4096) continue;
            const long rowbytes = (((long)cols * 3 + 15 + 15) / 16) * 16;
            const long bytes = (long)rows * rowbytes;
            if (bytes > need) need = bytes;
        }
    need += need / 16;  // slack for tiles between the probes (perspective)
    need = (need + 1023) / 1024 * 1024;
    return need > MAX_SRC_SMEM ? 0 : (int)need;
}

}  // namespace mk

using namespace mk;

extern "C" int mk_mapper_footprint(int canvas_h, int canvas_w, int src_h, int src_w, const double *H_host, int interp,
                                   int32_t *region_host)
{
    if (!H_host || !region_host || canvas_h <= 0 || canvas_w <= 0 || src_h <= 0 || src_w <= 0) {
        set_error("mk_mapper_footprint: bad argument");
        return MK_ERR_ARG;
    }
    footprint(canvas_h, canvas_w, src_h, src_w, H_host, interp, region_host);
    return MK_OK;
}

extern "C" int mk_mapper_update(uint8_t *canvas, size_t canvas_pitch, uint8_t *cmask, size_t cmask_pitch, float *wacc,
                                size_t wacc_pitch, int canvas_h, int canvas_w, const uint8_t *src, int src_h, int src_w,
                                size_t src_pitch, const uint8_t *roi, size_t roi_pitch, const double *H_host, double alpha,
                                int interp, int blend, float feather_ramp, int32_t *region_host, void *stream)
{
    if (!canvas || !cmask || !src || !H_host || canvas_h <= 0 || canvas_w <= 0 || src_h <= 0 || src_w <= 0) {
        set_error("mk_mapper_update: null pointer or empty shape");
        return MK_ERR_ARG;
    }
    if (src_h > 32767 || src_w > 32767) {
        set_error("mk_mapper_update: source larger than 32767 px (cv2 remap uses short coordinates)");
        return MK_ERR_ARG;
    }
    if (interp != MK_INTER_LINEAR && interp != MK_INTER_CUBIC) { set_error("mk_mapper_update: bad interp %d", interp); return MK_ERR_ARG; }
    if (blend != MK_BLEND_ALPHA && blend != MK_BLEND_FEATHER) { set_error("mk_mapper_update: bad blend %d", blend); return MK_ERR_ARG; }
    if (blend == MK_BLEND_ALPHA && !(alpha >= 0.0 && alpha <= 1.0)) { set_error("mk_mapper_update: alpha outside [0,1]"); return MK_ERR_ARG; }
    if (blend == MK_BLEND_FEATHER && (!wacc || !(feather_ramp > 0.f) || roi)) {
        set_error("mk_mapper_update: feather mode needs wacc, ramp > 0 and no ROI");
        return MK_ERR_ARG;
    }
    if (canvas_pitch < (size_t)canvas_w * 3 || cmask_pitch < (size_t)canvas_w || src_pitch < (size_t)src_w * 3 ||
        (roi && roi_pitch < (size_t)src_w) || (blend == MK_BLEND_FEATHER && wacc_pitch < (size_t)canvas_w * 4)) {
        set_error("mk_mapper_update: pitch smaller than a row");
        return MK_ERR_ARG;
    }
    if (!aligned(canvas, 16) || !aligned(cmask, 16) || (canvas_pitch & 15) || (cmask_pitch & 15) || !aligned(src, 4) ||
        (src_pitch & 3) || (roi && (!aligned(roi, 4) || (roi_pitch & 3))) ||
        (blend == MK_BLEND_FEATHER && (!aligned(wacc, 16) || (wacc_pitch & 15)))) {
        set_error("mk_mapper_update: alignment (canvas/cmask/wacc 16 B + pitch %% 16, src/roi 4 B + pitch %% 4)");
        return MK_ERR_ALIGN;
    }
    int32_t rect[4];
    footprint(canvas_h, canvas_w, src_h, src_w, H_host, interp, rect);
    if (region_host) { region_host[0] = rect[0]; region_host[1] = rect[1]; region_host[2] = rect[2]; region_host[3] = rect[3]; }
    if (rect[2] <= rect[0]) return MK_OK;  // footprint misses the canvas: nothing to do

    MapperParams P{};
    P.canvas = canvas; P.cmask = cmask; P.wacc = wacc;
    P.cpitch = canvas_pitch; P.mpitch = cmask_pitch; P.wpitch = wacc_pitch;
    P.Hc = canvas_h; P.Wc = canvas_w;
    P.f0.src = src; P.f0.roi = roi; P.f0.spitch = src_pitch; P.f0.rpitch = roi_pitch; P.f0.h = src_h; P.f0.w = src_w;
    make_invmap(H_host, P.f0.M);
    P.alpha = (float)alpha; P.beta = (float)(1.0 - alpha); P.ramp = feather_ramp; P.magic = 0x4B000000u;
    {
        DeviceTables tb;
        const int st = device_tables(tb);
        if (st != MK_OK) return st;
        P.cubic = tb.cubic; P.lut = tb.lut;
    }
    P.tx0 = rect[0] / TW; P.ty0 = rect[1] / TH;
    const int tx1 = (rect[2] + TW - 1) / TW, ty1 = (rect[3] + TH - 1) / TH;
    dim3 grid(tx1 - P.tx0, ty1 - P.ty0);
    P.f0.src_aligned = (aligned(src, 16) && (src_pitch & 15) == 0) ? 1 : 0;
    P.src_smem_bytes = P.f0.src_aligned ? estimate_src_smem(P.f0.M, rect, interp) : 0;
    if (blend == MK_BLEND_ALPHA && !roi) {      // persistent strip kernel (mk_strip.cu) whenever it applies
        bool done = false;
        const int rc = launch_strip_update(P, H_host, interp, rect, (cudaStream_t)stream, &done);
        if (rc != MK_OK) return rc;
        if (done) return MK_OK;
    }
    const cudaError_t le = launch_mapper_any(interp, blend, roi != nullptr, false, grid, (cudaStream_t)stream, P);
    if (le != cudaSuccess) return cuda_fail(le, "cudaFuncSetAttribute(mapper_update_kernel)");
    MK_LAUNCH_CHECK("mapper_update_kernel");
    return MK_OK;
}

extern "C" int mk_mapper_update_batch(uint8_t *canvas, size_t canvas_pitch, uint8_t *cmask, size_t cmask_pitch, float *wacc,
                                      size_t wacc_pitch, int canvas_h, int canvas_w, int nframes, const uint8_t *const *src_host,
                                      const int *src_h_host, const int *src_w_host, const size_t *src_pitch_host,
                                      const double *H_host, double alpha, int interp, int blend, float feather_ramp, void *stream)
{
    if (!canvas || !cmask || nframes < 0 || canvas_h <= 0 || canvas_w <= 0 || (nframes && (!src_host || !src_h_host || !src_w_host || !src_pitch_host || !H_host))) {
        set_error("mk_mapper_update_batch: null pointer or empty shape");
        return MK_ERR_ARG;
    }
    if (nframes == 0) return MK_OK;
    if (interp != MK_INTER_LINEAR && interp != MK_INTER_CUBIC) { set_error("mk_mapper_update_batch: bad interp %d", interp); return MK_ERR_ARG; }
    if (blend != MK_BLEND_ALPHA && blend != MK_BLEND_FEATHER) { set_error("mk_mapper_update_batch: bad blend %d", blend); return MK_ERR_ARG; }
    if (blend == MK_BLEND_ALPHA && !(alpha >= 0.0 && alpha <= 1.0)) { set_error("mk_mapper_update_batch: alpha outside [0,1]"); return MK_ERR_ARG; }
    if (blend == MK_BLEND_FEATHER && (!wacc || !(feather_ramp > 0.f))) { set_error("mk_mapper_update_batch: feather mode needs wacc and ramp > 0"); return MK_ERR_ARG; }
    if (canvas_pitch < (size_t)canvas_w * 3 || cmask_pitch < (size_t)canvas_w || (blend == MK_BLEND_FEATHER && wacc_pitch < (size_t)canvas_w * 4)) {
        set_error("mk_mapper_update_batch: pitch smaller than a row");
        return MK_ERR_ARG;
    }
    if (!aligned(canvas, 16) || !aligned(cmask, 16) || (canvas_pitch & 15) || (cmask_pitch & 15) ||
        (blend == MK_BLEND_FEATHER && (!aligned(wacc, 16) || (wacc_pitch & 15)))) {
        set_error("mk_mapper_update_batch: alignment (canvas/cmask/wacc 16 B + pitch %% 16)");
        return MK_ERR_ALIGN;
    }
    if (blend == MK_BLEND_ALPHA) {
        // alpha blending: a train of strip-kernel launches in sequence order.  The persistent strip kernel (mk_strip.cu) is not
        // HBM bound, so reading every canvas tile once (the gather launch below) buys nothing, and back to back its launches
        // cost less per frame (18.6 us vs 23.2 us for 1080p frames, bench `batched_compositing`); frames it cannot take
        // fall back to the tile kernel one by one inside mk_mapper_update.
        for (int k = 0; k < nframes; ++k) {
            const int rc = mk_mapper_update(canvas, canvas_pitch, cmask, cmask_pitch, nullptr, 0, canvas_h, canvas_w, src_host[k], src_h_host[k], src_w_host[k],
                                            src_pitch_host[k], nullptr, 0, H_host + 9 * (size_t)k, alpha, interp, blend, feather_ramp, nullptr, stream);
            if (rc != MK_OK) return rc;
        }
        return MK_OK;
    }
    // feather: per-frame descriptors + the (tile -> ordered frame list) table
    const int ntx = (canvas_w + TW - 1) / TW, nty = (canvas_h + TH - 1) / TH;
    std::vector<FrameDesc> frames((size_t)nframes);
    std::vector<std::array<int, 4>> rects((size_t)nframes);
    int smem_need = 0;
    bool all_aligned = true;
    std::vector<int> count((size_t)ntx * nty, 0);
    for (int k = 0; k < nframes; ++k) {
        const int h = src_h_host[k], w = src_w_host[k];
        const size_t sp = src_pitch_host[k];
        if (!src_host[k] || h <= 0 || w <= 0 || h > 32767 || w > 32767 || sp < (size_t)w * 3 || !aligned(src_host[k], 4) || (sp & 3)) {
            set_error("mk_mapper_update_batch: frame %d: bad pointer, shape, pitch or alignment", k);
            return MK_ERR_ARG;
        }
        FrameDesc &f = frames[k];
        f.src = src_host[k]; f.roi = nullptr; f.spitch = sp; f.rpitch = 0; f.h = h; f.w = w; f.pad_ = 0;
        f.src_aligned = (aligned(src_host[k], 16) && (sp & 15) == 0) ? 1 : 0;
        all_aligned = all_aligned && f.src_aligned;
        make_invmap(H_host + 9 * (size_t)k, f.M);
        int32_t rect[4];
        footprint(canvas_h, canvas_w, h, w, H_host + 9 * (size_t)k, interp, rect);
        rects[k] = { rect[0], rect[1], rect[2], rect[3] };
        if (rect[2] <= rect[0]) continue;
        if (f.src_aligned) smem_need = std::max(smem_need, estimate_src_smem(f.M, rect, interp));
        for (int ty = rect[1] / TH; ty < (rect[3] + TH - 1) / TH; ++ty)
            for (int tx = rect[0] / TW; tx < (rect[2] + TW - 1) / TW; ++tx) ++count[(size_t)ty * ntx + tx];
    }
    std::vector<int> tile_xy, tile_ptr(1, 0), slot((size_t)ntx * nty, -1);
    for (int ty = 0; ty < nty; ++ty)
        for (int tx = 0; tx < ntx; ++tx) {
            const int c = count[(size_t)ty * ntx + tx];
            if (!c) continue;
            slot[(size_t)ty * ntx + tx] = tile_ptr.back();
            tile_xy.push_back(tx); tile_xy.push_back(ty);
            tile_ptr.push_back(tile_ptr.back() + c);
        }
    const int ntiles = (int)tile_xy.size() / 2;
    if (ntiles == 0) return MK_OK;
    std::vector<int> tile_frame((size_t)tile_ptr.back());
    for (int k = 0; k < nframes; ++k) {          // frames are appended in order: every tile sees them in sequence order
        const auto &r = rects[k];
        if (r[2] <= r[0]) continue;
        for (int ty = r[1] / TH; ty < (r[3] + TH - 1) / TH; ++ty)
            for (int tx = r[0] / TW; tx < (r[2] + TW - 1) / TW; ++tx) tile_frame[(size_t)slot[(size_t)ty * ntx + tx]++] = k;
    }
    // one stream-ordered device scratch: [frames][tile_ptr][tile_xy][tile_frame]
    const size_t b_frames = (sizeof(FrameDesc) * (size_t)nframes + 15) & ~(size_t)15, b_ptr = (sizeof(int) * tile_ptr.size() + 15) & ~(size_t)15,
                 b_xy = (sizeof(int) * tile_xy.size() + 15) & ~(size_t)15, b_tf = sizeof(int) * tile_frame.size();
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *scratch = nullptr;
    MK_CUDA_TRY(cudaMallocAsync(&scratch, b_frames + b_ptr + b_xy + b_tf + 16, st));
    cudaError_t ce = cudaMemcpyAsync(scratch, frames.data(), sizeof(FrameDesc) * (size_t)nframes, cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(scratch + b_frames, tile_ptr.data(), sizeof(int) * tile_ptr.size(), cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(scratch + b_frames + b_ptr, tile_xy.data(), sizeof(int) * tile_xy.size(), cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(scratch + b_frames + b_ptr + b_xy, tile_frame.data(), b_tf, cudaMemcpyHostToDevice, st);
    if (ce != cudaSuccess) { cudaFreeAsync(scratch, st); return cuda_fail(ce, "mk_mapper_update_batch: table upload"); }
    MapperParams P{};
    P.canvas = canvas; P.cmask = cmask; P.wacc = wacc;
    P.cpitch = canvas_pitch; P.mpitch = cmask_pitch; P.wpitch = wacc_pitch;
    P.Hc = canvas_h; P.Wc = canvas_w;
    P.alpha = (float)alpha; P.beta = (float)(1.0 - alpha); P.ramp = feather_ramp; P.magic = 0x4B000000u;
    {
        DeviceTables tb;
        const int rc = device_tables(tb);
        if (rc != MK_OK) { cudaFreeAsync(scratch, st); return rc; }
        P.cubic = tb.cubic; P.lut = tb.lut;
    }
    P.frames = reinterpret_cast<const FrameDesc *>(scratch);
    P.tile_ptr = reinterpret_cast<const int *>(scratch + b_frames);
    P.tile_xy = reinterpret_cast<const int *>(scratch + b_frames + b_ptr);
    P.tile_frame = reinterpret_cast<const int *>(scratch + b_frames + b_ptr + b_xy);
    P.src_smem_bytes = smem_need;
    const cudaError_t le = launch_mapper_any(interp, blend, false, true, dim3((unsigned)ntiles), st, P);
    cudaFreeAsync(scratch, st);      // stream-ordered: released after the kernel
    if (le != cudaSuccess) return cuda_fail(le, "cudaFuncSetAttribute(mapper_update_kernel)");
    MK_LAUNCH_CHECK("mapper_update_kernel(batch)");
    return MK_OK;
}

extern "C" int mk_warp_perspective_u8(const uint8_t *src, int src_h, int src_w, int channels, size_t src_pitch,
                                      const double *H_host, uint8_t *dst, int dst_h, int dst_w, size_t dst_pitch, int interp,
                                      void *stream)
{
    if (!src || !dst || !H_host || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0 || (channels != 1 && channels != 3)) {
        set_error("mk_warp_perspective_u8: bad argument (channels must be 1 or 3)");
        return MK_ERR_ARG;
    }
    if (src_h > 32767 || src_w > 32767) { set_error("mk_warp_perspective_u8: source larger than 32767 px"); return MK_ERR_ARG; }
    if (interp != MK_INTER_LINEAR && interp != MK_INTER_CUBIC) { set_error("mk_warp_perspective_u8: bad interp"); return MK_ERR_ARG; }
    if (src_pitch < (size_t)src_w * channels || dst_pitch < (size_t)dst_w * channels) { set_error("mk_warp_perspective_u8: pitch smaller than a row"); return MK_ERR_ARG; }
    if (!aligned(dst, 16) || (dst_pitch & 15) || !aligned(src, 4) || (src_pitch & 3)) {
        set_error("mk_warp_perspective_u8: alignment (dst 16 B + pitch %% 16, src 4 B + pitch %% 4)");
        return MK_ERR_ALIGN;
    }
    WarpParams P;
    P.src = src; P.dst = dst; P.cubic = nullptr; P.spitch = src_pitch; P.dpitch = dst_pitch;
    P.h = src_h; P.w = src_w; P.Hd = dst_h; P.Wd = dst_w;
    make_invmap(H_host, P.M);
    const uint2 *lut = nullptr;
    {
        DeviceTables tb;
        const int st = device_tables(tb);
        if (st != MK_OK) return st;
        P.cubic = tb.cubic; lut = tb.lut;
    }
    dim3 grid((dst_w + TW - 1) / TW, (dst_h + TH - 1) / TH);
    cudaStream_t st = (cudaStream_t)stream;
    if (channels == 3) {          // source footprint staged in shared memory by TMA (mk_strip.cu) whenever it applies
        bool done = false;
        const int rc = launch_warp_staged(src, src_h, src_w, src_pitch, P.M, dst, dst_h, dst_w, dst_pitch, interp, P.cubic, lut, st, &done);
        if (rc != MK_OK) return rc;
        if (done) return MK_OK;
    }
    if (interp == MK_INTER_LINEAR) {
        if (channels == 3) warp_kernel<MK_INTER_LINEAR, 3><<<grid, NTHREADS, 0, st>>>(P);
        else warp_kernel<MK_INTER_LINEAR, 1><<<grid, NTHREADS, 0, st>>>(P);
    } else {
        if (channels == 3) warp_kernel<MK_INTER_CUBIC, 3><<<grid, NTHREADS, 0, st>>>(P);
        else warp_kernel<MK_INTER_CUBIC, 1><<<grid, NTHREADS, 0, st>>>(P);
    }
    MK_LAUNCH_CHECK("warp_kernel");
    return MK_OK;
}
