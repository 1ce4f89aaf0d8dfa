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
#include <mutex>

namespace mk {

constexpr int TW = 64;   // tile width  (== BLOCK_W)
constexpr int TH = 32;   // tile height
constexpr int HALO = 2;  // 5x5 erosion radius
constexpr int EH = TH + 2 * HALO;
constexpr int NTHREADS = 256;
constexpr int ROWB = TW * 3;       // canvas bytes per tile row
constexpr int ROWV = ROWB / 16;    // 16-byte vectors per tile row

struct MapperParams {
    uint8_t *canvas, *cmask;
    float *wacc;
    const uint8_t *src, *roi;
    const short *cubic;
    unsigned long long cpitch, mpitch, wpitch, spitch, rpitch;
    int Hc, Wc, h, w;
    int tx0, ty0;
    InvMap M;
    float alpha, beta, ramp;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// cv2.addWeighted on uint8: both scalars narrowed to fp32, rint(fmaf(c, a, w*b)), saturate.
__device__ __forceinline__ uint32_t blend_u8(uint32_t c, uint32_t w, float a, float b)
{
    const float t = __fmaf_rn((float)c, a, __fmul_rn((float)w, b));
    return (uint32_t)min(255, max(0, __float2int_rn(t)));
}

__device__ __forceinline__ uint32_t blend_px(uint32_t c, uint32_t w, float a, float b)
{
    return blend_u8(c & 255u, w & 255u, a, b) | (blend_u8((c >> 8) & 255u, (w >> 8) & 255u, a, b) << 8) |
           (blend_u8((c >> 16) & 255u, (w >> 16) & 255u, a, b) << 16);
}

// feather: canvas = sat_u8(rint((canvas*Wacc + warped*g) / (Wacc + g)))  
__device__ __forceinline__ uint32_t feather_u8(uint32_t c, uint32_t w, float acc, float g, float s)
{
    const float t = __fdiv_rn(__fadd_rn(__fmul_rn((float)c, acc), __fmul_rn((float)w, g)), s);
    return (uint32_t)min(255, max(0, __float2int_rn(t)));
}

template <int INTERP, int BLEND, bool ROI>
__global__ void __launch_bounds__(NTHREADS) mapper_update_kernel(const __grid_constant__ MapperParams p)
{
    __shared__ __align__(16) uint8_t s_canvas[TH][ROWB];
    __shared__ uint32_t s_warp[TH][TW];
    __shared__ double s_row[3][EH][3];
    __shared__ unsigned long long s_m0[EH];
    __shared__ unsigned long long s_h[EH];
    __shared__ unsigned long long s_v[TH];
    __shared__ uint8_t s_side[EH][4];
    // ROI mode keeps the (non-binary) warped ROI values: erosion = 5x5 minimum
    __shared__ uint8_t s_mval[ROI ? EH : 1][ROI ? TW + 2 * HALO + 4 : 1];
    __shared__ uint8_t s_mh[ROI ? EH : 1][ROI ? TW : 1];
    __shared__ float s_wgt[BLEND == MK_BLEND_FEATHER ? TH : 1][BLEND == MK_BLEND_FEATHER ? TW : 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tileX = (p.tx0 + (int)blockIdx.x) * TW, tileY = (p.ty0 + (int)blockIdx.y) * TH;
    const int nvec = min(ROWV, ((p.Wc - tileX) * 3 + 15) >> 4);  // vectors holding canvas pixels

    // ---- stage the canvas tile (async, overlaps phase 1) ---------------------------------
    for (int v = tid; v < TH * ROWV; v += NTHREADS) {
        const int r = v / ROWV, c = v - r * ROWV, y = tileY + r;
        if (y < p.Hc && c < nvec)
            cp_async16(&s_canvas[r][c * 16], p.canvas + (size_t)y * p.cpitch + (size_t)tileX * 3 + c * 16);
    }

    // ---- per-row coordinate bases for the left / centre / right 64-px blocks --------------
    if (tid < 3 * EH) {
        const int b = tid / EH, r = tid - b * EH;
        double X0, Y0, W0;
        row_base(p.M, tileX + (b - 1) * BLOCK_W, tileY - HALO + r, X0, Y0, W0);
        s_row[b][r][0] = X0; s_row[b][r][1] = Y0; s_row[b][r][2] = W0;
    }
    __syncthreads();

    // ---- phase 1: warp tile + vertical halo; lanes = consecutive columns -------------------
    {
        const int cx = (warp & 1) * 32 + lane, x = tileX + cx;
        const double dx1 = (double)cx;
        const double mx = dmul(p.M.m[0], dx1), my = dmul(p.M.m[3], dx1), mw = dmul(p.M.m[6], dx1);
#pragma unroll 1
        for (int k = 0; k < EH / 4; ++k) {
            const int r = (warp >> 1) + 4 * k, y = tileY - HALO + r;
            uint32_t px = 0, mv = 255;
            float wg = 0.f;
            if (y >= 0 && y < p.Hc && x < p.Wc) {
                int X, Y;
                quantize(p.M, s_row[1][r][0], s_row[1][r][1], s_row[1][r][2], mx, my, mw, X, Y);
                px = sample<INTERP, 3>(p.src, p.spitch, p.h, p.w, X, Y, p.cubic);
                if constexpr (ROI) mv = sample<INTERP, 1>(p.roi, p.rpitch, p.h, p.w, X, Y, p.cubic);
                else mv = px ? 255u : 0u;
                if constexpr (BLEND == MK_BLEND_FEATHER) {
                    const float u = (float)X * 0.03125f, v = (float)Y * 0.03125f;
                    const float e = fminf(fminf(__fadd_rn(u, 0.5f), __fsub_rn(__fsub_rn((float)p.w, 0.5f), u)),
                                          fminf(__fadd_rn(v, 0.5f), __fsub_rn(__fsub_rn((float)p.h, 0.5f), v)));
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
            quantize(p.M, s_row[b][r][0], s_row[b][r][1], s_row[b][r][2], dmul(p.M.m[0], dx1), dmul(p.M.m[3], dx1),
                     dmul(p.M.m[6], dx1), X, Y);
            if constexpr (ROI) mv = sample<INTERP, 1>(p.roi, p.rpitch, p.h, p.w, X, Y, p.cubic);
            else mv = sample<INTERP, 3>(p.src, p.spitch, p.h, p.w, X, Y, p.cubic) ? 255u : 0u;
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
        for (int i = tid; i < EH * TW; i += NTHREADS) {
            const int r = i / TW, c = i - r * TW;
            const uint8_t *q = &s_mval[r][c];
            s_mh[r][c] = (uint8_t)min(min(min((int)q[0], (int)q[1]), min((int)q[2], (int)q[3])), (int)q[4]);
        }
        __syncthreads();
        for (int i = tid; i < TH * TW; i += NTHREADS) {
            const int r = i / TW, c = i - r * TW;
            const int m = min(min(min((int)s_mh[r][c], (int)s_mh[r + 1][c]), min((int)s_mh[r + 2][c], (int)s_mh[r + 3][c])),
                              (int)s_mh[r + 4][c]);
            s_mval[r][c] = (uint8_t)m;   // rows r < TH of s_mval are no longer needed as input: s_mh holds them
        }
        __syncthreads();
        if (tid < TH) {
            unsigned long long bits = 0;
            for (int c = 0; c < TW; ++c) bits |= (unsigned long long)(s_mval[tid][c] != 0) << c;
            s_v[tid] = bits;
        }
    }
    cp_async_commit_wait_all();
    __syncthreads();

    // ---- phase 2: blend; thread = 4 consecutive pixels = 3 aligned canvas words --------------
#pragma unroll 1
    for (int item = tid; item < TH * (TW / 4); item += NTHREADS) {
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
        uint32_t mk4 = *mp;
        float4 wa;
        float *wp = nullptr;
        if constexpr (BLEND == MK_BLEND_FEATHER) {
            wp = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(p.wacc) + (size_t)y * p.wpitch) + x;
            wa = *reinterpret_cast<const float4 *>(wp);
        }
        bool touched = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!((mb >> j) & 1u)) continue;
            const uint32_t wv = s_warp[r][4 * g + j];
            if constexpr (BLEND == MK_BLEND_ALPHA) {
                const bool old = ((mk4 >> (8 * j)) & 255u) > 1u;
                px[j] = old ? blend_px(px[j], wv, p.alpha, p.beta) : wv;
                if constexpr (ROI) mk4 |= (uint32_t)s_mval[r][4 * g + j] << (8 * j);
                else mk4 |= 255u << (8 * j);
                touched = true;
            } else {
                const float gw = s_wgt[r][4 * g + j];
                if (gw > 0.f) {
                    float *accp = (j == 0) ? &wa.x : (j == 1) ? &wa.y : (j == 2) ? &wa.z : &wa.w;
                    const float acc = *accp, s = __fadd_rn(acc, gw);
                    px[j] = feather_u8(px[j] & 255u, wv & 255u, acc, gw, s) |
                            (feather_u8((px[j] >> 8) & 255u, (wv >> 8) & 255u, acc, gw, s) << 8) |
                            (feather_u8((px[j] >> 16) & 255u, (wv >> 16) & 255u, acc, gw, s) << 16);
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
            if constexpr (BLEND == MK_BLEND_FEATHER) *reinterpret_cast<float4 *>(wp) = wa;
        }
    }
    __syncthreads();

    // ---- write back the rows that changed, as whole 16-byte vectors ---------------------------
    for (int v = tid; v < TH * ROWV; v += NTHREADS) {
        const int r = v / ROWV, c = v - r * ROWV, y = tileY + r;
        if (y < p.Hc && c < nvec && s_v[r] != 0ull)
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

static const short *cubic_table_device(int *status)
{
    // one table per device, created on first use (read-only afterwards)
    static std::mutex mu;
    static short *tabs[64] = { nullptr };
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { *status = MK_ERR_CUDA; return nullptr; }
    std::lock_guard<std::mutex> lk(mu);
    if (!tabs[dev]) {
        short host[1024 * 16];
        build_cubic_table(host);
        short *d = nullptr;
        if (cudaMalloc(&d, sizeof host) != cudaSuccess || cudaMemcpy(d, host, sizeof host, cudaMemcpyHostToDevice) != cudaSuccess) {
            *status = cuda_fail(cudaGetLastError(), "cubic table upload");
            return nullptr;
        }
        tabs[dev] = d;
    }
    *status = MK_OK;
    return tabs[dev];
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

template <int INTERP, int BLEND>
static void launch_mapper(bool roi, dim3 grid, cudaStream_t st, const MapperParams &P)
{
    if (roi) mapper_update_kernel<INTERP, BLEND, true><<<grid, NTHREADS, 0, st>>>(P);
    else mapper_update_kernel<INTERP, BLEND, false><<<grid, NTHREADS, 0, st>>>(P);
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

    MapperParams P;
    P.canvas = canvas; P.cmask = cmask; P.wacc = wacc; P.src = src; P.roi = roi; P.cubic = nullptr;
    P.cpitch = canvas_pitch; P.mpitch = cmask_pitch; P.wpitch = wacc_pitch; P.spitch = src_pitch; P.rpitch = roi_pitch;
    P.Hc = canvas_h; P.Wc = canvas_w; P.h = src_h; P.w = src_w;
    make_invmap(H_host, P.M);
    P.alpha = (float)alpha; P.beta = (float)(1.0 - alpha); P.ramp = feather_ramp;
    if (interp == MK_INTER_CUBIC) {
        int st = MK_OK;
        P.cubic = cubic_table_device(&st);
        if (!P.cubic) return st;
    }
    P.tx0 = rect[0] / TW; P.ty0 = rect[1] / TH;
    const int tx1 = (rect[2] + TW - 1) / TW, ty1 = (rect[3] + TH - 1) / TH;
    dim3 grid(tx1 - P.tx0, ty1 - P.ty0);
    cudaStream_t st = (cudaStream_t)stream;
    if (interp == MK_INTER_LINEAR) {
        if (blend == MK_BLEND_ALPHA) launch_mapper<MK_INTER_LINEAR, MK_BLEND_ALPHA>(roi != nullptr, grid, st, P);
        else launch_mapper<MK_INTER_LINEAR, MK_BLEND_FEATHER>(false, grid, st, P);
    } else {
        if (blend == MK_BLEND_ALPHA) launch_mapper<MK_INTER_CUBIC, MK_BLEND_ALPHA>(roi != nullptr, grid, st, P);
        else launch_mapper<MK_INTER_CUBIC, MK_BLEND_FEATHER>(false, grid, st, P);
    }
    MK_LAUNCH_CHECK("mapper_update_kernel");
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
    if (interp == MK_INTER_CUBIC) {
        int st = MK_OK;
        P.cubic = cubic_table_device(&st);
        if (!P.cubic) return st;
    }
    dim3 grid((dst_w + TW - 1) / TW, (dst_h + TH - 1) / TH);
    cudaStream_t st = (cudaStream_t)stream;
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
