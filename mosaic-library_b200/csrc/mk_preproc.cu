// mk_preproc.cu -- row N4: the per-frame pre-processing either side of the hot path, on the device, so a
// frame that is already in HBM (decoder output, peer memory) never returns to the host before it is matched and composited:
//
//   mk_cvt_color_u8   cv2.cvtColor for the four conversions of preprocessing.make_gray / make_bgr
//                     (src/mosaicking/preprocessing.py:314-346; called from mosaic.py:66,112,312 and registration.py:100-149)
//   mk_remap_u8       cv2.remap(image, xmap, ymap, interp) with float32 maps, BORDER_CONSTANT 0 -- the undistortion of
//                     preprocessing.DistortionMapper.apply (preprocessing.py:76-98)
//
// Both are pure streaming kernels (HBM bound): 16 pixels per thread with 16-byte accesses for the colour conversions, one
// pixel per lane with coalesced map reads for the remap.  Arithmetic restated from OpenCV 4.x, checked on all 2^24 BGR triples:
//   gray = (B * 3735 + G * 19235 + R * 9798 + 2^14) >> 15          (the 15-bit coefficients of cv::cvtColor's u8 path)
//   remap: X = rint(xmap * 32), Y = rint(ymap * 32) in fp32, then exactly warpPerspective's sampler (mk_sample.cuh)
#include "mk_mapper_common.cuh"

namespace mk {

constexpr int CVT_BGR2GRAY = 0, CVT_BGRA2GRAY = 1, CVT_GRAY2BGR = 2, CVT_BGRA2BGR = 3;

__device__ __forceinline__ uint32_t gray_of(uint32_t b, uint32_t g, uint32_t r) { return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15; }

// 16 pixels per thread.  Fast path: the 16-pixel group lies inside the row and both rows are 16-byte aligned (pitches % 16 == 0
// and 16-byte aligned bases, checked on the host): whole uint4 loads and stores.  The last group of a row goes pixel by pixel.
template <int CODE>
__global__ void __launch_bounds__(256) cvt_color_kernel(const uint8_t *__restrict__ src, size_t spitch, uint8_t *__restrict__ dst, size_t dpitch,
                                                        int h, int w, int groups, int vec_ok)
{
    constexpr int SC = (CODE == CVT_BGR2GRAY) ? 3 : (CODE == CVT_GRAY2BGR) ? 1 : 4;      // source / destination channels
    constexpr int DC = (CODE == CVT_BGR2GRAY || CODE == CVT_BGRA2GRAY) ? 1 : 3;
    const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int y = (int)(gid / groups), gx = (int)(gid - (long)y * groups);
    if (y >= h) return;
    const int x0 = gx * 16;
    const uint8_t *s = src + (size_t)y * spitch + (size_t)x0 * SC;
    uint8_t *d = dst + (size_t)y * dpitch + (size_t)x0 * DC;
    if (vec_ok && x0 + 16 <= w) {
        uint32_t in[4 * SC];
#pragma unroll
        for (int k = 0; k < SC; ++k) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(s) + k);
            in[4 * k] = v.x; in[4 * k + 1] = v.y; in[4 * k + 2] = v.z; in[4 * k + 3] = v.w;
        }
        auto byte_at = [&](int i) -> uint32_t { return (in[i >> 2] >> (8 * (i & 3))) & 255u; };   // i is a compile-time constant after unrolling
        uint32_t out[4 * DC];
#pragma unroll
        for (int k = 0; k < 4 * DC; ++k) out[k] = 0;
#pragma unroll
        for (int px = 0; px < 16; ++px) {
            if constexpr (DC == 1) {
                const uint32_t g = gray_of(byte_at(px * SC), byte_at(px * SC + 1), byte_at(px * SC + 2));
                out[px >> 2] |= g << (8 * (px & 3));
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const uint32_t v = (CODE == CVT_GRAY2BGR) ? byte_at(px) : byte_at(px * 4 + c);
                    const int o = px * 3 + c;
                    out[o >> 2] |= v << (8 * (o & 3));
                }
            }
        }
#pragma unroll
        for (int k = 0; k < DC; ++k) reinterpret_cast<uint4 *>(d)[k] = make_uint4(out[4 * k], out[4 * k + 1], out[4 * k + 2], out[4 * k + 3]);
    } else {
        for (int px = 0; px < 16 && x0 + px < w; ++px) {
            if constexpr (DC == 1) d[px] = (uint8_t)gray_of(s[px * SC], s[px * SC + 1], s[px * SC + 2]);
            else if constexpr (CODE == CVT_GRAY2BGR) { d[px * 3] = d[px * 3 + 1] = d[px * 3 + 2] = s[px]; }
            else { d[px * 3] = s[px * 4]; d[px * 3 + 1] = s[px * 4 + 1]; d[px * 3 + 2] = s[px * 4 + 2]; }
        }
    }
}

struct RemapParams {
    const uint8_t *src;
    uint8_t *dst;
    const float *mapx, *mapy;
    const short *cubic;
    unsigned long long spitch, dpitch, mpitch;     // bytes
    int h, w, Hd, Wd;
};

// lanes = consecutive destination columns (coalesced map reads), 8 rows per CTA.  The coordinates go to 1/32 px exactly like
// cv::remap converts float maps (cvRound(x * INTER_TAB_SIZE) in fp32, round-half-even), then the sampler of the warp kernels.
template <int INTERP, int CN>
__global__ void __launch_bounds__(256) remap_kernel(const RemapParams p)
{
    const int x = (int)blockIdx.x * 32 + (threadIdx.x & 31), y = (int)blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= p.Wd || y >= p.Hd) return;
    const float fx = __ldg(reinterpret_cast<const float *>(reinterpret_cast<const uint8_t *>(p.mapx) + (size_t)y * p.mpitch) + x);
    const float fy = __ldg(reinterpret_cast<const float *>(reinterpret_cast<const uint8_t *>(p.mapy) + (size_t)y * p.mpitch) + x);
    const int X = __float2int_rn(__fmul_rn(fx, 32.f)), Y = __float2int_rn(__fmul_rn(fy, 32.f));
    const uint32_t v = sample<INTERP, CN>(p.src, (size_t)p.spitch, p.h, p.w, X, Y, p.cubic);
    uint8_t *d = p.dst + (size_t)y * p.dpitch + (size_t)x * CN;
    if constexpr (CN == 1) d[0] = (uint8_t)v;
    else { d[0] = (uint8_t)v; d[1] = (uint8_t)(v >> 8); d[2] = (uint8_t)(v >> 16); }
}

}  // namespace mk

using namespace mk;

extern "C" int mk_cvt_color_u8(const uint8_t *src, int h, int w, size_t src_pitch, int code, uint8_t *dst, size_t dst_pitch, void *stream)
{
    if (!src || !dst || h <= 0 || w <= 0 || code < 0 || code > 3) { set_error("mk_cvt_color_u8: bad argument"); return MK_ERR_ARG; }
    const int sc = (code == CVT_BGR2GRAY) ? 3 : (code == CVT_GRAY2BGR) ? 1 : 4, dc = (code == CVT_BGR2GRAY || code == CVT_BGRA2GRAY) ? 1 : 3;
    if (src_pitch < (size_t)w * sc || dst_pitch < (size_t)w * dc) { set_error("mk_cvt_color_u8: pitch smaller than a row"); return MK_ERR_ARG; }
    const int vec_ok = aligned(src, 16) && aligned(dst, 16) && !(src_pitch & 15) && !(dst_pitch & 15);
    const int groups = (w + 15) / 16;
    const long threads = (long)groups * h;
    const unsigned blocks = (unsigned)((threads + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    switch (code) {
    case CVT_BGR2GRAY: cvt_color_kernel<CVT_BGR2GRAY><<<blocks, 256, 0, st>>>(src, src_pitch, dst, dst_pitch, h, w, groups, vec_ok); break;
    case CVT_BGRA2GRAY: cvt_color_kernel<CVT_BGRA2GRAY><<<blocks, 256, 0, st>>>(src, src_pitch, dst, dst_pitch, h, w, groups, vec_ok); break;
    case CVT_GRAY2BGR: cvt_color_kernel<CVT_GRAY2BGR><<<blocks, 256, 0, st>>>(src, src_pitch, dst, dst_pitch, h, w, groups, vec_ok); break;
    default: cvt_color_kernel<CVT_BGRA2BGR><<<blocks, 256, 0, st>>>(src, src_pitch, dst, dst_pitch, h, w, groups, vec_ok); break;
    }
    MK_LAUNCH_CHECK("cvt_color_kernel");
    return MK_OK;
}

extern "C" int mk_remap_u8(const uint8_t *src, int src_h, int src_w, int channels, size_t src_pitch, const float *map_x, const float *map_y,
                           size_t map_pitch, uint8_t *dst, int dst_h, int dst_w, size_t dst_pitch, int interp, void *stream)
{
    if (!src || !dst || !map_x || !map_y || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0 || (channels != 1 && channels != 3)) {
        set_error("mk_remap_u8: bad argument (channels must be 1 or 3)");
        return MK_ERR_ARG;
    }
    if (src_h > 32767 || src_w > 32767) { set_error("mk_remap_u8: source larger than 32767 px"); return MK_ERR_ARG; }
    if (interp != MK_INTER_LINEAR && interp != MK_INTER_CUBIC) { set_error("mk_remap_u8: bad interp"); return MK_ERR_ARG; }
    if (src_pitch < (size_t)src_w * channels || dst_pitch < (size_t)dst_w * channels || map_pitch < (size_t)dst_w * 4) {
        set_error("mk_remap_u8: pitch smaller than a row");
        return MK_ERR_ARG;
    }
    if (!aligned(src, 4) || (src_pitch & 3) || !aligned(map_x, 4) || !aligned(map_y, 4) || (map_pitch & 3)) {
        set_error("mk_remap_u8: alignment (src and maps 4 B, pitches %% 4)");
        return MK_ERR_ALIGN;
    }
    RemapParams P;
    P.src = src; P.dst = dst; P.mapx = map_x; P.mapy = map_y; P.cubic = nullptr;
    P.spitch = src_pitch; P.dpitch = dst_pitch; P.mpitch = map_pitch; P.h = src_h; P.w = src_w; P.Hd = dst_h; P.Wd = dst_w;
    {
        DeviceTables tb;
        const int rc = device_tables(tb);
        if (rc != MK_OK) return rc;
        P.cubic = tb.cubic;
    }
    dim3 grid((dst_w + 31) / 32, (dst_h + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (interp == MK_INTER_LINEAR) {
        if (channels == 3) remap_kernel<MK_INTER_LINEAR, 3><<<grid, 256, 0, st>>>(P);
        else remap_kernel<MK_INTER_LINEAR, 1><<<grid, 256, 0, st>>>(P);
    } else {
        if (channels == 3) remap_kernel<MK_INTER_CUBIC, 3><<<grid, 256, 0, st>>>(P);
        else remap_kernel<MK_INTER_CUBIC, 1><<<grid, 256, 0, st>>>(P);
    }
    MK_LAUNCH_CHECK("remap_kernel");
    return MK_OK;
}
