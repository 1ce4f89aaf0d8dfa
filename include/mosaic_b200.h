/*
 * mosaic_b200.h -- C ABI of libmosaic_b200.so: the B200 (sm_100a) replacement for the
 * arithmetic behind mosaic-library's per-frame registration + compositing hot path.
 *
 * The reference has no FFI of its own: its hot path is three OpenCV call sites reached
 * from Python (paths relative to the mosaic-library checkout).  Every entry point below
 * names the call site it replaces; INTEGRATION.md shows the ctypes stub a maintainer adds.
 *
 * Conventions
 *   - plain C, no torch / C++ types.  All data pointers are DEVICE pointers owned by the
 *     caller; `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - every call is asynchronous on `stream`, keeps no global mutable state besides
 *     read-only tables, and returns MK_OK or a negative mk_status; mk_last_error() gives a
 *     thread-local message for the last failure on the calling thread.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails
 *     with MK_ERR_CUDA.
 */
#ifndef MOSAIC_B200_H
#define MOSAIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

#define MK_ABI_VERSION 1

typedef enum mk_status {
    MK_OK = 0,
    MK_ERR_ARG = -1,      /* bad shape / null pointer / unsupported size            */
    MK_ERR_ALIGN = -2,    /* pointer or pitch alignment requirement not met          */
    MK_ERR_CUDA = -3,     /* CUDA runtime error (message in mk_last_error)           */
    MK_ERR_WORKSPACE = -4 /* workspace too small (see mk_knn2_workspace_bytes)       */
} mk_status;

typedef enum mk_norm {
    MK_NORM_HAMMING = 0,  /* cv2.NORM_HAMMING on uint8 rows (north-star ORB metric)  */
    MK_NORM_L2_U8 = 1,    /* cv2.NORM_L2 on uint8 rows  (what registration.py:231 does for ORB) */
    MK_NORM_L2_F32 = 2,   /* cv2.NORM_L2 on float32 rows (SIFT)                      */
    MK_NORM_HAMMING_TC = 3 /* MK_NORM_HAMMING, same results, optional engine: bit-expanded rows (<= 32 bytes) through the
                             tcgen05 fp8 GEMM when the problem is large; otherwise the XOR+POPC kernel */
} mk_norm;

typedef enum mk_interp {
    MK_INTER_LINEAR = 1,  /* cv2.INTER_LINEAR: 1/32-px fixed point, north-star kernel */
    MK_INTER_CUBIC = 2    /* cv2.INTER_CUBIC: the flag mosaic.py:114,118 passes       */
} mk_interp;

typedef enum mk_blend {
    MK_BLEND_ALPHA = 0,   /* Mapper._update_cpu semantics (mosaic.py:115-138)         */
    MK_BLEND_FEATHER = 1  /* analytic feather weights + fp32 weight plane (extension) */
} mk_blend;

typedef enum mk_model {
    MK_MODEL_AFFINE = 0,      /* cv2.estimateAffine2D: 6 degrees of freedom, 3-point samples      */
    MK_MODEL_PERSPECTIVE = 1  /* cv2.findHomography:   8 degrees of freedom, 4-point samples      */
} mk_model;

int mk_abi_version(void);
const char *mk_version_string(void);
const char *mk_last_error(void);
/* Number of kernels this library has launched in the calling process (bench "gpu_launches"). */
uint64_t mk_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * Brute-force k=2 nearest neighbours + Lowe ratio test.
 * Replaces  Matcher.knn_match -> cv2.BFMatcher.knnMatch(query, train, k=2)
 *           (src/mosaicking/registration.py:233-242)  and the ratio test of
 *           src/mosaicking/mosaic.py:430  (m[0].distance < ratio * m[1].distance, float64).
 *
 *   q, t      row-major descriptors, `row_stride` BYTES between rows (multiple of 16; rows
 *             zero-padded up to a multiple of 16 bytes), `dim` = bytes per row for the uint8
 *             norms (<= 64) or floats per row for MK_NORM_L2_F32 (<= 256).
 *   idx       int32 [nq][2] train indices, ascending distance, ties -> lowest train index
 *             (cv2's strict-< insertion); -1 where fewer than 2 train rows exist.
 *   dist      float32 [nq][2] distances exactly as cv2 reports them (Hamming count as float,
 *             sqrtf of the squared L2 sum); FLT_MAX where idx = -1.
 *   keep      uint8 [nq] ratio-test flag (may be NULL); ratio <= 0 disables the test (keep=0).
 *   n_keep    int32 [1] device counter of kept rows (may be NULL); overwritten.
 *   workspace scratch of at least mk_knn2_workspace_bytes(..) bytes, 256-byte aligned.
 * ------------------------------------------------------------------------------------- */
size_t mk_knn2_workspace_bytes(int norm, int nq, int nt, int dim);
/* 1 if mk_knn2 routes this problem through the tcgen05 / TMA descriptor GEMM: L2 norms with dim <= 128, or
 * MK_NORM_HAMMING_TC with dim <= 32 bytes (bit-expanded fp8 rows, fp16 accumulators), and enough pairs to
 * amortise the extra launch (>= 256 K).  Float rows must also turn out to be integer-valued in [0,255] --
 * OpenCV SIFT is -- otherwise the bit-exact SIMT kernel runs, decided on the device.  MK_NORM_HAMMING always
 * takes the XOR+POPC kernel; both Hamming codes return identical results. */
int mk_knn2_uses_tensor_cores(int norm, int nq, int nt, int dim);

int mk_knn2(int norm, const void *q, int nq, const void *t, int nt, int dim, size_t row_stride,
            int32_t *idx, float *dist, double ratio, uint8_t *keep, int32_t *n_keep,
            void *workspace, size_t workspace_bytes, void *stream);

/* Batched form: many (query set, train set) pairs in one launch -- consecutive-frame pairs of
 * SequentialMosaic.match_features (mosaic.py:402-438) or all-pairs loop closure.
 *   desc       all descriptor rows of all sets, concatenated (same dim / row_stride); total_rows of them
 *   set_off    int32 [nsets+1] DEVICE: first row of each set
 *   pairs      int32 [npairs][2] DEVICE: (query set, train set)
 *   out_off    int32 [npairs+1] DEVICE: first output row of each pair (prefix sum of nq)
 *   max_nq / max_nt  upper bounds of the set sizes (host side, sizes the grid)
 *   idx/dist/keep     outputs over sum(nq) rows;  n_keep int32 [npairs] (may be NULL) */
size_t mk_knn2_batched_workspace_bytes(int norm, int npairs, int max_nq, int max_nt, int dim, int total_rows);

int mk_knn2_batched(int norm, const void *desc, int total_rows, int dim, size_t row_stride, const int32_t *set_off,
                    const int32_t *pairs, int npairs, const int32_t *out_off, int max_nq, int max_nt,
                    int32_t *idx, float *dist, double ratio, uint8_t *keep, int32_t *n_keep,
                    void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------
 * Inverse-homography warp.  Replaces
 *   cv2.warpPerspective(src, H, (dst_w, dst_h), None, interp, BORDER_CONSTANT, 0)
 *   (src/mosaicking/mosaic.py:114,118; transformations.py:116).
 * H is the 3x3 row-major float64 FORWARD map (frame -> canvas) in HOST memory, as the
 * reference passes it; it is inverted on the host exactly like cv::invert does.
 * uint8, channels = 1 or 3.  src/dst 16-byte aligned, pitches multiples of 16 (dst) / 4 (src);
 * src dims <= 32767 (cv2's own limit, short source coordinates).
 * ------------------------------------------------------------------------------------- */
int mk_warp_perspective_u8(const uint8_t *src, int src_h, int src_w, int channels, size_t src_pitch,
                           const double *H_host, uint8_t *dst, int dst_h, int dst_w, size_t dst_pitch,
                           int interp, void *stream);

/* ---------------------------------------------------------------------------------------
 * Fused warp + validity mask + 5x5 erosion + blend + canvas / mask accumulation: one
 * Mapper.update step (src/mosaicking/mosaic.py:159-172 -> _update_cpu :111-138) in a single
 * pass over the warped footprint instead of ~8 full-canvas passes.
 *
 *   canvas   uint8 [canvas_h][canvas_w][3] BGR, in place       (Mapper._canvas)
 *   cmask    uint8 [canvas_h][canvas_w], in place              (Mapper._canvas_mask)
 *   wacc     float [canvas_h][canvas_w] weight plane, MK_BLEND_FEATHER only (else NULL)
 *   src      uint8 [src_h][src_w][3] BGR frame
 *   roi      uint8 [src_h][src_w] keypoint-hull mask (mosaic.py:140-157) or NULL
 *   alpha    Mapper alpha_blend in [0,1]  (MK_BLEND_ALPHA);  feather_ramp in source px
 *   region   if non-NULL (HOST int32[4]) receives the canvas rectangle x0,y0,x1,y1 visited
 * Alignment: canvas/cmask/wacc 16-byte aligned with pitches (in bytes) multiples of 16;
 * src/roi 4-byte aligned with pitches multiples of 4.
 * ------------------------------------------------------------------------------------- */
int mk_mapper_update(uint8_t *canvas, size_t canvas_pitch, uint8_t *cmask, size_t cmask_pitch,
                     float *wacc, size_t wacc_pitch, int canvas_h, int canvas_w,
                     const uint8_t *src, int src_h, int src_w, size_t src_pitch,
                     const uint8_t *roi, size_t roi_pitch,
                     const double *H_host, double alpha, int interp, int blend, float feather_ramp,
                     int32_t *region_host, void *stream);

/* Batched form: `nframes` device-resident frames are composited in sequence order by ONE call -- a back-to-back train of
 * strip-kernel launches for MK_BLEND_ALPHA (that kernel is not HBM bound, so the train is the faster form), one "gather"
 * launch for MK_BLEND_FEATHER (every canvas tile is read once, updated by all the frames that reach it and written once).
 * The result is identical to nframes successive mk_mapper_update calls -- what SequentialMosaic.generate does for a tile
 * (src/mosaicking/mosaic.py:678-686).  Arrays with the _host suffix live in HOST memory: src_host[k] is the device
 * pointer of frame k, H_host holds nframes row-major 3x3 float64 matrices.  No ROI masks in this form. */
int mk_mapper_update_batch(uint8_t *canvas, size_t canvas_pitch, uint8_t *cmask, size_t cmask_pitch,
                           float *wacc, size_t wacc_pitch, int canvas_h, int canvas_w, int nframes,
                           const uint8_t *const *src_host, const int *src_h_host, const int *src_w_host,
                           const size_t *src_pitch_host, const double *H_host, double alpha, int interp, int blend,
                           float feather_ramp, void *stream);

/* ---------------------------------------------------------------------------------------
 * Host <-> device helpers for callers that hold HOST buffers (the reference passes numpy arrays):
 * thin wrappers over cudaMemcpy2DAsync so that a binding needs nothing but this library.
 * Asynchronous when the host buffer is page-locked, staged by the driver otherwise.
 * ------------------------------------------------------------------------------------- */
int mk_upload_2d(void *dst_dev, size_t dst_pitch, const void *src_host, size_t src_pitch, size_t row_bytes,
                 size_t rows, void *stream);
int mk_download_2d(void *dst_host, size_t dst_pitch, const void *src_dev, size_t src_pitch, size_t row_bytes,
                   size_t rows, void *stream);

/* Stream ordering without a host round trip, for callers that run matching and compositing of a frame side by side
 * (the reference's loop is single-threaded and serial, mosaic.py:402-438 / 678-686): mk_stream_follow records `event` on
 * `src_stream` and makes `dst_stream` wait for it -- everything queued on dst afterwards runs after everything queued on
 * src before.  With MK_STREAM_NONE as dst the event is only recorded, with MK_STREAM_NONE as src it is only waited for
 * (record now, join later).  Events come from mk_event_create (timing disabled) and are reusable. */
#define MK_STREAM_NONE ((void *)~(uintptr_t)0)
void *mk_event_create(void);
void mk_event_destroy(void *event);
int mk_stream_follow(void *event, void *src_stream, void *dst_stream);

/* Spatial partition of the GPU for the two halves of a frame step.  Matching (one CTA per SM, 180 KB of shared memory) and
 * compositing (5 small CTAs per SM) cannot share an SM, so on two ordinary streams the later kernel mostly waits for the
 * earlier one; with a green context each (CUDA driver API, sm_90+) they run side by side from the first microsecond.
 * mk_sm_partition_create splits the SMs of the current device into a first group of at least `min_sms_first` SMs (rounded up
 * by the driver, multiples of 8) and the rest, and returns one non-blocking stream in each; work launched on those streams
 * stays inside its group.  mk_knn2 sizes its grid for the SM count of the stream it is given.  Returns MK_ERR_CUDA when the
 * driver cannot provide green contexts (callers then just use ordinary streams). */
int mk_sm_partition_create(int min_sms_first, void **partition, void **stream_first, void **stream_rest, int *sms_first,
                           int *sms_rest);
void mk_sm_partition_destroy(void *partition);

/* ---------------------------------------------------------------------------------------
 * Device-side registration glue: what the reference does on the host between matching and
 * compositing.  All pointers are DEVICE pointers unless named *_host; asynchronous on `stream`.
 *
 * mk_gather_matches   SequentialMosaic.registration (src/mosaicking/mosaic.py:449-457): the keypoints of the ratio-test
 *   survivors in match order, src = kp_query[i] ("kp", the later frame), dst = kp_train[idx[i][0]] ("kp_prev").
 *   kp_* are float32 (x, y) pairs; idx / keep are mk_knn2's outputs; out_* hold up to nq rows; out_query (may be NULL)
 *   receives the query index of every gathered pair; count int32[1].
 * mk_ransac_transform HomographyEstimation.__call__ (src/mosaicking/registration.py:399-419) as called at mosaic.py:459:
 *   cv2.findHomography(src, dst, RANSAC, thresh) for MK_MODEL_PERSPECTIVE, cv2.estimateAffine2D(src, dst, RANSAC, thresh)
 *   for MK_MODEL_AFFINE.  `iters` hypotheses are scored in parallel (cv2's default cap is 2000), the best one is re-fitted
 *   to its inliers.  H_out double[9] row-major (affine: last row 0 0 1; all zeros when no model was found, where cv2
 *   returns None), inliers uint8[n] (0/1, the winning hypothesis' inliers like cv2's mask), n_inliers int32[1].
 *   workspace: mk_ransac_workspace_bytes() bytes, 8-byte aligned.  n_dev is reserved and must be NULL.
 *   Randomised like cv2's: results agree with cv2 up to points within rounding of the threshold (tests state the bounds).
 * mk_chain_homographies  core/interface.py:97-113 (_propagate_homographies): H_abs[0] = H[0], H_abs[k] = H_abs[k-1] @ H[k],
 *   n row-major 3x3 float64 matrices, left to right like itertools.accumulate(H, np.matmul).
 * ------------------------------------------------------------------------------------- */
int mk_gather_matches(const float *kp_query, const float *kp_train, const int32_t *idx, const uint8_t *keep, int nq,
                      float *out_src, float *out_dst, int32_t *out_query, int32_t *count, void *stream);
size_t mk_ransac_workspace_bytes(void);
int mk_ransac_transform(int model, const float *src_xy, const float *dst_xy, int n, const int32_t *n_dev, double thresh,
                        int iters, uint64_t seed, double *H_out, uint8_t *inliers, int32_t *n_inliers, void *workspace,
                        void *stream);
int mk_chain_homographies(const double *H, int n, double *H_abs, void *stream);

/* ---------------------------------------------------------------------------------------
 * Peer memory over NVLink / NVSwitch for the tile-sharded canvas (one process per GPU; src/mosaicking/mosaic.py:595-622 routes a
 * frame to every tile its footprint overlaps).  The rank that holds a frame exports its allocation, the ranks owning the other
 * overlapped tiles map it and pass the mapped pointer as `src` of mk_mapper_update: the compositing kernel then pulls exactly the
 * source rectangles its tiles need out of the ingest GPU's HBM (2-D TMA boxes over NVLink) -- no copy of the whole frame, no
 * collective, the transfer overlaps the arithmetic tile by tile.
 *   mk_ipc_export   handle_out[64] + byte offset of dev_ptr inside its allocation (cudaIpcGetMemHandle on the allocation base)
 *   mk_ipc_open     maps the allocation in the calling process (peer access enabled lazily); pointer = *base_out + offset.
 *                   Open a given handle once per process and mk_ipc_close it before the exporter frees the memory.
 * Ordering is the callers' business: the exporter must have finished writing the frame, and must keep it, while peers read it.
 * ------------------------------------------------------------------------------------- */
int mk_ipc_export(const void *dev_ptr, uint8_t *handle_out, uint64_t *offset_out);
int mk_ipc_open(const uint8_t *handle, void **base_out);
int mk_ipc_close(void *base);
/* Whole-frame pull over NVLink by the COPY ENGINE (no SM, no NCCL kernel next to the persistent compositing kernel):
 * cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault) on `stream`, src typically a pointer obtained through mk_ipc_open. */
int mk_copy_peer(void *dst_dev, const void *src_dev, size_t bytes, void *stream);

/* ---------------------------------------------------------------------------------------
 * Per-frame pre-processing on the device (row N4): a frame that is already in device memory is
 * converted / undistorted there instead of on the host.
 *
 * mk_cvt_color_u8 replaces the cv2.cvtColor calls of preprocessing.make_gray / make_bgr
 *   (src/mosaicking/preprocessing.py:314-346; used at mosaic.py:66,112,312, registration.py:100-149):
 *   code MK_CVT_BGR2GRAY / MK_CVT_BGRA2GRAY (3 or 4 -> 1 channel, gray = (B*3735 + G*19235 + R*9798 + 2^14) >> 15),
 *        MK_CVT_GRAY2BGR / MK_CVT_BGRA2BGR (1 or 4 -> 3 channels).  Bit-identical to cv2 for uint8.
 *   16-byte aligned src/dst with pitches % 16 == 0 take the vector path; anything else works too.
 *
 * mk_remap_u8 replaces cv2.remap(image, xmap, ymap, interp) with float32 maps and BORDER_CONSTANT 0
 *   (preprocessing.DistortionMapper.apply, preprocessing.py:76-98; the maps come from
 *   cv2.initUndistortRectifyMap once per camera).  map_x / map_y: float [dst_h][dst_w], pitch in bytes.
 *   uint8, channels = 1 or 3; src and maps 4-byte aligned, pitches multiples of 4; src dims <= 32767.
 * ------------------------------------------------------------------------------------- */
typedef enum mk_cvt { MK_CVT_BGR2GRAY = 0, MK_CVT_BGRA2GRAY = 1, MK_CVT_GRAY2BGR = 2, MK_CVT_BGRA2BGR = 3 } mk_cvt;
int mk_cvt_color_u8(const uint8_t *src, int h, int w, size_t src_pitch, int code, uint8_t *dst, size_t dst_pitch, void *stream);
int mk_remap_u8(const uint8_t *src, int src_h, int src_w, int channels, size_t src_pitch, const float *map_x, const float *map_y,
                size_t map_pitch, uint8_t *dst, int dst_h, int dst_w, size_t dst_pitch, int interp, void *stream);

/* Algorithmic bytes of one mk_mapper_update: src + 2*3*A + 2*A (+ 8*A feather),
 * A = area of the canvas rectangle visited.  Pure host arithmetic. */
int mk_mapper_footprint(int canvas_h, int canvas_w, int src_h, int src_w, const double *H_host,
                        int interp, int32_t *region_host);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* MOSAIC_B200_H */
