// mk_ransac.cu -- device-side registration glue: what sits between matching and compositing in the
// reference's frame loop and becomes the end-to-end limiter once K1-K4 run on the GPU.
//
//   mk_gather_matches     SequentialMosaic.registration, src/mosaicking/mosaic.py:449-457: keypoints of the ratio-test survivors,
//                         kp[match.queryIdx] / kp_prev[match.trainIdx], in match order (ascending query index)
//   mk_ransac_transform   HomographyEstimation.__call__, src/mosaicking/registration.py:399-419, as called at mosaic.py:459:
//                         cv2.findHomography(src, dst, RANSAC, thr) or cv2.estimateAffine2D(src, dst, RANSAC, thr)
//   mk_chain_homographies core/interface.py:97-113 (_propagate_homographies): itertools.accumulate(H, np.matmul)
//
// RANSAC here is the textbook algorithm cv2 implements, parallel instead of sequential: every warp scores one minimal-sample
// hypothesis (3 or 4 correspondences drawn by a counter-based hash, solved in fp64 by lane 0, inliers counted by the 32 lanes),
// the best (count, lowest hypothesis index) wins through one 64-bit atomicMax, and a single-block pass re-fits the model to
// its inliers (affine: centred normal equations; perspective: Gauss-Newton on the reprojection error in Hartley-normalised
// coordinates, what cv2's LM refinement minimises).  cv2 draws its samples from its own RNG and stops adaptively, so the
// results agree with cv2 as two runs of a randomised estimator do: same inlier set up to points within rounding of the
// threshold, same model up to the fit noise (tolerances stated in tests/test_gpu_parity.py).  Deterministic for a given seed.
#include "mk_common.cuh"

#include <math.h>

#include <algorithm>

namespace mk {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;

__device__ __forceinline__ uint32_t rs_hash(uint64_t seed, uint32_t hyp, uint32_t j)
{
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * ((uint64_t)hyp * 8u + j + 1u);     // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 16);
}

// k distinct indices in [0, n) for hypothesis `hyp` (n >= k)
template <int KS>
__device__ __forceinline__ void rs_sample(uint64_t seed, uint32_t hyp, int n, int *ix)
{
#pragma unroll
    for (int j = 0; j < KS; ++j) {
        int v = (int)(rs_hash(seed, hyp, j) % (uint32_t)(n - j));
        // map into the indices not taken yet (ascending insertion keeps it exact and branch-light)
        int taken[KS];
#pragma unroll
        for (int t = 0; t < KS; ++t) taken[t] = (t < j) ? ix[t] : 0x7fffffff;
        // sort the j taken values (j <= 3)
#pragma unroll
        for (int a = 0; a < KS; ++a)
#pragma unroll
            for (int b = a + 1; b < KS; ++b)
                if (taken[b] < taken[a]) { const int tmp = taken[a]; taken[a] = taken[b]; taken[b] = tmp; }
#pragma unroll
        for (int t = 0; t < KS; ++t)
            if (t < j && v >= taken[t]) ++v;
        ix[j] = v;
    }
}

// 3 correspondences -> affine map (row-major 3x3, last row 0 0 1).  false when the three source points are (nearly) collinear.
__device__ bool rs_affine3(const double s[3][2], const double d[3][2], double *M)
{
    const double x0 = s[0][0], y0 = s[0][1], x1 = s[1][0] - x0, y1 = s[1][1] - y0, x2 = s[2][0] - x0, y2 = s[2][1] - y0;
    const double det = x1 * y2 - x2 * y1;
    const double scale = fabs(x1) + fabs(y1) + fabs(x2) + fabs(y2);
    if (!(fabs(det) > 1e-9 * scale * scale) || !(scale > 0.0)) return false;
    const double inv = 1.0 / det;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const double u1 = d[1][c] - d[0][c], u2 = d[2][c] - d[0][c];
        const double a = (u1 * y2 - u2 * y1) * inv, b = (x1 * u2 - x2 * u1) * inv;
        M[3 * c + 0] = a; M[3 * c + 1] = b; M[3 * c + 2] = d[0][c] - a * x0 - b * y0;
    }
    M[6] = 0.0; M[7] = 0.0; M[8] = 1.0;
    return true;
}

// dense solve of an N x N system (Gaussian elimination with partial pivoting), A row-major with the right-hand side in column N
template <int N>
__device__ bool rs_solve(double A[N][N + 1], double *x)
{
    for (int c = 0; c < N; ++c) {
        int piv = c;
        double best = fabs(A[c][c]);
        for (int r = c + 1; r < N; ++r)
            if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
        if (!(best > 1e-12)) return false;
        if (piv != c)
            for (int k = c; k <= N; ++k) { const double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
        const double inv = 1.0 / A[c][c];
        for (int r = c + 1; r < N; ++r) {
            const double f = A[r][c] * inv;
            if (f != 0.0)
                for (int k = c; k <= N; ++k) A[r][k] -= f * A[c][k];
        }
    }
    for (int r = N - 1; r >= 0; --r) {
        double v = A[r][N];
        for (int k = r + 1; k < N; ++k) v -= A[r][k] * x[k];
        x[r] = v / A[r][r];
    }
    return true;
}

// 4 correspondences -> homography with h33 = 1.  Points are shifted to the first source / destination point first (conditioning).
__device__ bool rs_homography4(const double s[4][2], const double d[4][2], double *M)
{
    // no three source (or destination) points collinear: the check cv2's RANSAC applies to its subsets
    for (int set = 0; set < 2; ++set) {
        const double(*p)[2] = set ? d : s;
        for (int a = 0; a < 4; ++a) {
            const int i = (a + 1) & 3, j = (a + 2) & 3, k = (a + 3) & 3;      // the triple that leaves out point a
            const double ax = p[j][0] - p[i][0], ay = p[j][1] - p[i][1], bx = p[k][0] - p[i][0], by = p[k][1] - p[i][1];
            const double cr = ax * by - ay * bx, sc = (fabs(ax) + fabs(ay)) * (fabs(bx) + fabs(by));
            if (!(fabs(cr) > 1e-6 * sc) || !(sc > 0.0)) return false;
        }
    }
    double A[8][9];
    const double sx = s[0][0], sy = s[0][1], dx = d[0][0], dy = d[0][1];
    for (int i = 0; i < 4; ++i) {
        const double x = s[i][0] - sx, y = s[i][1] - sy, u = d[i][0] - dx, v = d[i][1] - dy;
        double *r0 = A[2 * i], *r1 = A[2 * i + 1];
        r0[0] = x; r0[1] = y; r0[2] = 1; r0[3] = 0; r0[4] = 0; r0[5] = 0; r0[6] = -u * x; r0[7] = -u * y; r0[8] = u;
        r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = x; r1[4] = y; r1[5] = 1; r1[6] = -v * x; r1[7] = -v * y; r1[8] = v;
    }
    double h[8];
    if (!rs_solve<8>(A, h)) return false;
    // undo the shifts: H = T(d0) * Hc * T(-s0)
    const double a = h[0], b = h[1], c = h[2], e = h[3], f = h[4], g = h[5], p6 = h[6], p7 = h[7];
    const double w = 1.0 - p6 * sx - p7 * sy;
    double H[9];
    H[0] = a + dx * p6; H[1] = b + dx * p7; H[2] = c - a * sx - b * sy + dx * w;
    H[3] = e + dy * p6; H[4] = f + dy * p7; H[5] = g - e * sx - f * sy + dy * w;
    H[6] = p6; H[7] = p7; H[8] = w;
    if (!(fabs(H[8]) > 1e-12)) return false;
    const double inv = 1.0 / H[8];
    for (int i = 0; i < 9; ++i) M[i] = H[i] * inv;
    return true;
}

template <int MODEL>   // 0 affine, 1 perspective
__device__ __forceinline__ bool rs_model_from_sample(const float2 *src, const float2 *dst, int n, uint64_t seed, uint32_t hyp, double *M)
{
    constexpr int KS = MODEL ? 4 : 3;
    int ix[KS];
    rs_sample<KS>(seed, hyp, n, ix);
    double s[KS][2], d[KS][2];
#pragma unroll
    for (int j = 0; j < KS; ++j) {
        const float2 a = src[ix[j]], b = dst[ix[j]];
        s[j][0] = a.x; s[j][1] = a.y; d[j][0] = b.x; d[j][1] = b.y;
    }
    if constexpr (MODEL) return rs_homography4(s, d, M);
    else return rs_affine3(s, d, M);
}

__device__ __forceinline__ bool rs_is_inlier(const double *M, float2 a, float2 b, double thr2)
{
    const double x = a.x, y = a.y;
    const double w = M[6] * x + M[7] * y + M[8];
    const double iw = (w != 0.0) ? 1.0 / w : 0.0;
    const double ex = (M[0] * x + M[1] * y + M[2]) * iw - (double)b.x, ey = (M[3] * x + M[4] * y + M[5]) * iw - (double)b.y;
    return ex * ex + ey * ey <= thr2;
}

template <int MODEL>
__global__ void __launch_bounds__(RS_THREADS) ransac_score_kernel(const float2 *__restrict__ src, const float2 *__restrict__ dst, int n, double thr2,
                                                                   int iters, uint64_t seed, unsigned long long *best)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long mine = 0ull;
    for (uint32_t hyp = blockIdx.x * RS_WARPS + warp; hyp < (uint32_t)iters; hyp += gridDim.x * RS_WARPS) {
        double M[9];
        int ok = 0;
        if (lane == 0) ok = rs_model_from_sample<MODEL>(src, dst, n, seed, hyp, M) ? 1 : 0;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (!ok) continue;
#pragma unroll
        for (int i = 0; i < 9; ++i) M[i] = __shfl_sync(0xffffffffu, M[i], 0);
        int cnt = 0;
        for (int i = lane; i < n; i += 32) cnt += rs_is_inlier(M, src[i], dst[i], thr2) ? 1 : 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        const unsigned long long key = ((unsigned long long)(uint32_t)cnt << 32) | (unsigned long long)(0xFFFFFFFFu - hyp);   // ties -> lowest hypothesis
        mine = key > mine ? key : mine;
    }
    if (lane == 0 && mine) atomicMax(best, mine);
}

__device__ double rs_block_sum(double v, double *scratch)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) t += scratch[w];
    return t;
}

// One block: inliers of the winning hypothesis, then the least-squares re-fit on them.
template <int MODEL>
__global__ void __launch_bounds__(RS_THREADS) ransac_refit_kernel(const float2 *__restrict__ src, const float2 *__restrict__ dst, int n, double thr2,
                                                                   uint64_t seed, const unsigned long long *best, double *H_out,
                                                                   uint8_t *inliers, int32_t *n_inliers)
{
    __shared__ double s_M[9];
    __shared__ double s_red[RS_WARPS];
    __shared__ double s_acc[48];
    __shared__ int s_ok;
    const int tid = threadIdx.x;
    const unsigned long long key = *best;
    const uint32_t hyp = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
    constexpr int KS = MODEL ? 4 : 3;
    if (tid == 0) {
        double M[9];
        s_ok = (key != 0ull && rs_model_from_sample<MODEL>(src, dst, n, seed, hyp, M)) ? 1 : 0;
        for (int i = 0; i < 9; ++i) s_M[i] = M[i];
    }
    __syncthreads();
    if (!s_ok) {       // no valid hypothesis (fewer points than the minimal sample, all samples degenerate): cv2 returns None
        for (int i = tid; i < n; i += RS_THREADS) inliers[i] = 0;
        if (tid < 9) H_out[tid] = 0.0;
        if (tid == 0) *n_inliers = 0;
        return;
    }
    double M[9];
    for (int i = 0; i < 9; ++i) M[i] = s_M[i];
    // inlier mask of the winning hypothesis (what cv2 reports) + centroid / scale of the inliers for the normalisation
    double cnt = 0, sx = 0, sy = 0, su = 0, sv = 0;
    for (int i = tid; i < n; i += RS_THREADS) {
        const float2 a = src[i], b = dst[i];
        const bool in = rs_is_inlier(M, a, b, thr2);
        inliers[i] = in ? 1 : 0;
        if (in) { cnt += 1; sx += a.x; sy += a.y; su += b.x; sv += b.y; }
    }
    cnt = rs_block_sum(cnt, s_red); sx = rs_block_sum(sx, s_red); sy = rs_block_sum(sy, s_red);
    su = rs_block_sum(su, s_red); sv = rs_block_sum(sv, s_red);
    if (tid == 0) *n_inliers = (int)cnt;
    if (cnt < KS) {                                            // cannot happen for a valid hypothesis; keep the sample model
        if (tid < 9) H_out[tid] = M[tid];
        return;
    }
    const double mx = sx / cnt, my = sy / cnt, mu = su / cnt, mv = sv / cnt;
    double ds = 0, dd = 0;
    for (int i = tid; i < n; i += RS_THREADS)
        if (inliers[i]) {
            const float2 a = src[i], b = dst[i];
            ds += sqrt((a.x - mx) * (a.x - mx) + (a.y - my) * (a.y - my));
            dd += sqrt((b.x - mu) * (b.x - mu) + (b.y - mv) * (b.y - mv));
        }
    ds = rs_block_sum(ds, s_red); dd = rs_block_sum(dd, s_red);
    const double ks = (ds > 0) ? 1.41421356237309505 * cnt / ds : 1.0, kd = (dd > 0) ? 1.41421356237309505 * cnt / dd : 1.0;   // Hartley
    // model in normalised coordinates: Mn = Td * M * Ts^-1,  Ts = [ks 0 -ks mx; 0 ks -ks my; 0 0 1],  Td likewise
    double Mn[9];
    {
        // M * Ts^-1 : Ts^-1 = [1/ks 0 mx; 0 1/ks my; 0 0 1]
        double T[9];
        for (int r = 0; r < 3; ++r) {
            T[3 * r + 0] = M[3 * r + 0] / ks; T[3 * r + 1] = M[3 * r + 1] / ks;
            T[3 * r + 2] = M[3 * r + 0] * mx + M[3 * r + 1] * my + M[3 * r + 2];
        }
        for (int c = 0; c < 3; ++c) {
            Mn[0 + c] = kd * (T[0 + c] - mu * T[6 + c]);
            Mn[3 + c] = kd * (T[3 + c] - mv * T[6 + c]);
            Mn[6 + c] = T[6 + c];
        }
        const double inv = 1.0 / Mn[8];
        for (int i = 0; i < 9; ++i) Mn[i] *= inv;
    }
    if constexpr (MODEL == 0) {
        // affine least squares: minimise sum |A p + t - q|^2 -> normal equations on [x y 1] (normalised coordinates)
        double acc[12] = { 0 };   // Sxx Sxy Sx Syy Sy S1 | Sxu Syu Su | Sxv Syv Sv
        for (int i = tid; i < n; i += RS_THREADS)
            if (inliers[i]) {
                const float2 a = src[i], b = dst[i];
                const double x = ks * (a.x - mx), y = ks * (a.y - my), u = kd * (b.x - mu), v = kd * (b.y - mv);
                acc[0] += x * x; acc[1] += x * y; acc[2] += x; acc[3] += y * y; acc[4] += y; acc[5] += 1.0;
                acc[6] += x * u; acc[7] += y * u; acc[8] += u; acc[9] += x * v; acc[10] += y * v; acc[11] += v;
            }
        for (int k = 0; k < 12; ++k) { const double t = rs_block_sum(acc[k], s_red); if (tid == 0) s_acc[k] = t; }
        __syncthreads();
        if (tid == 0) {
            bool ok = true;
            for (int c = 0; c < 2 && ok; ++c) {
                double A[3][4] = { { s_acc[0], s_acc[1], s_acc[2], s_acc[6 + 3 * c] }, { s_acc[1], s_acc[3], s_acc[4], s_acc[7 + 3 * c] },
                                   { s_acc[2], s_acc[4], s_acc[5], s_acc[8 + 3 * c] } };
                double x[3];
                ok = rs_solve<3>(A, x);
                if (ok) { Mn[3 * c] = x[0]; Mn[3 * c + 1] = x[1]; Mn[3 * c + 2] = x[2]; }
            }
            Mn[6] = 0; Mn[7] = 0; Mn[8] = 1;
            for (int i = 0; i < 9; ++i) s_M[i] = Mn[i];
        }
        __syncthreads();
    } else {
        // Gauss-Newton on the reprojection error sum |proj(h, p) - q|^2 over the inliers, h33 = 1 (10 iterations, like cv2's LM cap)
        if (tid == 0) for (int i = 0; i < 9; ++i) s_M[i] = Mn[i];
        __syncthreads();
        for (int it = 0; it < 10; ++it) {
            double h[8];
            for (int i = 0; i < 8; ++i) h[i] = s_M[i];
            double JtJ[36], Jtr[8];
            for (int i = 0; i < 36; ++i) JtJ[i] = 0.0;
            for (int i = 0; i < 8; ++i) Jtr[i] = 0.0;
            for (int i = tid; i < n; i += RS_THREADS)
                if (inliers[i]) {
                    const float2 a = src[i], b = dst[i];
                    const double x = ks * (a.x - mx), y = ks * (a.y - my), u = kd * (b.x - mu), v = kd * (b.y - mv);
                    const double w = h[6] * x + h[7] * y + 1.0, iw = 1.0 / w;
                    const double pu = (h[0] * x + h[1] * y + h[2]) * iw, pv = (h[3] * x + h[4] * y + h[5]) * iw;
                    const double ru = pu - u, rv = pv - v;
                    const double Ju[8] = { x * iw, y * iw, iw, 0, 0, 0, -x * pu * iw, -y * pu * iw };
                    const double Jv[8] = { 0, 0, 0, x * iw, y * iw, iw, -x * pv * iw, -y * pv * iw };
                    int k = 0;
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
#pragma unroll
                        for (int c = r; c < 8; ++c) JtJ[k++] += Ju[r] * Ju[c] + Jv[r] * Jv[c];
                        Jtr[r] += Ju[r] * ru + Jv[r] * rv;
                    }
                }
            for (int k = 0; k < 36; ++k) { const double t = rs_block_sum(JtJ[k], s_red); if (tid == 0) s_acc[k] = t; }
            for (int k = 0; k < 8; ++k) { const double t = rs_block_sum(Jtr[k], s_red); if (tid == 0) s_acc[36 + k] = t; }
            __syncthreads();
            if (tid == 0) {
                double A[8][9];
                int k = 0;
                for (int r = 0; r < 8; ++r)
                    for (int c = r; c < 8; ++c) { A[r][c] = s_acc[k]; A[c][r] = s_acc[k]; ++k; }
                for (int r = 0; r < 8; ++r) { A[r][r] *= 1.0 + 1e-9; A[r][8] = -s_acc[36 + r]; }
                double dlt[8];
                if (rs_solve<8>(A, dlt))
                    for (int i = 0; i < 8; ++i) s_M[i] += dlt[i];
            }
            __syncthreads();
        }
    }
    // denormalise: H = Td^-1 * Mn * Ts
    if (tid == 0) {
        double Mr[9];
        for (int i = 0; i < 9; ++i) Mr[i] = s_M[i];
        double T[9];
        for (int r = 0; r < 3; ++r) {          // Mn * Ts
            T[3 * r + 0] = Mr[3 * r + 0] * ks; T[3 * r + 1] = Mr[3 * r + 1] * ks;
            T[3 * r + 2] = Mr[3 * r + 2] - ks * (Mr[3 * r + 0] * mx + Mr[3 * r + 1] * my);
        }
        double H[9];
        for (int c = 0; c < 3; ++c) {          // Td^-1 = [1/kd 0 mu; 0 1/kd mv; 0 0 1]
            H[0 + c] = T[0 + c] / kd + mu * T[6 + c];
            H[3 + c] = T[3 + c] / kd + mv * T[6 + c];
            H[6 + c] = T[6 + c];
        }
        const double inv = (MODEL && H[8] != 0.0) ? 1.0 / H[8] : 1.0;
        for (int i = 0; i < 9; ++i) H_out[i] = H[i] * inv;
        if (!MODEL) { H_out[6] = 0.0; H_out[7] = 0.0; H_out[8] = 1.0; }
    }
}

// ordered compaction of the ratio-test survivors (one block): out_src[j] = kp_query[i], out_dst[j] = kp_train[idx[i][0]]
__global__ void __launch_bounds__(1024) gather_matches_kernel(const float2 *__restrict__ kp_query, const float2 *__restrict__ kp_train,
                                                               const int32_t *__restrict__ idx, const uint8_t *__restrict__ keep, int nq,
                                                               float2 *out_src, float2 *out_dst, int32_t *out_query, int32_t *count)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < nq; i0 += 1024) {
        const int i = i0 + tid;
        const int k = (i < nq && keep[i] && idx[2 * i] >= 0) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, k);
        const int pre = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (k) {
            const int j = off + pre;
            out_src[j] = kp_query[i];
            out_dst[j] = kp_train[idx[2 * i]];
            if (out_query) out_query[j] = i;
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < 32; ++w) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
    if (tid == 0) *count = s_base;
}

// itertools.accumulate(H, np.matmul): out[0] = H[0], out[k] = out[k-1] @ H[k]  (fp64, left to right like the reference)
__global__ void chain_kernel(const double *__restrict__ H, int n, double *out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double A[9];
    for (int i = 0; i < 9; ++i) { A[i] = H[i]; out[i] = A[i]; }
    for (int k = 1; k < n; ++k) {
        const double *B = H + 9 * (size_t)k;
        double C[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                C[3 * r + c] = __dadd_rn(__dadd_rn(__dmul_rn(A[3 * r], B[c]), __dmul_rn(A[3 * r + 1], B[3 + c])), __dmul_rn(A[3 * r + 2], B[6 + c]));
        for (int i = 0; i < 9; ++i) { A[i] = C[i]; out[9 * (size_t)k + i] = C[i]; }
    }
}

}  // namespace mk

using namespace mk;

extern "C" int mk_gather_matches(const float *kp_query, const float *kp_train, const int32_t *idx, const uint8_t *keep, int nq,
                                 float *out_src, float *out_dst, int32_t *out_query, int32_t *count, void *stream)
{
    if (!kp_query || !kp_train || !idx || !keep || !out_src || !out_dst || !count || nq < 0) { set_error("mk_gather_matches: bad argument"); return MK_ERR_ARG; }
    gather_matches_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(kp_query), reinterpret_cast<const float2 *>(kp_train), idx, keep, nq,
                                                                 reinterpret_cast<float2 *>(out_src), reinterpret_cast<float2 *>(out_dst), out_query, count);
    MK_LAUNCH_CHECK("gather_matches_kernel");
    return MK_OK;
}

extern "C" size_t mk_ransac_workspace_bytes(void) { return 64; }

extern "C" int mk_ransac_transform(int model, const float *src_xy, const float *dst_xy, int n, const int32_t *n_dev, double thresh, int iters,
                                   uint64_t seed, double *H_out, uint8_t *inliers, int32_t *n_inliers, void *workspace, void *stream)
{
    if ((model != MK_MODEL_AFFINE && model != MK_MODEL_PERSPECTIVE) || !src_xy || !dst_xy || n < 0 || !H_out || !inliers || !n_inliers || !workspace ||
        !(thresh > 0.0) || iters <= 0 || n_dev) {
        set_error("mk_ransac_transform: bad argument (model 0|1, n >= 0, thresh > 0, iters > 0, n_dev reserved = NULL)");
        return MK_ERR_ARG;
    }
    if (!aligned(src_xy, 8) || !aligned(dst_xy, 8) || !aligned(H_out, 8) || !aligned(workspace, 8)) { set_error("mk_ransac_transform: alignment (8 B)"); return MK_ERR_ALIGN; }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *best = reinterpret_cast<unsigned long long *>(workspace);
    MK_CUDA_TRY(cudaMemsetAsync(best, 0, sizeof(unsigned long long), st));
    const float2 *s = reinterpret_cast<const float2 *>(src_xy), *d = reinterpret_cast<const float2 *>(dst_xy);
    const double thr2 = thresh * thresh;
    const int minimal = model == MK_MODEL_PERSPECTIVE ? 4 : 3;
    if (n >= minimal) {
        const int blocks = std::min((iters + RS_WARPS - 1) / RS_WARPS, 4 * std::max(1, stream_sm_count(st)));
        if (model == MK_MODEL_PERSPECTIVE) ransac_score_kernel<1><<<blocks, RS_THREADS, 0, st>>>(s, d, n, thr2, iters, seed, best);
        else ransac_score_kernel<0><<<blocks, RS_THREADS, 0, st>>>(s, d, n, thr2, iters, seed, best);
        MK_LAUNCH_CHECK("ransac_score_kernel");
    }
    if (model == MK_MODEL_PERSPECTIVE) ransac_refit_kernel<1><<<1, RS_THREADS, 0, st>>>(s, d, n, thr2, seed, best, H_out, inliers, n_inliers);
    else ransac_refit_kernel<0><<<1, RS_THREADS, 0, st>>>(s, d, n, thr2, seed, best, H_out, inliers, n_inliers);
    MK_LAUNCH_CHECK("ransac_refit_kernel");
    return MK_OK;
}

extern "C" int mk_chain_homographies(const double *H, int n, double *H_abs, void *stream)
{
    if (!H || !H_abs || n <= 0) { set_error("mk_chain_homographies: bad argument"); return MK_ERR_ARG; }
    chain_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(H, n, H_abs);
    MK_LAUNCH_CHECK("chain_kernel");
    return MK_OK;
}
