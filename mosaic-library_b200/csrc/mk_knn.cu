// mk_knn.cu -- brute-force k=2 nearest neighbours + Lowe ratio test (K1) for sm_100a.
//
// Replaces Matcher.knn_match -> cv2.BFMatcher.knnMatch(k=2) (src/mosaicking/registration.py:233-242)
// and the ratio test of src/mosaicking/mosaic.py:430.  cv::batchDistance inserts candidates into a
// sorted top-k list with strict comparisons while scanning train rows in ascending order, so equal
// distances keep the LOWER train index.  Here every candidate is a single integer key
// (score << idx_bits | train index): min / second-min over keys reproduces exactly that order no
// matter how the train rows are partitioned over lanes, warps and CTAs.
//
//   MK_NORM_HAMMING  XOR + POPC on 128-bit packed rows; one query per lane held in registers, train
//                    rows fetched with warp-uniform 128-bit loads; 32-bit keys.
//   MK_NORM_L2_U8    exact integer sum of squared byte differences (VABSDIFF4 + IDP4A), 64-bit keys,
//                    distance = sqrtf((float)sum) like cv2 (what registration.py:231 does for ORB).
//   MK_NORM_L2_F32   SIMT fp32 kernel that accumulates in cv::normL2Sqr_'s own order (4 lanes x 4
//                    accumulators, mul then add, fold upper onto lower) => bit-identical distances.
//                    The tensor-core path for integer-valued SIFT lives in mk_knn_tc.cu.
//
// Grid = (query blocks, train splits, pairs).  CTAs of one query block publish partial top-2 keys;
// the last one to finish (ticket counter) merges them and writes idx / dist / keep.
#include "mk_common.cuh"

#include <float.h>

namespace mk {

constexpr int KNN_THREADS = 256;
constexpr int KNN_WARPS = KNN_THREADS / 32;
constexpr int QB = 32;  // queries per CTA in the register-resident kernels

struct KnnParams {
    const uint8_t *q, *t;                     // single pair: separate bases; batched: both = desc
    const int32_t *set_off, *pairs, *out_off; // batched tables (device) or null
    int nq, nt;                               // single-pair sizes / batched upper bounds
    int dim;                                  // bytes (u8 norms) or floats (f32)
    unsigned long long stride;                // bytes between rows
    int splits;
    int nqb;                                  // query blocks per pair (grid.x)
    unsigned long long *part;                 // [pair][qblock][split][rows][2] partial keys
    int *tickets;                             // [pair][qblock]
    int32_t *idx;
    float *dist;
    uint8_t *keep;
    int32_t *n_keep;
    double ratio;
    const int *guard;   // if non-null the kernel only runs when the tensor-core path did not produce the result
    const int32_t *tc_count;  // guarded runs: the tensor-core kernel's survivor counts, published to n_keep only if its result stands
};

// Guarded SIMT pass (float32 rows that went through the tensor-core kernel first).  The TC kernel counted its survivors
// into a scratch counter, never into n_keep: if its result stands, one thread per pair publishes that count and the pass
// is over; otherwise n_keep is still zero and this pass counts the survivors of the recomputed call itself.
__device__ __forceinline__ bool knn_guard_skip(const KnnParams &p)
{
    if (!p.guard) return false;
    // guarded passes are launched with programmatic stream serialization: their CTAs may be resident before the tensor-core
    // kernel has finished, and may read its verdict only after this wait (a no-op under a plain launch)
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    if (!knn_tc_result_ok(p.guard)) return false;
    if (p.tc_count && p.n_keep && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) p.n_keep[blockIdx.z] = p.tc_count[blockIdx.z];
    return true;
}

// tensor-core path (mk_knn_tc.cu)
size_t knn_tc_workspace_bytes(int norm, int nq, int nt, int dim);
size_t knn_tc_batched_workspace_bytes(int norm, int npairs, int max_nq, int max_nt, int dim, int total_rows);
bool knn_tc_supported(int norm, int nq, int nt, int dim);
int knn_tc_run(int norm, const void *q, int nq, const void *t, int nt, int dim, size_t stride, const int32_t *set_off,
               const int32_t *pairs, int npairs, const int32_t *out_off, int total_rows, int32_t *idx, float *dist, double ratio,
               uint8_t *keep, int32_t *n_keep, void *workspace, size_t workspace_bytes, cudaStream_t st, const int **flags_out,
               const int32_t **tc_count_out, int *simt_tickets, size_t simt_ticket_bytes);

struct Problem {
    const uint8_t *q, *t;
    int nq, nt, out0;
};

__device__ __forceinline__ Problem load_problem(const KnnParams &p, int z)
{
    Problem pr;
    if (p.pairs) {
        const int qs = p.pairs[2 * z], ts = p.pairs[2 * z + 1];
        const int q0 = p.set_off[qs], t0 = p.set_off[ts];
        pr.q = p.q + (size_t)q0 * p.stride;
        pr.t = p.t + (size_t)t0 * p.stride;
        pr.nq = p.set_off[qs + 1] - q0;
        pr.nt = p.set_off[ts + 1] - t0;
        pr.out0 = p.out_off[z];
    } else {
        pr.q = p.q; pr.t = p.t; pr.nq = p.nq; pr.nt = p.nt; pr.out0 = 0;
    }
    return pr;
}

__device__ __forceinline__ uint4 ldg128(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }

// ---------------------------------------------------------------------------------------------
// Metric policies for the register-resident kernel (NV = 16-byte vectors per row)
// ---------------------------------------------------------------------------------------------
template <int NV>
struct Hamming {
    using Key = uint32_t;
    static constexpr int IDX_BITS = 22;  // distance <= 512 needs 10 bits
    static constexpr Key NONE = 0xFFFFFFFFu;
    // 3:2 carry-save compression before counting: POPC runs on the 16-lane XU pipe (the kernel's limiter), LOP3 on the
    // ALU pipe that has slack, so every 8 XOR words are folded by three full adders into 2 "ones" + 3 "twos" words:
    // 5 POPC instead of 8 per 256 bits, exact.
    static __device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }
    static __device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a | b)); }
    static __device__ __forceinline__ uint32_t count8(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4,
                                                      uint32_t x5, uint32_t x6, uint32_t x7)
    {
        const uint32_t s0 = xor3(x0, x1, x2), c0 = maj3(x0, x1, x2);
        const uint32_t s1 = xor3(x3, x4, x5), c1 = maj3(x3, x4, x5);
        const uint32_t s2 = xor3(s0, s1, x6), c2 = maj3(s0, s1, x6);
        return __popc(s2) + __popc(x7) + 2u * (__popc(c0) + __popc(c1) + __popc(c2));   // a 4th adder (4 POPC) makes the ALU pipe the limiter: measured slower
    }
    static __device__ __forceinline__ uint32_t score(const uint4 (&a)[NV], const uint4 (&b)[NV])
    {
        uint32_t s = 0;
        if constexpr (NV % 2 == 0) {
#pragma unroll
            for (int v = 0; v < NV; v += 2)
                s += count8(a[v].x ^ b[v].x, a[v].y ^ b[v].y, a[v].z ^ b[v].z, a[v].w ^ b[v].w, a[v + 1].x ^ b[v + 1].x,
                            a[v + 1].y ^ b[v + 1].y, a[v + 1].z ^ b[v + 1].z, a[v + 1].w ^ b[v + 1].w);
        } else {
#pragma unroll
            for (int v = 0; v < NV; ++v)
                s += __popc(a[v].x ^ b[v].x) + __popc(a[v].y ^ b[v].y) + __popc(a[v].z ^ b[v].z) + __popc(a[v].w ^ b[v].w);
        }
        return s;
    }
    static __device__ __forceinline__ Key make(uint32_t s, int j) { return s * (1u << IDX_BITS) + (uint32_t)j; }   // one IMAD (FMA pipe)
    static __device__ __forceinline__ void decode(Key k, int &idx, float &d)
    {
        if (k == NONE) { idx = -1; d = FLT_MAX; }
        else { idx = (int)(k & ((1u << IDX_BITS) - 1u)); d = (float)(k >> IDX_BITS); }
    }
};

template <int NV>
struct L2U8 {
    using Key = unsigned long long;
    static constexpr Key NONE = ~0ull;
    static __device__ __forceinline__ uint32_t sq(uint32_t a, uint32_t b, uint32_t acc)
    {
        const uint32_t d = __vabsdiffu4(a, b);
        return __dp4a(d, d, acc);
    }
    static __device__ __forceinline__ uint32_t score(const uint4 (&a)[NV], const uint4 (&b)[NV])
    {
        uint32_t s = 0;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            s = sq(a[v].x, b[v].x, s); s = sq(a[v].y, b[v].y, s);
            s = sq(a[v].z, b[v].z, s); s = sq(a[v].w, b[v].w, s);
        }
        return s;
    }
    static __device__ __forceinline__ Key make(uint32_t s, int j) { return ((Key)s << 32) | (uint32_t)j; }
    static __device__ __forceinline__ void decode(Key k, int &idx, float &d)
    {
        if (k == NONE) { idx = -1; d = FLT_MAX; }
        else { idx = (int)(uint32_t)k; d = __fsqrt_rn((float)(uint32_t)(k >> 32)); }
    }
};

// keys of the fp32 kernel: (bits of sqrtf(sum) << 32) | idx -- cv2 compares the rooted distances
struct F32Key {
    using Key = unsigned long long;
    static constexpr Key NONE = ~0ull;
    static __device__ __forceinline__ void decode(Key k, int &idx, float &d)
    {
        if (k == NONE) { idx = -1; d = FLT_MAX; }
        else { idx = (int)(uint32_t)k; d = __uint_as_float((uint32_t)(k >> 32)); }
    }
};

template <class Key>
__device__ __forceinline__ void top2_push(Key &k1, Key &k2, Key key)
{
    const Key mx = key > k1 ? key : k1;
    k1 = key < k1 ? key : k1;
    k2 = mx < k2 ? mx : k2;
}

// Publish this CTA's top-2 keys for `rows` queries; the last CTA of the query block merges all
// splits and writes the final outputs.  s_k1/s_k2 hold the CTA result for row = threadIdx.x < rows.
template <class Policy>
__device__ __forceinline__ void finish_block(const KnnParams &p, const Problem &pr, int z, int qblk, int rows_per_block,
                                             typename Policy::Key k1, typename Policy::Key k2, bool row_valid, int row,
                                             int *s_flag)
{
    using Key = typename Policy::Key;
    const int tid = threadIdx.x;
    const bool is_row_thread = tid < rows_per_block;
    if (p.splits > 1) {
        const size_t blk = ((size_t)z * p.nqb + qblk);
        unsigned long long *base = p.part + (blk * p.splits) * rows_per_block * 2;
        if (is_row_thread) {
            unsigned long long *mine = base + ((size_t)blockIdx.y * rows_per_block + tid) * 2;
            mine[0] = (unsigned long long)k1; mine[1] = (unsigned long long)k2;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const int t = atomicAdd(&p.tickets[blk], 1);
            *s_flag = (t == p.splits - 1);
        }
        __syncthreads();
        if (!*s_flag) return;
        __threadfence();
        if (is_row_thread) {
            k1 = Policy::NONE; k2 = Policy::NONE;
            for (int s = 0; s < p.splits; ++s) {
                const unsigned long long *o = base + ((size_t)s * rows_per_block + tid) * 2;
                const Key a = (Key)__ldcg(o), b = (Key)__ldcg(o + 1);
                if (a != Policy::NONE) top2_push(k1, k2, a);
                if (b != Policy::NONE) top2_push(k1, k2, b);
            }
        }
        if (tid == 0) p.tickets[blk] = 0;  // leave the counter clean for the next call
    }
    bool kept = false;
    if (is_row_thread && row_valid) {
        int i1, i2; float d1, d2;
        Policy::decode(k1, i1, d1);
        Policy::decode(k2, i2, d2);
        const size_t o = (size_t)(pr.out0 + row);
        p.idx[2 * o] = i1; p.idx[2 * o + 1] = i2;
        p.dist[2 * o] = d1; p.dist[2 * o + 1] = d2;
        // mosaic.py:430: len(m) == 2 and m[0].distance < ratio * m[1].distance  (float64 compare)
        kept = (i2 >= 0) && (p.ratio > 0.0) && ((double)d1 < p.ratio * (double)d2);
        if (p.keep) p.keep[o] = kept ? 1 : 0;
    }
    if (p.n_keep && tid < ((rows_per_block + 31) & ~31)) {
        const unsigned b = __ballot_sync(0xffffffffu, kept);
        if ((tid & 31) == 0 && b) atomicAdd(&p.n_keep[z], __popc(b));
    }
}

template <class Metric, int NV>
__global__ void __launch_bounds__(KNN_THREADS) knn2_reg_kernel(const __grid_constant__ KnnParams p)
{
    using Key = typename Metric::Key;
    __shared__ Key s_part[KNN_WARPS][QB][2];
    __shared__ int s_flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int z = blockIdx.z, qblk = blockIdx.x;
    if (knn_guard_skip(p)) return;
    const Problem pr = load_problem(p, z);
    if (qblk * QB >= pr.nq) return;
    const int qi = qblk * QB + lane;
    const bool qvalid = qi < pr.nq;

    uint4 qv[NV];
    {
        const uint8_t *row = pr.q + (size_t)(qvalid ? qi : pr.nq - 1) * p.stride;
#pragma unroll
        for (int v = 0; v < NV; ++v) qv[v] = ldg128(row + 16 * v);
    }
    // contiguous train range of this CTA, then of this warp
    const int chunk = (pr.nt + p.splits - 1) / p.splits;
    const int c_begin = min(pr.nt, (int)blockIdx.y * chunk), c_end = min(pr.nt, c_begin + chunk);
    const int sub = (c_end - c_begin + KNN_WARPS - 1) / KNN_WARPS;
    const int w_begin = min(c_end, c_begin + warp * sub), w_end = min(c_end, w_begin + sub);

    Key k1 = Metric::NONE, k2 = Metric::NONE;
    const uint8_t *trow = pr.t + (size_t)w_begin * p.stride;
#pragma unroll 4
    for (int j = w_begin; j < w_end; ++j, trow += p.stride) {
        uint4 tv[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) tv[v] = ldg128(trow + 16 * v);
        top2_push(k1, k2, Metric::make(Metric::score(qv, tv), j));
    }
    s_part[warp][lane][0] = k1; s_part[warp][lane][1] = k2;
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int w = 1; w < KNN_WARPS; ++w) {
            const Key a = s_part[w][lane][0], b = s_part[w][lane][1];
            if (a != Metric::NONE) top2_push(k1, k2, a);
            if (b != Metric::NONE) top2_push(k1, k2, b);
        }
    }
    finish_block<Metric>(p, pr, z, qblk, QB, k1, k2, qvalid, qi, &s_flag);
}

// ---------------------------------------------------------------------------------------------
// fp32 L2, SIMT, bit-identical to cv::normL2Sqr_ (baseline SSE build of cv2 4.13):
//   sum over k of (a-b)^2 with lane = k % 4, accumulator = (k / 4) % 4, 16 elements per step,
//   lanes[l] = ((acc0 + acc1) + acc2) + acc3,  d = (l0 + l2) + (l1 + l3),  scalar tail,  sqrtf.
// CTA = 64 queries x 64-train tiles; thread = 4 x 4 pairs; the whole K extent sits in shared
// memory k-major so every chain is walked with two broadcast / conflict-free LDS.128 per k.
// ---------------------------------------------------------------------------------------------
constexpr int FT = 64;  // tile edge (queries and trains)

__device__ __forceinline__ void f32_chain(const float *__restrict__ Qs, const float *__restrict__ Ts, int ty, int tx,
                                          int kstart, int Dm, float (&acc)[4][4])
{
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k = kstart; k < Dm; k += 16) {
        const float4 a = *reinterpret_cast<const float4 *>(Qs + (size_t)k * FT + ty * 4);
        const float4 b = *reinterpret_cast<const float4 *>(Ts + (size_t)k * FT + tx * 4);
        const float av[4] = { a.x, a.y, a.z, a.w }, bv[4] = { b.x, b.y, b.z, b.w };
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d = __fsub_rn(av[i], bv[j]);
                acc[i][j] = __fadd_rn(__fmul_rn(d, d), acc[i][j]);
            }
    }
}

// one SIMD lane l of cv's loop: ((acc0 + acc1) + acc2) + acc3
__device__ __forceinline__ void f32_lane(const float *__restrict__ Qs, const float *__restrict__ Ts, int ty, int tx, int l,
                                         int Dm, float (&lanev)[4][4])
{
    f32_chain(Qs, Ts, ty, tx, l, Dm, lanev);
#pragma unroll 1
    for (int v = 1; v < 4; ++v) {
        float acc[4][4];
        f32_chain(Qs, Ts, ty, tx, 4 * v + l, Dm, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) lanev[i][j] = __fadd_rn(lanev[i][j], acc[i][j]);
    }
}

__global__ void __launch_bounds__(KNN_THREADS) knn2_f32_kernel(const __grid_constant__ KnnParams p)
{
    using Key = F32Key::Key;
    extern __shared__ __align__(16) float s_f32[];
    __shared__ Key s_keys[16][FT][2];
    __shared__ int s_flag;
    const int D = p.dim;
    float *Qs = s_f32;                    // [D][FT]
    float *Ts = s_f32 + (size_t)D * FT;   // [D][FT]

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int z = blockIdx.z, qblk = blockIdx.x;
    if (knn_guard_skip(p)) return;
    const Problem pr = load_problem(p, z);
    if (qblk * FT >= pr.nq) return;
    const int q0 = qblk * FT;
    const int fstride = (int)(p.stride >> 2);

    for (int i = tid; i < FT * D; i += KNN_THREADS) {  // queries, transposed to k-major
        const int r = i / D, k = i - r * D;
        const int qi = min(q0 + r, pr.nq - 1);
        Qs[(size_t)k * FT + r] = __ldg(reinterpret_cast<const float *>(pr.q) + (size_t)qi * fstride + k);
    }
    const int chunk = (((pr.nt + p.splits - 1) / p.splits) + FT - 1) / FT * FT;
    const int c_begin = min(pr.nt, (int)blockIdx.y * chunk), c_end = min(pr.nt, c_begin + chunk);
    const int Dm = D & ~15;

    Key k1[4], k2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { k1[i] = F32Key::NONE; k2[i] = F32Key::NONE; }

    for (int t0 = c_begin; t0 < c_end; t0 += FT) {
        __syncthreads();
        for (int i = tid; i < FT * D; i += KNN_THREADS) {
            const int r = i / D, k = i - r * D;
            const int ti = min(t0 + r, pr.nt - 1);
            Ts[(size_t)k * FT + r] = __ldg(reinterpret_cast<const float *>(pr.t) + (size_t)ti * fstride + k);
        }
        __syncthreads();
        float tot[4][4], tmp[4][4];
        f32_lane(Qs, Ts, ty, tx, 0, Dm, tot);
        f32_lane(Qs, Ts, ty, tx, 2, Dm, tmp);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) tot[i][j] = __fadd_rn(tot[i][j], tmp[i][j]);   // l0 + l2
        float s13[4][4];
        f32_lane(Qs, Ts, ty, tx, 1, Dm, s13);
        f32_lane(Qs, Ts, ty, tx, 3, Dm, tmp);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) tot[i][j] = __fadd_rn(tot[i][j], __fadd_rn(s13[i][j], tmp[i][j]));
        for (int k = Dm; k < D; ++k) {  // scalar tail: d += t*t
            const float4 a = *reinterpret_cast<const float4 *>(Qs + (size_t)k * FT + ty * 4);
            const float4 b = *reinterpret_cast<const float4 *>(Ts + (size_t)k * FT + tx * 4);
            const float av[4] = { a.x, a.y, a.z, a.w }, bv[4] = { b.x, b.y, b.z, b.w };
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float d = __fsub_rn(av[i], bv[j]);
                    tot[i][j] = __fadd_rn(tot[i][j], __fmul_rn(d, d));
                }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int tj = t0 + tx * 4 + j;
            if (tj >= c_end) continue;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float s = __fsqrt_rn(tot[i][j]);
                if (s < FLT_MAX) top2_push(k1[i], k2[i], ((Key)__float_as_uint(s) << 32) | (uint32_t)tj);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { s_keys[tx][ty * 4 + i][0] = k1[i]; s_keys[tx][ty * 4 + i][1] = k2[i]; }
    __syncthreads();
    Key a1 = F32Key::NONE, a2 = F32Key::NONE;
    if (tid < FT) {
        for (int w = 0; w < 16; ++w) {
            const Key a = s_keys[w][tid][0], b = s_keys[w][tid][1];
            if (a != F32Key::NONE) top2_push(a1, a2, a);
            if (b != F32Key::NONE) top2_push(a1, a2, b);
        }
    }
    finish_block<F32Key>(p, pr, z, qblk, FT, a1, a2, q0 + tid < pr.nq, q0 + tid, &s_flag);
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
struct KnnPlan {
    int rows, nqb, splits;
    size_t ticket_bytes, part_bytes;
};

static KnnPlan make_plan(int norm, int npairs, int max_nq, int max_nt)
{
    KnnPlan pl;
    pl.rows = (norm == MK_NORM_L2_F32) ? FT : QB;
    pl.nqb = max_nq > 0 ? (max_nq + pl.rows - 1) / pl.rows : 0;
    const long base = (long)npairs * pl.nqb;
    int max_splits, target;
    if (norm == MK_NORM_L2_F32) { max_splits = max_nt / (2 * FT); target = 148 * 2; }
    else { max_splits = max_nt / (KNN_WARPS * 32); target = 148 * 8; }
    if (max_splits < 1) max_splits = 1;
    long want = base > 0 ? (target + base - 1) / base : 1;
    pl.splits = (int)(want < 1 ? 1 : (want > max_splits ? max_splits : want));
    pl.ticket_bytes = (((size_t)(base > 0 ? base : 1) * sizeof(int)) + 255) & ~(size_t)255;
    pl.part_bytes = pl.splits > 1 ? (size_t)base * pl.splits * pl.rows * 2 * sizeof(unsigned long long) : 0;
    return pl;
}

static int check_knn_args(int norm, int dim, size_t stride)
{
    if (norm != MK_NORM_HAMMING && norm != MK_NORM_L2_U8 && norm != MK_NORM_L2_F32) { set_error("mk_knn2: unknown norm %d", norm); return MK_ERR_ARG; }
    if (dim <= 0) { set_error("mk_knn2: dim must be positive"); return MK_ERR_ARG; }
    if (stride & 15) { set_error("mk_knn2: row_stride must be a multiple of 16 bytes"); return MK_ERR_ALIGN; }
    if (norm == MK_NORM_L2_F32) {
        if (dim > 256) { set_error("mk_knn2: float descriptors wider than 256 are not supported"); return MK_ERR_ARG; }
        if (stride < (size_t)dim * 4) { set_error("mk_knn2: row_stride smaller than a row"); return MK_ERR_ARG; }
    } else {
        if (dim > 64) { set_error("mk_knn2: binary/uint8 descriptors wider than 64 bytes are not supported"); return MK_ERR_ARG; }
        if (stride < (size_t)((dim + 15) & ~15)) { set_error("mk_knn2: rows must be zero-padded to a multiple of 16 bytes"); return MK_ERR_ARG; }
    }
    return MK_OK;
}

template <class Metric, int NV>
static void launch_reg(dim3 grid, cudaStream_t st, const KnnParams &P)
{
    knn2_reg_kernel<Metric, NV><<<grid, KNN_THREADS, 0, st>>>(P);
}

static int run_knn(int norm, KnnParams &P, int npairs, int max_nq, int max_nt, void *workspace, size_t workspace_bytes,
                   cudaStream_t st, bool zero_n_keep = true, bool tickets_zeroed = false)
{
    const KnnPlan pl = make_plan(norm, npairs, max_nq, max_nt);
    if (workspace_bytes < pl.ticket_bytes + pl.part_bytes || !workspace) { set_error("mk_knn2: workspace too small (%zu < %zu)", workspace_bytes, pl.ticket_bytes + pl.part_bytes); return MK_ERR_WORKSPACE; }
    if (!aligned(workspace, 16)) { set_error("mk_knn2: workspace must be 16-byte aligned"); return MK_ERR_ALIGN; }
    if (max_nt >= (1 << 22) && norm == MK_NORM_HAMMING) { set_error("mk_knn2: more than 4M train rows per set"); return MK_ERR_ARG; }
    P.tickets = reinterpret_cast<int *>(workspace);
    P.part = reinterpret_cast<unsigned long long *>(reinterpret_cast<uint8_t *>(workspace) + pl.ticket_bytes);
    P.splits = pl.splits; P.nqb = pl.nqb;
    if (P.n_keep && zero_n_keep) MK_CUDA_TRY(cudaMemsetAsync(P.n_keep, 0, sizeof(int32_t) * npairs, st));
    if (pl.nqb == 0 || npairs == 0) return MK_OK;
    if (pl.splits > 1 && !tickets_zeroed) MK_CUDA_TRY(cudaMemsetAsync(P.tickets, 0, pl.ticket_bytes, st));
    dim3 grid(pl.nqb, pl.splits, npairs);
    if (norm == MK_NORM_L2_F32) {
        const size_t smem = (size_t)2 * FT * P.dim * sizeof(float);
        static std::atomic<int> configured{-1};   // device the attribute was last set on (cudaFuncSetAttribute is per device)
        if (smem > 48 * 1024) {
            int dev = 0;
            cudaGetDevice(&dev);
            if (configured.load(std::memory_order_acquire) != dev) {
                MK_CUDA_TRY(cudaFuncSetAttribute(knn2_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * FT * 256 * sizeof(float))));
                configured.store(dev, std::memory_order_release);
            }
        }
        if (P.guard) {       // guarded pass behind the tensor-core kernel: launch latency hidden under that kernel's tail
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = grid; cfg.blockDim = dim3(KNN_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            MK_CUDA_TRY(cudaLaunchKernelEx(&cfg, knn2_f32_kernel, P));
        } else {
            knn2_f32_kernel<<<grid, KNN_THREADS, smem, st>>>(P);
        }
    } else {
        const int nv = (P.dim + 15) / 16;
        if (norm == MK_NORM_HAMMING) {
            switch (nv) {
                case 1: launch_reg<Hamming<1>, 1>(grid, st, P); break;
                case 2: launch_reg<Hamming<2>, 2>(grid, st, P); break;
                case 3: launch_reg<Hamming<3>, 3>(grid, st, P); break;
                default: launch_reg<Hamming<4>, 4>(grid, st, P); break;
            }
        } else {
            switch (nv) {
                case 1: launch_reg<L2U8<1>, 1>(grid, st, P); break;
                case 2: launch_reg<L2U8<2>, 2>(grid, st, P); break;
                case 3: launch_reg<L2U8<3>, 3>(grid, st, P); break;
                default: launch_reg<L2U8<4>, 4>(grid, st, P); break;
            }
        }
    }
    MK_LAUNCH_CHECK("knn2 kernel");
    return MK_OK;
}

}  // namespace mk

using namespace mk;

// MK_NORM_HAMMING_TC is MK_NORM_HAMMING with the optional tensor-core engine: everywhere except the engine choice it is Hamming.
static inline int base_norm(int norm) { return norm == MK_NORM_HAMMING_TC ? MK_NORM_HAMMING : norm; }

// Problems large enough to amortise the extra launches go through the tensor-core GEMM.
static bool use_tc(int norm, int nq, int nt, int dim)
{
    return knn_tc_supported(norm, nq, nt, dim) && (long)nq * nt >= 256L * 1024;
}

extern "C" size_t mk_knn2_workspace_bytes(int norm, int nq, int nt, int dim)
{
    const KnnPlan pl = make_plan(base_norm(norm), 1, nq, nt);
    size_t bytes = pl.ticket_bytes + pl.part_bytes;
    if (use_tc(norm, nq, nt, dim)) bytes = ((bytes + 255) & ~(size_t)255) + knn_tc_workspace_bytes(norm, nq, nt, dim);
    return bytes;
}

extern "C" int mk_knn2_uses_tensor_cores(int norm, int nq, int nt, int dim) { return use_tc(norm, nq, nt, dim) ? 1 : 0; }

static bool use_tc_batched(int norm, int npairs, int max_nq, int max_nt, int dim)
{
    return knn_tc_supported(norm, max_nq, max_nt, dim) && (long)npairs * max_nq * max_nt >= 256L * 1024;
}

extern "C" size_t mk_knn2_batched_workspace_bytes(int norm, int npairs, int max_nq, int max_nt, int dim, int total_rows)
{
    const KnnPlan pl = make_plan(base_norm(norm), npairs, max_nq, max_nt);
    size_t bytes = pl.ticket_bytes + pl.part_bytes;
    if (use_tc_batched(norm, npairs, max_nq, max_nt, dim))
        bytes = ((bytes + 255) & ~(size_t)255) + knn_tc_batched_workspace_bytes(norm, npairs, max_nq, max_nt, dim, total_rows);
    return bytes;
}

extern "C" int mk_knn2(int norm, const void *q, int nq, const void *t, int nt, int dim, size_t row_stride, int32_t *idx,
                       float *dist, double ratio, uint8_t *keep, int32_t *n_keep, void *workspace, size_t workspace_bytes,
                       void *stream)
{
    const int bn = base_norm(norm);
    int rc = check_knn_args(bn, dim, row_stride);
    if (rc != MK_OK) return rc;
    if (nq < 0 || nt < 0 || (nq > 0 && (!q || !idx || !dist)) || (nt > 0 && !t)) { set_error("mk_knn2: null pointer or negative size"); return MK_ERR_ARG; }
    if ((q && !aligned(q, 16)) || (t && !aligned(t, 16))) { set_error("mk_knn2: descriptor base must be 16-byte aligned"); return MK_ERR_ALIGN; }
    KnnParams P{};
    P.q = static_cast<const uint8_t *>(q); P.t = static_cast<const uint8_t *>(t);
    P.nq = nq; P.nt = nt; P.dim = dim; P.stride = row_stride;
    P.idx = idx; P.dist = dist; P.keep = keep; P.n_keep = n_keep; P.ratio = ratio;
    if (use_tc(norm, nq, nt, dim)) {
        // [SIMT scratch][tensor-core scratch]; the SIMT kernel only runs if the prep pass found values that
        // bf16 cannot hold exactly (never the case for uint8 rows, so it is not even launched for them)
        const KnnPlan pl = make_plan(bn, 1, nq, nt);
        const size_t simt_bytes = ((pl.ticket_bytes + pl.part_bytes) + 255) & ~(size_t)255;
        if (workspace_bytes < simt_bytes || !workspace) { set_error("mk_knn2: workspace too small"); return MK_ERR_WORKSPACE; }
        const int *flags = nullptr;
        const int32_t *tc_count = nullptr;
        // the guarded SIMT pass needs zeroed tickets at the front of the workspace: the TC prep kernel zeroes them (no memset node)
        const bool guarded = !(norm == MK_NORM_HAMMING_TC || (norm == MK_NORM_L2_U8 && dim <= 64));
        rc = knn_tc_run(norm, q, nq, t, nt, dim, row_stride, nullptr, nullptr, 1, nullptr, 0, idx, dist, ratio, keep, n_keep,
                        static_cast<uint8_t *>(workspace) + simt_bytes, workspace_bytes - simt_bytes, (cudaStream_t)stream, &flags, &tc_count,
                        static_cast<int *>(workspace), guarded && pl.splits > 1 ? pl.ticket_bytes : 0);
        // uint8 rows: dim*255^2 < 2^22, bit counts <= 256: the TC result always stands
        if (rc != MK_OK || norm == MK_NORM_HAMMING_TC || (norm == MK_NORM_L2_U8 && dim <= 64)) return rc;
        P.guard = flags; P.tc_count = tc_count;
        return run_knn(bn, P, 1, nq, nt, workspace, simt_bytes, (cudaStream_t)stream, /*zero_n_keep=*/false, /*tickets_zeroed=*/true);
    }
    return run_knn(bn, P, 1, nq, nt, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int mk_knn2_batched(int norm, const void *desc, int total_rows, int dim, size_t row_stride, const int32_t *set_off,
                               const int32_t *pairs, int npairs, const int32_t *out_off, int max_nq, int max_nt, int32_t *idx,
                               float *dist, double ratio, uint8_t *keep, int32_t *n_keep, void *workspace,
                               size_t workspace_bytes, void *stream)
{
    const int bn = base_norm(norm);
    int rc = check_knn_args(bn, dim, row_stride);
    if (rc != MK_OK) return rc;
    if (npairs < 0 || max_nq < 0 || max_nt < 0 || (npairs > 0 && (!desc || !set_off || !pairs || !out_off || !idx || !dist))) {
        set_error("mk_knn2_batched: null pointer or negative size");
        return MK_ERR_ARG;
    }
    if (npairs > 65535) { set_error("mk_knn2_batched: at most 65535 pairs per call"); return MK_ERR_ARG; }
    if (desc && !aligned(desc, 16)) { set_error("mk_knn2_batched: descriptor base must be 16-byte aligned"); return MK_ERR_ALIGN; }
    KnnParams P{};
    P.q = static_cast<const uint8_t *>(desc); P.t = P.q;
    P.set_off = set_off; P.pairs = pairs; P.out_off = out_off;
    P.nq = max_nq; P.nt = max_nt; P.dim = dim; P.stride = row_stride;
    P.idx = idx; P.dist = dist; P.keep = keep; P.n_keep = n_keep; P.ratio = ratio;
    if (total_rows > 0 && use_tc_batched(norm, npairs, max_nq, max_nt, dim)) {
        const KnnPlan pl = make_plan(bn, npairs, max_nq, max_nt);
        const size_t simt_bytes = ((pl.ticket_bytes + pl.part_bytes) + 255) & ~(size_t)255;
        if (workspace_bytes < simt_bytes || !workspace) { set_error("mk_knn2_batched: workspace too small"); return MK_ERR_WORKSPACE; }
        const int *flags = nullptr;
        const int32_t *tc_count = nullptr;
        rc = knn_tc_run(norm, desc, max_nq, nullptr, max_nt, dim, row_stride, set_off, pairs, npairs, out_off, total_rows, idx, dist,
                        ratio, keep, n_keep, static_cast<uint8_t *>(workspace) + simt_bytes, workspace_bytes - simt_bytes,
                        (cudaStream_t)stream, &flags, &tc_count, static_cast<int *>(workspace),
                        !(norm == MK_NORM_HAMMING_TC || (norm == MK_NORM_L2_U8 && dim <= 64)) && pl.splits > 1 ? pl.ticket_bytes : 0);
        if (rc != MK_OK || norm == MK_NORM_HAMMING_TC || (norm == MK_NORM_L2_U8 && dim <= 64)) return rc;
        P.guard = flags; P.tc_count = tc_count;
        return run_knn(bn, P, npairs, max_nq, max_nt, workspace, simt_bytes, (cudaStream_t)stream, /*zero_n_keep=*/false, /*tickets_zeroed=*/true);
    }
    return run_knn(bn, P, npairs, max_nq, max_nt, workspace, workspace_bytes, (cudaStream_t)stream);
}
