// mk_knn_tc.cu -- L2 kNN-2 as a tcgen05 / TMA descriptor GEMM (K2-TC) for sm_100a.
//
// L2 matching is a dense contraction: d^2(q,t) = |q|^2 + |t|^2 - 2 q.t .  OpenCV's SIFT descriptors
// are integer-valued floats in [0,255] and the reference matches ORB bytes under NORM_L2 as well
// (src/mosaicking/registration.py:230-242), so bf16 operands are exact, every partial sum is an
// integer < 2^24 and the fp32 accumulator in tensor memory holds q.t EXACTLY: the distances equal
// cv2's sqrtf((float)sum) bit for bit.
//
//   prep kernel   rows -> bf16 [n][KP] (KP = K padded to 64), |row|^2 in fp32, "all integers in
//                 [0,255]" flag.  Rows that are not integer-valued make the TC kernel return at once
//                 and the bit-exact SIMT kernel (mk_knn.cu) takes over -- decided on the device.
//   TC kernel     CTA = 128 queries (UMMA M=128), loops over train tiles of 256 rows (UMMA N=256).
//                 warp 0   TMA producer  : cp.async.bulk.tensor 2-D boxes of 64 x rows bf16, 128-B swizzle
//                 warp 1   MMA issuer    : one thread issues tcgen05.mma (kind::f16, K=16 per instr),
//                                          accumulators 128 x 256 fp32 in TMEM, two accumulator stages
//                 warps 2-5 epilogue     : tcgen05.ld 32x32b -> one query row per thread, d^2 = (|q|^2 - 2S) + |t|^2,
//                                          pre-filter on d^2, exact top-2 on (sqrtf bits << 32 | index) keys
//                 Train splits are merged by the ticket scheme of mk_knn.cu (finish_block).
#include "mk_common.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <float.h>
#include <stdlib.h>

#include <mutex>

namespace mk {

// ---- shared with mk_knn.cu --------------------------------------------------------------------
struct KnnParams;  // defined in mk_knn.cu (same translation-unit-independent layout declared below)

constexpr int TC_M = 128;          // queries per CTA
constexpr int TC_N = 256;          // train rows per tile
constexpr int TC_KA = 64;          // bf16 elements per 128-byte swizzle atom
constexpr int TC_STAGES = 2;       // B tile ring when a row is two swizzle atoms (256 B); twice as deep for 128-byte rows
constexpr int TC_MAX_STAGES = 2 * TC_STAGES;
constexpr int TC_EPI_WARPS = 16;   // two teams of 8 warps; team t scans the tiles i = t (mod 2), i.e. accumulator stage t
constexpr int TC_TEAMS = 2;
constexpr int TC_GROUPS = TC_EPI_WARPS / 4;      // warp groups (4 warps = 128 TMEM lanes each): [team][column half]
constexpr int TC_CSPLIT = TC_GROUPS / TC_TEAMS;  // column groups per tile
constexpr int TC_CW = TC_N / TC_CSPLIT;         // columns scanned per epilogue thread and tile (128: 7 key bits)
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32;
constexpr int TC_TEAM_THREADS = TC_EPI_THREADS / TC_TEAMS;
constexpr int TC_THREADS = 64 + TC_EPI_THREADS;    // warps: 0 TMA, 1 MMA, 2.. epilogue
constexpr int TC_MAX_KP = 128;     // padded K supported (SIFT 128)
constexpr int TC_TABS = 4;         // ring of per-tile column tables (written two tiles ahead by the epilogue threads)
constexpr int TC_MAX_CLUSTER = 8;  // train splits merged inside one thread-block cluster (portable cluster size); more splits merge through global memory

struct TcParams {
    const void *qb, *tb;            // operand copies: bf16 [n][kp] (L2) or fp8-e4m3 0/1 bytes [n][kp] (Hamming on bit-expanded rows)
    const float *qn, *tn;           // squared norms
    int *flags;                     // see mk_common.cuh: flags[0] inputs not exact in bf16, flags[3] result needs the SIMT recompute
    int nq, nt, kp, dim;            // single pair: sizes; batched: upper bounds
    int row_bytes, atom_elems;      // operand row = row_bytes (128 or 256) = kp * element size; elements per 128-byte swizzle atom
    int hamming;                    // 1: scores are Hamming counts (distance = count), 0: squared L2 (distance = sqrtf)
    const int32_t *set_off, *pairs, *out_off;   // batched tables (device) or null: rows of qb == rows of tb == all sets
    int splits, nqb;
    int cluster_merge;              // 1: the CTAs of one query block (blockIdx.y = split) form a thread-block cluster and merge through rank 0's shared memory
    unsigned long long *part;
    int *tickets;
    int32_t *idx;
    float *dist;
    uint8_t *keep;
    int32_t *n_keep;
    double ratio;
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_addr(bar)) : "memory");
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
        "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                     smem_addr(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), descriptor
// version 1 (Blackwell), layout type 2 = SWIZZLE_128B.  Advancing K by 16 bf16 = +32 B on the start address.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address, 16-byte units, bits [0,14)
    d |= (uint64_t)1 << 16;                            // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset = 1024 B
    d |= (uint64_t)1 << 46;                            // descriptor version
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}

// instruction descriptor: D = F32, A = B = BF16, both K-major, N = 256, M = 128
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

// same geometry, A = B = E4M3 (format code 0): 0/1 bytes of bit-expanded binary descriptors, K = 32 per instruction,
// D = F16 (format code 0): q.t <= 256 is exact in fp16 and two accumulators share one 32-bit TMEM column, which halves the
// tensor-memory read (64 B / cycle / SM, the bound of this kernel) per candidate
constexpr uint32_t TC_IDESC_F8 = ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
template <int KIND> __host__ __device__ constexpr int tc_acc_cols() { return TC_N; }   // TMEM columns of one accumulator stage (fp16 accumulators still take a 32-bit column each)

template <int KIND>   // 0: kind::f16 (bf16 operands), 1: kind::f8f6f4 (e4m3 operands); both consume 32 bytes of K per instruction
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate)
{
    if constexpr (KIND == 0)
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(TC_IDESC_F8), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_addr(bar)) : "memory");
}

#define TC_LD32(taddr, r)                                                                                                      \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                     \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                      \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"                    \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),   \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),        \
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),      \
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                    \
                 : "r"(taddr))

// the same load with .pack::16b: register j = (column 2j, column 2j + 1) low halves, i.e. 64 fp16 accumulators per thread
#define TC_LD32_PACK16(taddr, r)                                                                                               \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.pack::16b.x32.b32 "                                                           \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                      \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"                    \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),   \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),        \
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),      \
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                    \
                 : "r"(taddr))

// tcgen05.wait::ld tied to the registers of one load, so the compiler cannot move their uses above the wait
#define TC_WAIT_LD(r)                                                                                                          \
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"                                                                            \
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),   \
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),        \
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),      \
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])                    \
                 :: "memory")

__device__ __forceinline__ int4 lds128(uint32_t saddr)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}

// ---- epilogue scans: 32 accumulator columns of one query row against the per-column table of the tile ------------------------
// KIND 0 (L2): S = q.t is an exact integer < 2^23, so bits(S + 2^23) = 0x4B000000 + S and
//     key = -256 * bits + (|t|^2 * 128 + col)  ==  (|t|^2 - 2 S) * 128 + col      (mod 2^32; 256 * 0x4B000000 == 0)
// is ONE IMAD per element; it orders candidates by (d^2, column) because |q|^2 is a per-row constant, it fits int32
// (|.| <= 2^30 + 127), and min / second-min over keys needs no branch, hence no warp divergence.
template <bool PARTIAL>
__device__ __forceinline__ void scan32_i32(const uint32_t (&r)[32], uint32_t tsaddr, int &b1, int &b2)
{
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const int4 t4 = lds128(tsaddr + j * 4);
        const int tv[4] = { t4.x, t4.y, t4.z, t4.w };
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int bits = __float_as_int(__fadd_rn(__uint_as_float(r[j + u]), 8388608.0f));
            int key = bits * -256 + tv[u];
            if (PARTIAL) key = (tv[u] == 0x7fffffff) ? 0x7fffffff : key;   // tail columns: another set (batched) or zero fill
            b2 = min(b2, max(b1, key));
            b1 = min(b1, key);
        }
    }
}

// KIND 1 (Hamming): S <= 256 and popc(t) - 2 S lies in [-256, 256], so a candidate fits 16 bits:
//     key16 = (popc(t) + 256 - 2 S) << 6 | pair       pair = 0..63, the thread's 128 columns as 64 (even, odd) column pairs
// The fp16 accumulators of columns 2p and 2p+1 arrive in one register.  HADD2 with 1024 turns both into 0x6400 + S, one
// IMAD forms both keys (T2 - 128 * bits2: 128 * 0x6400 vanishes mod 2^16 and the low half's borrow of 0x32 is pre-added
// to the table's high half) and VIMNMX.U16x2 keeps the two best of each half.  0xFFFF marks columns beyond the train set.
template <bool PARTIAL>
__device__ __forceinline__ void scan32_u16(const uint32_t (&r)[32], uint32_t tsaddr, uint32_t &b1, uint32_t &b2)
{
    const __half2 magic = __float2half2_rn(1024.0f);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const int4 t4 = lds128(tsaddr + j * 4);
        const uint32_t tv[4] = { (uint32_t)t4.x, (uint32_t)t4.y, (uint32_t)t4.z, (uint32_t)t4.w };
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const __half2 s2 = __hadd2(*reinterpret_cast<const __half2 *>(&r[j + u]), magic);
            uint32_t key = *reinterpret_cast<const uint32_t *>(&s2) * 0xFFFFFF80u + tv[u];
            if (PARTIAL)
                key |= ((tv[u] & 0xFFFFu) == 0xFFFFu ? 0xFFFFu : 0u) | ((tv[u] >> 16) == 0x0031u ? 0xFFFF0000u : 0u);
            b2 = __vminu2(b2, __vmaxu2(b1, key));
            b1 = __vminu2(b1, key);
        }
    }
}

// One tile for one thread: 128 candidates = four (fp32) or two (packed fp16) tcgen05.ld of 32 columns, the next load in
// flight while the previous one is scanned.
template <int KIND, bool PARTIAL>
__device__ __forceinline__ void scan_tile(uint32_t taddr, uint32_t tsaddr, uint32_t &b1, uint32_t &b2)
{
    constexpr uint32_t CH = 128;                      // table bytes per 32-column load
    uint32_t ra[32], rb[32];
    auto scan = [&](const uint32_t (&r)[32], uint32_t ts) {
        if constexpr (KIND == 1) scan32_u16<PARTIAL>(r, ts, b1, b2);
        else scan32_i32<PARTIAL>(r, ts, reinterpret_cast<int &>(b1), reinterpret_cast<int &>(b2));
    };
    if constexpr (KIND == 1) {
        TC_LD32_PACK16(taddr, ra);
        TC_WAIT_LD(ra);
        TC_LD32_PACK16(taddr + 64u, rb);
        scan(ra, tsaddr);
        TC_WAIT_LD(rb);
        scan(rb, tsaddr + CH);
    } else {
        TC_LD32(taddr, ra);
        TC_WAIT_LD(ra);
        TC_LD32(taddr + 32u, rb);
        scan(ra, tsaddr);
        TC_WAIT_LD(rb);
        TC_LD32(taddr + 64u, ra);
        scan(rb, tsaddr + CH);
        TC_WAIT_LD(ra);
        TC_LD32(taddr + 96u, rb);
        scan(ra, tsaddr + 2 * CH);
        TC_WAIT_LD(rb);
        scan(rb, tsaddr + 3 * CH);
    }
}
static_assert(TC_CW == 128, "scan_tile walks 128 candidates per thread");

struct TcKey {
    using Key = unsigned long long;
    static constexpr Key NONE = ~0ull;
    static __device__ __forceinline__ void decode(Key k, int &idx, float &d)
    {
        if (k == NONE) { idx = -1; d = FLT_MAX; }
        else { idx = (int)(uint32_t)k; d = __fsqrt_rn((float)(uint32_t)(k >> 32)); }   // cv2: sqrtf((float)sum)
    }
};

// layout of the dynamic shared memory (1024-byte aligned base)
struct TcSmem {
    static constexpr int A_BYTES = TC_M * 128;                  // one k-half of A (128 rows x 128 B)
    static constexpr int B_BYTES = TC_N * 128;                  // one k-half of one B stage
    static __host__ __device__ constexpr int a_off(int half) { return half * A_BYTES; }
    static __host__ __device__ constexpr int b_slot(int slot) { return 2 * A_BYTES + slot * B_BYTES; }          // 2 * TC_STAGES slots of one atom
    static __host__ __device__ constexpr int tn_off(int slot) { return 2 * A_BYTES + TC_STAGES * 2 * B_BYTES + slot * TC_N * 4; }   // TC_TABS slots
    static constexpr int keys_off = 2 * A_BYTES + TC_STAGES * 2 * B_BYTES + TC_TABS * TC_N * 4;      // [TC_GROUPS][128][2] u64
    static constexpr int bar_off = keys_off + TC_GROUPS * TC_M * 2 * 8;
    static constexpr int part_off = bar_off + 256;              // [TC_MAX_CLUSTER - 1][128][2] u64: best two of the other splits of the cluster (written remotely)
    static constexpr int total = part_off + (TC_MAX_CLUSTER - 1) * TC_M * 2 * 8;
};

template <int KIND>
__global__ void __launch_bounds__(TC_THREADS, 1)
knn2_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t, const __grid_constant__ TcParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + TcSmem::bar_off);
    unsigned long long *bar_a = bars + 0, *bar_bfull = bars + 1, *bar_bempty = bars + 1 + TC_MAX_STAGES,
                       *bar_tfull = bars + 1 + 2 * TC_MAX_STAGES, *bar_tempty = bars + 3 + 2 * TC_MAX_STAGES,
                       *bar_tab = bars + 5 + 2 * TC_MAX_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 5 + TC_TABS + 2 * TC_MAX_STAGES);
    static_assert((6 + TC_TABS + 2 * TC_MAX_STAGES) * 8 <= 256, "barrier block");
    __shared__ int s_flag;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int qblk = blockIdx.x, split = blockIdx.y, z = blockIdx.z;
    // float rows are followed by the guarded SIMT pass (mk_knn.cu), which only reads this kernel's verdict: let its CTAs be
    // scheduled as soon as there is room (they wait in griddepcontrol.wait until this grid has completed)
    if (KIND == 0) asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    // problem of this CTA: rows [qrow0, qrow0 + nq) of the query matrix against rows [trow0, trow0 + nt) of the train matrix
    int qrow0 = 0, trow0 = 0, nq = p.nq, nt = p.nt, out0 = 0;
    if (p.pairs) {
        const int qs = p.pairs[2 * z], ts = p.pairs[2 * z + 1];
        qrow0 = p.set_off[qs]; nq = p.set_off[qs + 1] - qrow0;
        trow0 = p.set_off[ts]; nt = p.set_off[ts + 1] - trow0;
        out0 = p.out_off[z];
    }
    const int q0 = qblk * TC_M;
    if (q0 >= nq) return;
    const int nhalf = p.row_bytes / 128;                 // 128-byte swizzle atoms per operand row: 1 or 2
    const int nstages = TC_MAX_STAGES / nhalf;           // B ring depth: the same shared memory holds 2 tiles of 64 KB or 4 of 32 KB
    // train tiles of this split
    const int ntiles = (nt + TC_N - 1) / TC_N;
    const int per = (ntiles + p.splits - 1) / p.splits;
    const int tile_lo = min(ntiles, split * per), tile_hi = min(ntiles, tile_lo + per);
    const int my_tiles = tile_hi - tile_lo;

    if (tid == 0) {
        mbar_init(bar_a, 1);
        for (int s = 0; s < TC_MAX_STAGES; ++s) { mbar_init(bar_bfull + s, 1); mbar_init(bar_bempty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_tfull + s, 1); mbar_init(bar_tempty + s, TC_TEAM_THREADS); }
        for (int s = 0; s < TC_TABS; ++s) mbar_init(bar_tab + s, KIND == 1 ? TC_N / 2 : TC_N);   // one arrival per table entry
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {   // allocate all 512 TMEM columns (2 accumulator stages x 256)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_addr(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    // Launched with programmatic stream serialization: everything above overlapped the tail of the prep kernel.  Its outputs
    // (operand copies, norms, zeroed flags and tickets) may only be touched below this line.
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    // the first loads need nothing but the mbarriers thread 0 has just initialised, so they start before the block-wide barrier
    const int first_loads = min(my_tiles, nstages);
    if (tid == 0) {
        mbar_expect_tx(bar_a, (uint32_t)(nhalf * TcSmem::A_BYTES));
        for (int h = 0; h < nhalf; ++h) tma_load_2d(smem + TcSmem::a_off(h), &map_q, h * p.atom_elems, qrow0 + q0, bar_a);
        for (int i = 0; i < first_loads; ++i) {
            mbar_expect_tx(bar_bfull + i, (uint32_t)(nhalf * TcSmem::B_BYTES));
            for (int h = 0; h < nhalf; ++h)
                tma_load_2d(smem + TcSmem::b_slot(i * nhalf + h), &map_t, h * p.atom_elems, trow0 + (tile_lo + i) * TC_N, bar_bfull + i);
        }
    }
    const bool inexact = KIND == 0 && !knn_tc_inputs_ok(p.flags);   // bf16 would round: the guarded SIMT kernel handles this call
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (inexact) {                           // (the loads above are harmless; they complete into shared memory nobody reads)
        if (tid == 0) {
            mbar_wait(bar_a, 0);
            for (int i = 0; i < first_loads; ++i) mbar_wait(bar_bfull + i, 0);
        }
        __syncthreads();
        if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base) : "memory");
        return;
    }

    if (warp == 0) {
        // ===================== TMA producer (the first ring fill was issued above) =====================
        if (lane == 0) {
            for (int i = first_loads; i < my_tiles; ++i) {
                const int s = i % nstages, use = i / nstages;
                mbar_wait(bar_bempty + s, (use - 1) & 1);
                mbar_expect_tx(bar_bfull + s, (uint32_t)(nhalf * TcSmem::B_BYTES));
                for (int h = 0; h < nhalf; ++h)
                    tma_load_2d(smem + TcSmem::b_slot(s * nhalf + h), &map_t, h * p.atom_elems, trow0 + (tile_lo + i) * TC_N, bar_bfull + s);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            mbar_wait(bar_a, 0);
            for (int i = 0; i < my_tiles; ++i) {
                const int s = i % nstages, use = i / nstages;
                const int acc = i & 1, acc_use = i >> 1;
                if (acc_use > 0) mbar_wait(bar_tempty + acc, (acc_use - 1) & 1);   // epilogue drained this accumulator
                mbar_wait(bar_bfull + s, use & 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * tc_acc_cols<KIND>());
                uint32_t accumulate = 0;
                for (int h = 0; h < nhalf; ++h) {
                    const uint64_t ad = make_sw128_desc(smem_addr(smem + TcSmem::a_off(h)));
                    const uint64_t bd = make_sw128_desc(smem_addr(smem + TcSmem::b_slot(s * nhalf + h)));
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma<KIND>(d_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), accumulate);   // +32 B per K step
                        accumulate = 1;
                    }
                }
                umma_commit(bar_bempty + s);      // smem stage free once these MMAs retire
                umma_commit(bar_tfull + acc);     // accumulator ready for the epilogue
            }
        }
    } else {
        // ===================== epilogue: one query row x one 128-column half per thread (see scan32_*) =====================
        // Two teams of 8 warps: team t owns accumulator stage t, i.e. the tiles i = t (mod 2).  While one team waits for its
        // accumulator or its first tcgen05.ld, the other one is in the middle of a scan, so every scheduler keeps both kinds of
        // warps and the MMA of tile i + 1 runs under the scans of tiles i - 1 and i.
        const int quarter = warp & 3;                      // TMEM lane quarter this warp may access (warp id % 4)
        const int group = (warp - 2) >> 2;                 // 0..3 = [team][column half]
        const int team = group / TC_CSPLIT, chalf = group % TC_CSPLIT;
        const int row = quarter * 32 + lane;
        const int qi = q0 + row;
        const int qn = (qi < nq) ? (int)p.qn[qrow0 + qi] : 0;    // exact integer
        unsigned long long k1 = TcKey::NONE, k2 = TcKey::NONE;   // running (d^2 << 32 | train index)
        auto fold = [&](uint32_t d, uint32_t col) {
            const unsigned long long key = ((unsigned long long)d << 32) | col;
            const unsigned long long mx = key > k1 ? key : k1; k1 = key < k1 ? key : k1; k2 = mx < k2 ? mx : k2;
        };
        // Per-tile column tables, two slots per team.  The first TC_N (Hamming: TC_N / 2) threads of a team each own one entry;
        // the entry of the team's NEXT tile (i + 2) is LOADED before tile i is scanned and STORED after it, so the global-load
        // latency hides behind the scan.  The slot of tile i + 2 was last read for tile i - 2, which every team mate left
        // before tile i could start (tile i's MMA waited for that release), and the readers of tile i + 2 wait for tile i's
        // release anyway: the table barrier adds no coupling of its own.
        const int et = (tid - 64) % TC_TEAM_THREADS;       // index within the team
        const bool tab_writer = et < (KIND == 1 ? TC_N / 2 : TC_N);
        auto slot_of = [&](int i) { return (i & 1) * 2 + ((i >> 1) & 1); };
        // raw loads first (table_load), arithmetic on them only when the entry is stored (table_store): nothing between the
        // two touches the loaded registers, so the scan never waits on the global load
        const int pid = et % (TC_CW / 2);                  // Hamming: pair owned by this thread ...
        const int pcol = (et / (TC_CW / 2)) * TC_CW + 2 * pid;    // ... = columns pcol (low half), pcol + 1 (high half)
        auto table_load = [&](int i, float &ta, float &tb) {
            const int col0 = (tile_lo + i) * TC_N;
            if constexpr (KIND == 1) {
                ta = (col0 + pcol < nt) ? __ldg(p.tn + trow0 + col0 + pcol) : -1.0f;
                tb = (col0 + pcol + 1 < nt) ? __ldg(p.tn + trow0 + col0 + pcol + 1) : -1.0f;
            } else {
                ta = (col0 + et < nt) ? __ldg(p.tn + trow0 + col0 + et) : -1.0f;
                tb = 0.0f;
            }
        };
        auto table_store = [&](int i, float ta, float tb) {
            int v;
            if constexpr (KIND == 1) {       // (popc(t) + 256) << 6 | pair per half, 0xFFFF beyond nt; +0x32 on the high half (scan32_u16)
                const uint32_t lo = ta >= 0.0f ? ((((uint32_t)ta + 256u) << 6) | (uint32_t)pid) : 0xFFFFu;
                const uint32_t hi = tb >= 0.0f ? ((((uint32_t)tb + 256u) << 6) | (uint32_t)pid) : 0xFFFFu;
                v = (int)(lo + ((hi + 0x32u) << 16));
            } else {                         // |t|^2 * 128 + (col & 127); INT_MAX beyond nt
                v = ta >= 0.0f ? ((int)ta * 128 + (et % TC_CW)) : 0x7fffffff;
            }
            reinterpret_cast<int *>(smem + TcSmem::tn_off(slot_of(i)))[et] = v;
            mbar_arrive(bar_tab + slot_of(i));
        };
        if (tab_writer && team < my_tiles) {
            float ta, tb;
            table_load(team, ta, tb);
            table_store(team, ta, tb);
        }
        // Hamming: running best two of each 16-bit half as 32-bit keys (count + 256) << 22 | tile << 6 | pair, ordered like
        // (count, column) within the half; unpacked once after the last tile
        uint32_t run[4] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu };   // [half][rank]
        for (int i = team; i < my_tiles; i += TC_TEAMS) {
            const int acc = team, acc_use = i >> 1;
            const int col0 = (tile_lo + i) * TC_N;
            const bool partial = col0 + TC_N > nt;        // warp-uniform, last tile of a set only
            const uint32_t tsaddr = smem_addr(smem + TcSmem::tn_off(slot_of(i))) + (uint32_t)chalf * (uint32_t)(KIND == 1 ? TC_CW * 2 : TC_CW * 4);   // one table word per TMEM column
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * TC_CSPLIT + chalf) * (tc_acc_cols<KIND>() / TC_CSPLIT));
            const bool prefetch = tab_writer && i + TC_TEAMS < my_tiles;
            float ta = 0.0f, tb = 0.0f;
            if (prefetch) table_load(i + TC_TEAMS, ta, tb);
            mbar_wait(bar_tab + slot_of(i), (i >> 2) & 1);
            mbar_wait(bar_tfull + acc, acc_use & 1);
            tc_fence_after();
            uint32_t b1 = KIND == 1 ? 0xFFFFFFFFu : 0x7fffffffu, b2 = b1;
            if (partial) scan_tile<KIND, true>(taddr, tsaddr, b1, b2);
            else scan_tile<KIND, false>(taddr, tsaddr, b1, b2);
            tc_fence_before();
            mbar_arrive(bar_tempty + acc);                 // the team releases its accumulator stage
            if (prefetch) table_store(i + TC_TEAMS, ta, tb);
            // fold the tile's best candidates into the running keys
            if constexpr (KIND == 1) {
                const uint32_t tl = (uint32_t)i << 6;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const uint32_t b = r ? b2 : b1;
                        const uint32_t key = (h ? (b & 0xFFC00000u) : ((b << 16) & 0xFFC00000u)) | tl | ((h ? (b >> 16) : b) & 63u);
                        const uint32_t mx = max(run[2 * h], key);
                        run[2 * h] = min(run[2 * h], key);
                        run[2 * h + 1] = min(run[2 * h + 1], mx);
                    }
                }
            } else {                                       // key >> 7 = d^2 - |q|^2, key & 127 = column in this half
                const uint32_t cbase = (uint32_t)(col0 + chalf * TC_CW);
                if (b1 != 0x7fffffffu) fold((uint32_t)(((int)b1 >> 7) + qn), cbase + (b1 & 127u));
                if (b2 != 0x7fffffffu) fold((uint32_t)(((int)b2 >> 7) + qn), cbase + (b2 & 127u));
            }
        }
        if constexpr (KIND == 1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t r = run[e], dpart = r >> 22, pr = r & 63u, tile = (r >> 6) & 0xFFFFu;
                if (dpart != 1023u)
                    fold(dpart - 256u + (uint32_t)qn,
                         (uint32_t)((tile_lo + (int)tile) * TC_N + chalf * TC_CW) + 2u * pr + (uint32_t)(e >> 1));
            }
        }
        unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(smem + TcSmem::keys_off);
        s_keys[(group * TC_M + row) * 2] = k1; s_keys[(group * TC_M + row) * 2 + 1] = k2;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base) : "memory");
    }

    // ---- merge the train splits and write the outputs (same ticket scheme as mk_knn.cu) ----------
    unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(smem + TcSmem::keys_off);
    unsigned long long k1 = TcKey::NONE, k2 = TcKey::NONE;
    const bool is_row_thread = tid < TC_M;
    if (is_row_thread) {                                   // merge the column groups of this row
        k1 = s_keys[tid * 2]; k2 = s_keys[tid * 2 + 1];
#pragma unroll
        for (int g = 1; g < TC_GROUPS; ++g) {
            const unsigned long long a = s_keys[(g * TC_M + tid) * 2], b = s_keys[(g * TC_M + tid) * 2 + 1];
            if (a != TcKey::NONE) { const unsigned long long mx = a > k1 ? a : k1; k1 = a < k1 ? a : k1; k2 = mx < k2 ? mx : k2; }
            if (b != TcKey::NONE) { const unsigned long long mx = b > k1 ? b : k1; k1 = b < k1 ? b : k1; k2 = mx < k2 ? mx : k2; }
        }
    }
    if (p.splits > 1 && p.cluster_merge) {
        // The splits of this query block are the CTAs of one cluster (cluster dims (1, splits, 1): rank == split).  Everybody but
        // rank 0 stores its best two per row straight into rank 0's shared memory (st.shared::cluster over the SM-to-SM network),
        // one cluster barrier publishes them, rank 0 merges and writes the outputs: no global-memory round trip, no fence, no
        // ticket (the global path below costs store -> __threadfence -> atomic -> load, about 2 us on the critical path).
        if (split != 0 && is_row_thread) {
            const uint32_t local = smem_addr(smem + TcSmem::part_off + ((split - 1) * TC_M + tid) * 16);
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(remote) : "r"(local), "r"(0));
            asm volatile("st.shared::cluster.v2.u64 [%0], {%1, %2};\n" ::"r"(remote), "l"(k1), "l"(k2) : "memory");
        }
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        if (split != 0) return;
        if (is_row_thread) {
            const unsigned long long *o = reinterpret_cast<const unsigned long long *>(smem + TcSmem::part_off) + tid * 2;
            for (int s = 1; s < p.splits; ++s, o += TC_M * 2) {
                const unsigned long long a = o[0], b = o[1];
                if (a != TcKey::NONE) { const unsigned long long mx = a > k1 ? a : k1; k1 = a < k1 ? a : k1; k2 = mx < k2 ? mx : k2; }
                if (b != TcKey::NONE) { const unsigned long long mx = b > k1 ? b : k1; k1 = b < k1 ? b : k1; k2 = mx < k2 ? mx : k2; }
            }
        }
    } else if (p.splits > 1) {
        const size_t blk = (size_t)z * p.nqb + qblk;
        unsigned long long *base = p.part + (blk * p.splits) * TC_M * 2;
        if (is_row_thread) {
            unsigned long long *mine = base + ((size_t)split * TC_M + tid) * 2;
            mine[0] = k1; mine[1] = k2;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) s_flag = (atomicAdd(&p.tickets[blk], 1) == p.splits - 1);
        __syncthreads();
        if (!s_flag) return;
        __threadfence();
        if (is_row_thread) {
            k1 = TcKey::NONE; k2 = TcKey::NONE;
            for (int s = 0; s < p.splits; ++s) {
                const unsigned long long *o = base + ((size_t)s * TC_M + tid) * 2;
                const unsigned long long a = __ldcg(o), b = __ldcg(o + 1);
                if (a != TcKey::NONE) { const unsigned long long mx = a > k1 ? a : k1; k1 = a < k1 ? a : k1; k2 = mx < k2 ? mx : k2; }
                if (b != TcKey::NONE) { const unsigned long long mx = b > k1 ? b : k1; k1 = b < k1 ? b : k1; k2 = mx < k2 ? mx : k2; }
            }
        }
        if (tid == 0) p.tickets[blk] = 0;
    }
    bool kept = false;
    const int qi = q0 + tid;
    if (is_row_thread && qi < nq) {
        const size_t o = (size_t)out0 + qi;
        int i1, i2; float d1, d2;
        TcKey::decode(k1, i1, d1);
        TcKey::decode(k2, i2, d2);
        if (p.hamming) {                                   // the score is the Hamming count itself
            if (k1 != TcKey::NONE) d1 = (float)(uint32_t)(k1 >> 32);
            if (k2 != TcKey::NONE) d2 = (float)(uint32_t)(k2 >> 32);
        }
        if (!p.hamming && ((k1 != TcKey::NONE && (uint32_t)(k1 >> 32) >= 4194304u) || (k2 != TcKey::NONE && (uint32_t)(k2 >> 32) >= 4194304u)))
            atomicOr(p.flags + 3, 1);                      // sqrtf no longer injective: let the SIMT kernel redo the call
        p.idx[2 * o] = i1; p.idx[2 * o + 1] = i2;
        p.dist[2 * o] = d1; p.dist[2 * o + 1] = d2;
        kept = (i2 >= 0) && (p.ratio > 0.0) && ((double)d1 < p.ratio * (double)d2);   // mosaic.py:430
        if (p.keep) p.keep[o] = kept ? 1 : 0;
    }
    if (p.n_keep && tid < TC_M) {
        const unsigned b = __ballot_sync(0xffffffffu, kept);
        if (lane == 0 && b) atomicAdd(p.n_keep + z, __popc(b));
    }
}

// ---- prep: rows -> bf16 [n][kp] (zero padded), squared norms, exactness flag; one launch for both sets ----------
template <typename T>
__global__ void __launch_bounds__(256) knn_tc_prep_kernel(const T *__restrict__ q, int nq, const T *__restrict__ t, int nt,
                                                          size_t stride_bytes, int dim, int kp, __nv_bfloat16 *__restrict__ qb,
                                                          __nv_bfloat16 *__restrict__ tb, float *__restrict__ qn,
                                                          float *__restrict__ tn, int *flags, int32_t *n_keep, int epoch, int nzero,
                                                          int *zero2, int nzero2)
{
    asm volatile("griddepcontrol.launch_dependents;\n");     // the GEMM kernel may start its prologue now (it waits before reading)
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    int row = gid >> 5;
    const int lane = threadIdx.x & 31;
    if (row == 0 && lane == 0 && n_keep) *n_keep = 0;
    // No memset node in front of this kernel: flags[0] is only ever SET to this call's epoch (below, by whoever finds a bad value),
    // flags[1] = the epoch, so "inputs not exact" reads flags[0] == flags[1] (mk_common.cuh) and nothing has to be zeroed before
    // it can be set; flags[2..] (flags[3] is raised by the TC kernel), the TC tickets / counters behind them and the SIMT
    // fallback's tickets are zeroed here, grid-wide.
    if (gid == 0) flags[1] = epoch;
    for (int i = 2 + gid; i < nzero; i += gridDim.x * blockDim.x) flags[i] = 0;
    for (int i = gid; i < nzero2; i += gridDim.x * blockDim.x) zero2[i] = 0;
    if (row >= nq + nt) return;
    const bool is_q = row < nq;
    if (!is_q) row -= nq;
    const T *src = reinterpret_cast<const T *>(reinterpret_cast<const uint8_t *>(is_q ? q : t) + (size_t)row * stride_bytes);
    __nv_bfloat16 *dst = (is_q ? qb : tb) + (size_t)row * kp;
    float acc = 0.f;
    bool bad = false;
    for (int k = lane; k < kp; k += 32) {
        float v = 0.f;
        if (k < dim) v = (float)src[k];
        if (!(v >= 0.f && v <= 255.f && v == truncf(v))) bad = true;     // also catches NaN
        dst[k] = __float2bfloat16_rn(v);
        acc = __fadd_rn(acc, __fmul_rn(v, v));                            // integers: exact in any order
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) (is_q ? qn : tn)[row] = acc;
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicExch(flags, epoch);
}

// ---- prep for Hamming: bits -> e4m3 bytes (1.0 = 0x38 / 0), popcounts as "norms" -----------------------------------------
// One thread per 16 output bytes (two descriptor bytes), kp / 16 threads per row.
__device__ __forceinline__ uint32_t expand4(uint32_t nib)   // bits 0..3 -> four bytes of 0x38 / 0x00
{
    return ((nib * 0x00204081u) & 0x01010101u) * 0x38u;
}
__global__ void __launch_bounds__(256) knn_tc_prep_bits_kernel(const uint8_t *__restrict__ q, int nq, const uint8_t *__restrict__ t, int nt,
                                                               size_t stride_bytes, int nbytes, int kp, uint8_t *__restrict__ qb,
                                                               uint8_t *__restrict__ tb, float *__restrict__ qn, float *__restrict__ tn,
                                                               int32_t *n_keep, int *zero, int nzero)
{
    asm volatile("griddepcontrol.launch_dependents;\n");     // the GEMM kernel may start its prologue now (it waits before reading)
    const int tpr = kp >> 4;                                 // threads per row: 8 or 16
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    int row = gid / tpr;
    const int part = gid % tpr;
    if (gid == 0 && n_keep) *n_keep = 0;
    if (gid == 0) { zero[0] = 0; zero[1] = 1; }              // flags[0] != flags[1]: bit rows are always exact (mk_common.cuh)
    for (int i = 2 + gid; i < nzero; i += gridDim.x * blockDim.x) zero[i] = 0;   // flags + tickets of the TC kernel (no memset node)
    const bool live = row < nq + nt;
    const bool is_q = row < nq;
    if (!is_q) row -= nq;
    uint32_t v = 0;
    if (live && 2 * part < nbytes) {
        const uint8_t *src = (is_q ? q : t) + (size_t)row * stride_bytes + 2 * part;
        v = src[0];
        if (2 * part + 1 < nbytes) v |= (uint32_t)src[1] << 8;
    }
    if (live) {
        uint4 o;
        o.x = expand4(v & 15u); o.y = expand4((v >> 4) & 15u); o.z = expand4((v >> 8) & 15u); o.w = expand4(v >> 12);
        *reinterpret_cast<uint4 *>((is_q ? qb : tb) + (size_t)row * kp + 16 * part) = o;
    }
    int cnt = __popc(v);
    for (int o = tpr >> 1; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (live && part == 0) (is_q ? qn : tn)[row] = (float)cnt;
}

// ---- host -------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// esz = 2: bf16 rows, 64 elements per 128-byte atom; esz = 1: e4m3 bytes, 128 elements per atom
static int make_map(CUtensorMap *map, const void *base, int rows, int kp, int box_rows, int esz)
{
    // a frame loop asks for the same few maps over and over (the operand copies live in a persistent workspace): keep the
    // last encodings per thread, cuTensorMapEncodeTiled costs more than a microsecond each
    struct Entry { const void *base; int rows, kp, box_rows, esz; CUtensorMap map; };
    static thread_local Entry cache[8];
    static thread_local int next = 0;
    for (const Entry &e : cache)
        if (e.base == base && e.rows == rows && e.kp == kp && e.box_rows == box_rows && e.esz == esz && base) { *map = e.map; return MK_OK; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return MK_ERR_CUDA; }
    const cuuint64_t dims[2] = { (cuuint64_t)kp, (cuuint64_t)rows };
    const cuuint64_t strides[1] = { (cuuint64_t)kp * esz };
    const cuuint32_t box[2] = { (cuuint32_t)(128 / esz), (cuuint32_t)box_rows };
    const cuuint32_t estr[2] = { 1, 1 };
    CUresult r = fn(map, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return MK_ERR_CUDA; }
    Entry &slot = cache[next];
    next = (next + 1) % 8;
    slot.base = base; slot.rows = rows; slot.kp = kp; slot.box_rows = box_rows; slot.esz = esz; slot.map = *map;
    return MK_OK;
}

struct TcPlan {
    int kp, esz, nqb, splits;
    size_t off_qb, off_tb, off_qn, off_tn, off_flags, off_tickets, off_cnt, off_part, total;
};

// rows_q / rows_t: rows of the bf16 copies (batched: both = all rows of all sets and the copies are shared)
static TcPlan tc_plan(int norm, int npairs, int max_nq, int max_nt, int dim, long rows_q, long rows_t, bool shared, int sms = 148)
{
    TcPlan pl;
    if (norm == MK_NORM_HAMMING_TC) { pl.esz = 1; pl.kp = dim * 8 <= 128 ? 128 : 256; }   // one e4m3 byte per descriptor bit
    else { pl.esz = 2; pl.kp = dim <= 64 ? 64 : 128; }
    pl.nqb = (max_nq + TC_M - 1) / TC_M;
    const int ntiles = (max_nt + TC_N - 1) / TC_N;
    // one CTA per SM (165 KB of shared memory).  Cost model: waves x (tiles per CTA + 4); the constant stands for the A tile,
    // the prologue and the merge, so long scans are preferred.
    int best = 1;
    double best_cost = 1e300;
    for (int s = 1; s <= ntiles && s <= 64; ++s) {
        const int per = (ntiles + s - 1) / s;
        const int s_eff = (ntiles + per - 1) / per;
        const long ctas = (long)npairs * pl.nqb * s_eff;
        const long waves = (ctas + sms - 1) / sms;
        const double cost = (double)waves * (per + 4.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = s_eff; }
    }
    pl.splits = pl.nqb > 0 ? best : 1;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
    pl.off_qb = take((size_t)rows_q * pl.kp * pl.esz);
    pl.off_tb = shared ? pl.off_qb : take((size_t)rows_t * pl.kp * pl.esz);
    pl.off_qn = take((size_t)rows_q * 4);
    pl.off_tn = shared ? pl.off_qn : take((size_t)rows_t * 4);
    pl.off_flags = take(256);
    const size_t blocks = (size_t)(npairs > 0 ? npairs : 1) * (pl.nqb > 0 ? pl.nqb : 1);
    pl.off_tickets = take(blocks * 4);
    pl.off_cnt = take((size_t)(npairs > 0 ? npairs : 1) * 4);   // float rows: survivor counts of the TC kernel (zeroed with the tickets)
    pl.off_part = take(pl.splits > 1 ? blocks * pl.splits * TC_M * 2 * 8 : 0);
    pl.total = o;
    return pl;
}

size_t knn_tc_workspace_bytes(int norm, int nq, int nt, int dim) { return tc_plan(norm, 1, nq, nt, dim, nq, nt, false).total; }
size_t knn_tc_batched_workspace_bytes(int norm, int npairs, int max_nq, int max_nt, int dim, int total_rows)
{
    return tc_plan(norm, npairs, max_nq, max_nt, dim, total_rows, total_rows, true).total;
}

bool knn_tc_supported(int norm, int nq, int nt, int dim)
{
    if (norm == MK_NORM_HAMMING_TC) return dim <= 32 && nq >= 1 && nt >= 2 && nt < (1 << 22);   // 256 bits = two 128-byte atoms of e4m3 bytes
    if (norm != MK_NORM_L2_F32 && norm != MK_NORM_L2_U8) return false;
    if (dim > TC_MAX_KP || nq < 1 || nt < 2) return false;
    return true;
}

// Launches prep + the TC kernel.  Single pair: q / t are the two row sets.  Batched: q = all rows of all sets (t unused),
// set_off / pairs / out_off are the device tables of mk_knn2_batched.  *flags_out lets the caller guard the SIMT fallback.
int knn_tc_run(int norm, const void *q, int nq, const void *t, int nt, int dim, size_t stride, const int32_t *set_off,
               const int32_t *pairs, int npairs, const int32_t *out_off, int total_rows, int32_t *idx, float *dist, double ratio,
               uint8_t *keep, int32_t *n_keep, void *workspace, size_t workspace_bytes, cudaStream_t st, const int **flags_out,
               const int32_t **tc_count_out, int *simt_tickets, size_t simt_ticket_bytes)
{
    const bool batched = pairs != nullptr;
    const int sms = stream_sm_count(st);     // splits (hence the partial-result scratch) never exceed what 148 SMs ask for
    TcPlan pl = batched ? tc_plan(norm, npairs, nq, nt, dim, total_rows, total_rows, true, sms) : tc_plan(norm, 1, nq, nt, dim, nq, nt, false, sms);
    if (pl.total > workspace_bytes)          // the workspace was sized for the whole device: that plan always fits
        pl = batched ? tc_plan(norm, npairs, nq, nt, dim, total_rows, total_rows, true) : tc_plan(norm, 1, nq, nt, dim, nq, nt, false);
    if (workspace_bytes < pl.total) { set_error("mk_knn2: workspace too small for the tensor-core path (%zu < %zu)", workspace_bytes, pl.total); return MK_ERR_WORKSPACE; }
    uint8_t *ws = static_cast<uint8_t *>(workspace);
    if (!aligned(ws, 256)) { set_error("mk_knn2: workspace must be 256-byte aligned for the tensor-core path"); return MK_ERR_ALIGN; }
    void *qb = ws + pl.off_qb, *tb = ws + pl.off_tb;
    float *qn = reinterpret_cast<float *>(ws + pl.off_qn), *tn = reinterpret_cast<float *>(ws + pl.off_tn);
    int *flags = reinterpret_cast<int *>(ws + pl.off_flags), *tickets = reinterpret_cast<int *>(ws + pl.off_tickets);
    const int np = batched ? npairs : 1;
    // flags and tickets are adjacent in the workspace: one memset (the Hamming prep kernel, which sets no flag, zeroes them itself)
    // float rows may be recomputed by the guarded SIMT kernel: the TC kernel then counts survivors into `cnt`, not n_keep
    int32_t *cnt = (norm == MK_NORM_L2_F32 && n_keep) ? reinterpret_cast<int32_t *>(ws + pl.off_cnt) : nullptr;
    const size_t zero_bytes = cnt ? (pl.off_cnt - pl.off_flags) + (size_t)np * 4 : (size_t)((uint8_t *)tickets - (uint8_t *)flags) + (size_t)np * pl.nqb * 4;
    static std::atomic<int> epoch_counter{0};
    int epoch = epoch_counter.fetch_add(1) + 1;
    if (epoch == 0) epoch = epoch_counter.fetch_add(1) + 1;
    const int nzero2 = (int)(simt_ticket_bytes / 4);
    if (n_keep && batched) MK_CUDA_TRY(cudaMemsetAsync(n_keep, 0, sizeof(int32_t) * np, st));
    const int rows_q = batched ? total_rows : nq, rows_t = batched ? 0 : nt;
    const int wpb = 8, rows = rows_q + rows_t;
    static const bool exp_skip_prep = getenv("MK_EXP_SKIP_PREP") != nullptr;   // timing experiment only: operands of the previous call
    if (exp_skip_prep) { /* nothing */ }
    else if (norm == MK_NORM_HAMMING_TC)
        knn_tc_prep_bits_kernel<<<(int)(((long)rows * (pl.kp >> 4) + 255) / 256), 256, 0, st>>>(static_cast<const uint8_t *>(q), rows_q, static_cast<const uint8_t *>(t), rows_t, stride,
                                                                        dim, pl.kp, static_cast<uint8_t *>(qb), static_cast<uint8_t *>(tb), qn, tn,
                                                                        batched ? nullptr : n_keep, flags, (int)(zero_bytes / 4));
    else if (norm == MK_NORM_L2_F32)
        knn_tc_prep_kernel<float><<<(rows + wpb - 1) / wpb, 256, 0, st>>>(static_cast<const float *>(q), rows_q, static_cast<const float *>(t), rows_t,
                                                                          stride, dim, pl.kp, static_cast<__nv_bfloat16 *>(qb), static_cast<__nv_bfloat16 *>(tb), qn, tn, flags, batched ? nullptr : n_keep,
                                                                          epoch, (int)(zero_bytes / 4), simt_tickets, nzero2);
    else
        knn_tc_prep_kernel<uint8_t><<<(rows + wpb - 1) / wpb, 256, 0, st>>>(static_cast<const uint8_t *>(q), rows_q, static_cast<const uint8_t *>(t), rows_t,
                                                                            stride, dim, pl.kp, static_cast<__nv_bfloat16 *>(qb), static_cast<__nv_bfloat16 *>(tb), qn, tn, flags, batched ? nullptr : n_keep,
                                                                            epoch, (int)(zero_bytes / 4), simt_tickets, nzero2);
    MK_LAUNCH_CHECK("knn_tc_prep_kernel");
    alignas(64) CUtensorMap map_q, map_t;
    int rc = make_map(&map_q, qb, rows_q, pl.kp, TC_M, pl.esz);
    if (rc != MK_OK) return rc;
    rc = make_map(&map_t, tb, batched ? total_rows : nt, pl.kp, TC_N, pl.esz);
    if (rc != MK_OK) return rc;
    TcParams P;
    P.qb = qb; P.tb = tb; P.qn = qn; P.tn = tn; P.flags = flags;
    P.nq = nq; P.nt = nt; P.kp = pl.kp; P.dim = dim;
    P.row_bytes = pl.kp * pl.esz; P.atom_elems = 128 / pl.esz; P.hamming = norm == MK_NORM_HAMMING_TC;
    P.splits = pl.splits; P.nqb = pl.nqb;
    static const bool no_cluster = getenv("MK_NO_CLUSTER_MERGE") != nullptr;   // diagnosis / A-B switch: merge the splits through global memory
    P.cluster_merge = (pl.splits > 1 && pl.splits <= TC_MAX_CLUSTER && !no_cluster) ? 1 : 0;
    P.set_off = set_off; P.pairs = pairs; P.out_off = out_off;
    P.part = reinterpret_cast<unsigned long long *>(ws + pl.off_part); P.tickets = tickets;
    P.idx = idx; P.dist = dist; P.keep = keep; P.n_keep = cnt ? cnt : n_keep; P.ratio = ratio;
    static std::atomic<int> configured{-1};
    int dev = 0;
    cudaGetDevice(&dev);
    const int smem = TcSmem::total + 1024;
    if (configured.load() != dev) {
        MK_CUDA_TRY(cudaFuncSetAttribute(knn2_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        MK_CUDA_TRY(cudaFuncSetAttribute(knn2_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured.store(dev);
    }
    dim3 grid(pl.nqb, pl.splits, np);
    // programmatic dependent launch: the GEMM kernel's prologue (barriers, TMEM allocation) runs under the prep kernel
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    static const bool no_pdl = getenv("MK_NO_PDL") != nullptr;   // diagnosis switch: plain stream-ordered launch
    if (!no_pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (P.cluster_merge) {                   // the splits of a query block = one cluster (see the merge in the kernel)
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = (unsigned)pl.splits; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    cudaError_t le = P.hamming ? cudaLaunchKernelEx(&cfg, knn2_tc_kernel<1>, map_q, map_t, P) : cudaLaunchKernelEx(&cfg, knn2_tc_kernel<0>, map_q, map_t, P);
    if (le != cudaSuccess && P.cluster_merge) {
        // a cluster of `splits` one-CTA-per-SM blocks does not fit this stream's SMs (a small green-context partition, MIG):
        // same kernel without the cluster, the splits merge through global memory
        cudaGetLastError();
        P.cluster_merge = 0;
        cfg.numAttrs = na - 1;
        le = P.hamming ? cudaLaunchKernelEx(&cfg, knn2_tc_kernel<1>, map_q, map_t, P) : cudaLaunchKernelEx(&cfg, knn2_tc_kernel<0>, map_q, map_t, P);
    }
    MK_CUDA_TRY(le);
    MK_LAUNCH_CHECK("knn2_tc_kernel");
    *flags_out = flags;
    if (tc_count_out) *tc_count_out = cnt;
    return MK_OK;
}

}  // namespace mk
