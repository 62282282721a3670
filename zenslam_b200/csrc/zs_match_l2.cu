// zs_match_l2.cu -- L2 top-2 for 128-dim integer-valued descriptors (cv::SIFT) on the 5th-gen tensor cores.
//
// Reference: cv::BFMatcher(NORM_L2) behind zenslam::matcher for non-binary descriptors
// (zenslam_core/source/matching/matcher.cpp:60-80, matching_utils.cpp:63-95, keypoint_tracker.cpp:22); SURVEY A.6.
//
// cv::SIFT descriptors are integers 0..255 stored as float, so ||a-b||^2 = |a|^2 + |b|^2 - 2 a.b is an exact
// integer (< 2^24) and the dense contraction A.B^T is the one GEMM-shaped piece of the whole front-end:
//   * operands are the u8-quantised rows (128 bytes = exactly one 128-byte swizzle row, K-major),
//   * one CTA = one 128-query x 128-train tile: TMA (SWIZZLE_128B) stages both operand tiles in shared memory,
//     a single elected thread issues 4 x tcgen05.mma.kind::i8 (M128 N128 K32, u8 x u8 -> s32 accumulators in TMEM),
//     tcgen05.commit signals an mbarrier, four epilogue warps read their TMEM lane quarter with tcgen05.ld and
//     keep a stable top-2 per query (ascending train index, strict '<': OpenCV's tie rule),
//   * the train dimension is split across CTAs (grid = q-tiles x t-splits x pairs) so that even one 2000 x 2000
//     problem fills the machine; a small merge kernel folds the per-split top-2 in ascending split order.
// Integer MMA makes exactness a property of the instruction, not of rounding analysis.
#include <cuda.h>
#include <stdlib.h>

#include "zs_common.cuh"

#define L2TC_M 128
#define L2TC_N 128
#define L2TC_K 128           // bytes per descriptor row
#define L2TC_THREADS 192     // warp 0: TMEM alloc + TMA, warp 1: MMA issue, warps 2..5: epilogue

__device__ __forceinline__ uint32_t l2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void l2_mbar_init(uint32_t bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void l2_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void l2_mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// polling with back-off for the single-lane producer / MMA roles: they share their scheduler with an epilogue warp, and a
// tight try_wait loop steals its issue slots (ncu: ~20 % of the kernel's samples sat in those loops)
template <int NS = 64>
__device__ __forceinline__ void l2_mbar_wait_backoff(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(NS);
    }
}
__device__ __forceinline__ void l2_tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major; 1),
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024), [46,48) version = 1, [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t l2_smem_desc(uint32_t addr)
{
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 (2) at [4,6), a/b format U8 (0) at [7,10)/[10,13),
// a/b K-major (0) at 15/16, N >> 3 at [17,23), M >> 4 at [24,29)
#define L2TC_IDESC ((2u << 4) | ((uint32_t)(L2TC_N >> 3) << 17) | ((uint32_t)(L2TC_M >> 4) << 24))

struct l2tc_args {
    const int* nq; const int* nt;
    int cap_q, cap_t, splits;
    int4* part;            // [pairs][cap_q][splits] = (i0, i1, d0, d1); only splits below ceil(nt/128) are written
};

struct l2_top2 { int d0, d1, i0, i1; };
__device__ __forceinline__ void l2_top2_update(l2_top2& t, int d, int j)
{
    if (d < t.d0) { t.d1 = t.d0; t.i1 = t.i0; t.d0 = d; t.i0 = j; }
    else if (d < t.d1) { t.d1 = d; t.i1 = j; }
}

// grid: (ceil(cap_q/128), splits, pairs); dynamic smem: 1024 (alignment slack) + 16 KB A + 16 KB B
__global__ void __launch_bounds__(L2TC_THREADS) k_l2_tc_tile(const __grid_constant__ CUtensorMap map_q,
                                                             const __grid_constant__ CUtensorMap map_t, l2tc_args a)
{
    extern __shared__ uint8_t l2_smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[2];       // [0] operands landed, [1] accumulator ready
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) int s_tnorm[L2TC_N];
    const int pair = blockIdx.z, q0 = blockIdx.x * L2TC_M, t0 = blockIdx.y * L2TC_N;
    const int n_q = min(a.nq[pair], a.cap_q), n_t = min(a.nt[pair], a.cap_t);
    if (q0 >= n_q || t0 >= n_t) return;              // uniform per CTA: nothing allocated yet
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sm = (uint8_t*)(((uintptr_t)l2_smem_raw + 1023) & ~(uintptr_t)1023);     // swizzle atoms are 1024-byte aligned
    uint8_t* sA = sm; uint8_t* sB = sm + L2TC_M * L2TC_K;
    const uint32_t bar_full = l2_smem_u32(&s_bar[0]), bar_acc = l2_smem_u32(&s_bar[1]);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(l2_smem_u32(&s_tmem)), "r"(L2TC_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else if (warp == 1 && lane == 0) {
        l2_mbar_init(bar_full, 1);
        l2_mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp == 0) {
        if (lane == 0) {
            // rows past the pair's own count are other pairs' rows or zero fill: masked in the epilogue
            l2_mbar_expect_tx(bar_full, (L2TC_M + L2TC_N) * L2TC_K);
            l2_tma_load_2d(l2_smem_u32(sA), &map_q, 0, pair * a.cap_q + q0, bar_full);
            l2_tma_load_2d(l2_smem_u32(sB), &map_t, 0, pair * a.cap_t + t0, bar_full);
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            l2_mbar_wait(bar_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t da = l2_smem_desc(l2_smem_u32(sA)), db = l2_smem_desc(l2_smem_u32(sB));
#pragma unroll
            for (int k = 0; k < L2TC_K / 32; ++k) {
                // advancing K inside the 128-byte swizzle span = advancing the start address by 32 bytes (>> 4 = 2)
                const uint64_t dak = da + (uint64_t)(2 * k), dbk = db + (uint64_t)(2 * k);
                const uint32_t acc = k > 0 ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(dak), "l"(dbk), "r"(L2TC_IDESC), "r"(acc) : "memory");
            }
            // commit: arrives on bar_acc when the MMAs above have completed (implies fence::before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_acc) : "memory");
        }
        __syncwarp();
    } else {
        // ---- epilogue warps 2..5: TMEM lane quarter = warp % 4
        const int e = threadIdx.x - 64;                       // 0..127
        l2_mbar_wait(bar_full, 0);
        // |b|^2 of train row e and |a|^2 of this thread's query row, straight from the swizzled tiles (a row's
        // sixteen-byte chunks are permuted inside its own 128 bytes, which a norm does not care about)
        {
            const uint4* rb = (const uint4*)(sB + (size_t)e * L2TC_K);
            int s = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 v = rb[c ^ (e & 7)];
                s = (int)__dp4a(v.x, v.x, (unsigned)s); s = (int)__dp4a(v.y, v.y, (unsigned)s); s = (int)__dp4a(v.z, v.z, (unsigned)s); s = (int)__dp4a(v.w, v.w, (unsigned)s);
            }
            s_tnorm[e] = s;
        }
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;                  // accumulator row == TMEM lane
        int qn = 0;
        {
            const uint4* ra = (const uint4*)(sA + (size_t)row * L2TC_K);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 v = ra[c ^ (row & 7)];
                qn = (int)__dp4a(v.x, v.x, (unsigned)qn); qn = (int)__dp4a(v.y, v.y, (unsigned)qn); qn = (int)__dp4a(v.z, v.z, (unsigned)qn); qn = (int)__dp4a(v.w, v.w, (unsigned)qn);
            }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");       // s_tnorm complete (epilogue warps only)
        l2_mbar_wait(bar_acc, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // running top-2 over packed keys.  key = d2 * 128 + column orders by distance first, column second -- the stable rule
        // (smaller train index wins a tie) -- and d2 = |a|^2 + |b|^2 - 2 a.b < 2^23.  |a|^2 is the same for every column of
        // a row, so it is left out of the loop: key' = (|b|^2 * 128 + column) - 256 * dot is ONE IMAD per element (the
        // bracket is a per-column table in shared memory), followed by min / min / max for the two best.  Signed compare:
        // key' lies in (-2^31, 2^30).
        int k0 = 0x7fffffff, k1 = 0x7fffffff;
        const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16);
        const int ncol = min(L2TC_N, n_t - t0);
        s_tnorm[e] = s_tnorm[e] * 128 + e;                   // own entry only; published by the barrier below
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
        for (int c0 = 0; c0 < L2TC_N; c0 += 32) {
            uint32_t v[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr + (uint32_t)c0) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int4* tn4 = (const int4*)&s_tnorm[c0];
            if (c0 + 32 <= ncol) {                            // warp-uniform: a full chunk needs no per-column test
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const int4 tn = tn4[j4];
                    const int kk[4] = { tn.x - 256 * (int)v[4 * j4], tn.y - 256 * (int)v[4 * j4 + 1], tn.z - 256 * (int)v[4 * j4 + 2],
                                        tn.w - 256 * (int)v[4 * j4 + 3] };
#pragma unroll
                    for (int q = 0; q < 4; ++q) { k1 = max(k0, min(k1, kk[q])); k0 = min(k0, kk[q]); }
                }
            } else if (c0 < ncol) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c0 + j < ncol) {
                        const int key = s_tnorm[c0 + j] - 256 * (int)v[j];
                        k1 = max(k0, min(k1, key)); k0 = min(k0, key);
                    }
            }
        }
        if (q0 + row < n_q) {
            const int f0 = k0 + qn * 128, f1 = k1 + qn * 128;          // put |a|^2 back: (d2 << 7) | column
            const int i0 = k0 != 0x7fffffff ? t0 + (f0 & 127) : -1, i1 = k1 != 0x7fffffff ? t0 + (f1 & 127) : -1;
            a.part[((size_t)pair * a.cap_q + q0 + row) * a.splits + blockIdx.y] =
                make_int4(i0, i1, i0 >= 0 ? (f0 >> 7) : 0x7fffffff, i1 >= 0 ? (f1 >> 7) : 0x7fffffff);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(L2TC_N) : "memory");
    }
}

// fold the per-split top-2 in ascending split (= ascending train index) order; strict '<' keeps OpenCV's tie rule
__global__ void __launch_bounds__(256) k_l2_tc_merge(l2tc_args a, int* __restrict__ o_idx, int* __restrict__ o_dist)
{
    const int pair = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_q = min(a.nq[pair], a.cap_q), n_t = min(a.nt[pair], a.cap_t);
    if (q >= n_q) return;
    const int4* p = a.part + ((size_t)pair * a.cap_q + q) * a.splits;
    l2_top2 best = { 0x7fffffff, 0x7fffffff, -1, -1 };
    const int valid = (n_t + L2TC_N - 1) / L2TC_N;
    for (int s = 0; s < valid; ++s) {
        const int4 v = p[s];
        if (v.x >= 0) l2_top2_update(best, v.z, v.x);
        if (v.y >= 0) l2_top2_update(best, v.w, v.y);
    }
    const size_t o = 2 * ((size_t)pair * a.cap_q + q);
    o_idx[o] = best.i0; o_idx[o + 1] = best.i1; o_dist[o] = best.d0; o_dist[o + 1] = best.d1;
}

// ------------------------------------------------------------------------------------------------------
// Persistent, pipelined variant: a CTA keeps its 128-query tile in shared memory and walks `tpc` consecutive train
// tiles.  Three roles run concurrently and hand tiles to one another through mbarriers:
//   warp 0 (one lane)   TMA producer: train tile i+1 streams into the other half of a 2-stage ring while
//   warp 1 (one lane)   the MMA issuer runs the 4 x tcgen05.mma of tile i into TMEM accumulator (i & 1), and
//   warps 2..5          the epilogue drains accumulator (i-1 & 1) with tcgen05.ld and updates the running top-2.
// tcgen05.commit releases both the smem stage (back to the producer) and the accumulator (on to the epilogue);
// the epilogue releases the accumulator back to the MMA issuer as soon as its tcgen05.ld have landed.
// Row norms come from a small pre-pass (k_l2_row_norms) so the epilogue never touches the operand tiles.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_l2_row_norms(const uint8_t* __restrict__ rows, size_t nrows, int* __restrict__ norms)
{
    const size_t r = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= nrows) return;
    const uint32_t v = ((const uint32_t*)(rows + r * L2TC_K))[threadIdx.x & 31];
    int s = (int)__dp4a(v, v, 0u);
    s = __reduce_add_sync(0xffffffffu, s);
    if ((threadIdx.x & 31) == 0) norms[r] = s;
}

struct l2p_args {
    const int* nq; const int* nt;
    int cap_q, cap_t, splits, tpc;
    const int* qnorm; const int* tnorm;        // [pairs*cap_q], [pairs*cap_t]
    int4* part;
};

__device__ __forceinline__ void l2_mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

#define L2P_BSTAGES 4
// back-off of the two single-lane roles while they wait (ns): both have slack -- the operand ring is 4 tiles deep and the
// MMA of tile i+2 only has to land before the epilogue finishes tile i+1 -- and every poll they issue competes with the
// epilogue warps of lane quarters 0 and 1 for the same schedulers and the same ALU pipe
#ifndef L2P_SLEEP_TMA
#define L2P_SLEEP_TMA 512
#endif
#ifndef L2P_SLEEP_MMA
#define L2P_SLEEP_MMA 256
#endif
__device__ __forceinline__ int l2_min3(int a, int b, int c)
{
    int d;
    asm("min.s32 %0, %1, %2;\n\tmin.s32 %0, %0, %3;" : "=&r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned l2_min3u(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm("min.u32 %0, %1, %2;\n\tmin.u32 %0, %0, %3;" : "=&r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// smallest of N keys as four independent min3 chains (N = 16 or 32)
template <int N>
__device__ __forceinline__ int l2_tree_min(const int (&k)[N])
{
    int m[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int* g = &k[q * (N / 4)];
        int v = l2_min3(g[0], g[1], g[2]);
#pragma unroll
        for (int j = 3; j + 1 < N / 4; j += 2) v = l2_min3(v, g[j], g[j + 1]);
        if ((N / 4) % 2 == 0) v = min(v, g[N / 4 - 1]);
        m[q] = v;
    }
    return min(l2_min3(m[0], m[1], m[2]), m[3]);
}
// smallest of the N values (unsigned)(k - base): with base = smallest key + 1 the smallest key itself wraps to 2^32 - 1 and
// the result is (second smallest key) - base
template <int N>
__device__ __forceinline__ unsigned l2_tree_min_rebased(const int (&k)[N], unsigned base)
{
    unsigned m[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int* g = &k[q * (N / 4)];
        unsigned v = l2_min3u((unsigned)g[0] - base, (unsigned)g[1] - base, (unsigned)g[2] - base);
#pragma unroll
        for (int j = 3; j + 1 < N / 4; j += 2) v = l2_min3u(v, (unsigned)g[j] - base, (unsigned)g[j + 1] - base);
        if ((N / 4) % 2 == 0) v = min(v, (unsigned)g[N / 4 - 1] - base);
        m[q] = v;
    }
    return min(l2_min3u(m[0], m[1], m[2]), m[3]);
}
// CG = epilogue warps per TMEM lane quarter (the 128 accumulator columns of a tile are split into CG groups of
// 128/CG); threads = 64 + 128 CG.  One epilogue warp per scheduler (CG = 1) leaves the drain latency-bound -- a
// warp cannot issue its dependent min/max chain back to back -- so the MMA issuer idles on acc_empty; with CG = 2 or
// 4 every scheduler interleaves 4 to 8 epilogue warps (2 CTAs per SM) and the drain approaches the issue rate.
// grid: (ceil(cap_q/128), splits, pairs); dynamic smem: 1024 slack + 16 KB A + L2P_BSTAGES x 16 KB B
// KH = 128-byte K halves per descriptor row: 1 = 128-d u8 rows (SIFT), 2 = 256-d rows -- binary descriptors (ORB) expanded to
// one 0 / 1 byte per bit, for which |a - b|^2 IS the Hamming distance: the query tile then holds two swizzled 16 KB halves, a
// train tile passes through the operand ring as two stages, and the second half's four MMAs accumulate onto the first's.
template <int CG, bool TREE = false, int KH = 1>
__global__ void __launch_bounds__(64 + 128 * CG, 2) k_l2_tc_persist(const __grid_constant__ CUtensorMap map_q,
                                                                     const __grid_constant__ CUtensorMap map_t, l2p_args a)
{
    extern __shared__ uint8_t l2_smem_raw[];
    // barriers: 0 full_a | 1..NB full_b | 1+NB..2NB empty_b | 1+2NB, 2+2NB acc_full | 3+2NB, 4+2NB acc_empty
    constexpr int NB = L2P_BSTAGES;
    constexpr int NCOL = L2TC_N / CG;                // accumulator columns per epilogue warp
    constexpr int CW = CG >= 4 ? 16 : 32;            // columns per tcgen05.ld (register budget: 56 at 576 threads x 2 CTAs)
    __shared__ __align__(8) uint64_t s_bar[5 + 2 * NB];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) int s_tn[4 * CG][NCOL];             // per-warp column key table (private: no CTA barrier)
    __shared__ __align__(16) int4 s_best[CG > 1 ? CG - 1 : 1][L2TC_M];
    const int pair = blockIdx.z, q0 = blockIdx.x * L2TC_M;
    const int n_q = min(a.nq[pair], a.cap_q), n_t = min(a.nt[pair], a.cap_t);
    const int tiles_total = (n_t + L2TC_N - 1) / L2TC_N, tile_begin = blockIdx.y * a.tpc;
    if (q0 >= n_q || tile_begin >= tiles_total) return;          // uniform per CTA
    const int ntile = min(a.tpc, tiles_total - tile_begin);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sm = (uint8_t*)(((uintptr_t)l2_smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = sm; uint8_t* sB0 = sm + KH * L2TC_M * L2TC_K;
    const uint32_t bar0 = l2_smem_u32(&s_bar[0]);
#define BAR(i) (bar0 + 8u * (uint32_t)(i))

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(l2_smem_u32(&s_tmem)), "r"(2 * L2TC_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else if (warp == 1 && lane == 0) {
        for (int i = 0; i < 3 + 2 * NB; ++i) l2_mbar_init(BAR(i), 1);
        l2_mbar_init(BAR(3 + 2 * NB), 4 * CG); l2_mbar_init(BAR(4 + 2 * NB), 4 * CG);   // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp == 0) {
        if (lane == 0) {
            l2_mbar_expect_tx(BAR(0), KH * L2TC_M * L2TC_K);
#pragma unroll
            for (int h = 0; h < KH; ++h)
                l2_tma_load_2d(l2_smem_u32(sA + h * (L2TC_M * L2TC_K)), &map_q, h * L2TC_K, pair * a.cap_q + q0, BAR(0));
            for (int r = 0; r < ntile * KH; ++r) {                               // ring item r = (train tile r / KH, K half r % KH)
                const int sb = r % NB, pb = (r / NB) & 1;
                l2_mbar_wait_backoff<L2P_SLEEP_TMA>(BAR(1 + NB + sb), pb ^ 1);   // stage free (passes at once the first time round)
                l2_mbar_expect_tx(BAR(1 + sb), L2TC_N * L2TC_K);
                l2_tma_load_2d(l2_smem_u32(sB0 + sb * (L2TC_N * L2TC_K)), &map_t, (r % KH) * L2TC_K,
                               pair * a.cap_t + (tile_begin + r / KH) * L2TC_N, BAR(1 + sb));
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            l2_mbar_wait(BAR(0), 0);
            for (int i = 0; i < ntile; ++i) {
                const int st = i & 1, ph = (i >> 1) & 1;          // accumulator ring (2 deep)
                l2_mbar_wait_backoff<L2P_SLEEP_MMA>(BAR(3 + 2 * NB + st), ph ^ 1);   // accumulator drained
                const uint32_t d = tmem + (uint32_t)(st * L2TC_N);
#pragma unroll
                for (int h = 0; h < KH; ++h) {
                    const int r = i * KH + h;
                    const int sb = r % NB, pb = (r / NB) & 1;     // operand ring (NB deep: covers the TMA round trip)
                    l2_mbar_wait_backoff<L2P_SLEEP_MMA>(BAR(1 + sb), pb);        // this half of the train tile landed
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = l2_smem_desc(l2_smem_u32(sA + h * (L2TC_M * L2TC_K)));
                    const uint64_t db = l2_smem_desc(l2_smem_u32(sB0 + sb * (L2TC_N * L2TC_K)));
#pragma unroll
                    for (int k = 0; k < L2TC_K / 32; ++k) {
                        const uint64_t dak = da + (uint64_t)(2 * k), dbk = db + (uint64_t)(2 * k);
                        const uint32_t acc = (h > 0 || k > 0) ? 1u : 0u;
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(d), "l"(dak), "l"(dbk), "r"(L2TC_IDESC), "r"(acc) : "memory");
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(BAR(1 + NB + sb)) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(BAR(1 + 2 * NB + st)) : "memory");
            }
        }
        __syncwarp();
    } else {
        // epilogue warp w = warp - 2: TMEM lane quarter = warp % 4 (hardware rule), column group cg = w / 4
        const int ew = warp - 2, quarter = warp & 3, cg = ew >> 2, row = quarter * 32 + lane;
        const int qn = (q0 + row < n_q) ? a.qnorm[(size_t)pair * a.cap_q + q0 + row] : 0;
        l2_top2 best = { 0x7fffffff, 0x7fffffff, -1, -1 };
        int* tn = s_tn[ew];
        // the train norms of tile i+1 are fetched while tile i is being reduced; lane l owns columns l, l+32, .. of
        // the warp's column group
        const int* tnp = a.tnorm + (size_t)pair * a.cap_t;
        int tn_next[NCOL / 32 > 0 ? NCOL / 32 : 1];
#pragma unroll
        for (int u = 0; u < NCOL / 32; ++u) {
            const int col = tile_begin * L2TC_N + cg * NCOL + u * 32 + lane;
            tn_next[u] = col < n_t ? tnp[col] : 0;
        }
        for (int i = 0; i < ntile; ++i) {
            const int st = i & 1, ph = (i >> 1) & 1;
            const int t0 = (tile_begin + i) * L2TC_N, ncol = min(NCOL, n_t - t0 - cg * NCOL);   // may be <= 0
            // per-column key table of this warp's columns: |b|^2 * 128 + column (see k_l2_tc_tile)
            __syncwarp();                                        // every lane is done with the previous tile's table
#pragma unroll
            for (int u = 0; u < NCOL / 32; ++u) {
                tn[u * 32 + lane] = tn_next[u] * 128 + cg * NCOL + u * 32 + lane;
                const int col = t0 + L2TC_N + cg * NCOL + u * 32 + lane;
                if (i + 1 < ntile) tn_next[u] = col < n_t ? tnp[col] : 0;
            }
            __syncwarp();
            l2_mbar_wait(BAR(1 + 2 * NB + st), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // four independent (best, second) chains -- element j feeds chain j % 4 -- so the min / max updates of
            // consecutive elements do not wait on one another; keys are unique per column, so the order in which
            // candidates are folded does not matter
            int c0k[4] = { 0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff }, c1k[4] = { 0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff };
            const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(st * L2TC_N + cg * NCOL);
#pragma unroll 1
            for (int c0 = 0; c0 < NCOL; c0 += CW) {
                uint32_t v[CW];
                if (CW == 32)
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                                   "=r"(v[16 % CW]), "=r"(v[17 % CW]), "=r"(v[18 % CW]), "=r"(v[19 % CW]), "=r"(v[20 % CW]), "=r"(v[21 % CW]),
                                   "=r"(v[22 % CW]), "=r"(v[23 % CW]), "=r"(v[24 % CW]), "=r"(v[25 % CW]), "=r"(v[26 % CW]), "=r"(v[27 % CW]),
                                   "=r"(v[28 % CW]), "=r"(v[29 % CW]), "=r"(v[30 % CW]), "=r"(v[31 % CW])
                                 : "r"(taddr + (uint32_t)c0) : "memory");
                else
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                                 : "r"(taddr + (uint32_t)c0) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c0 + CW >= NCOL) {
                    // the last chunk is in registers: hand the accumulator back to the MMA issuer before the arithmetic
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) l2_mbar_arrive(BAR(3 + 2 * NB + st));
                }
                const int4* tn4 = (const int4*)&tn[c0];
                if (TREE && c0 + CW <= ncol) {
                    // the chunk's two smallest keys by two min3 trees (keys are unique: the column sits in their low bits): the
                    // smallest, then the smallest of the keys re-based on it as unsigned numbers (ptxas fuses most of the
                    // subtractions into VIADDMNMX) -- tree-shaped dependencies and ~2.1 operations per element instead of 2.5
                    // on four serial chains
                    int kk[CW];
#pragma unroll
                    for (int j4 = 0; j4 < CW / 4; ++j4) {
                        const int4 t4 = tn4[j4];
                        kk[4 * j4] = t4.x - 256 * (int)v[4 * j4]; kk[4 * j4 + 1] = t4.y - 256 * (int)v[4 * j4 + 1];
                        kk[4 * j4 + 2] = t4.z - 256 * (int)v[4 * j4 + 2]; kk[4 * j4 + 3] = t4.w - 256 * (int)v[4 * j4 + 3];
                    }
                    const int m0 = l2_tree_min<CW>(kk);
                    const unsigned base = (unsigned)m0 + 1u;
                    const int m1 = (int)(l2_tree_min_rebased<CW>(kk, base) + base);
                    c1k[0] = l2_min3(c1k[0], m1, max(c0k[0], m0)); c0k[0] = min(c0k[0], m0);
                } else if (c0 + CW <= ncol) {
#pragma unroll
                    for (int j4 = 0; j4 < CW / 4; ++j4) {
                        const int4 t4 = tn4[j4];
                        const int kk[4] = { t4.x - 256 * (int)v[4 * j4], t4.y - 256 * (int)v[4 * j4 + 1], t4.z - 256 * (int)v[4 * j4 + 2],
                                            t4.w - 256 * (int)v[4 * j4 + 3] };
                        // two candidates per chain update: lo/hi of the pair, then second' = min3(second, hi, max(best, lo)),
                        // best' = min(best, lo) -- 5 ALU-pipe instructions per 2 elements instead of 6 (the epilogue is bound
                        // by the half-rate ALU pipe that executes VIMNMX, not by issue slots or the tensor pipe)
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int lo = min(kk[2 * q], kk[2 * q + 1]), hi = max(kk[2 * q], kk[2 * q + 1]);
                            const int ch = (2 * j4 + q) & 3;
                            c1k[ch] = l2_min3(c1k[ch], hi, max(c0k[ch], lo)); c0k[ch] = min(c0k[ch], lo);
                        }
                    }
                } else if (c0 < ncol) {
#pragma unroll
                    for (int j = 0; j < CW; ++j)
                        if (c0 + j < ncol) {
                            const int key = tn[c0 + j] - 256 * (int)v[j];
                            c1k[j & 3] = max(c0k[j & 3], min(c1k[j & 3], key)); c0k[j & 3] = min(c0k[j & 3], key);
                        }
                }
            }
            // the two smallest of the eight chain values
            int k0 = c0k[0], k1 = c1k[0];
#pragma unroll
            for (int q = 1; q < 4; ++q) {
                k1 = max(k0, min(k1, c0k[q])); k0 = min(k0, c0k[q]);
                k1 = min(k1, c1k[q]);                        // c1k[q] >= c0k[q] >= new k0: it can only replace the second best
            }
            // fold the tile's two best into the running top-2 (a warp sees its columns in ascending train order; strict '<')
            if (k0 != 0x7fffffff) { const int f = k0 + qn * 128; l2_top2_update(best, f >> 7, t0 + (f & 127)); }
            if (k1 != 0x7fffffff) { const int f = k1 + qn * 128; l2_top2_update(best, f >> 7, t0 + (f & 127)); }
        }
        // fold the CG column groups of a row: train indices interleave between groups, so ties are broken explicitly
        // (smaller index wins -- the order cv::BFMatcher's ascending scan with strict '<' produces)
        if (CG > 1) {
            if (cg > 0) s_best[cg - 1][row] = make_int4(best.i0, best.i1, best.d0, best.d1);
            asm volatile("bar.sync 1, %0;" ::"n"(128 * CG) : "memory");
            if (cg == 0) {
#pragma unroll
                for (int g = 0; g < CG - 1; ++g) {
                    const int4 o = s_best[g][row];
                    const int od[2] = { o.z, o.w }, oi[2] = { o.x, o.y };
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (oi[u] < 0) continue;
                        if (od[u] < best.d0 || (od[u] == best.d0 && oi[u] < best.i0)) {
                            best.d1 = best.d0; best.i1 = best.i0; best.d0 = od[u]; best.i0 = oi[u];
                        } else if (od[u] < best.d1 || (od[u] == best.d1 && oi[u] < best.i1)) {
                            best.d1 = od[u]; best.i1 = oi[u];
                        }
                    }
                }
            }
        }
        if (cg == 0 && q0 + row < n_q)
            a.part[((size_t)pair * a.cap_q + q0 + row) * a.splits + blockIdx.y] = make_int4(best.i0, best.i1, best.d0, best.d1);
    }
#undef BAR
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * L2TC_N) : "memory");
    }
}

// merge for the persistent kernel: split s covers train tiles [s*tpc, (s+1)*tpc)
__global__ void __launch_bounds__(256) k_l2p_merge(l2p_args a, int* __restrict__ o_idx, int* __restrict__ o_dist)
{
    const int pair = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_q = min(a.nq[pair], a.cap_q), n_t = min(a.nt[pair], a.cap_t);
    if (q >= n_q) return;
    const int4* p = a.part + ((size_t)pair * a.cap_q + q) * a.splits;
    l2_top2 best = { 0x7fffffff, 0x7fffffff, -1, -1 };
    const int tiles_total = (n_t + L2TC_N - 1) / L2TC_N;
    const int valid = (tiles_total + a.tpc - 1) / a.tpc;
    for (int s = 0; s < valid; ++s) {
        const int4 v = p[s];
        if (v.x >= 0) l2_top2_update(best, v.z, v.x);
        if (v.y >= 0) l2_top2_update(best, v.w, v.y);
    }
    const size_t o = 2 * ((size_t)pair * a.cap_q + q);
    o_idx[o] = best.i0; o_idx[o + 1] = best.i1; o_dist[o] = best.d0; o_dist[o + 1] = best.d1;
}

static zs_status l2_make_map(CUtensorMap* m, const uint8_t* base, size_t rows, int row_bytes = L2TC_K)
{
    zs_encode_tiled_fn enc = zs_get_encode_tiled();
    if (!enc) { zs_set_error("cuTensorMapEncodeTiled is not available from this driver"); return ZS_ERR_CUDA; }
    const cuuint64_t dims[2] = { (cuuint64_t)row_bytes, (cuuint64_t)rows };
    const cuuint64_t strides[1] = { (cuuint64_t)row_bytes };
    const cuuint32_t box[2] = { L2TC_K, L2TC_M }, es[2] = { 1, 1 };
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { zs_set_error("cuTensorMapEncodeTiled(L2 descriptors) failed: %d", (int)r); return ZS_ERR_CUDA; }
    return ZS_OK;
}

// q8 / t8: [pairs][cap][128] u8 (16-byte aligned); idx / dist: [pairs][cap_q][2] (squared distances as int);
// part: zs_l2_tensor_part_ints() ints of scratch for the per-split partial top-2
zs_status zs_l2_tensor_top2(zs_context* ctx, const uint8_t* q8, const int* nq, const uint8_t* t8, const int* nt, int pairs,
                            int cap_q, int cap_t, int dim, int* idx, int* dist, void* part, const int* qnorm, const int* tnorm)
{
    ZS_REQUIRE(dim == L2TC_K || dim == 2 * L2TC_K, "tensor-core path needs 128- or 256-byte descriptor rows");
    ZS_REQUIRE(((uintptr_t)q8 % 16) == 0 && ((uintptr_t)t8 % 16) == 0, "descriptor arrays must be 16-byte aligned");
    CUtensorMap mq, mt;
    zs_status st = l2_make_map(&mq, q8, (size_t)pairs * cap_q, dim);
    if (st != ZS_OK) return st;
    if ((st = l2_make_map(&mt, t8, (size_t)pairs * cap_t, dim)) != ZS_OK) return st;
    const int q_tiles = zs_div_up(cap_q, L2TC_M), t_tiles = zs_div_up(cap_t, L2TC_N);
    if (dim == 2 * L2TC_K) {
        // 256-byte rows (expanded binary descriptors): two K halves; the caller always supplies the norms (bit counts)
        ZS_REQUIRE(qnorm && tnorm, "256-byte rows need their norms");
        int splits = (int)((4LL * ctx->sm_count + (long long)q_tiles * pairs - 1) / ((long long)q_tiles * pairs));
        splits = splits < 1 ? 1 : splits > t_tiles ? t_tiles : splits;
        l2p_args b;
        b.nq = nq; b.nt = nt; b.cap_q = cap_q; b.cap_t = cap_t;
        b.tpc = zs_div_up(t_tiles, splits); b.splits = zs_div_up(t_tiles, b.tpc);
        b.part = (int4*)part; b.qnorm = qnorm; b.tnorm = tnorm;
        const size_t smem = 1024 + (size_t)(2 * L2TC_M + L2P_BSTAGES * L2TC_N) * L2TC_K;
        static bool attr_h = false;
        if (!attr_h) {
            ZS_CUDA(cudaFuncSetAttribute((k_l2_tc_persist<2, true, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_h = true;
        }
        k_l2_tc_persist<2, true, 2><<<dim3(q_tiles, b.splits, pairs), 64 + 256, smem, ctx->stream>>>(mq, mt, b);
        ZS_LAUNCH_CHECK(ctx);
        k_l2p_merge<<<dim3(zs_div_up(cap_q, 256), pairs), 256, 0, ctx->stream>>>(b, idx, dist);
        ZS_LAUNCH_CHECK(ctx);
        return ZS_OK;
    }
    if (!ctx->sw.l2_one_tile) {
        // persistent kernel: enough CTAs for two waves of 2 CTAs per SM, otherwise as many train tiles per CTA as possible
        int splits = (int)((4LL * ctx->sm_count + (long long)q_tiles * pairs - 1) / ((long long)q_tiles * pairs));
        if (ctx->sw.l2_splits > 0) splits = ctx->sw.l2_splits;
        splits = splits < 1 ? 1 : splits > t_tiles ? t_tiles : splits;
        l2p_args b;
        b.nq = nq; b.nt = nt; b.cap_q = cap_q; b.cap_t = cap_t;
        b.tpc = zs_div_up(t_tiles, splits); b.splits = zs_div_up(t_tiles, b.tpc);
        b.part = (int4*)part;
        int* norms = (int*)part + 4 * (size_t)pairs * cap_q * t_tiles;           // behind the (worst-case) partial area
        b.qnorm = qnorm ? qnorm : norms; b.tnorm = tnorm ? tnorm : norms + (size_t)pairs * cap_q;
        if (!qnorm) {                          // u8 rows from the caller: the f32 -> u8 pass that leaves the norms did not run
            k_l2_row_norms<<<zs_div_up(pairs * cap_q, 8), 256, 0, ctx->stream>>>(q8, (size_t)pairs * cap_q, norms);
            ZS_LAUNCH_CHECK(ctx);
        }
        if (!tnorm) {
            k_l2_row_norms<<<zs_div_up(pairs * cap_t, 8), 256, 0, ctx->stream>>>(t8, (size_t)pairs * cap_t, norms + (size_t)pairs * cap_q);
            ZS_LAUNCH_CHECK(ctx);
        }
        const size_t smem = 1024 + (size_t)(L2TC_M + L2P_BSTAGES * L2TC_N) * L2TC_K;
        static bool attr_p = false;
        if (!attr_p) {
            ZS_CUDA(cudaFuncSetAttribute(k_l2_tc_persist<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ZS_CUDA(cudaFuncSetAttribute(k_l2_tc_persist<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ZS_CUDA(cudaFuncSetAttribute(k_l2_tc_persist<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ZS_CUDA(cudaFuncSetAttribute((k_l2_tc_persist<1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ZS_CUDA(cudaFuncSetAttribute((k_l2_tc_persist<2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ZS_CUDA(cudaFuncSetAttribute((k_l2_tc_persist<4, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_p = true;
        }
        const int cgsel = ctx->sw.l2_epi_groups > 0 ? ctx->sw.l2_epi_groups : 2;   // 1, 2 or 4 epilogue warps per TMEM lane quarter
        const dim3 grid(q_tiles, b.splits, pairs);
        // default: two epilogue warps per lane quarter, chunk top-2 by two min3 trees (whole 64-pair call 0.1329 -> 0.1283 ms against
        // the four serial chains, ZS_L2_CHAINS)
        const bool tree = !ctx->sw.l2_chains;
        if (cgsel == 1) { if (tree) k_l2_tc_persist<1, true><<<grid, 64 + 128, smem, ctx->stream>>>(mq, mt, b); else k_l2_tc_persist<1><<<grid, 64 + 128, smem, ctx->stream>>>(mq, mt, b); }
        else if (cgsel == 2) { if (tree) k_l2_tc_persist<2, true><<<grid, 64 + 256, smem, ctx->stream>>>(mq, mt, b); else k_l2_tc_persist<2><<<grid, 64 + 256, smem, ctx->stream>>>(mq, mt, b); }
        else { if (tree) k_l2_tc_persist<4, true><<<grid, 64 + 512, smem, ctx->stream>>>(mq, mt, b); else k_l2_tc_persist<4><<<grid, 64 + 512, smem, ctx->stream>>>(mq, mt, b); }
        ZS_LAUNCH_CHECK(ctx);
        k_l2p_merge<<<dim3(zs_div_up(cap_q, 256), pairs), 256, 0, ctx->stream>>>(b, idx, dist);
        ZS_LAUNCH_CHECK(ctx);
        return ZS_OK;
    }
    l2tc_args a;
    a.nq = nq; a.nt = nt; a.cap_q = cap_q; a.cap_t = cap_t; a.splits = zs_div_up(cap_t, L2TC_N);
    ZS_REQUIRE(part && ((uintptr_t)part % 16) == 0, "partial top-2 scratch must be 16-byte aligned");
    a.part = (int4*)part;                                   // zs_l2_tensor_part_ints() ints of caller scratch
    const size_t smem = 1024 + (size_t)(L2TC_M + L2TC_N) * L2TC_K;
    static bool attr_set = false;
    if (!attr_set) {
        ZS_CUDA(cudaFuncSetAttribute(k_l2_tc_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    k_l2_tc_tile<<<dim3(zs_div_up(cap_q, L2TC_M), a.splits, pairs), L2TC_THREADS, smem, ctx->stream>>>(mq, mt, a);
    ZS_LAUNCH_CHECK(ctx);
    // unmatched rows keep (-1, INT_MAX) like the CUDA-core kernel; rows >= nq are not touched
    k_l2_tc_merge<<<dim3(zs_div_up(cap_q, 256), pairs), 256, 0, ctx->stream>>>(a, idx, dist);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

size_t zs_l2_tensor_part_ints(int pairs, int cap_q, int cap_t)
{
    // per-split partial top-2 (worst case: one split per train tile) + the row norms of both sides
    return 4 * (size_t)pairs * cap_q * zs_div_up(cap_t, L2TC_N) + (size_t)pairs * ((size_t)cap_q + cap_t) + 64;
}
