// zs_match.cu -- brute-force descriptor matching: cv::BFMatcher knnMatch(k=2) + Lowe ratio, and
// match() with crossCheck=true, Hamming norm.
// Reference: zenslam::matcher (zenslam_core/source/matching/matcher.cpp:60-80) and utils::create_matcher
// (zenslam_core/source/matching/matching_utils.cpp:63-95); tie rules in SURVEY A.5.
//
// Popcount tiles: one thread owns a query descriptor in registers (8 x u32); the block stages tiles of
// train descriptors in shared memory and every thread sweeps the tile with broadcast 128-bit reads,
// XOR + POPC, keeping a running top-2.  Train indices are visited in ascending order with strict '<'
// updates, which is exactly OpenCV's stable tie rule (smaller train index wins).  Integer-ALU bound.
#include <stdlib.h>

#include <algorithm>

#include "zs_common.cuh"

#define MATCH_THREADS 128
#define MATCH_TILE 128

struct top2 { int d0, d1, i0, i1; };

__device__ __forceinline__ void top2_update(top2& t, int d, int j)
{
    if (d < t.d0) { t.d1 = t.d0; t.i1 = t.i0; t.d0 = d; t.i0 = j; }
    else if (d < t.d1) { t.d1 = d; t.i1 = j; }
}

// grid: (ceil(cap_q / 128), splits, pairs).  With splits > 1 (large rectangular problems: new keypoints against the
// landmark map, keypoint_tracker.cpp:199-291) block y sweeps train rows [y*chunk, (y+1)*chunk) and writes a partial
// top-2 (int4 i0,i1,d0,d1) per (query, split); k_top2_merge folds the partials in ascending split order.
// The running (best, second) of a query as packed keys (distance << 22 | train row): distances are at most 256 and a
// launch has fewer than 2^22 train rows (checked by the launcher), so a key orders by distance first and train row second --
// the smaller train row wins a tie, OpenCV's rule -- and the update is three VIMNMX instead of a compare /
// select chain of eight.
#define HAMMING_ROW_BITS 22
#define HAMMING_NONE 0x7fffffff
struct top2k { int k0, k1; };
__device__ __forceinline__ void top2k_update(top2k& t, int key)
{
    const int hi = max(t.k0, key);
    t.k0 = min(t.k0, key);
    t.k1 = min(t.k1, hi);
}

// Key of one (query, train row) distance.  Popcount of a 256-bit XOR with 4 POPC instead of 8: POPC issues at a quarter of
// the LOP3 rate on sm_100 (the first kernel was POPC-bound: 1.28 G POPC per 128-pair batch at ~16 per clock per SM), so seven
// of the eight words go through four full adders (LOP3 0x96 = a^b^c, 0xe8 = majority) and leave one word of weight 1 (plus
// the eighth word, counted as it is), one of weight 2 and one of weight 4.  Round 1 folded the eighth word and the carries
// through three more half adders into (ones, twos, fours, eights): the same four POPC for six more LOP3, on the ALU pipe
// that bounds the kernel (ncu: ALU 85 %, XU 52 %).  The weights and the row index are applied by IMADs (FMA pipe, idle).
__device__ __forceinline__ int hamming_key(const uint4& qa, const uint4& qb, const uint4& a, const uint4& b, int row)
{
    const unsigned w0 = qa.x ^ a.x, w1 = qa.y ^ a.y, w2 = qa.z ^ a.z, w3 = qa.w ^ a.w;
    const unsigned w4 = qb.x ^ b.x, w5 = qb.y ^ b.y, w6 = qb.z ^ b.z, w7 = qb.w ^ b.w;
    const unsigned s1 = w0 ^ w1 ^ w2, c1 = (w0 & w1) | (w2 & (w0 | w1));
    const unsigned s2 = w3 ^ w4 ^ w5, c2 = (w3 & w4) | (w5 & (w3 | w4));
    const unsigned s3 = s1 ^ s2 ^ w6, c3 = (s1 & s2) | (w6 & (s1 | s2));
    const unsigned s5 = c1 ^ c2 ^ c3, c5 = (c1 & c2) | (c3 & (c1 | c2));
    int key = row;
    key = __popc(s3) * (1 << HAMMING_ROW_BITS) + key;
    key = __popc(w7) * (1 << HAMMING_ROW_BITS) + key;
    key = __popc(s5) * (2 << HAMMING_ROW_BITS) + key;
    key = __popc(c5) * (4 << HAMMING_ROW_BITS) + key;
    return key;
}

// Each thread owns HQ query descriptors in registers (8 x u32 each): one pair of broadcast 128-bit shared loads of a train
// row then feeds HQ distances, so the shared-memory pipe (the limiter of the one-query-per-thread form: ncu showed
// mio_throttle as the top stall) carries a quarter of the traffic per distance.
#define HAMMING_THREADS 128
template <int HQ, bool CSA>
__global__ void __launch_bounds__(HAMMING_THREADS) k_hamming_top2(
    const uint8_t* __restrict__ q, const int* __restrict__ nq, size_t q_stride,
    const uint8_t* __restrict__ t, const int* __restrict__ nt, size_t t_stride,
    int* __restrict__ o_idx, int* __restrict__ o_dist, int out_stride /* ints per pair */,
    int chunk, int splits, int cap_q, int4* __restrict__ part)
{
    __shared__ uint4 tile[MATCH_TILE * 2];
    const int pair = blockIdx.z;
    const int n_q = nq[pair], n_t = nt[pair];
    const int q0 = blockIdx.x * (HAMMING_THREADS * HQ);
    if (q0 >= n_q) return;
    const int t_begin = blockIdx.y * chunk, t_end = min(n_t, t_begin + chunk);
    if (splits > 1 && t_begin >= n_t) return;             // the merge only reads splits that exist
    const uint4* qp = (const uint4*)(q + (size_t)pair * q_stride);
    const uint4* tp = (const uint4*)(t + (size_t)pair * t_stride);
    // query k of this thread is row q0 + k*HAMMING_THREADS + threadIdx.x (consecutive threads read consecutive rows)
    uint4 qa[HQ], qb[HQ];
    top2k best[HQ];
#pragma unroll
    for (int k = 0; k < HQ; ++k) {
        const int qi = q0 + k * HAMMING_THREADS + threadIdx.x;
        qa[k] = make_uint4(0, 0, 0, 0); qb[k] = qa[k];
        if (qi < n_q) { qa[k] = qp[2 * qi]; qb[k] = qp[2 * qi + 1]; }
        best[k].k0 = best[k].k1 = HAMMING_NONE;
    }
    for (int base = t_begin; base < t_end; base += MATCH_TILE) {
        const int m = min(MATCH_TILE, t_end - base);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * m; i += HAMMING_THREADS) tile[i] = tp[2 * base + i];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < m; ++j) {
            const uint4 a = tile[2 * j], b = tile[2 * j + 1];
#pragma unroll
            for (int k = 0; k < HQ; ++k) {
                int key;
                if (CSA) key = hamming_key(qa[k], qb[k], a, b, base + j);
                else key = (__popc(qa[k].x ^ a.x) + __popc(qa[k].y ^ a.y) + __popc(qa[k].z ^ a.z) + __popc(qa[k].w ^ a.w) +
                            __popc(qb[k].x ^ b.x) + __popc(qb[k].y ^ b.y) + __popc(qb[k].z ^ b.z) + __popc(qb[k].w ^ b.w)) *
                               (1 << HAMMING_ROW_BITS) + (base + j);
                top2k_update(best[k], key);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < HQ; ++k) {
        const int qi = q0 + k * HAMMING_THREADS + threadIdx.x;
        if (qi >= n_q) continue;
        const int k0 = best[k].k0, k1 = best[k].k1;
        const int i0 = k0 == HAMMING_NONE ? -1 : k0 & ((1 << HAMMING_ROW_BITS) - 1), d0 = k0 == HAMMING_NONE ? HAMMING_NONE : k0 >> HAMMING_ROW_BITS;
        const int i1 = k1 == HAMMING_NONE ? -1 : k1 & ((1 << HAMMING_ROW_BITS) - 1), d1 = k1 == HAMMING_NONE ? HAMMING_NONE : k1 >> HAMMING_ROW_BITS;
        if (splits > 1) {
            part[((size_t)pair * cap_q + qi) * splits + blockIdx.y] = make_int4(i0, i1, d0, d1);
        } else {
            int* oi = o_idx + (size_t)pair * out_stride + 2 * qi;
            int* od = o_dist + (size_t)pair * out_stride + 2 * qi;
            oi[0] = i0; oi[1] = i1; od[0] = d0; od[1] = d1;
        }
    }
}

// fold per-split partial top-2 in ascending split (= ascending train index) order; strict '<' keeps the tie rule
__global__ void __launch_bounds__(256) k_top2_merge(const int4* __restrict__ part, const int* __restrict__ nq, const int* __restrict__ nt,
                                                    int cap_q, int chunk, int splits, int* __restrict__ o_idx, int* __restrict__ o_dist)
{
    const int pair = blockIdx.y, qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= min(nq[pair], cap_q)) return;
    const int valid = min(splits, (nt[pair] + chunk - 1) / chunk);
    const int4* p = part + ((size_t)pair * cap_q + qi) * splits;
    top2 best = { 0x7fffffff, 0x7fffffff, -1, -1 };
    for (int s = 0; s < valid; ++s) {
        const int4 v = p[s];
        if (v.x >= 0) top2_update(best, v.z, v.x);
        if (v.y >= 0) top2_update(best, v.w, v.y);
    }
    const size_t o = 2 * ((size_t)pair * cap_q + qi);
    o_idx[o] = best.i0; o_idx[o + 1] = best.i1; o_dist[o] = best.d0; o_dist[o + 1] = best.d1;
}

// knn2 epilogue: int distances -> float, ratio gate of matcher.cpp:70
// bad (L2 only, may be null): device flag set when the descriptors were not integers in 0..255 -- every row then reports "no
// neighbour" instead of a distance computed from wrapped bytes
__global__ void k_knn2_finish(const int* __restrict__ idx_in, const int* __restrict__ dist_in, const int* __restrict__ nq,
                              int cap_q, double ratio, int l2, int* __restrict__ idx, float* __restrict__ dist,
                              uint8_t* __restrict__ pass, const int* __restrict__ bad)
{
    const int pair = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap_q) return;
    const size_t o = (size_t)pair * cap_q + i;
    if (i >= nq[pair] || (bad && *bad)) {
        idx[2 * o] = idx[2 * o + 1] = -1; dist[2 * o] = dist[2 * o + 1] = 0.f;
        if (pass) pass[o] = 0;
        return;
    }
    const int i0 = idx_in[2 * o], i1 = idx_in[2 * o + 1];
    float d0 = 0.f, d1 = 0.f;
    if (i0 >= 0) d0 = l2 ? sqrtf((float)dist_in[2 * o]) : (float)dist_in[2 * o];
    if (i1 >= 0) d1 = l2 ? sqrtf((float)dist_in[2 * o + 1]) : (float)dist_in[2 * o + 1];
    idx[2 * o] = i0; idx[2 * o + 1] = i1; dist[2 * o] = d0; dist[2 * o + 1] = d1;
    if (pass) pass[o] = (i0 >= 0 && i1 >= 0 && (double)d0 < ratio * (double)d1) ? 1 : 0;
}

// cross-check epilogue: keep (i, fwd[i]) iff bwd[fwd[i]] == i
__global__ void k_cross_finish(const int* __restrict__ fwd_idx, const int* __restrict__ fwd_dist,
                               const int* __restrict__ bwd_idx, const int* __restrict__ nq, int cap_q, int cap_t, int l2,
                               int* __restrict__ idx, float* __restrict__ dist, const int* __restrict__ bad)
{
    const int pair = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap_q) return;
    const size_t o = (size_t)pair * cap_q + i;
    int out = -1; float d = 0.f;
    if (i < nq[pair] && !(bad && *bad)) {
        const int f = fwd_idx[2 * o];
        if (f >= 0 && bwd_idx[2 * ((size_t)pair * cap_t + f)] == i) {
            out = f;
            d = l2 ? sqrtf((float)fwd_dist[2 * o]) : (float)fwd_dist[2 * o];
        }
    }
    idx[o] = out; dist[o] = d;
}

// L2 top-2 on integer-valued descriptors stored as u8 (dim bytes per row, dim % 4 == 0, dim <= 128).
// d2 = sum (a-b)^2 computed exactly in int32 with dp4a: |a|^2 + |b|^2 - 2 a.b.
// (CUDA-core form; the tcgen05 kernel in zs_match_l2.cu is the production path for large problems.)
template <int NW>
__global__ void __launch_bounds__(MATCH_THREADS) k_l2u8_top2(
    const uint8_t* __restrict__ q, const int* __restrict__ nq, size_t q_stride,
    const uint8_t* __restrict__ t, const int* __restrict__ nt, size_t t_stride,
    int* __restrict__ o_idx, int* __restrict__ o_dist, int out_stride)
{
    constexpr int TILE = 32;
    __shared__ uint32_t tile[TILE * NW];
    __shared__ int tnorm[TILE];
    const int pair = blockIdx.y;
    const int n_q = nq[pair], n_t = nt[pair];
    const int qi = blockIdx.x * MATCH_THREADS + threadIdx.x;
    if (blockIdx.x * MATCH_THREADS >= n_q) return;
    const uint32_t* qp = (const uint32_t*)(q + (size_t)pair * q_stride);
    const uint32_t* tp = (const uint32_t*)(t + (size_t)pair * t_stride);
    uint32_t qr[NW];
    int qn = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) { qr[w] = (qi < n_q) ? qp[(size_t)qi * NW + w] : 0u; qn = (int)__dp4a(qr[w], qr[w], (unsigned)qn); }
    top2 best = { 0x7fffffff, 0x7fffffff, -1, -1 };
    for (int base = 0; base < n_t; base += TILE) {
        const int m = min(TILE, n_t - base);
        __syncthreads();
        for (int i = threadIdx.x; i < m * NW; i += MATCH_THREADS) tile[i] = tp[(size_t)base * NW + i];
        __syncthreads();
        if (threadIdx.x < m) {
            int s = 0;
            for (int w = 0; w < NW; ++w) { const uint32_t v = tile[threadIdx.x * NW + w]; s = (int)__dp4a(v, v, (unsigned)s); }
            tnorm[threadIdx.x] = s;
        }
        __syncthreads();
        for (int j = 0; j < m; ++j) {
            int dot = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) dot = (int)__dp4a(qr[w], tile[j * NW + w], (unsigned)dot);
            top2_update(best, qn + tnorm[j] - 2 * dot, base + j);
        }
    }
    if (qi < n_q) {
        int* oi = o_idx + (size_t)pair * out_stride + 2 * qi;
        int* od = o_dist + (size_t)pair * out_stride + 2 * qi;
        oi[0] = best.i0; oi[1] = best.i1; od[0] = best.d0; od[1] = best.d1;
    }
}

// float (integer-valued, 0..255) -> u8, flagging anything else; four values per thread (dim is a multiple of 4 and rows
// are 16-byte aligned whenever the caller's stride is a multiple of 4 floats, which the launcher checks)
// norms (dim == 128 only, else null): a row is then exactly one warp of this kernel, which also leaves the row's squared
// norm for the tensor-core kernel's epilogue -- one pass over the floats instead of a conversion and a norm kernel
__global__ void __launch_bounds__(256) k_f32_to_u8(const float* __restrict__ src, const int* __restrict__ n, size_t src_stride, int cap,
                                                   int dim, uint8_t* __restrict__ dst, int* __restrict__ bad, int vec,
                                                   int* __restrict__ norms)
{
    const int pair = blockIdx.y;
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const size_t total = (size_t)min(n[pair], cap) * dim;
    if (i >= total) return;
    const float* s = src + (size_t)pair * src_stride + i;
    float v[4];
    if (vec) { const float4 f = *(const float4*)s; v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w; }
    else { v[0] = s[0]; v[1] = s[1]; v[2] = s[2]; v[3] = s[3]; }
    uint32_t out = 0; bool wrong = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = __float2int_rn(v[k]);
        wrong |= ((float)r != v[k] || r < 0 || r > 255);
        out |= (uint32_t)(r & 255) << (8 * k);
    }
    if (wrong) { atomicExch(bad, 1); atomicExch(bad + 1, 1); }     // [0]: this call (reset when the next one starts); [1]: sticky, for zs_context_async_error
    *(uint32_t*)(dst + (size_t)pair * cap * dim + i) = out;
    if (norms) {                                   // dim == 128: total is a multiple of 128 values = whole warps get here
        const int s2 = __reduce_add_sync(0xffffffffu, (int)__dp4a(out, out, 0u));
        if ((threadIdx.x & 31) == 0) norms[(size_t)pair * cap + i / 128] = s2;
    }
}

// Train-side splitting: a block sweeps `chunk` train rows.  Large maps are cut into 2048-row chunks; small problems are
// cut too when the grid would otherwise leave the machine short of warps (the kernel is latency-bound per warp).
#define HAMMING_SPLIT_CHUNK 2048
static int hamming_splits(const zs_context* ctx, int pairs, int cap_q, int cap_t)
{
    if (cap_t > 2 * HAMMING_SPLIT_CHUNK) return zs_div_up(cap_t, HAMMING_SPLIT_CHUNK);
    if (ctx->sw.hamming_splits > 0) return ctx->sw.hamming_splits;
    const long long blocks = (long long)zs_div_up(cap_q, 128) * pairs;
    int sp = (int)((148LL * 32 + blocks - 1) / blocks);
    const int max_sp = cap_t / 256 > 0 ? cap_t / 256 : 1;
    sp = sp < 1 ? 1 : sp > 8 ? 8 : sp;
    return sp > max_sp ? max_sp : sp;
}
// ---- Hamming on the tensor cores (late round 2).  For 0 / 1 vectors |a - b|^2 = |a| + |b| - 2 a.b is the Hamming distance, so the
// 256-bit descriptors, expanded to one byte per bit, go through the tcgen05 kind::i8 kernel of the L2 matcher (zs_match_l2.cu,
// two 128-byte K halves per row) with their bit counts as the "norms": same exact integers, same tie rule (smaller train row).
// Taken for problems large enough to pay for the expansion pass (ZS_HAMMING_TENSOR_MIN distances per call, default 12 M;
// ZS_HAMMING_NO_TENSOR keeps the CUDA-core kernel).
zs_status zs_l2_tensor_top2(zs_context* ctx, const uint8_t* q8, const int* nq, const uint8_t* t8, const int* nt, int pairs,
                            int cap_q, int cap_t, int dim, int* idx, int* dist, void* part, const int* qnorm, const int* tnorm);   // zs_match_l2.cu
size_t zs_l2_tensor_part_ints(int pairs, int cap_q, int cap_t);

// one warp per descriptor row: lane l expands byte l into eight 0 / 1 bytes; rows at or beyond the pair's count become zeros
__global__ void __launch_bounds__(256) k_bits_expand(const uint8_t* __restrict__ d, const int* __restrict__ n, size_t pair_stride, int cap,
                                                     size_t rows, uint8_t* __restrict__ out, int* __restrict__ norms)
{
    const size_t r = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    const int pair = (int)(r / cap), i = (int)(r - (size_t)pair * cap);
    unsigned b = 0;
    if (i < min(n[pair], cap)) b = d[(size_t)pair * pair_stride + (size_t)i * 32 + lane];
    uint2 o;
    o.x = (b & 1u) | ((b & 2u) << 7) | ((b & 4u) << 14) | ((b & 8u) << 21);
    o.y = ((b >> 4) & 1u) | ((b & 32u) << 3) | ((b & 64u) << 10) | ((b & 128u) << 17);
    *(uint2*)(out + r * 256 + 8 * lane) = o;
    const int c = __reduce_add_sync(0xffffffffu, __popc(b));
    if (lane == 0) norms[r] = c;
}

static bool hamming_use_tensor(const zs_context* ctx, int pairs, int cap_q, int cap_t)
{
    if (ctx->sw.hamming_no_tensor) return false;
    // measured crossover (profiles/r2_hamming_tensor.txt): below ~2^23.5 distances a call is launch-bound either way (32 - 38 us)
    // and the CUDA-core kernel's two launches beat the tensor path's five
    const long long min_work = ctx->sw.hamming_tensor_min > 0 ? ctx->sw.hamming_tensor_min : 12000000LL;
    return (long long)pairs * cap_q * cap_t >= min_work;
}
// ints of scratch behind `part` for the tensor path: partial top-2 + both expanded sides + their bit counts (+ alignment slack)
static size_t hamming_tensor_ints(int pairs, int cap_q, int cap_t)
{
    return zs_l2_tensor_part_ints(pairs, cap_q, cap_t) + 64 * (size_t)pairs * ((size_t)cap_q + cap_t) + (size_t)pairs * ((size_t)cap_q + cap_t) + 256;
}
// both sides expanded behind `part`: [q8 | t8 | bit counts of q | of t | partial top-2 of the tensor kernel]
struct ham_expanded { uint8_t* q8; uint8_t* t8; int* qn; int* tn; void* l2part; };
static zs_status hamming_expand(zs_context* ctx, const uint8_t* q, const int* nq, size_t qs, const uint8_t* t, const int* nt, size_t ts,
                                int pairs, int cap_q, int cap_t, void* part, ham_expanded* e)
{
    const size_t rq = (size_t)pairs * cap_q, rt = (size_t)pairs * cap_t;
    ZS_REQUIRE(rq < (1u << 30) && rt < (1u << 30), "too many descriptor rows for one call");
    e->q8 = (uint8_t*)(((uintptr_t)part + 127) & ~(uintptr_t)127); e->t8 = e->q8 + rq * 256;
    e->qn = (int*)(e->t8 + rt * 256); e->tn = e->qn + rq;
    e->l2part = (void*)(((uintptr_t)(e->tn + rt) + 15) & ~(uintptr_t)15);
    k_bits_expand<<<(unsigned)((rq + 7) / 8), 256, 0, ctx->stream>>>(q, nq, qs, cap_q, rq, e->q8, e->qn);
    ZS_LAUNCH_CHECK(ctx);
    k_bits_expand<<<(unsigned)((rt + 7) / 8), 256, 0, ctx->stream>>>(t, nt, ts, cap_t, rt, e->t8, e->tn);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}
static zs_status hamming_top2_tensor(zs_context* ctx, const uint8_t* q, const int* nq, size_t qs, const uint8_t* t, const int* nt, size_t ts,
                                     int pairs, int cap_q, int cap_t, int* idx, int* dist, void* part)
{
    ham_expanded e;
    zs_status st = hamming_expand(ctx, q, nq, qs, t, nt, ts, pairs, cap_q, cap_t, part, &e);
    if (st != ZS_OK) return st;
    return zs_l2_tensor_top2(ctx, e.q8, nq, e.t8, nt, pairs, cap_q, cap_t, 256, idx, dist, e.l2part, e.qn, e.tn);
}

static int hamming_chunk(int cap_t, int splits) { return splits > 1 ? (zs_div_up(cap_t, splits) + 127) / 128 * 128 : 0x7fffffff; }
// ints of scratch the partial top-2 of one direction need (0 when the train side is not split)
static size_t hamming_part_ints(const zs_context* ctx, int pairs, int cap_q, int cap_t)
{
    if (hamming_use_tensor(ctx, pairs, cap_q, cap_t)) return hamming_tensor_ints(pairs, cap_q, cap_t);
    const int sp = hamming_splits(ctx, pairs, cap_q, cap_t);
    return sp > 1 ? 4 * (size_t)pairs * cap_q * sp : 0;
}

static zs_status hamming_top2(zs_context* ctx, const uint8_t* q, const int* nq, size_t qs, const uint8_t* t, const int* nt,
                              size_t ts, int pairs, int cap_q, int cap_t, int* idx, int* dist, void* part)
{
    ZS_REQUIRE(cap_t < (1 << HAMMING_ROW_BITS), "more than 2^22 train rows per pair");
    if (hamming_use_tensor(ctx, pairs, cap_q, cap_t)) return hamming_top2_tensor(ctx, q, nq, qs, t, nt, ts, pairs, cap_q, cap_t, idx, dist, part);
    const int splits = hamming_splits(ctx, pairs, cap_q, cap_t);
    const int chunk = cap_t > 2 * HAMMING_SPLIT_CHUNK ? HAMMING_SPLIT_CHUNK : hamming_chunk(cap_t, splits);
    const int variant = ctx->sw.hamming_variant;
#define HAM_LAUNCH(HQ_, CSA_)                                                                                                   \
    k_hamming_top2<HQ_, CSA_><<<dim3(zs_div_up(cap_q, HAMMING_THREADS * HQ_), splits, pairs), HAMMING_THREADS, 0, ctx->stream>>>( \
        q, nq, qs, t, nt, ts, idx, dist, 2 * cap_q, chunk, splits, cap_q, (int4*)part)
    switch (variant) {
    case 2: HAM_LAUNCH(2, false); break;
    case 3: HAM_LAUNCH(2, true); break;
    case 4: HAM_LAUNCH(4, true); break;
    case 5: HAM_LAUNCH(1, false); break;
    default: HAM_LAUNCH(1, true); break;
    }
#undef HAM_LAUNCH
    ZS_LAUNCH_CHECK(ctx);
    if (splits > 1) {
        k_top2_merge<<<dim3(zs_div_up(cap_q, 256), pairs), 256, 0, ctx->stream>>>((const int4*)part, nq, nt, cap_q, chunk, splits, idx, dist);
        ZS_LAUNCH_CHECK(ctx);
    }
    return ZS_OK;
}

extern "C" zs_status zs_match_hamming_knn2(zs_context* ctx, const uint8_t* d_q, const int* d_nq, size_t q_stride,
                                           const uint8_t* d_t, const int* d_nt, size_t t_stride, int pairs, int cap_q,
                                           int cap_t, double ratio, int* d_idx, float* d_dist, uint8_t* d_pass)
{
    ZS_REQUIRE(ctx && d_q && d_nq && d_t && d_nt && d_idx && d_dist, "null argument");
    ZS_REQUIRE(pairs >= 0 && cap_q > 0 && cap_t > 0, "bad sizes");
    ZS_REQUIRE(((uintptr_t)d_q % 16) == 0 && ((uintptr_t)d_t % 16) == 0 && q_stride % 16 == 0 && t_stride % 16 == 0,
               "descriptor arrays must be 16-byte aligned");
    if (pairs == 0) return ZS_OK;
    void* s;
    zs_status st = zs_scratch(ctx, sizeof(int) * (4 * (size_t)cap_q * pairs + hamming_part_ints(ctx, pairs, cap_q, cap_t)), &s);
    if (st != ZS_OK) return st;
    int* ti = (int*)s; int* td = ti + 2 * (size_t)cap_q * pairs;
    st = hamming_top2(ctx, d_q, d_nq, q_stride, d_t, d_nt, t_stride, pairs, cap_q, cap_t, ti, td, td + 2 * (size_t)cap_q * pairs);
    if (st != ZS_OK) return st;
    k_knn2_finish<<<dim3(zs_div_up(cap_q, 256), pairs), 256, 0, ctx->stream>>>(ti, td, d_nq, cap_q, ratio, 0, d_idx, d_dist, d_pass, nullptr);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

extern "C" zs_status zs_match_hamming_cross(zs_context* ctx, const uint8_t* d_q, const int* d_nq, size_t q_stride,
                                            const uint8_t* d_t, const int* d_nt, size_t t_stride, int pairs, int cap_q,
                                            int cap_t, int* d_idx, float* d_dist)
{
    ZS_REQUIRE(ctx && d_q && d_nq && d_t && d_nt && d_idx && d_dist, "null argument");
    ZS_REQUIRE(pairs >= 0 && cap_q > 0 && cap_t > 0, "bad sizes");
    ZS_REQUIRE(((uintptr_t)d_q % 16) == 0 && ((uintptr_t)d_t % 16) == 0 && q_stride % 16 == 0 && t_stride % 16 == 0,
               "descriptor arrays must be 16-byte aligned");
    if (pairs == 0) return ZS_OK;
    void* s;
    zs_status st = zs_scratch(ctx, sizeof(int) * (4 * ((size_t)cap_q + cap_t) * pairs +
                                                  std::max(hamming_part_ints(ctx, pairs, cap_q, cap_t), hamming_part_ints(ctx, pairs, cap_t, cap_q))), &s);
    if (st != ZS_OK) return st;
    int* fi = (int*)s; int* fd = fi + 2 * (size_t)cap_q * pairs;
    int* bi = fd + 2 * (size_t)cap_q * pairs; int* bd = bi + 2 * (size_t)cap_t * pairs;
    void* part = bd + 2 * (size_t)cap_t * pairs;          // 16-byte aligned: every piece before it is a multiple of 16 bytes
    if (hamming_use_tensor(ctx, pairs, cap_q, cap_t)) {            // both directions from ONE expansion of the two sides
        ZS_REQUIRE(cap_t < (1 << HAMMING_ROW_BITS) && cap_q < (1 << HAMMING_ROW_BITS), "more than 2^22 rows per pair");
        ham_expanded e;
        if ((st = hamming_expand(ctx, d_q, d_nq, q_stride, d_t, d_nt, t_stride, pairs, cap_q, cap_t, part, &e)) != ZS_OK) return st;
        if ((st = zs_l2_tensor_top2(ctx, e.q8, d_nq, e.t8, d_nt, pairs, cap_q, cap_t, 256, fi, fd, e.l2part, e.qn, e.tn)) != ZS_OK) return st;
        if ((st = zs_l2_tensor_top2(ctx, e.t8, d_nt, e.q8, d_nq, pairs, cap_t, cap_q, 256, bi, bd, e.l2part, e.tn, e.qn)) != ZS_OK) return st;
    } else {
        st = hamming_top2(ctx, d_q, d_nq, q_stride, d_t, d_nt, t_stride, pairs, cap_q, cap_t, fi, fd, part);
        if (st != ZS_OK) return st;
        st = hamming_top2(ctx, d_t, d_nt, t_stride, d_q, d_nq, q_stride, pairs, cap_t, cap_q, bi, bd, part);
        if (st != ZS_OK) return st;
    }
    k_cross_finish<<<dim3(zs_div_up(cap_q, 256), pairs), 256, 0, ctx->stream>>>(fi, fd, bi, d_nq, cap_q, cap_t, 0, d_idx, d_dist, nullptr);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

// ---- L2 (integer-valued descriptors: cv::SIFT) ---------------------------------------------------------
// qnorm / tnorm: squared row norms [pairs*cap] when the caller has them already (the f32 -> u8 pass leaves them), else null
zs_status zs_l2_tensor_top2(zs_context* ctx, const uint8_t* q8, const int* nq, const uint8_t* t8, const int* nt, int pairs,
                            int cap_q, int cap_t, int dim, int* idx, int* dist, void* part, const int* qnorm, const int* tnorm);   // zs_match_l2.cu
size_t zs_l2_tensor_part_ints(int pairs, int cap_q, int cap_t);                                  // ints of `part` scratch

struct l2_operands { const uint8_t* q8; const uint8_t* t8; const int* qnorm; const int* tnorm; int* extra; const int* bad; };

// Float rows -> u8 rows (+ squared norms for 128-d rows) in context scratch.  Nothing here waits for the device: values that
// are not integers in 0..255 raise ctx->d_async_err[0], which the finish kernels turn into "no neighbour" rows and which
// zs_context_async_error() / the host entries report as ZS_ERR_UNSUPPORTED.
static zs_status l2_prepare(zs_context* ctx, const float* d_q, const int* d_nq, size_t q_stride, const float* d_t,
                            const int* d_nt, size_t t_stride, int pairs, int cap_q, int cap_t, int dim, size_t extra_ints,
                            l2_operands* op)
{
    ZS_REQUIRE(dim > 0 && dim % 4 == 0 && dim <= 128, "dim must be a multiple of 4, at most 128");
    const size_t qb = ((size_t)pairs * cap_q * dim + 255) / 256 * 256, tb = ((size_t)pairs * cap_t * dim + 255) / 256 * 256;
    const size_t nb = ((size_t)pairs * ((size_t)cap_q + cap_t) * sizeof(int) + 255) / 256 * 256;
    void* s;
    zs_status st = zs_scratch(ctx, qb + tb + nb + extra_ints * sizeof(int), &s);
    if (st != ZS_OK) return st;
    uint8_t* q8 = (uint8_t*)s; uint8_t* t8 = q8 + qb;
    int* norms = (int*)(t8 + tb);
    op->q8 = q8; op->t8 = t8; op->extra = (int*)((uint8_t*)norms + nb); op->bad = ctx->d_async_err;
    ZS_CUDA(cudaMemsetAsync(ctx->d_async_err, 0, sizeof(int), ctx->stream));   // the per-call flag; the sticky one stays
    const bool fuse = dim == 128;
    op->qnorm = fuse ? norms : nullptr; op->tnorm = fuse ? norms + (size_t)pairs * cap_q : nullptr;
    const int vq = ((uintptr_t)d_q % 16 == 0 && q_stride % 4 == 0) ? 1 : 0, vt = ((uintptr_t)d_t % 16 == 0 && t_stride % 4 == 0) ? 1 : 0;
    k_f32_to_u8<<<dim3(zs_div_up(zs_div_up(cap_q * dim, 4), 256), pairs), 256, 0, ctx->stream>>>(d_q, d_nq, q_stride, cap_q, dim, q8, ctx->d_async_err, vq,
                                                                                                   fuse ? norms : nullptr);
    ZS_LAUNCH_CHECK(ctx);
    k_f32_to_u8<<<dim3(zs_div_up(zs_div_up(cap_t * dim, 4), 256), pairs), 256, 0, ctx->stream>>>(d_t, d_nt, t_stride, cap_t, dim, t8, ctx->d_async_err, vt,
                                                                                                   fuse ? norms + (size_t)pairs * cap_q : nullptr);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

static zs_status l2_top2(zs_context* ctx, const uint8_t* q8, const int* nq, const uint8_t* t8, const int* nt, int pairs,
                         int cap_q, int cap_t, int dim, int* idx, int* dist, void* part, const int* qnorm, const int* tnorm)
{
    // 128-dim (SIFT) rows take the tcgen05 path; ZS_L2_NO_TENSOR=1 selects the CUDA-core dp4a kernel instead
    // (used by tools/bench_l2.py and the tests to cross-check the two implementations)
    if (dim == 128 && !ctx->sw.l2_no_tensor) return zs_l2_tensor_top2(ctx, q8, nq, t8, nt, pairs, cap_q, cap_t, dim, idx, dist, part, qnorm, tnorm);
    const dim3 grid(zs_div_up(cap_q, MATCH_THREADS), pairs);
#define L2_CASE(NW)                                                                                          \
    case NW:                                                                                                 \
        k_l2u8_top2<NW><<<grid, MATCH_THREADS, 0, ctx->stream>>>(q8, nq, (size_t)cap_q * dim, t8, nt,         \
                                                                 (size_t)cap_t * dim, idx, dist, 2 * cap_q); \
        break;
    switch (dim / 4) {
        L2_CASE(1) L2_CASE(2) L2_CASE(4) L2_CASE(8) L2_CASE(16) L2_CASE(32)
    default:
        zs_set_error("L2 descriptor dim %d not supported (4, 8, 16, 32, 64, 128)", dim);
        return ZS_ERR_UNSUPPORTED;
    }
#undef L2_CASE
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

// CUDA-core L2 top-2, exported for tests that cross-check the tensor-core kernel
zs_status zs_l2_cuda_core_top2(zs_context* ctx, const uint8_t* q8, const int* nq, const uint8_t* t8, const int* nt, int pairs,
                               int cap_q, int cap_t, int dim, int* idx, int* dist)
{
    ZS_REQUIRE(dim == 128, "dim");
    k_l2u8_top2<32><<<dim3(zs_div_up(cap_q, MATCH_THREADS), pairs), MATCH_THREADS, 0, ctx->stream>>>(
        q8, nq, (size_t)cap_q * dim, t8, nt, (size_t)cap_t * dim, idx, dist, 2 * cap_q);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

// the two matchers on prepared operands (u8 rows in device memory, dense [pairs][cap][dim])
static zs_status l2_knn2_run(zs_context* ctx, const l2_operands& op, const int* d_nq, const int* d_nt, int pairs, int cap_q, int cap_t,
                             int dim, double ratio, int* d_idx, float* d_dist, uint8_t* d_pass)
{
    int* ti = op.extra; int* td = ti + 2 * (size_t)cap_q * pairs;
    void* part = td + 2 * (size_t)cap_q * pairs;          // 16-byte aligned: every piece before it is a multiple of 16 bytes
    zs_status st = l2_top2(ctx, op.q8, d_nq, op.t8, d_nt, pairs, cap_q, cap_t, dim, ti, td, part, op.qnorm, op.tnorm);
    if (st != ZS_OK) return st;
    k_knn2_finish<<<dim3(zs_div_up(cap_q, 256), pairs), 256, 0, ctx->stream>>>(ti, td, d_nq, cap_q, ratio, 1, d_idx, d_dist, d_pass, op.bad);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

static zs_status l2_cross_run(zs_context* ctx, const l2_operands& op, const int* d_nq, const int* d_nt, int pairs, int cap_q, int cap_t,
                              int dim, int* d_idx, float* d_dist)
{
    int* fi = op.extra; int* fd = fi + 2 * (size_t)cap_q * pairs;
    int* bi = fd + 2 * (size_t)cap_q * pairs; int* bd = bi + 2 * (size_t)cap_t * pairs;
    void* part = bd + 2 * (size_t)cap_t * pairs;
    zs_status st = l2_top2(ctx, op.q8, d_nq, op.t8, d_nt, pairs, cap_q, cap_t, dim, fi, fd, part, op.qnorm, op.tnorm);
    if (st != ZS_OK) return st;
    st = l2_top2(ctx, op.t8, d_nt, op.q8, d_nq, pairs, cap_t, cap_q, dim, bi, bd, part, op.tnorm, op.qnorm);
    if (st != ZS_OK) return st;
    k_cross_finish<<<dim3(zs_div_up(cap_q, 256), pairs), 256, 0, ctx->stream>>>(fi, fd, bi, d_nq, cap_q, cap_t, 1, d_idx, d_dist, op.bad);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

static size_t l2_knn2_extra(int pairs, int cap_q, int cap_t) { return 4 * (size_t)cap_q * pairs + zs_l2_tensor_part_ints(pairs, cap_q, cap_t); }
static size_t l2_cross_extra(int pairs, int cap_q, int cap_t)
{
    return 4 * ((size_t)cap_q + cap_t) * pairs + std::max(zs_l2_tensor_part_ints(pairs, cap_q, cap_t), zs_l2_tensor_part_ints(pairs, cap_t, cap_q));
}

extern "C" zs_status zs_match_l2_knn2(zs_context* ctx, const float* d_q, const int* d_nq, size_t q_stride, const float* d_t,
                                      const int* d_nt, size_t t_stride, int pairs, int cap_q, int cap_t, int dim,
                                      double ratio, int* d_idx, float* d_dist, uint8_t* d_pass)
{
    ZS_REQUIRE(ctx && d_q && d_nq && d_t && d_nt && d_idx && d_dist, "null argument");
    ZS_REQUIRE(pairs >= 0 && cap_q > 0 && cap_t > 0, "bad sizes");
    if (pairs == 0) return ZS_OK;
    l2_operands op;
    zs_status st = l2_prepare(ctx, d_q, d_nq, q_stride, d_t, d_nt, t_stride, pairs, cap_q, cap_t, dim, l2_knn2_extra(pairs, cap_q, cap_t), &op);
    if (st != ZS_OK) return st;
    return l2_knn2_run(ctx, op, d_nq, d_nt, pairs, cap_q, cap_t, dim, ratio, d_idx, d_dist, d_pass);
}

extern "C" zs_status zs_match_l2_cross(zs_context* ctx, const float* d_q, const int* d_nq, size_t q_stride, const float* d_t,
                                       const int* d_nt, size_t t_stride, int pairs, int cap_q, int cap_t, int dim,
                                       int* d_idx, float* d_dist)
{
    ZS_REQUIRE(ctx && d_q && d_nq && d_t && d_nt && d_idx && d_dist, "null argument");
    ZS_REQUIRE(pairs >= 0 && cap_q > 0 && cap_t > 0, "bad sizes");
    if (pairs == 0) return ZS_OK;
    l2_operands op;
    zs_status st = l2_prepare(ctx, d_q, d_nq, q_stride, d_t, d_nt, t_stride, pairs, cap_q, cap_t, dim, l2_cross_extra(pairs, cap_q, cap_t), &op);
    if (st != ZS_OK) return st;
    return l2_cross_run(ctx, op, d_nq, d_nt, pairs, cap_q, cap_t, dim, d_idx, d_dist);
}

// u8 rows straight from the caller (cv::SIFT with descriptorType CV_8U, or float rows narrowed where they were produced): no
// conversion pass, a quarter of the bytes on the way to the device.  Rows are dense [pairs][cap][dim], 16-byte aligned.
static zs_status l2_u8_operands(zs_context* ctx, const uint8_t* d_q, const uint8_t* d_t, int dim, size_t extra_ints, l2_operands* op)
{
    ZS_REQUIRE(dim > 0 && dim % 4 == 0 && dim <= 128, "dim must be a multiple of 4, at most 128");
    ZS_REQUIRE(((uintptr_t)d_q % 16) == 0 && ((uintptr_t)d_t % 16) == 0, "descriptor arrays must be 16-byte aligned");
    void* s;
    zs_status st = zs_scratch(ctx, extra_ints * sizeof(int), &s);
    if (st != ZS_OK) return st;
    op->q8 = d_q; op->t8 = d_t; op->qnorm = op->tnorm = nullptr; op->extra = (int*)s; op->bad = nullptr;
    return ZS_OK;
}

extern "C" zs_status zs_match_l2_knn2_u8(zs_context* ctx, const uint8_t* d_q, const int* d_nq, const uint8_t* d_t, const int* d_nt,
                                         int pairs, int cap_q, int cap_t, int dim, double ratio, int* d_idx, float* d_dist,
                                         uint8_t* d_pass)
{
    ZS_REQUIRE(ctx && d_q && d_nq && d_t && d_nt && d_idx && d_dist, "null argument");
    ZS_REQUIRE(pairs >= 0 && cap_q > 0 && cap_t > 0, "bad sizes");
    if (pairs == 0) return ZS_OK;
    l2_operands op;
    zs_status st = l2_u8_operands(ctx, d_q, d_t, dim, l2_knn2_extra(pairs, cap_q, cap_t), &op);
    if (st != ZS_OK) return st;
    return l2_knn2_run(ctx, op, d_nq, d_nt, pairs, cap_q, cap_t, dim, ratio, d_idx, d_dist, d_pass);
}

extern "C" zs_status zs_match_l2_cross_u8(zs_context* ctx, const uint8_t* d_q, const int* d_nq, const uint8_t* d_t, const int* d_nt,
                                          int pairs, int cap_q, int cap_t, int dim, int* d_idx, float* d_dist)
{
    ZS_REQUIRE(ctx && d_q && d_nq && d_t && d_nt && d_idx && d_dist, "null argument");
    ZS_REQUIRE(pairs >= 0 && cap_q > 0 && cap_t > 0, "bad sizes");
    if (pairs == 0) return ZS_OK;
    l2_operands op;
    zs_status st = l2_u8_operands(ctx, d_q, d_t, dim, l2_cross_extra(pairs, cap_q, cap_t), &op);
    if (st != ZS_OK) return st;
    return l2_cross_run(ctx, op, d_nq, d_nt, pairs, cap_q, cap_t, dim, d_idx, d_dist);
}
