// zs_tracker.cu -- keypoint_tracker::track (zenslam_core/source/tracking/keypoint_tracker.cpp:41-105) for one stereo
// sequence, one call per stereo frame, all bookkeeping on the device:
//
//   temporal forward-backward KLT of both cameras' keypoints (:47-51)            zs_klt_launch, 2 jobs
//   detection in the cells the tracked left keypoints leave free (:53-57)        occupancy -> zs_fast_grid_detect -> zs_orb_compute
//   stereo track L -> R of the left keypoints the right camera lacks (:59-67)    index-set difference -> zs_klt_launch
//   detection in the right image behind the occupancy of ALL right keypoints     (:69-73)
//   stereo track R -> L of the right keypoints the left camera lacks (:75-83)
//
// zenslam::map<keypoint> is a std::map keyed by keypoint::index (types/map.h:24-100): here each camera's map is a
// structure of arrays kept SORTED by index (index, xy, response, descriptor), so "values in key order" is array order,
// `contains` is a binary search and `add` without overwrite is an append of indices known to be absent, followed by one
// sort per frame.  New keypoints take sequential indices from a device-side counter in detection order (left image
// first), exactly like keypoint::index_next.  Not done here (injected on the host in the reference too): landmark
// projection for the initial flow, assign_landmark_indices, filter_epipolar's RANSAC.
#include <stdlib.h>

#include "zs_common.cuh"

#define TRK_THREADS 1024

struct trk_map {                 // one camera's keypoint map
    int* idx; float* xy; float* resp; uint8_t* desc; int* n;     // [cap], [cap][2], [cap], [cap][32], [1]
};

struct zs_tracker {
    zs_context* ctx;
    zs_tracker_options opt;
    int cap, cells, gw, gh;
    uint64_t frame;
    zs_pyramid* pyr;             // 4 slots: (frame & 1) * 2 + camera
    uint8_t* dev; size_t dev_bytes;
    trk_map prev[2], cur[2];     // fields of the two cameras are contiguous ([2][cap]): one KLT launch serves both
    int* slots;                  // [2][8] job slot tables, one per frame parity (written once):
                                 //   0,1 temporal prev L/R | 2,3 temporal next L/R | 4,5 stereo L->R from/to | 6,7 stereo R->L from/to
    // CUDA graph of the per-frame launch sequence (pyramids ... sort), one per frame parity; uploads and result copies
    // stay outside.  Captured the second time a parity comes round (the first run sizes the context scratch).
    bool graph_ok; cudaGraphExec_t gexec[2]; void* g_scratch[2]; uint64_t g_launches[2]; int runs[2];
    float* t_pts; uint8_t* t_status; float* t_err; uint8_t* t_keep;   // [2][cap] KLT outputs
    uint8_t* occ;                // [cells]
    float* raw_xy; float* raw_resp; int* raw_n;   // grid candidates before ORB's border filter [cells]
    float* det_xy; float* det_resp; int* det_n; uint8_t* det_desc;   // after ORB::compute
    int* sel; int* sel_n; float* sel_pts;         // positions of the entries to stereo-track, their count, their points
    int* pred_idx; float* pred_xy; int* pred_n;   // [2][cap], [2][cap][2], [2]: initial-flow predictions for the next frame, by index
    int* marks;                  // [2][2] per camera: map sizes before the appends that start a new sorted run (see k_trk_sort)
    int* next_index;             // device copy of keypoint::index_next
    int* overflow;               // set when a map would exceed cap
};

// ordered compaction of the temporally tracked keypoints of camera blockIdx.x: cur = {prev[i] : keep[i]} with the new positions
__global__ void __launch_bounds__(TRK_THREADS) k_trk_compact(trk_map p0, trk_map p1, trk_map c0, trk_map c1, const float* __restrict__ t_pts,
                                                             const uint8_t* __restrict__ keep, int cap)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const trk_map p = blockIdx.x ? p1 : p0, c = blockIdx.x ? c1 : c0;
    const float* pts = t_pts + (size_t)blockIdx.x * cap * 2;
    const uint8_t* kp = keep + (size_t)blockIdx.x * cap;
    const int n = min(*p.n, cap), lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += TRK_THREADS) {
        const int i = base + threadIdx.x;
        const bool f = i < n && kp[i];
        const uint32_t m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_sums[warp] = __popc(m);
        __syncthreads();
        int woff = 0, total = 0;
        for (int k = 0; k < TRK_THREADS / 32; ++k) { const int s = warp_sums[k]; if (k < warp) woff += s; total += s; }
        if (f) {
            const int o = carry + woff + __popc(m & ((1u << lane) - 1));
            c.idx[o] = p.idx[i]; c.xy[2 * o] = pts[2 * i]; c.xy[2 * o + 1] = pts[2 * i + 1]; c.resp[o] = p.resp[i];
            const uint4* s = (const uint4*)(p.desc + (size_t)i * 32);
            uint4* d = (uint4*)(c.desc + (size_t)o * 32);
            d[0] = s[0]; d[1] = s[1];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *c.n = carry;
}

// initial flow of the temporal tracks (keypoint_tracker.cpp:361-373): the predicted position where the host supplied one
// for the keypoint's index (landmark projection), the keypoint's own position otherwise; the predictions are consumed
__global__ void __launch_bounds__(TRK_THREADS) k_trk_init_flow(trk_map p0, trk_map p1, int cap, const int* __restrict__ pred_idx,
                                                               const float* __restrict__ pred_xy, int* __restrict__ pred_n,
                                                               float* __restrict__ t_pts)
{
    const trk_map p = blockIdx.x ? p1 : p0;
    const int* pi = pred_idx + (size_t)blockIdx.x * cap;
    const float* px = pred_xy + (size_t)blockIdx.x * cap * 2;
    float* out = t_pts + (size_t)blockIdx.x * cap * 2;
    const int n = min(*p.n, cap), np = min(pred_n[blockIdx.x], cap);
    for (int i = threadIdx.x; i < n; i += TRK_THREADS) {
        const int key = p.idx[i];
        int lo = 0, hi = np;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (pi[mid] < key) lo = mid + 1; else hi = mid; }
        const bool hit = lo < np && pi[lo] == key;
        out[2 * i] = hit ? px[2 * lo] : p.xy[2 * i];
        out[2 * i + 1] = hit ? px[2 * lo + 1] : p.xy[2 * i + 1];
    }
    __syncthreads();
    if (threadIdx.x == 0) pred_n[blockIdx.x] = 0;
}

// occupied[int(pt.x) / cw][int(pt.y) / ch] (keypoint_detector_grid.cpp:47-64: truncating cast, then integer division)
__global__ void __launch_bounds__(TRK_THREADS) k_trk_occupancy(trk_map m, int cap, int gw, int gh, int cw, int ch, uint8_t* __restrict__ occ)
{
    for (int i = threadIdx.x; i < gw * gh; i += TRK_THREADS) occ[i] = 0;
    __syncthreads();
    const int n = min(*m.n, cap);
    for (int i = threadIdx.x; i < n; i += TRK_THREADS) {
        const int gx = (int)m.xy[2 * i] / cw, gy = (int)m.xy[2 * i + 1] / ch;
        if (gx >= 0 && gx < gw && gy >= 0 && gy < gh) occ[gy * gw + gx] = 1;
    }
}

// map.add(detected): new keypoints get index_next, index_next + 1, .. in detection order (keypoint_detector_grid.cpp:142-147)
__global__ void __launch_bounds__(TRK_THREADS) k_trk_append_detected(trk_map m, int cap, const float* __restrict__ dxy, const float* __restrict__ dresp,
                                                                     const uint8_t* __restrict__ ddesc, const int* __restrict__ dn,
                                                                     int* __restrict__ next_index, int* __restrict__ overflow,
                                                                     int* __restrict__ mark)
{
    const int n = *m.n, k = *dn, first = *next_index;
    if (mark && threadIdx.x == 0) *mark = n;             // a new sorted run starts here
    for (int i = threadIdx.x; i < k; i += TRK_THREADS) {
        const int o = n + i;
        if (o >= cap) continue;
        m.idx[o] = first + i; m.xy[2 * o] = dxy[2 * i]; m.xy[2 * o + 1] = dxy[2 * i + 1]; m.resp[o] = dresp[i];
        const uint4* s = (const uint4*)(ddesc + (size_t)i * 32);
        uint4* d = (uint4*)(m.desc + (size_t)o * 32);
        d[0] = s[0]; d[1] = s[1];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (n + k > cap) *overflow = 1;
        *m.n = min(n + k, cap);
        *next_index = first + k;
    }
}

// the values of `a` whose index `b` does not contain, in key order (map::values_unmatched): positions + points for KLT.
// `b` is searched in its first nb_sorted entries, which are sorted by index.
__global__ void __launch_bounds__(TRK_THREADS) k_trk_unmatched(trk_map a, trk_map b, int nb_sorted_is_all, const int* __restrict__ nb_sorted, int cap,
                                                               int* __restrict__ sel, int* __restrict__ sel_n, float* __restrict__ sel_pts)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int na = min(*a.n, cap), nb = nb_sorted_is_all ? min(*b.n, cap) : *nb_sorted;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < na; base += TRK_THREADS) {
        const int i = base + threadIdx.x;
        bool f = false;
        if (i < na) {
            const int key = a.idx[i];
            int lo = 0, hi = nb;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (b.idx[mid] < key) lo = mid + 1; else hi = mid; }
            f = !(lo < nb && b.idx[lo] == key);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_sums[warp] = __popc(m);
        __syncthreads();
        int woff = 0, total = 0;
        for (int k = 0; k < TRK_THREADS / 32; ++k) { const int s = warp_sums[k]; if (k < warp) woff += s; total += s; }
        if (f) {
            const int o = carry + woff + __popc(m & ((1u << lane) - 1));
            sel[o] = i; sel_pts[2 * o] = a.xy[2 * i]; sel_pts[2 * o + 1] = a.xy[2 * i + 1];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *sel_n = carry;
}

// dst.add(stereo-tracked): the kept ones of the selected entries of src, with their tracked positions, in order
__global__ void __launch_bounds__(TRK_THREADS) k_trk_append_tracked(trk_map src, trk_map dst, int cap, const int* __restrict__ sel,
                                                                    const int* __restrict__ sel_n, const float* __restrict__ t_pts,
                                                                    const uint8_t* __restrict__ keep, int* __restrict__ overflow,
                                                                    int* __restrict__ mark)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int ns = *sel_n, n0 = *dst.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry = 0; if (mark) *mark = n0; }
    __syncthreads();
    for (int base = 0; base < ns; base += TRK_THREADS) {
        const int i = base + threadIdx.x;
        const bool f = i < ns && keep[i];
        const uint32_t m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_sums[warp] = __popc(m);
        __syncthreads();
        int woff = 0, total = 0;
        for (int k = 0; k < TRK_THREADS / 32; ++k) { const int s = warp_sums[k]; if (k < warp) woff += s; total += s; }
        if (f) {
            const int o = n0 + carry + woff + __popc(m & ((1u << lane) - 1));
            if (o < cap) {
                const int j = sel[i];
                dst.idx[o] = src.idx[j]; dst.xy[2 * o] = t_pts[2 * i]; dst.xy[2 * o + 1] = t_pts[2 * i + 1]; dst.resp[o] = src.resp[j];
                const uint4* s = (const uint4*)(src.desc + (size_t)j * 32);
                uint4* d = (uint4*)(dst.desc + (size_t)o * 32);
                d[0] = s[0]; d[1] = s[1];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (n0 + carry > cap) *overflow = 1;
        *dst.n = min(n0 + carry, cap);
    }
}

// Sort a map by index into `out`.  A map is at most three runs that are each ascending already -- [0, m0) the temporal
// tracks (+ the left camera's detections, whose new indices exceed every older one), [m0, m1) the keypoints tracked over
// from the other camera (selected in key order), [m1, n) the right camera's detections -- so an element's rank is its
// offset in its own run plus a lower-bound search in the other runs (indices are unique): O(n log n) in one block
// instead of the O(n^2) counting sort this replaced (149 us per map at n = 2 600).
__device__ __forceinline__ int trk_lower_bound(const int* __restrict__ a, int lo, int hi, int key)
{
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(TRK_THREADS) k_trk_sort(trk_map m, trk_map out, int cap, const int* __restrict__ marks)
{
    const int n = min(*m.n, cap);
    const int m0 = min(max(marks[0], 0), n), m1 = min(max(marks[1], m0), n);
    const int start[4] = { 0, m0, m1, n };
    for (int i = threadIdx.x; i < n; i += TRK_THREADS) {
        const int key = m.idx[i];
        const int run = i < m0 ? 0 : i < m1 ? 1 : 2;
        int r = i - start[run];
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (q != run) r += trk_lower_bound(m.idx, start[q], start[q + 1], key) - start[q];
        out.idx[r] = key; out.xy[2 * r] = m.xy[2 * i]; out.xy[2 * r + 1] = m.xy[2 * i + 1]; out.resp[r] = m.resp[i];
        const uint4* s = (const uint4*)(m.desc + (size_t)i * 32);
        uint4* d = (uint4*)(out.desc + (size_t)r * 32);
        d[0] = s[0]; d[1] = s[1];
    }
    if (threadIdx.x == 0) *out.n = n;
}

__global__ void k_trk_set_mark(const int* __restrict__ n, int* __restrict__ mark)
{
    if (threadIdx.x == 0) *mark = *n;
}

static inline size_t trk_al(size_t v) { return (v + 255) / 256 * 256; }

extern "C" zs_status zs_tracker_create(zs_context* ctx, const zs_tracker_options* opt, zs_tracker** out)
{
    ZS_REQUIRE(ctx && opt && out, "null argument");
    ZS_REQUIRE(opt->width > 0 && opt->height > 0 && opt->cell_w > 0 && opt->cell_h > 0, "bad geometry");
    ZS_CUDA(cudaSetDevice(ctx->device));
    ZS_REQUIRE((opt->width / opt->cell_w) * (opt->height / opt->cell_h) > 0, "cells larger than the image");
    zs_tracker* t = (zs_tracker*)calloc(1, sizeof(zs_tracker));
    t->ctx = ctx; t->opt = *opt;
    t->gw = opt->width / opt->cell_w; t->gh = opt->height / opt->cell_h; t->cells = t->gw * t->gh;
    // a camera's map holds its own detections (at most one per cell and frame) plus the keypoints tracked over from the
    // other camera, and tracked keypoints may share a cell: in steady state it settles near twice the cell count; the
    // default leaves a factor of two above that, and overflow is reported, not hidden
    t->cap = opt->capacity > 0 ? opt->capacity : 4 * t->cells + 64;
    zs_status st = zs_pyramid_create(ctx, opt->width, opt->height, 4, opt->klt_win_w, opt->klt_win_h, opt->klt_max_level, &t->pyr);
    if (st != ZS_OK) { free(t); return st; }
    const size_t cap = t->cap, cells = t->cells;
    size_t off = 0;
    size_t o_map[2][5];                                  // prev, cur: idx | xy | resp | desc | n, each [2 cameras][cap]
    for (int m = 0; m < 2; ++m) {
        o_map[m][0] = off; off += trk_al(sizeof(int) * 2 * cap);
        o_map[m][1] = off; off += trk_al(sizeof(float) * 4 * cap);
        o_map[m][2] = off; off += trk_al(sizeof(float) * 2 * cap);
        o_map[m][3] = off; off += trk_al(64 * cap);
        o_map[m][4] = off; off += 256;
    }
#define TCARVE(name, bytes) const size_t o_##name = off; off += trk_al(bytes);
    TCARVE(slots, sizeof(int) * 16) TCARVE(t_pts, sizeof(float) * 4 * cap) TCARVE(t_status, 2 * cap) TCARVE(t_err, sizeof(float) * 2 * cap)
    TCARVE(t_keep, 2 * cap) TCARVE(occ, cells) TCARVE(raw_xy, sizeof(float) * 2 * cells) TCARVE(raw_resp, sizeof(float) * cells)
    TCARVE(raw_n, 256) TCARVE(det_xy, sizeof(float) * 2 * cells) TCARVE(det_resp, sizeof(float) * cells) TCARVE(det_n, 256)
    TCARVE(det_desc, 32 * cells) TCARVE(sel, sizeof(int) * cap) TCARVE(sel_n, 256) TCARVE(sel_pts, sizeof(float) * 2 * cap)
    TCARVE(next_index, 256) TCARVE(overflow, 256) TCARVE(marks, 256)
    TCARVE(pred_idx, sizeof(int) * 2 * cap) TCARVE(pred_xy, sizeof(float) * 4 * cap) TCARVE(pred_n, 256)
#undef TCARVE
    cudaError_t e = cudaMalloc((void**)&t->dev, off);
    if (e != cudaSuccess) { zs_pyramid_destroy(t->pyr); free(t); return zs_cuda_fail(e, "cudaMalloc(tracker)", __FILE__, __LINE__); }
    t->dev_bytes = off;
    e = cudaMemsetAsync(t->dev, 0, off, ctx->stream);
    if (e != cudaSuccess) { cudaFree(t->dev); zs_pyramid_destroy(t->pyr); free(t); return zs_cuda_fail(e, "cudaMemset(tracker)", __FILE__, __LINE__); }
    for (int m = 0; m < 2; ++m)
        for (int cam = 0; cam < 2; ++cam) {
            trk_map& mp = m == 0 ? t->prev[cam] : t->cur[cam];
            mp.idx = (int*)(t->dev + o_map[m][0]) + (size_t)cam * cap; mp.xy = (float*)(t->dev + o_map[m][1]) + (size_t)cam * cap * 2;
            mp.resp = (float*)(t->dev + o_map[m][2]) + (size_t)cam * cap; mp.desc = t->dev + o_map[m][3] + (size_t)cam * cap * 32;
            mp.n = (int*)(t->dev + o_map[m][4]) + cam;
        }
#define TBIND(name, type) t->name = (type*)(t->dev + o_##name);
    TBIND(slots, int) TBIND(t_pts, float) TBIND(t_status, uint8_t) TBIND(t_err, float) TBIND(t_keep, uint8_t) TBIND(occ, uint8_t)
    TBIND(raw_xy, float) TBIND(raw_resp, float) TBIND(raw_n, int) TBIND(det_xy, float) TBIND(det_resp, float) TBIND(det_n, int)
    TBIND(det_desc, uint8_t) TBIND(sel, int) TBIND(sel_n, int) TBIND(sel_pts, float) TBIND(next_index, int) TBIND(overflow, int) TBIND(marks, int) TBIND(pred_idx, int) TBIND(pred_xy, float) TBIND(pred_n, int)
#undef TBIND
    {
        int hs[16];
        for (int par = 0; par < 2; ++par) {
            const int cs = par * 2, ps = 2 - cs;
            int* q = hs + 8 * par;
            q[0] = ps; q[1] = ps + 1; q[2] = cs; q[3] = cs + 1; q[4] = cs; q[5] = cs + 1; q[6] = cs + 1; q[7] = cs;
        }
        e = cudaMemcpyAsync(t->slots, hs, sizeof(hs), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { cudaFree(t->dev); zs_pyramid_destroy(t->pyr); free(t); return zs_cuda_fail(e, "tracker slot tables", __FILE__, __LINE__); }
    }
    t->graph_ok = !getenv("ZS_FE_NO_GRAPH");
    if (opt->first_index > 0) {
        const int fi = opt->first_index;
        e = cudaMemcpyAsync(t->next_index, &fi, sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { cudaFree(t->dev); zs_pyramid_destroy(t->pyr); free(t); return zs_cuda_fail(e, "tracker init", __FILE__, __LINE__); }
    }
    *out = t;
    return ZS_OK;
}

extern "C" void zs_tracker_destroy(zs_tracker* t)
{
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    for (int par = 0; par < 2; ++par) if (t->gexec[par]) cudaGraphExecDestroy(t->gexec[par]);
    if (t->pyr) zs_pyramid_destroy(t->pyr);
    if (t->dev) cudaFree(t->dev);
    free(t);
}

extern "C" int zs_tracker_capacity(const zs_tracker* t) { return t ? t->cap : 0; }

extern "C" zs_status zs_tracker_set_predictions(zs_tracker* t, int camera, const int* index, const float* xy, int n)
{
    ZS_REQUIRE(t && (camera == 0 || camera == 1), "bad argument");
    ZS_REQUIRE(n >= 0 && n <= t->cap && (n == 0 || (index && xy)), "bad prediction list");
    zs_context* ctx = t->ctx;
    ZS_CUDA(cudaSetDevice(ctx->device));
    for (int i = 1; i < n; ++i) ZS_REQUIRE(index[i - 1] < index[i], "prediction indices must be strictly ascending");
    if (n > 0) {
        ZS_CUDA(cudaMemcpyAsync(t->pred_idx + (size_t)camera * t->cap, index, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
        ZS_CUDA(cudaMemcpyAsync(t->pred_xy + (size_t)camera * t->cap * 2, xy, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
    }
    ZS_CUDA(cudaMemcpyAsync(t->pred_n + camera, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));     // the host arrays may go away after the call
    return ZS_OK;
}

// detection of one camera behind the occupancy of its current map (keypoint_tracker.cpp:53-57 / 69-73)
static zs_status trk_detect(zs_tracker* t, int cam, int slot, int* d_mark)
{
    zs_context* ctx = t->ctx;
    const zs_tracker_options& o = t->opt;
    k_trk_occupancy<<<1, TRK_THREADS, 0, ctx->stream>>>(t->cur[cam], t->cap, t->gw, t->gh, o.cell_w, o.cell_h, t->occ);
    ZS_LAUNCH_CHECK(ctx);
    zs_status st = zs_fast_grid_detect(ctx, t->pyr, slot, 1, o.cell_w, o.cell_h, o.fast_threshold, t->occ, t->raw_xy, t->raw_resp, t->raw_n,
                                       t->cells);
    if (st != ZS_OK) return st;
    if ((st = zs_orb_compute(ctx, t->pyr, slot, 1, t->raw_xy, t->raw_resp, nullptr, t->raw_n, t->cells, t->det_xy, t->det_resp, nullptr,
                             t->det_n, t->det_desc)) != ZS_OK) return st;
    k_trk_append_detected<<<1, TRK_THREADS, 0, ctx->stream>>>(t->cur[cam], t->cap, t->det_xy, t->det_resp, t->det_desc, t->det_n, t->next_index,
                                                              t->overflow, d_mark);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

// stereo track of the keypoints of camera `from` that camera `to` lacks (keypoint_tracker.cpp:59-67 / 75-83); `to`'s map is
// sorted by index at that point
static zs_status trk_stereo(zs_tracker* t, int from, int to, const int* d_slot_from, const int* d_slot_to, const zs_lk_params* prm,
                            int* d_mark)
{
    zs_context* ctx = t->ctx;
    k_trk_unmatched<<<1, TRK_THREADS, 0, ctx->stream>>>(t->cur[from], t->cur[to], 1, nullptr, t->cap, t->sel, t->sel_n, t->sel_pts);
    ZS_LAUNCH_CHECK(ctx);
    zs_status st = zs_klt_launch(ctx, t->pyr, d_slot_from, d_slot_to, t->sel_pts, t->t_pts, t->sel_n, nullptr, 1, t->cap, prm, t->t_status,
                                 t->t_err, 1, t->opt.klt_threshold, t->t_keep);
    if (st != ZS_OK) return st;
    k_trk_append_tracked<<<1, TRK_THREADS, 0, ctx->stream>>>(t->cur[from], t->cur[to], t->cap, t->sel, t->sel_n, t->t_pts, t->t_keep, t->overflow, d_mark);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

// everything between the uploads and the result copies for a frame of parity `par`
static zs_status trk_frame_body(zs_tracker* t, int par)
{
    zs_context* ctx = t->ctx;
    const zs_tracker_options& o = t->opt;
    const int cap = t->cap, cs = par * 2;
    const int* sl = t->slots + 8 * par;
    zs_status st;
    if ((st = zs_pyramid_build(ctx, t->pyr, cs, 2)) != ZS_OK) return st;
    zs_lk_params prm;
    prm.win_w = o.klt_win_w; prm.win_h = o.klt_win_h; prm.max_level = o.klt_max_level; prm.max_iters = 99; prm.epsilon = 0.001;
    prm.flags = ZS_LK_GET_MIN_EIGENVALS; prm.min_eig_threshold = 1e-4;
    // 1. temporal tracks of both cameras (:47-51, the overload with initial flow :343-434), one launch with two jobs;
    //    frame 0 has no previous keypoints (n = 0).  OPTFLOW_USE_INITIAL_FLOW with the keypoint's own position is what the
    //    plain call starts from, so the flag is always on and the graph is the same with and without predictions.
    k_trk_init_flow<<<2, TRK_THREADS, 0, ctx->stream>>>(t->prev[0], t->prev[1], cap, t->pred_idx, t->pred_xy, t->pred_n, t->t_pts);
    ZS_LAUNCH_CHECK(ctx);
    zs_lk_params prm_init = prm;
    prm_init.flags |= ZS_LK_USE_INITIAL_FLOW;
    if ((st = zs_klt_launch(ctx, t->pyr, sl, sl + 2, t->prev[0].xy, t->t_pts, t->prev[0].n, nullptr, 2, cap, &prm_init, t->t_status, t->t_err, 1,
                            o.klt_threshold, t->t_keep)) != ZS_OK) return st;
    k_trk_compact<<<2, TRK_THREADS, 0, ctx->stream>>>(t->prev[0], t->prev[1], t->cur[0], t->cur[1], t->t_pts, t->t_keep, cap);
    ZS_LAUNCH_CHECK(ctx);
    // 2. new left keypoints in the free cells (:53-57)
    if ((st = trk_detect(t, 0, cs, nullptr)) != ZS_OK) return st;          // same run: new indices exceed every tracked one
    // 3. left keypoints the right camera lacks: L -> R (:59-67); the right map holds only its temporal tracks, sorted
    if ((st = trk_stereo(t, 0, 1, sl + 4, sl + 5, &prm, t->marks + 2)) != ZS_OK) return st;
    // 4. new right keypoints behind the occupancy of everything the right map now holds (:69-73)
    if ((st = trk_detect(t, 1, cs + 1, t->marks + 3)) != ZS_OK) return st;
    // 5. right keypoints the left camera lacks: R -> L (:75-83); the left map (tracks + detections) is still sorted
    if ((st = trk_stereo(t, 1, 0, sl + 6, sl + 7, &prm, t->marks + 0)) != ZS_OK) return st;
    // 6. key order for the output and for the next frame's searches; the sorted maps become `prev`
    // left map: [tracks + detections | from the right camera | -] ; right map: [tracks | from the left camera | detections]
    k_trk_set_mark<<<1, 32, 0, ctx->stream>>>(t->cur[0].n, t->marks + 1);     // the left map has no third run
    ZS_LAUNCH_CHECK(ctx);
    for (int cam = 0; cam < 2; ++cam) {
        k_trk_sort<<<1, TRK_THREADS, 0, ctx->stream>>>(t->cur[cam], t->prev[cam], cap, t->marks + 2 * cam);
        ZS_LAUNCH_CHECK(ctx);
    }
    return ZS_OK;
}

// eager the first time a parity is seen, then captured once and replayed (same scheme as zs_frontend_run)
static zs_status trk_frame(zs_tracker* t, int par)
{
    zs_context* ctx = t->ctx;
    if (!t->graph_ok || t->runs[par]++ == 0) return trk_frame_body(t, par);
    if (t->gexec[par] && t->g_scratch[par] != ctx->scratch) { cudaGraphExecDestroy(t->gexec[par]); t->gexec[par] = nullptr; }
    if (!t->gexec[par]) {
        void* scratch_before = ctx->scratch;
        const uint64_t l0 = ctx->launches;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            t->graph_ok = false;
            return trk_frame_body(t, par);
        }
        const zs_status st = trk_frame_body(t, par);
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
        const bool good = st == ZS_OK && e == cudaSuccess && g && ctx->scratch == scratch_before &&
                          cudaGraphInstantiate(&t->gexec[par], g, 0) == cudaSuccess;
        if (g) cudaGraphDestroy(g);
        if (!good) {
            cudaGetLastError();
            t->gexec[par] = nullptr; t->graph_ok = false;
            ctx->launches = l0;
            return trk_frame_body(t, par);           // nothing ran during the capture
        }
        t->g_launches[par] = ctx->launches - l0;
        ctx->launches = l0;
        t->g_scratch[par] = ctx->scratch;
    }
    ZS_CUDA(cudaGraphLaunch(t->gexec[par], ctx->stream));
    ctx->launches += t->g_launches[par];
    return ZS_OK;
}

extern "C" zs_status zs_tracker_track_host(zs_tracker* t, const uint8_t* left, const uint8_t* right, size_t pitch,
                                           const zs_tracker_results* res)
{
    ZS_REQUIRE(t && left && right && res, "null argument");
    ZS_REQUIRE(res->cap >= t->cap, "results.cap must be at least zs_tracker_capacity()");
    zs_context* ctx = t->ctx;
    ZS_CUDA(cudaSetDevice(ctx->device));
    const zs_tracker_options& o = t->opt;
    const int cap = t->cap;
    const int par = (int)(t->frame & 1), cs = par * 2;           // pyramid slots of the current stereo frame: cs, cs + 1
    zs_status st;
    if ((st = zs_pyramid_upload(ctx, t->pyr, left, pitch, pitch * o.height, cs, 1, 1)) != ZS_OK) return st;
    if ((st = zs_pyramid_upload(ctx, t->pyr, right, pitch, pitch * o.height, cs + 1, 1, 1)) != ZS_OK) return st;
    if ((st = trk_frame(t, par)) != ZS_OK) return st;
    // results: the four counters first, then only the live part of every array
    int h_n[2] = { 0, 0 }, h_over = 0, h_next = 0;
    ZS_CUDA(cudaMemcpyAsync(&h_n[0], t->prev[0].n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(&h_n[1], t->prev[1].n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(&h_over, t->overflow, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(&h_next, t->next_index, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int cam = 0; cam < 2; ++cam) {
        const size_t c = (size_t)(h_n[cam] < cap ? h_n[cam] : cap);
        if (c == 0) continue;
        if (res->index[cam]) ZS_CUDA(cudaMemcpyAsync(res->index[cam], t->prev[cam].idx, sizeof(int) * c, cudaMemcpyDeviceToHost, ctx->stream));
        if (res->xy[cam]) ZS_CUDA(cudaMemcpyAsync(res->xy[cam], t->prev[cam].xy, sizeof(float) * 2 * c, cudaMemcpyDeviceToHost, ctx->stream));
        if (res->response[cam]) ZS_CUDA(cudaMemcpyAsync(res->response[cam], t->prev[cam].resp, sizeof(float) * c, cudaMemcpyDeviceToHost, ctx->stream));
        if (res->desc[cam]) ZS_CUDA(cudaMemcpyAsync(res->desc[cam], t->prev[cam].desc, 32 * c, cudaMemcpyDeviceToHost, ctx->stream));
    }
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (res->n) { res->n[0] = h_n[0]; res->n[1] = h_n[1]; }
    if (res->next_index) *res->next_index = h_next;
    t->frame++;
    if (h_over) { zs_set_error("tracker capacity %d exceeded", cap); return ZS_ERR_CAPACITY; }
    return ZS_OK;
}
