// zs_tracker.cu -- keypoint_tracker::track (zenslam_core/source/tracking/keypoint_tracker.cpp:41-105) for S independent
// stereo sequences in lock-step, one call per stereo frame (of every sequence), all bookkeeping on the device:
//
//   temporal forward-backward KLT of both cameras' keypoints (:47-51, :343-434)  zs_klt_launch, 2S jobs
//   detection in the cells the tracked left keypoints leave free (:53-57)        occupancy -> zs_fast_grid_detect -> zs_orb_compute
//   stereo track L -> R of the left keypoints the right camera lacks (:59-67)    index-set difference -> zs_klt_launch, S jobs
//   detection in the right image behind the occupancy of ALL right keypoints     (:69-73)
//   stereo track R -> L of the right keypoints the left camera lacks (:75-83)
//
// zenslam::map<keypoint> is a std::map keyed by keypoint::index (types/map.h:24-100): here each camera's map is a
// structure of arrays kept SORTED by index (index, xy, response, descriptor), so "values in key order" is array order,
// `contains` is a binary search and `add` without overwrite is an append of indices known to be absent, followed by one
// merge of the sorted runs per frame.  New keypoints take sequential indices from a per-sequence device-side counter in
// detection order (left image first), exactly like keypoint::index_next.  Sequences never interact: a block (or a KLT job,
// or a detector image) belongs to one sequence.  Not done here (host-side in the reference too): the landmark projection
// behind the initial flow (its result comes in through zs_tracker_set_predictions) and filter_epipolar's RANSAC.
//
// assign_landmark_indices (:55,71,199-291) runs between each detection and its map add when the tracker was created with a
// landmark store (landmark_capacity > 0): radius pre-filter around the camera centre (zs_landmarks.cu), cross-checked Hamming
// match of the detected keypoints against the candidate landmarks (zs_match.cu), distance gate, and a map add that skips the
// landmark indices the map already holds.  Landmark indices are arbitrary keys, so in that mode each map is re-sorted right
// after its detection add (rank by binary search over its at most four ascending runs) and every search sees a sorted map.
#include <stdlib.h>

#include <unordered_set>
#include <vector>

#include "zs_common.cuh"

#define TRK_THREADS 1024

struct trk_map {                 // one camera's keypoint map of one sequence
    int* idx; float* xy; float* resp; uint8_t* desc; int* n;     // [cap], [cap][2], [cap], [cap][32], [1]
};

struct trk_maps {                // one generation (previous / current frame) of every map: fields are [S][2][cap]
    int* idx; float* xy; float* resp; uint8_t* desc; int* n; int cap;
    __host__ __device__ trk_map at(int seq, int cam) const
    {
        const size_t r = (size_t)seq * 2 + cam;
        trk_map m;
        m.idx = idx + r * cap; m.xy = xy + r * cap * 2; m.resp = resp + r * cap; m.desc = desc + r * cap * 32; m.n = n + r;
        return m;
    }
};

struct zs_tracker {
    zs_context* ctx;
    zs_tracker_options opt;
    int S, cap, cells, gw, gh;
    uint64_t frame;
    zs_pyramid* pyr;             // 4S slots: slot(parity, camera, sequence) = (parity * 2 + camera) * S + sequence
    uint8_t* dev; size_t dev_bytes;
    trk_maps prev, cur;
    int* slots;                  // [2 parities][6 S] job slot tables (written once):
                                 //   [0, 2S) temporal prev, job = 2 seq + cam | [2S, 4S) temporal next |
                                 //   [4S, 5S) slots of the left images | [5S, 6S) slots of the right images
    float* t_pts; uint8_t* t_status; float* t_err; uint8_t* t_keep;   // [2S][cap] KLT outputs
    uint8_t* occ;                // [S][cells]
    float* raw_xy; float* raw_resp; int* raw_n;   // [S][cells] grid candidates before ORB's border filter
    float* det_xy; float* det_resp; int* det_n; uint8_t* det_desc;   // after ORB::compute
    int* sel; int* sel_n; float* sel_pts;         // [S][cap] positions of the entries to stereo-track, their count, their points
    int* pred_idx; float* pred_xy; int* pred_n;   // [S][2][cap]: initial-flow predictions for the next frame, by index
    int* marks;                  // [S][2][2] map sizes before the appends that start a new sorted run (see k_trk_sort)
    int* next_index;             // [S] device copies of keypoint::index_next
    int* overflow;               // set when a map would exceed cap
    // landmark association: per-sequence landmark store in insertion order (system.points3d), [S][lm_cap] rows; 0 = off
    int lm_cap;
    double* lm_xyz; int* lm_index; uint8_t* lm_desc; int* lm_n; double* lm_center; int* lm_nt;
    int* lm_match; float* lm_dist;   // [S][cells]: cross-checked landmark row of every detected keypoint (-1: none), distance
    int* lm_runs;                    // [S][2][5]: boundaries of the ascending runs of a map before its mid-frame sort
    trk_maps tmp;                    // scratch generation the mid-frame sorts write into
    int lm_bucket, g_bucket[2];      // train-side row count the matcher is launched with (power of two >= the fullest store)
    std::vector<std::unordered_set<int>>* lm_seen;   // indices each store holds: the host side of map::operator+=
    std::vector<int>* lm_host_n;
    // CUDA graph of the per-frame launch sequence (pyramids ... sort), one per frame parity; uploads and result copies
    // stay outside.  Captured the second time a parity comes round (the first run sizes the context scratch).
    bool graph_ok; cudaGraphExec_t gexec[2]; void* g_scratch[2]; uint64_t g_launches[2]; int runs[2];
    // pipelined host path (zs_tracker_submit_host / zs_tracker_wait): two steps in flight on three streams --
    // copy-in: frames of step k+1 -> stage[(k+1)&1]; compute (the context's stream): unpack, step k, snapshot of the maps
    // into outbox[k&1]; copy-out: outbox[(k-1)&1] -> the caller's arrays.  Frames must be staged: step k reads the pyramid
    // slots of step k-1's parity, which is where step k+1's frames will live.
    cudaStream_t s_in, s_out;
    uint8_t* stage[2]; uint8_t* outbox[2]; size_t outbox_bytes;
    cudaEvent_t ev_in[2], ev_unpacked[2], ev_run[2], ev_out[2];
    bool busy[2]; uint64_t submitted, waited;
    int* h_tail[2];              // pinned: n [2S] | next_index [S] | overflow, per in-flight step
    zs_tracker_results pending[2];
};

__device__ __forceinline__ int trk_lower_bound(const int* __restrict__ a, int lo, int hi, int key)
{
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

__device__ __forceinline__ void trk_copy_desc(uint8_t* dst, const uint8_t* src)
{
    const uint4* s = (const uint4*)src;
    uint4* d = (uint4*)dst;
    d[0] = s[0]; d[1] = s[1];
}

// block-wide ordered compaction step: returns this thread's output position (valid when f) and advances `carry`
__device__ __forceinline__ int trk_scan_step(bool f, int* warp_sums, int* carry)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t m = __ballot_sync(0xffffffffu, f);
    if (lane == 0) warp_sums[warp] = __popc(m);
    __syncthreads();
    int woff = 0, total = 0;
    for (int k = 0; k < TRK_THREADS / 32; ++k) { const int s = warp_sums[k]; if (k < warp) woff += s; total += s; }
    const int pos = *carry + woff + __popc(m & ((1u << lane) - 1));
    __syncthreads();
    if (threadIdx.x == 0) *carry += total;
    __syncthreads();
    return pos;
}

// initial flow of the temporal tracks (keypoint_tracker.cpp:361-373): the predicted position where the host supplied one
// for the keypoint's index (landmark projection), the keypoint's own position otherwise; the predictions are consumed.
// grid: 2S blocks, block = job = 2 seq + cam
__global__ void __launch_bounds__(TRK_THREADS) k_trk_init_flow(trk_maps prev, const int* __restrict__ pred_idx, const float* __restrict__ pred_xy,
                                                               int* __restrict__ pred_n, float* __restrict__ t_pts)
{
    const int cap = prev.cap;
    const trk_map p = prev.at(blockIdx.x >> 1, blockIdx.x & 1);
    const int* pi = pred_idx + (size_t)blockIdx.x * cap;
    const float* px = pred_xy + (size_t)blockIdx.x * cap * 2;
    float* out = t_pts + (size_t)blockIdx.x * cap * 2;
    const int n = min(*p.n, cap), np = min(pred_n[blockIdx.x], cap);
    for (int i = threadIdx.x; i < n; i += TRK_THREADS) {
        const int key = p.idx[i];
        const int lo = trk_lower_bound(pi, 0, np, key);
        const bool hit = lo < np && pi[lo] == key;
        out[2 * i] = hit ? px[2 * lo] : p.xy[2 * i];
        out[2 * i + 1] = hit ? px[2 * lo + 1] : p.xy[2 * i + 1];
    }
    __syncthreads();
    if (threadIdx.x == 0) pred_n[blockIdx.x] = 0;
}

// ordered compaction of the temporally tracked keypoints: cur = {prev[i] : keep[i]} with the new positions.  grid: 2S
__global__ void __launch_bounds__(TRK_THREADS) k_trk_compact(trk_maps prev, trk_maps cur, const float* __restrict__ t_pts,
                                                             const uint8_t* __restrict__ keep)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int cap = prev.cap;
    const trk_map p = prev.at(blockIdx.x >> 1, blockIdx.x & 1), c = cur.at(blockIdx.x >> 1, blockIdx.x & 1);
    const float* pts = t_pts + (size_t)blockIdx.x * cap * 2;
    const uint8_t* kp = keep + (size_t)blockIdx.x * cap;
    const int n = min(*p.n, cap);
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += TRK_THREADS) {
        const int i = base + threadIdx.x;
        const bool f = i < n && kp[i];
        const int o = trk_scan_step(f, warp_sums, &carry);
        if (f) {
            c.idx[o] = p.idx[i]; c.xy[2 * o] = pts[2 * i]; c.xy[2 * o + 1] = pts[2 * i + 1]; c.resp[o] = p.resp[i];
            trk_copy_desc(c.desc + (size_t)o * 32, p.desc + (size_t)i * 32);
        }
    }
    if (threadIdx.x == 0) *c.n = carry;
}

// occupied[int(pt.x) / cw][int(pt.y) / ch] (keypoint_detector_grid.cpp:47-64: truncating cast, then integer division).  grid: S
__global__ void __launch_bounds__(TRK_THREADS) k_trk_occupancy(trk_maps cur, int cam, int gw, int gh, int cw, int ch, uint8_t* __restrict__ occ)
{
    const trk_map m = cur.at(blockIdx.x, cam);
    uint8_t* o = occ + (size_t)blockIdx.x * gw * gh;
    for (int i = threadIdx.x; i < gw * gh; i += TRK_THREADS) o[i] = 0;
    __syncthreads();
    const int n = min(*m.n, cur.cap);
    for (int i = threadIdx.x; i < n; i += TRK_THREADS) {
        const int gx = (int)m.xy[2 * i] / cw, gy = (int)m.xy[2 * i + 1] / ch;
        if (gx >= 0 && gx < gw && gy >= 0 && gy < gh) o[gy * gw + gx] = 1;
    }
}

// map.add(detected): new keypoints get index_next, index_next + 1, .. in detection order (keypoint_detector_grid.cpp:142-147).
// grid: S; det arrays are [S][cells]; mark (optional, stride 4 per sequence) = map size before the append
__global__ void __launch_bounds__(TRK_THREADS) k_trk_append_detected(trk_maps cur, int cam, int cells, const float* __restrict__ dxy,
                                                                     const float* __restrict__ dresp, const uint8_t* __restrict__ ddesc,
                                                                     const int* __restrict__ dn, int* __restrict__ next_index,
                                                                     int* __restrict__ overflow, int* __restrict__ mark)
{
    const int seq = blockIdx.x, cap = cur.cap;
    const trk_map m = cur.at(seq, cam);
    const float* xy = dxy + (size_t)seq * cells * 2; const float* rs = dresp + (size_t)seq * cells;
    const uint8_t* ds = ddesc + (size_t)seq * cells * 32;
    const int n = *m.n, k = dn[seq], first = next_index[seq];
    if (mark && threadIdx.x == 0) mark[4 * seq] = n;             // a new sorted run starts here
    for (int i = threadIdx.x; i < k; i += TRK_THREADS) {
        const int o = n + i;
        if (o >= cap) continue;
        m.idx[o] = first + i; m.xy[2 * o] = xy[2 * i]; m.xy[2 * o + 1] = xy[2 * i + 1]; m.resp[o] = rs[i];
        trk_copy_desc(m.desc + (size_t)o * 32, ds + (size_t)i * 32);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (n + k > cap) *overflow = 1;
        *m.n = min(n + k, cap);
        next_index[seq] = first + k;
    }
}

// the values of map `from` whose index map `to` does not contain, in key order (map::values_unmatched): positions + points
// for KLT.  `to` is sorted by index at that point.  grid: S
__global__ void __launch_bounds__(TRK_THREADS) k_trk_unmatched(trk_maps cur, int from, int to, int* __restrict__ sel, int* __restrict__ sel_n,
                                                               float* __restrict__ sel_pts)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int seq = blockIdx.x, cap = cur.cap;
    const trk_map a = cur.at(seq, from), b = cur.at(seq, to);
    int* s_out = sel + (size_t)seq * cap; float* p_out = sel_pts + (size_t)seq * cap * 2;
    const int na = min(*a.n, cap), nb = min(*b.n, cap);
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < na; base += TRK_THREADS) {
        const int i = base + threadIdx.x;
        bool f = false;
        if (i < na) {
            const int key = a.idx[i];
            const int lo = trk_lower_bound(b.idx, 0, nb, key);
            f = !(lo < nb && b.idx[lo] == key);
        }
        const int o = trk_scan_step(f, warp_sums, &carry);
        if (f) { s_out[o] = i; p_out[2 * o] = a.xy[2 * i]; p_out[2 * o + 1] = a.xy[2 * i + 1]; }
    }
    if (threadIdx.x == 0) sel_n[seq] = carry;
}

// to.add(stereo-tracked): the kept ones of the selected entries of `from`, with their tracked positions, in order.  grid: S
__global__ void __launch_bounds__(TRK_THREADS) k_trk_append_tracked(trk_maps cur, int from, int to, const int* __restrict__ sel,
                                                                    const int* __restrict__ sel_n, const float* __restrict__ t_pts,
                                                                    const uint8_t* __restrict__ keep, int* __restrict__ overflow,
                                                                    int* __restrict__ mark)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int seq = blockIdx.x, cap = cur.cap;
    const trk_map src = cur.at(seq, from), dst = cur.at(seq, to);
    const int* s_in = sel + (size_t)seq * cap; const float* pts = t_pts + (size_t)seq * cap * 2; const uint8_t* kp = keep + (size_t)seq * cap;
    const int ns = sel_n[seq], n0 = *dst.n;
    if (threadIdx.x == 0) { carry = 0; if (mark) mark[4 * seq] = n0; }
    __syncthreads();
    for (int base = 0; base < ns; base += TRK_THREADS) {
        const int i = base + threadIdx.x;
        const bool f = i < ns && kp[i];
        const int o = n0 + trk_scan_step(f, warp_sums, &carry);
        if (f && o < cap) {
            const int j = s_in[i];
            dst.idx[o] = src.idx[j]; dst.xy[2 * o] = pts[2 * i]; dst.xy[2 * o + 1] = pts[2 * i + 1]; dst.resp[o] = src.resp[j];
            trk_copy_desc(dst.desc + (size_t)o * 32, src.desc + (size_t)j * 32);
        }
    }
    if (threadIdx.x == 0) {
        if (n0 + carry > cap) *overflow = 1;
        *dst.n = min(n0 + carry, cap);
    }
}

// Sort every map by index into the other generation.  A map is at most three runs that are each ascending already --
// [0, m0) the temporal tracks (+ the left camera's detections, whose new indices exceed every older one), [m0, m1) the
// keypoints tracked over from the other camera (selected in key order), [m1, n) the right camera's detections -- so an
// element's rank is its offset in its own run plus a lower-bound search in the other runs (indices are unique):
// O(n log n) in one block instead of the O(n^2) counting sort this replaced (149 us per map at n = 2 600).
// grid: 2S; marks: [S][2][2], the left camera's second mark is unused (its map has no third run)
__global__ void __launch_bounds__(TRK_THREADS) k_trk_sort(trk_maps cur, trk_maps out, const int* __restrict__ marks)
{
    const int seq = blockIdx.x >> 1, cam = blockIdx.x & 1, cap = cur.cap;
    const trk_map m = cur.at(seq, cam), o = out.at(seq, cam);
    const int n = min(*m.n, cap);
    const int* mk = marks + 4 * seq + 2 * cam;
    const int m0 = min(max(mk[0], 0), n), m1 = cam == 0 ? n : min(max(mk[1], m0), n);
    const int start[4] = { 0, m0, m1, n };
    for (int i = threadIdx.x; i < n; i += TRK_THREADS) {
        const int key = m.idx[i];
        const int run = i < m0 ? 0 : i < m1 ? 1 : 2;
        int r = i - start[run];
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (q != run) r += trk_lower_bound(m.idx, start[q], start[q + 1], key) - start[q];
        o.idx[r] = key; o.xy[2 * r] = m.xy[2 * i]; o.xy[2 * r + 1] = m.xy[2 * i + 1]; o.resp[r] = m.resp[i];
        trk_copy_desc(o.desc + (size_t)r * 32, m.desc + (size_t)i * 32);
    }
    if (threadIdx.x == 0) *o.n = n;
}

// ---- landmark association ------------------------------------------------------------------------------------
// map.add(detected) after assign_landmark_indices (keypoint_tracker.cpp:55-57, 71-73, 283-287).  The detector stamps every
// new keypoint with index_next++ (keypoint_detector_grid.cpp:142-147); a keypoint whose cross-checked landmark match has
// distance <= landmark_match_distance then takes the landmark's index, and add() skips it when the map already holds
// that index.  Appends two ascending runs: the keypoints that keep their fresh index, then the accepted landmark-indexed
// ones ranked by index; records the run boundaries for k_trk_sort_runs.  grid: S
__global__ void __launch_bounds__(TRK_THREADS) k_trk_append_detected_lm(trk_maps cur, int cam, int cells, const float* __restrict__ dxy,
                                                                        const float* __restrict__ dresp, const uint8_t* __restrict__ ddesc,
                                                                        const int* __restrict__ dn, const int* __restrict__ lm_match,
                                                                        const float* __restrict__ lm_dist, const int* __restrict__ lm_index,
                                                                        int lm_cap, double max_dist, int* __restrict__ next_index,
                                                                        int* __restrict__ overflow, const int* __restrict__ marks,
                                                                        int* __restrict__ runs, int* __restrict__ stage_key,
                                                                        int* __restrict__ stage_src)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int seq = blockIdx.x, cap = cur.cap;
    const trk_map m = cur.at(seq, cam);
    const float* xy = dxy + (size_t)seq * cells * 2; const float* rs = dresp + (size_t)seq * cells;
    const uint8_t* ds = ddesc + (size_t)seq * cells * 32;
    const int* mt = lm_match + (size_t)seq * cells; const float* md = lm_dist + (size_t)seq * cells;
    const int* li = lm_index + (size_t)seq * lm_cap;
    int* skey = stage_key + (size_t)seq * cap; int* ssrc = stage_src + (size_t)seq * cap;
    const int n = min(*m.n, cap), k = min(dn[seq], cells), first = next_index[seq];
    // the map so far: left = its temporal tracks; right = [temporal tracks | keypoints tracked over from the left camera]
    const int b1 = cam == 0 ? n : min(max(marks[4 * seq + 2], 0), n);
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < k; base += TRK_THREADS) {               // fresh indices, detection order
        const int i = base + threadIdx.x;
        bool f = false;
        if (i < k) { const int r = mt[i]; f = !(r >= 0 && (double)md[i] <= max_dist); }
        const int o = n + trk_scan_step(f, warp_sums, &carry);
        if (f && o < cap) {
            m.idx[o] = first + i; m.xy[2 * o] = xy[2 * i]; m.xy[2 * o + 1] = xy[2 * i + 1]; m.resp[o] = rs[i];
            trk_copy_desc(m.desc + (size_t)o * 32, ds + (size_t)i * 32);
        }
    }
    const int nA = carry;
    __syncthreads();
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < k; base += TRK_THREADS) {               // landmark indices the map does not hold yet
        const int i = base + threadIdx.x;
        bool f = false;
        int key = -1;
        if (i < k) {
            const int r = mt[i];
            if (r >= 0 && (double)md[i] <= max_dist) {
                key = li[r];
                int lo = trk_lower_bound(m.idx, 0, b1, key);
                bool have = lo < b1 && m.idx[lo] == key;
                if (!have) { lo = trk_lower_bound(m.idx, b1, n, key); have = lo < n && m.idx[lo] == key; }
                f = !have;
            }
        }
        const int o = trk_scan_step(f, warp_sums, &carry);
        if (f && o < cap) { skey[o] = key; ssrc[o] = i; }
    }
    const int nB = min(carry, cap);
    __syncthreads();
    for (int j = threadIdx.x; j < nB; j += TRK_THREADS) {             // ranked by index (a mutual match: keys are unique)
        const int key = skey[j], i = ssrc[j];
        int r = 0;
        for (int q = 0; q < nB; ++q) r += skey[q] < key ? 1 : 0;
        const int o = n + nA + r;
        if (o < cap) {
            m.idx[o] = key; m.xy[2 * o] = xy[2 * i]; m.xy[2 * o + 1] = xy[2 * i + 1]; m.resp[o] = rs[i];
            trk_copy_desc(m.desc + (size_t)o * 32, ds + (size_t)i * 32);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int total = n + nA + nB;
        if (total > cap) *overflow = 1;
        *m.n = min(total, cap);
        next_index[seq] = first + k;
        int* rb = runs + ((size_t)seq * 2 + cam) * 5;
        rb[0] = 0; rb[1] = b1; rb[2] = n; rb[3] = min(n + nA, cap); rb[4] = min(total, cap);
    }
}

// one camera's map of every sequence sorted by index into `out`: an element's rank is its offset in its own ascending run
// plus a lower-bound search in the (at most three) other runs.  grid: S
__global__ void __launch_bounds__(TRK_THREADS) k_trk_sort_runs(trk_maps cur, trk_maps out, int cam, const int* __restrict__ runs)
{
    const int seq = blockIdx.x, cap = cur.cap;
    const trk_map m = cur.at(seq, cam), o = out.at(seq, cam);
    const int n = min(*m.n, cap);
    const int* rb = runs + ((size_t)seq * 2 + cam) * 5;
    int start[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) start[q] = min(max(rb[q], 0), n);
    for (int i = threadIdx.x; i < n; i += TRK_THREADS) {
        const int key = m.idx[i];
        const int run = i < start[1] ? 0 : i < start[2] ? 1 : i < start[3] ? 2 : 3;
        int r = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) r += q == run ? i - start[q] : trk_lower_bound(m.idx, start[q], start[q + 1], key) - start[q];
        o.idx[r] = key; o.xy[2 * r] = m.xy[2 * i]; o.xy[2 * r + 1] = m.xy[2 * i + 1]; o.resp[r] = m.resp[i];
        trk_copy_desc(o.desc + (size_t)r * 32, m.desc + (size_t)i * 32);
    }
    if (threadIdx.x == 0) *o.n = n;
}

// the sorted map back into the working generation; a sorted right map is one run for the final k_trk_sort.  grid: S
__global__ void __launch_bounds__(TRK_THREADS) k_trk_copy_map(trk_maps src, trk_maps dst, int cam, int* __restrict__ marks)
{
    const int seq = blockIdx.x, cap = src.cap;
    const trk_map s = src.at(seq, cam), d = dst.at(seq, cam);
    const int n = min(*s.n, cap);
    for (int i = threadIdx.x; i < n; i += TRK_THREADS) {
        d.idx[i] = s.idx[i]; d.xy[2 * i] = s.xy[2 * i]; d.xy[2 * i + 1] = s.xy[2 * i + 1]; d.resp[i] = s.resp[i];
        trk_copy_desc(d.desc + (size_t)i * 32, s.desc + (size_t)i * 32);
    }
    if (threadIdx.x == 0) {
        *d.n = n;
        if (cam == 1) { marks[4 * seq + 2] = n; marks[4 * seq + 3] = n; }
    }
}

static inline size_t trk_al(size_t v) { return (v + 255) / 256 * 256; }

extern "C" zs_status zs_tracker_create(zs_context* ctx, const zs_tracker_options* opt, zs_tracker** out)
{
    ZS_REQUIRE(ctx && opt && out, "null argument");
    ZS_REQUIRE(opt->width > 0 && opt->height > 0 && opt->cell_w > 0 && opt->cell_h > 0, "bad geometry");
    ZS_REQUIRE((opt->width / opt->cell_w) * (opt->height / opt->cell_h) > 0, "cells larger than the image");
    // GRID (not PARALLEL_GRID) sends a free cell where FAST finds nothing through cv::ORB::detect (keypoint_detector_grid.cpp:92-95),
    // which is not implemented.  That detector has a search area only in cells of at least 63 x 63 px; on its level 0 it finds
    // nothing FAST(threshold <= 20) has not found, and its level 1 needs a 76-px cell: inside these bounds the batched flow is
    // exact, outside it is refused (the per-call host entry checks the actual cells instead, zs_host.cu)
    if (!opt->parallel_grid && opt->cell_w >= 63 && opt->cell_h >= 63 && !(opt->cell_w < 76 && opt->cell_h < 76 && opt->fast_threshold <= 20)) {
        zs_set_error("GRID with %d x %d px cells and FAST threshold %d can reach the reference's ORB::detect fallback for empty cells "
                     "(keypoint_detector_grid.cpp:92-95), which is not implemented: use PARALLEL_GRID or cells <= 62 px",
                     opt->cell_w, opt->cell_h, opt->fast_threshold);
        return ZS_ERR_UNSUPPORTED;
    }
    ZS_REQUIRE(opt->sequences >= 0 && opt->sequences <= 4096, "sequences outside 0..4096");
    ZS_CUDA(cudaSetDevice(ctx->device));
    zs_tracker* t = (zs_tracker*)calloc(1, sizeof(zs_tracker));
    t->ctx = ctx; t->opt = *opt;
    t->S = opt->sequences > 0 ? opt->sequences : 1;
    t->gw = opt->width / opt->cell_w; t->gh = opt->height / opt->cell_h; t->cells = t->gw * t->gh;
    // a camera's map holds its own detections (at most one per cell and frame) plus the keypoints tracked over from the
    // other camera, and tracked keypoints may share a cell: in steady state it settles near twice the cell count; the
    // default leaves a factor of two above that, and overflow is reported, not hidden
    t->cap = opt->capacity > 0 ? opt->capacity : 4 * t->cells + 64;
    zs_status st = zs_pyramid_create(ctx, opt->width, opt->height, 4 * t->S, opt->klt_win_w, opt->klt_win_h, opt->klt_max_level, &t->pyr);
    if (st != ZS_OK) { free(t); return st; }
    const size_t S = t->S, cap = t->cap, cells = t->cells, R = 2 * S;      // R map rows per generation
    size_t off = 0;
    ZS_REQUIRE(opt->landmark_capacity >= 0 && opt->landmark_capacity <= (1 << 22), "landmark_capacity outside 0..4M");
    t->lm_cap = opt->landmark_capacity > 0 ? (opt->landmark_capacity + 127) / 128 * 128 : 0;
    const int gens = t->lm_cap > 0 ? 3 : 2;
    size_t o_map[3][5];                                  // prev, cur (, tmp): idx | xy | resp | desc | n, each [S][2][cap]
    for (int m = 0; m < gens; ++m) {
        o_map[m][0] = off; off += trk_al(sizeof(int) * R * cap);
        o_map[m][1] = off; off += trk_al(sizeof(float) * 2 * R * cap);
        o_map[m][2] = off; off += trk_al(sizeof(float) * R * cap);
        o_map[m][3] = off; off += trk_al(32 * R * cap);
        o_map[m][4] = off; off += trk_al(sizeof(int) * R);
    }
#define TCARVE(name, bytes) const size_t o_##name = off; off += trk_al(bytes);
    TCARVE(slots, sizeof(int) * 12 * S) TCARVE(t_pts, sizeof(float) * 2 * R * cap) TCARVE(t_status, R * cap) TCARVE(t_err, sizeof(float) * R * cap)
    TCARVE(t_keep, R * cap) TCARVE(occ, S * cells) TCARVE(raw_xy, sizeof(float) * 2 * S * cells) TCARVE(raw_resp, sizeof(float) * S * cells)
    TCARVE(raw_n, sizeof(int) * S) TCARVE(det_xy, sizeof(float) * 2 * S * cells) TCARVE(det_resp, sizeof(float) * S * cells)
    TCARVE(det_n, sizeof(int) * S) TCARVE(det_desc, 32 * S * cells) TCARVE(sel, sizeof(int) * S * cap) TCARVE(sel_n, sizeof(int) * S)
    TCARVE(sel_pts, sizeof(float) * 2 * S * cap) TCARVE(next_index, sizeof(int) * S) TCARVE(overflow, 256) TCARVE(marks, sizeof(int) * 4 * S)
    TCARVE(pred_idx, sizeof(int) * R * cap) TCARVE(pred_xy, sizeof(float) * 2 * R * cap) TCARVE(pred_n, sizeof(int) * R)
    const size_t LM = t->lm_cap;
    TCARVE(lm_xyz, sizeof(double) * 3 * S * LM) TCARVE(lm_index, sizeof(int) * S * LM) TCARVE(lm_desc, 32 * S * LM) TCARVE(lm_n, sizeof(int) * S)
    TCARVE(lm_center, sizeof(double) * 3 * S) TCARVE(lm_nt, sizeof(int) * S) TCARVE(lm_match, sizeof(int) * S * cells)
    TCARVE(lm_dist, sizeof(float) * S * cells) TCARVE(lm_runs, sizeof(int) * 10 * S)
#undef TCARVE
    cudaError_t e = cudaMalloc((void**)&t->dev, off);
    if (e != cudaSuccess) { zs_pyramid_destroy(t->pyr); free(t); return zs_cuda_fail(e, "cudaMalloc(tracker)", __FILE__, __LINE__); }
    t->dev_bytes = off;
    e = cudaMemsetAsync(t->dev, 0, off, ctx->stream);
    if (e != cudaSuccess) { cudaFree(t->dev); zs_pyramid_destroy(t->pyr); free(t); return zs_cuda_fail(e, "cudaMemset(tracker)", __FILE__, __LINE__); }
    for (int m = 0; m < gens; ++m) {
        trk_maps& g = m == 0 ? t->prev : m == 1 ? t->cur : t->tmp;
        g.idx = (int*)(t->dev + o_map[m][0]); g.xy = (float*)(t->dev + o_map[m][1]); g.resp = (float*)(t->dev + o_map[m][2]);
        g.desc = t->dev + o_map[m][3]; g.n = (int*)(t->dev + o_map[m][4]); g.cap = t->cap;
    }
#define TBIND(name, type) t->name = (type*)(t->dev + o_##name);
    TBIND(slots, int) TBIND(t_pts, float) TBIND(t_status, uint8_t) TBIND(t_err, float) TBIND(t_keep, uint8_t) TBIND(occ, uint8_t)
    TBIND(raw_xy, float) TBIND(raw_resp, float) TBIND(raw_n, int) TBIND(det_xy, float) TBIND(det_resp, float) TBIND(det_n, int)
    TBIND(det_desc, uint8_t) TBIND(sel, int) TBIND(sel_n, int) TBIND(sel_pts, float) TBIND(next_index, int) TBIND(overflow, int)
    TBIND(marks, int) TBIND(pred_idx, int) TBIND(pred_xy, float) TBIND(pred_n, int)
    TBIND(lm_xyz, double) TBIND(lm_index, int) TBIND(lm_desc, uint8_t) TBIND(lm_n, int) TBIND(lm_center, double) TBIND(lm_nt, int)
    TBIND(lm_match, int) TBIND(lm_dist, float) TBIND(lm_runs, int)
#undef TBIND
    if (t->lm_cap > 0) {
        t->lm_seen = new std::vector<std::unordered_set<int>>(S);
        t->lm_host_n = new std::vector<int>(S, 0);
        t->lm_bucket = t->lm_cap < 2048 ? t->lm_cap : 2048;
    }
    {
        // slot(parity, camera, sequence) = (parity * 2 + camera) * S + sequence: the S images of one camera and parity are
        // contiguous slots, which is what the batched pyramid / detector / descriptor calls address
        int* hs = (int*)malloc(sizeof(int) * 12 * S + sizeof(int) * S);
        for (int par = 0; par < 2; ++par) {
            int* q = hs + 6 * S * par;
            for (size_t sq = 0; sq < S; ++sq)
                for (int cam = 0; cam < 2; ++cam) {
                    q[2 * sq + cam] = (int)(((1 - par) * 2 + cam) * S + sq);               // temporal prev
                    q[2 * S + 2 * sq + cam] = (int)((par * 2 + cam) * S + sq);             // temporal next
                    q[4 * S + cam * S + sq] = (int)((par * 2 + cam) * S + sq);             // this frame's left / right slots
                }
        }
        int* first = hs + 12 * S;
        for (size_t sq = 0; sq < S; ++sq) first[sq] = opt->first_index > 0 ? opt->first_index : 0;
        e = cudaMemcpyAsync(t->slots, hs, sizeof(int) * 12 * S, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(t->next_index, first, sizeof(int) * S, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        free(hs);
        if (e != cudaSuccess) { cudaFree(t->dev); zs_pyramid_destroy(t->pyr); free(t); return zs_cuda_fail(e, "tracker tables", __FILE__, __LINE__); }
    }
    t->graph_ok = !ctx->sw.fe_no_graph;
    *out = t;
    return ZS_OK;
}

extern "C" void zs_tracker_destroy(zs_tracker* t)
{
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    for (int par = 0; par < 2; ++par) if (t->gexec[par]) cudaGraphExecDestroy(t->gexec[par]);
    if (t->s_in) {
        cudaStreamSynchronize(t->s_in); cudaStreamSynchronize(t->s_out);
        for (int b = 0; b < 2; ++b) {
            if (t->stage[b]) cudaFree(t->stage[b]);
            if (t->outbox[b]) cudaFree(t->outbox[b]);
            if (t->h_tail[b]) cudaFreeHost(t->h_tail[b]);
            cudaEventDestroy(t->ev_in[b]); cudaEventDestroy(t->ev_unpacked[b]); cudaEventDestroy(t->ev_run[b]); cudaEventDestroy(t->ev_out[b]);
        }
        cudaStreamDestroy(t->s_in); cudaStreamDestroy(t->s_out);
    }
    if (t->pyr) zs_pyramid_destroy(t->pyr);
    if (t->dev) cudaFree(t->dev);
    delete t->lm_seen;
    delete t->lm_host_n;
    free(t);
}

extern "C" int zs_tracker_capacity(const zs_tracker* t) { return t ? t->cap : 0; }
extern "C" int zs_tracker_sequences(const zs_tracker* t) { return t ? t->S : 0; }

extern "C" zs_status zs_tracker_set_predictions(zs_tracker* t, int sequence, int camera, const int* index, const float* xy, int n)
{
    ZS_REQUIRE(t && (camera == 0 || camera == 1) && sequence >= 0 && sequence < t->S, "bad argument");
    ZS_REQUIRE(n >= 0 && n <= t->cap && (n == 0 || (index && xy)), "bad prediction list");
    zs_context* ctx = t->ctx;
    ZS_CUDA(cudaSetDevice(ctx->device));
    for (int i = 1; i < n; ++i) ZS_REQUIRE(index[i - 1] < index[i], "prediction indices must be strictly ascending");
    const size_t row = (size_t)sequence * 2 + camera;
    if (n > 0) {
        ZS_CUDA(cudaMemcpyAsync(t->pred_idx + row * t->cap, index, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
        ZS_CUDA(cudaMemcpyAsync(t->pred_xy + row * t->cap * 2, xy, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
    }
    ZS_CUDA(cudaMemcpyAsync(t->pred_n + row, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));     // the host arrays may go away after the call
    return ZS_OK;
}

extern "C" zs_status zs_tracker_landmarks_add_host(zs_tracker* t, int sequence, const int* index, const double* xyz, const uint8_t* desc,
                                                   int n, int* n_added)
{
    ZS_REQUIRE(t && sequence >= 0 && sequence < t->S, "bad argument");
    ZS_REQUIRE(t->lm_cap > 0, "the tracker was created without a landmark store (landmark_capacity = 0)");
    ZS_REQUIRE(n >= 0 && (n == 0 || (index && xyz && desc)), "bad landmark list");
    if (n_added) *n_added = 0;
    if (n == 0) return ZS_OK;
    zs_context* ctx = t->ctx;
    ZS_CUDA(cudaSetDevice(ctx->device));
    // map::operator+=(const map&) (types/map.h:222-236): only indices the map does not hold yet are added, in the order given
    std::unordered_set<int>& seen = (*t->lm_seen)[sequence];
    std::vector<int> rows;
    rows.reserve(n);
    for (int i = 0; i < n; ++i) if (seen.find(index[i]) == seen.end()) { seen.insert(index[i]); rows.push_back(i); }
    const int have = (*t->lm_host_n)[sequence], add = (int)rows.size();
    if (have + add > t->lm_cap) {
        for (int i : rows) seen.erase(index[i]);
        zs_set_error("landmark store of sequence %d would hold %d landmarks, landmark_capacity is %d", sequence, have + add, t->lm_cap);
        return ZS_ERR_CAPACITY;
    }
    if (add == 0) return ZS_OK;
    void* pin;
    zs_status st = zs_pinned(ctx, (size_t)add * (24 + 4 + 32), &pin);
    if (st != ZS_OK) return st;
    double* hx = (double*)pin; int* hi = (int*)(hx + 3 * (size_t)add); uint8_t* hd = (uint8_t*)(hi + add);
    for (int j = 0; j < add; ++j) {
        const int i = rows[j];
        hx[3 * j] = xyz[3 * i]; hx[3 * j + 1] = xyz[3 * i + 1]; hx[3 * j + 2] = xyz[3 * i + 2];
        hi[j] = index[i];
        memcpy(hd + (size_t)j * 32, desc + (size_t)i * 32, 32);
    }
    const size_t row0 = (size_t)sequence * t->lm_cap + have;
    const int total = have + add;
    ZS_CUDA(cudaMemcpyAsync(t->lm_xyz + row0 * 3, hx, sizeof(double) * 3 * add, cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(t->lm_index + row0, hi, sizeof(int) * add, cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(t->lm_desc + row0 * 32, hd, (size_t)add * 32, cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(t->lm_n + sequence, &total, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    (*t->lm_host_n)[sequence] = total;
    // the matcher's train-side grid covers lm_bucket rows: grow it in powers of two (a change re-captures the step's graphs)
    while (t->lm_bucket < total) t->lm_bucket = t->lm_bucket * 2 < t->lm_cap ? t->lm_bucket * 2 : t->lm_cap;
    if (n_added) *n_added = add;
    return ZS_OK;
}

extern "C" int zs_tracker_landmarks_size(const zs_tracker* t, int sequence)
{
    return (t && t->lm_cap > 0 && sequence >= 0 && sequence < t->S) ? (*t->lm_host_n)[sequence] : 0;
}

extern "C" zs_status zs_tracker_set_camera_center(zs_tracker* t, int sequence, const double* center)
{
    ZS_REQUIRE(t && center && sequence >= 0 && sequence < t->S, "bad argument");
    ZS_REQUIRE(t->lm_cap > 0, "the tracker was created without a landmark store (landmark_capacity = 0)");
    zs_context* ctx = t->ctx;
    ZS_CUDA(cudaSetDevice(ctx->device));
    ZS_CUDA(cudaMemcpyAsync(t->lm_center + 3 * (size_t)sequence, center, sizeof(double) * 3, cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZS_OK;
}

// detection of one camera of every sequence behind the occupancy of its current map (keypoint_tracker.cpp:53-57 / 69-73)
static zs_status trk_detect(zs_tracker* t, int cam, int first_slot, int* d_mark)
{
    zs_context* ctx = t->ctx;
    const zs_tracker_options& o = t->opt;
    k_trk_occupancy<<<t->S, TRK_THREADS, 0, ctx->stream>>>(t->cur, cam, t->gw, t->gh, o.cell_w, o.cell_h, t->occ);
    ZS_LAUNCH_CHECK(ctx);
    zs_status st = zs_fast_grid_detect(ctx, t->pyr, first_slot, t->S, o.cell_w, o.cell_h, o.fast_threshold, t->occ, t->raw_xy, t->raw_resp,
                                       t->raw_n, t->cells);
    if (st != ZS_OK) return st;
    // PARALLEL_GRID: cv::cornerSubPix(win 5x5, 30 iterations, eps 0.01) on the selected corners before ORB::compute
    // (keypoint_detector_parallel.cpp:160-170)
    if (o.parallel_grid && (st = zs_corner_subpix(ctx, t->pyr, first_slot, t->S, t->raw_xy, t->raw_n, t->cells, 5, 5, 30, 0.01)) != ZS_OK)
        return st;
    if ((st = zs_orb_compute(ctx, t->pyr, first_slot, t->S, t->raw_xy, t->raw_resp, nullptr, t->raw_n, t->cells, t->det_xy, t->det_resp,
                             nullptr, t->det_n, t->det_desc)) != ZS_OK) return st;
    if (t->lm_cap > 0) {
        // assign_landmark_indices(detected, system.points3d, frame_0.pose.translation(), landmark_match_radius,
        // landmark_match_distance) (keypoint_tracker.cpp:55,71): candidates = radius search, then
        // cv::BFMatcher(NORM_HAMMING, true).match(detected, candidates); an empty store matches nothing (:208)
        if ((st = zs_lm_radius_count(ctx, t->lm_xyz, t->lm_n, t->lm_cap, t->lm_center, o.landmark_match_radius, t->S, t->lm_nt)) != ZS_OK)
            return st;
        if ((st = zs_match_hamming_cross(ctx, t->det_desc, t->det_n, (size_t)t->cells * 32, t->lm_desc, t->lm_nt, (size_t)t->lm_cap * 32, t->S,
                                         t->cells, t->lm_bucket, t->lm_match, t->lm_dist)) != ZS_OK) return st;
        k_trk_append_detected_lm<<<t->S, TRK_THREADS, 0, ctx->stream>>>(t->cur, cam, t->cells, t->det_xy, t->det_resp, t->det_desc, t->det_n,
                                                                        t->lm_match, t->lm_dist, t->lm_index, t->lm_cap,
                                                                        o.landmark_match_distance, t->next_index, t->overflow, t->marks,
                                                                        t->lm_runs, t->sel, (int*)t->sel_pts);
        ZS_LAUNCH_CHECK(ctx);
        k_trk_sort_runs<<<t->S, TRK_THREADS, 0, ctx->stream>>>(t->cur, t->tmp, cam, t->lm_runs);
        ZS_LAUNCH_CHECK(ctx);
        k_trk_copy_map<<<t->S, TRK_THREADS, 0, ctx->stream>>>(t->tmp, t->cur, cam, t->marks);
        ZS_LAUNCH_CHECK(ctx);
        return ZS_OK;
    }
    k_trk_append_detected<<<t->S, TRK_THREADS, 0, ctx->stream>>>(t->cur, cam, t->cells, t->det_xy, t->det_resp, t->det_desc, t->det_n,
                                                                 t->next_index, t->overflow, d_mark);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

// stereo track of the keypoints of camera `from` that camera `to` lacks (keypoint_tracker.cpp:59-67 / 75-83); `to`'s map is
// sorted by index at that point
static zs_status trk_stereo(zs_tracker* t, int from, int to, const int* d_slot_from, const int* d_slot_to, const zs_lk_params* prm,
                            int* d_mark)
{
    zs_context* ctx = t->ctx;
    k_trk_unmatched<<<t->S, TRK_THREADS, 0, ctx->stream>>>(t->cur, from, to, t->sel, t->sel_n, t->sel_pts);
    ZS_LAUNCH_CHECK(ctx);
    zs_status st = zs_klt_launch(ctx, t->pyr, d_slot_from, d_slot_to, t->sel_pts, t->t_pts, t->sel_n, nullptr, t->S, t->cap, prm, t->t_status,
                                 t->t_err, 1, t->opt.klt_threshold, t->t_keep);
    if (st != ZS_OK) return st;
    k_trk_append_tracked<<<t->S, TRK_THREADS, 0, ctx->stream>>>(t->cur, from, to, t->sel, t->sel_n, t->t_pts, t->t_keep, t->overflow, d_mark);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

// everything between the uploads and the result copies for a frame of parity `par`
static zs_status trk_frame_body(zs_tracker* t, int par)
{
    zs_context* ctx = t->ctx;
    const zs_tracker_options& o = t->opt;
    const int S = t->S, cap = t->cap, first_l = (par * 2) * S, first_r = (par * 2 + 1) * S;
    const int* sl = t->slots + 6 * S * par;
    zs_status st;
    // the overflow flag belongs to ONE step: it is cleared here and reported (ZS_ERR_CAPACITY) by the download / wait of this
    // step only -- a map that shrinks again stops failing.  A step that overflowed has truncated maps (keypoints past the
    // capacity are dropped, keypoint::index_next still advances past them).
    ZS_CUDA(cudaMemsetAsync(t->overflow, 0, sizeof(int), ctx->stream));
    if ((st = zs_pyramid_build(ctx, t->pyr, first_l, 2 * S)) != ZS_OK) return st;      // left and right slots are adjacent
    zs_lk_params prm;
    prm.win_w = o.klt_win_w; prm.win_h = o.klt_win_h; prm.max_level = o.klt_max_level; prm.max_iters = 99; prm.epsilon = 0.001;
    prm.flags = ZS_LK_GET_MIN_EIGENVALS; prm.min_eig_threshold = 1e-4;
    // 1. temporal tracks of both cameras (:47-51, the overload with initial flow :343-434), one launch with 2S jobs;
    //    frame 0 has no previous keypoints (n = 0).  OPTFLOW_USE_INITIAL_FLOW with the keypoint's own position is what the
    //    plain call starts from, so the flag is always on and the graph is the same with and without predictions.
    k_trk_init_flow<<<2 * S, TRK_THREADS, 0, ctx->stream>>>(t->prev, t->pred_idx, t->pred_xy, t->pred_n, t->t_pts);
    ZS_LAUNCH_CHECK(ctx);
    zs_lk_params prm_init = prm;
    prm_init.flags |= ZS_LK_USE_INITIAL_FLOW;
    if ((st = zs_klt_launch(ctx, t->pyr, sl, sl + 2 * S, t->prev.xy, t->t_pts, t->prev.n, nullptr, 2 * S, cap, &prm_init, t->t_status, t->t_err,
                            1, o.klt_threshold, t->t_keep)) != ZS_OK) return st;
    k_trk_compact<<<2 * S, TRK_THREADS, 0, ctx->stream>>>(t->prev, t->cur, t->t_pts, t->t_keep);
    ZS_LAUNCH_CHECK(ctx);
    // 2. new left keypoints in the free cells (:53-57); same sorted run: new indices exceed every tracked one
    if ((st = trk_detect(t, 0, first_l, nullptr)) != ZS_OK) return st;
    // 3. left keypoints the right camera lacks: L -> R (:59-67); the right map holds only its temporal tracks, sorted
    if ((st = trk_stereo(t, 0, 1, sl + 4 * S, sl + 5 * S, &prm, t->marks + 2)) != ZS_OK) return st;
    // 4. new right keypoints behind the occupancy of everything the right map now holds (:69-73)
    if ((st = trk_detect(t, 1, first_r, t->marks + 3)) != ZS_OK) return st;
    // 5. right keypoints the left camera lacks: R -> L (:75-83); the left map (tracks + detections) is still sorted
    if ((st = trk_stereo(t, 1, 0, sl + 5 * S, sl + 4 * S, &prm, t->marks + 0)) != ZS_OK) return st;
    // 6. key order for the output and for the next frame's searches; the sorted maps become `prev`
    //    left map: [tracks + detections | from the right camera] ; right map: [tracks | from the left camera | detections]
    k_trk_sort<<<2 * S, TRK_THREADS, 0, ctx->stream>>>(t->cur, t->prev, t->marks);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

// eager the first time a parity is seen, then captured once and replayed (same scheme as zs_frontend_run)
static zs_status trk_frame(zs_tracker* t, int par)
{
    zs_context* ctx = t->ctx;
    if (!t->graph_ok || t->runs[par]++ == 0) return trk_frame_body(t, par);
    if (t->gexec[par] && (t->g_scratch[par] != ctx->scratch || t->g_bucket[par] != t->lm_bucket)) {
        cudaGraphExecDestroy(t->gexec[par]); t->gexec[par] = nullptr;        // the captured launches bake scratch pointers and grid sizes in
    }
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
        t->g_bucket[par] = t->lm_bucket;
    }
    ZS_CUDA(cudaGraphLaunch(t->gexec[par], ctx->stream));
    ctx->launches += t->g_launches[par];
    return ZS_OK;
}

// one time step with the new frames already on the device (or on the host: src_is_host); results stay on the device
static zs_status trk_step(zs_tracker* t, const uint8_t* left, const uint8_t* right, size_t pitch, size_t stride, int src_is_host)
{
    zs_context* ctx = t->ctx;
    const int S = t->S;
    if (stride == 0) stride = pitch * t->opt.height;
    const int par = (int)(t->frame & 1);
    zs_status st;
    if ((st = zs_pyramid_upload(ctx, t->pyr, left, pitch, stride, (par * 2) * S, S, src_is_host)) != ZS_OK) return st;
    if ((st = zs_pyramid_upload(ctx, t->pyr, right, pitch, stride, (par * 2 + 1) * S, S, src_is_host)) != ZS_OK) return st;
    if ((st = trk_frame(t, par)) != ZS_OK) return st;
    t->frame++;
    return ZS_OK;
}

extern "C" zs_status zs_tracker_track(zs_tracker* t, const uint8_t* d_left, const uint8_t* d_right, size_t pitch, size_t stride)
{
    ZS_REQUIRE(t && d_left && d_right, "null argument");
    ZS_CUDA(cudaSetDevice(t->ctx->device));
    return trk_step(t, d_left, d_right, pitch, stride, 0);
}

extern "C" zs_status zs_tracker_download(zs_tracker* t, const zs_tracker_results* res)
{
    ZS_REQUIRE(t && res, "null argument");
    ZS_REQUIRE(res->cap >= t->cap, "results.cap must be at least zs_tracker_capacity()");
    zs_context* ctx = t->ctx;
    ZS_CUDA(cudaSetDevice(ctx->device));
    const int S = t->S, cap = t->cap;
    zs_status st;
    // the counters first, then only the live part of every array.  Host arrays: n [S][2], next_index [S],
    // per camera index [S][res->cap], xy [S][res->cap][2], response [S][res->cap], desc [S][res->cap][32]
    void* pin;
    if ((st = zs_pinned(ctx, sizeof(int) * (3 * (size_t)S + 1), &pin)) != ZS_OK) return st;
    int* h_n = (int*)pin; int* h_next = h_n + 2 * S; int* h_over = h_next + S;
    ZS_CUDA(cudaMemcpyAsync(h_n, t->prev.n, sizeof(int) * 2 * S, cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(h_next, t->next_index, sizeof(int) * S, cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(h_over, t->overflow, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    // one strided copy per field and camera: rows = sequences, row length = the longest live map of that camera
    const size_t rc = (size_t)res->cap;
    for (int cam = 0; cam < 2; ++cam) {
        size_t c = 0;
        for (int sq = 0; sq < S; ++sq) { const size_t v = (size_t)(h_n[2 * sq + cam] < cap ? h_n[2 * sq + cam] : cap); c = v > c ? v : c; }
        if (c == 0) continue;
        const trk_map m = t->prev.at(0, cam);
        const size_t dr = 2 * (size_t)cap;                 // device rows of one camera are two maps apart
        if (res->index[cam]) ZS_CUDA(cudaMemcpy2DAsync(res->index[cam], rc * sizeof(int), m.idx, dr * sizeof(int), c * sizeof(int), S, cudaMemcpyDeviceToHost, ctx->stream));
        if (res->xy[cam]) ZS_CUDA(cudaMemcpy2DAsync(res->xy[cam], rc * 2 * sizeof(float), m.xy, dr * 2 * sizeof(float), c * 2 * sizeof(float), S, cudaMemcpyDeviceToHost, ctx->stream));
        if (res->response[cam]) ZS_CUDA(cudaMemcpy2DAsync(res->response[cam], rc * sizeof(float), m.resp, dr * sizeof(float), c * sizeof(float), S, cudaMemcpyDeviceToHost, ctx->stream));
        if (res->desc[cam]) ZS_CUDA(cudaMemcpy2DAsync(res->desc[cam], rc * 32, m.desc, dr * 32, c * 32, S, cudaMemcpyDeviceToHost, ctx->stream));
    }
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (res->n) memcpy(res->n, h_n, sizeof(int) * 2 * S);
    if (res->next_index) memcpy(res->next_index, h_next, sizeof(int) * S);
    if (*h_over) { zs_set_error("tracker capacity %d exceeded", cap); return ZS_ERR_CAPACITY; }
    return ZS_OK;
}

extern "C" zs_status zs_tracker_track_host(zs_tracker* t, const uint8_t* left, const uint8_t* right, size_t pitch, size_t stride,
                                           const zs_tracker_results* res)
{
    ZS_REQUIRE(t && left && right && res, "null argument");
    ZS_REQUIRE(res->cap >= t->cap, "results.cap must be at least zs_tracker_capacity()");
    ZS_CUDA(cudaSetDevice(t->ctx->device));
    zs_status st = trk_step(t, left, right, pitch, stride, 1);
    if (st != ZS_OK) return st;
    return zs_tracker_download(t, res);
}

// filter_epipolar (keypoint_tracker.cpp:293-341) with the fundamental matrix supplied by the caller (the reference estimates
// it with cv::findFundamentalMat RANSAC on the matched points -- a CPU step -- right before this test): both maps keep only
// the keypoints whose index is in BOTH maps and whose epipolar error |pt0^T F pt1| is below the threshold.  Matx product
// order: (pt0^T F) first, each sum left to right in double.  One block per sequence; `prev` holds the maps of the last step.
__global__ void __launch_bounds__(TRK_THREADS) k_trk_filter_epipolar(trk_maps maps, trk_maps tmp, int seq, double f00, double f01, double f02,
                                                                     double f10, double f11, double f12, double f20, double f21, double f22,
                                                                     double threshold)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int cap = maps.cap;
    const trk_map a = maps.at(seq, 0), b = maps.at(seq, 1), oa = tmp.at(seq, 0), ob = tmp.at(seq, 1);
    const int na = min(*a.n, cap), nb = min(*b.n, cap);
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < na; base += TRK_THREADS) {
        const int i = base + threadIdx.x;
        bool f = false;
        int j = 0;
        if (i < na) {
            const int key = a.idx[i];
            j = trk_lower_bound(b.idx, 0, nb, key);
            if (j < nb && b.idx[j] == key) {
                const double x0 = (double)a.xy[2 * i], y0 = (double)a.xy[2 * i + 1], x1 = (double)b.xy[2 * j], y1 = (double)b.xy[2 * j + 1];
                const double r0 = __dadd_rn(__dadd_rn(__dmul_rn(x0, f00), __dmul_rn(y0, f10)), f20);      // pt0 = (x0, y0, 1)
                const double r1 = __dadd_rn(__dadd_rn(__dmul_rn(x0, f01), __dmul_rn(y0, f11)), f21);
                const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(x0, f02), __dmul_rn(y0, f12)), f22);
                const double err = __dadd_rn(__dadd_rn(__dmul_rn(r0, x1), __dmul_rn(r1, y1)), r2);
                f = fabs(err) < threshold;
            }
        }
        const int o = trk_scan_step(f, warp_sums, &carry);
        if (f) {
            oa.idx[o] = a.idx[i]; oa.xy[2 * o] = a.xy[2 * i]; oa.xy[2 * o + 1] = a.xy[2 * i + 1]; oa.resp[o] = a.resp[i];
            trk_copy_desc(oa.desc + (size_t)o * 32, a.desc + (size_t)i * 32);
            ob.idx[o] = b.idx[j]; ob.xy[2 * o] = b.xy[2 * j]; ob.xy[2 * o + 1] = b.xy[2 * j + 1]; ob.resp[o] = b.resp[j];
            trk_copy_desc(ob.desc + (size_t)o * 32, b.desc + (size_t)j * 32);
        }
    }
    __syncthreads();
    const int kept = carry;
    // copy the compacted maps back (tmp = the other generation, free between steps)
    for (int i = threadIdx.x; i < kept; i += TRK_THREADS) {
        a.idx[i] = oa.idx[i]; a.xy[2 * i] = oa.xy[2 * i]; a.xy[2 * i + 1] = oa.xy[2 * i + 1]; a.resp[i] = oa.resp[i];
        trk_copy_desc(a.desc + (size_t)i * 32, oa.desc + (size_t)i * 32);
        b.idx[i] = ob.idx[i]; b.xy[2 * i] = ob.xy[2 * i]; b.xy[2 * i + 1] = ob.xy[2 * i + 1]; b.resp[i] = ob.resp[i];
        trk_copy_desc(b.desc + (size_t)i * 32, ob.desc + (size_t)i * 32);
    }
    if (threadIdx.x == 0) { *a.n = kept; *b.n = kept; }
}

extern "C" zs_status zs_tracker_filter_epipolar(zs_tracker* t, int sequence, const double* F, double threshold)
{
    ZS_REQUIRE(t && F && sequence >= 0 && sequence < t->S, "bad argument");
    ZS_REQUIRE(t->submitted == t->waited, "steps are still in flight: wait for them first");
    zs_context* ctx = t->ctx;
    ZS_CUDA(cudaSetDevice(ctx->device));
    // with pt0 = (x0, y0, 1): the 1 * F(2, j) terms are added last, like Matx's k = 2 term
    k_trk_filter_epipolar<<<1, TRK_THREADS, 0, ctx->stream>>>(t->prev, t->cur, sequence, F[0], F[1], F[2], F[3], F[4], F[5], F[6], F[7], F[8], threshold);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

// ------------------------------------------------------------------------------------------------------
// Pipelined host path: see the comment in zs_tracker.  Host buffers should be pinned (cudaHostAlloc /
// cudaHostRegister) for the copies to overlap; pageable buffers work but serialise inside the driver.
// ------------------------------------------------------------------------------------------------------
static zs_status trk_pipeline_init(zs_tracker* t)
{
    if (t->s_in) return ZS_OK;
    const size_t S = t->S, cap = t->cap, R = 2 * S;
    ZS_CUDA(cudaStreamCreateWithFlags(&t->s_in, cudaStreamNonBlocking));
    ZS_CUDA(cudaStreamCreateWithFlags(&t->s_out, cudaStreamNonBlocking));
    // outbox: idx [R][cap] | xy [R][cap][2] | resp [R][cap] | desc [R][cap][32] | n [R] | next_index [S] | overflow
    t->outbox_bytes = trk_al(sizeof(int) * R * cap) + trk_al(sizeof(float) * 2 * R * cap) + trk_al(sizeof(float) * R * cap) +
                      trk_al(32 * R * cap) + trk_al(sizeof(int) * (R + S + 1));
    for (int b = 0; b < 2; ++b) {
        ZS_CUDA(cudaEventCreateWithFlags(&t->ev_in[b], cudaEventDisableTiming));
        ZS_CUDA(cudaEventCreateWithFlags(&t->ev_unpacked[b], cudaEventDisableTiming));
        ZS_CUDA(cudaEventCreateWithFlags(&t->ev_run[b], cudaEventDisableTiming));
        ZS_CUDA(cudaEventCreateWithFlags(&t->ev_out[b], cudaEventDisableTiming));
        ZS_CUDA(cudaMalloc((void**)&t->stage[b], 2 * S * (size_t)t->opt.width * t->opt.height));
        ZS_CUDA(cudaMalloc((void**)&t->outbox[b], t->outbox_bytes));
        ZS_CUDA(cudaMallocHost((void**)&t->h_tail[b], sizeof(int) * (R + S + 1)));
    }
    return ZS_OK;
}

extern "C" zs_status zs_tracker_wait(zs_tracker* t)
{
    ZS_REQUIRE(t, "null argument");
    ZS_REQUIRE(t->waited < t->submitted, "zs_tracker_wait: nothing in flight");
    const int b = (int)(t->waited & 1), S = t->S;
    ZS_CUDA(cudaEventSynchronize(t->ev_out[b]));
    const zs_tracker_results& res = t->pending[b];
    const int* tail = t->h_tail[b];
    if (res.n) memcpy(res.n, tail, sizeof(int) * 2 * S);
    if (res.next_index) memcpy(res.next_index, tail + 2 * S, sizeof(int) * S);
    t->busy[b] = false;
    t->waited++;
    if (tail[3 * S]) { zs_set_error("tracker capacity %d exceeded", t->cap); return ZS_ERR_CAPACITY; }
    return ZS_OK;
}

extern "C" int zs_tracker_in_flight(const zs_tracker* t) { return t ? (int)(t->submitted - t->waited) : 0; }

// res: where the maps of THIS step go; the arrays must stay valid until the zs_tracker_wait that returns this step
extern "C" zs_status zs_tracker_submit_host(zs_tracker* t, const uint8_t* left, const uint8_t* right, size_t pitch, size_t stride,
                                            const zs_tracker_results* res)
{
    ZS_REQUIRE(t && left && right && res, "null argument");
    ZS_REQUIRE(res->cap >= t->cap, "results.cap must be at least zs_tracker_capacity()");
    zs_context* ctx = t->ctx;
    ZS_CUDA(cudaSetDevice(ctx->device));
    zs_status st = trk_pipeline_init(t);
    if (st != ZS_OK) return st;
    const int b = (int)(t->submitted & 1);
    if (t->busy[b] && (st = zs_tracker_wait(t)) != ZS_OK) return st;      // at most two steps in flight
    const size_t S = t->S, cap = t->cap, R = 2 * S, w = t->opt.width, h = t->opt.height;
    if (stride == 0) stride = pitch * h;
    // copy-in: the staging buffer is free once the step that last used it has been unpacked
    ZS_CUDA(cudaStreamWaitEvent(t->s_in, t->ev_unpacked[b], 0));
    cudaMemcpy3DParms c;
    for (int cam = 0; cam < 2; ++cam) {
        memset(&c, 0, sizeof(c));
        c.srcPtr = make_cudaPitchedPtr((void*)(cam ? right : left), pitch, pitch, stride / pitch);
        c.dstPtr = make_cudaPitchedPtr((void*)(t->stage[b] + (size_t)cam * S * w * h), w, w, h);
        c.extent = make_cudaExtent(w, h, S);
        c.kind = cudaMemcpyHostToDevice;
        if (stride % pitch != 0) {
            for (size_t sq = 0; sq < S; ++sq)
                ZS_CUDA(cudaMemcpy2DAsync(t->stage[b] + ((size_t)cam * S + sq) * w * h, w, (cam ? right : left) + sq * stride, pitch, w, h,
                                          cudaMemcpyHostToDevice, t->s_in));
        } else {
            ZS_CUDA(cudaMemcpy3DAsync(&c, t->s_in));
        }
    }
    ZS_CUDA(cudaEventRecord(t->ev_in[b], t->s_in));
    // compute: unpack into this step's pyramid slots, run the step, snapshot the maps
    ZS_CUDA(cudaStreamWaitEvent(ctx->stream, t->ev_in[b], 0));
    ZS_CUDA(cudaStreamWaitEvent(ctx->stream, t->ev_out[b], 0));          // the outbox of two steps ago has been read
    const int par = (int)(t->frame & 1);
    if ((st = zs_pyramid_upload(ctx, t->pyr, t->stage[b], w, w * h, (par * 2) * (int)S, (int)S, 0)) != ZS_OK) return st;
    if ((st = zs_pyramid_upload(ctx, t->pyr, t->stage[b] + S * w * h, w, w * h, (par * 2 + 1) * (int)S, (int)S, 0)) != ZS_OK) return st;
    ZS_CUDA(cudaEventRecord(t->ev_unpacked[b], ctx->stream));
    if ((st = trk_frame(t, par)) != ZS_OK) return st;
    t->frame++;
    uint8_t* ob = t->outbox[b];
    const size_t o_idx = 0, o_xy = o_idx + trk_al(sizeof(int) * R * cap), o_resp = o_xy + trk_al(sizeof(float) * 2 * R * cap),
                 o_desc = o_resp + trk_al(sizeof(float) * R * cap), o_tail = o_desc + trk_al(32 * R * cap);
    const cudaMemcpyKind dd = cudaMemcpyDeviceToDevice;
    ZS_CUDA(cudaMemcpyAsync(ob + o_idx, t->prev.idx, sizeof(int) * R * cap, dd, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(ob + o_xy, t->prev.xy, sizeof(float) * 2 * R * cap, dd, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(ob + o_resp, t->prev.resp, sizeof(float) * R * cap, dd, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(ob + o_desc, t->prev.desc, 32 * R * cap, dd, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(ob + o_tail, t->prev.n, sizeof(int) * R, dd, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(ob + o_tail + sizeof(int) * R, t->next_index, sizeof(int) * S, dd, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(ob + o_tail + sizeof(int) * (R + S), t->overflow, sizeof(int), dd, ctx->stream));
    ZS_CUDA(cudaEventRecord(t->ev_run[b], ctx->stream));
    // copy-out: whole rows (the counts are not known on the host yet); rows of one camera are two maps apart
    ZS_CUDA(cudaStreamWaitEvent(t->s_out, t->ev_run[b], 0));
    const size_t rc = (size_t)res->cap;
    const cudaMemcpyKind dh = cudaMemcpyDeviceToHost;
    for (int cam = 0; cam < 2; ++cam) {
        const size_t r0 = (size_t)cam * cap;               // first element of this camera's first row
        if (res->index[cam]) ZS_CUDA(cudaMemcpy2DAsync(res->index[cam], rc * sizeof(int), ob + o_idx + r0 * sizeof(int), 2 * cap * sizeof(int), cap * sizeof(int), S, dh, t->s_out));
        if (res->xy[cam]) ZS_CUDA(cudaMemcpy2DAsync(res->xy[cam], rc * 2 * sizeof(float), ob + o_xy + r0 * 2 * sizeof(float), 4 * cap * sizeof(float), 2 * cap * sizeof(float), S, dh, t->s_out));
        if (res->response[cam]) ZS_CUDA(cudaMemcpy2DAsync(res->response[cam], rc * sizeof(float), ob + o_resp + r0 * sizeof(float), 2 * cap * sizeof(float), cap * sizeof(float), S, dh, t->s_out));
        if (res->desc[cam]) ZS_CUDA(cudaMemcpy2DAsync(res->desc[cam], rc * 32, ob + o_desc + r0 * 32, 64 * cap, 32 * cap, S, dh, t->s_out));
    }
    ZS_CUDA(cudaMemcpyAsync(t->h_tail[b], ob + o_tail, sizeof(int) * (R + S + 1), dh, t->s_out));
    ZS_CUDA(cudaEventRecord(t->ev_out[b], t->s_out));
    t->pending[b] = *res;
    t->busy[b] = true;
    t->submitted++;
    return ZS_OK;
}
