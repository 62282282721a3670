// zs_common.cuh -- shared declarations of libzenslam_cuda.so (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/zenslam_cuda.h"

#define ZS_MAX_LEVELS 8
#define ZS_LK_CACHE_SLOTS 6

// Device-side view of a zs_pyramid: S slots, `levels` levels.  Level l image plane of slot s starts at
// img[l] + s*img_slot[l]; interior pixel (x,y) at + (y + pad_y)*pitch[l] + pad_x + x.  The derivative
// plane has the same geometry with 4-byte (dx,dy int16) elements: der[l] + s*der_slot[l] (in elements).
struct zs_pyr_view {
    int levels, slots;
    int pad_x, pad_y;            // pad_x = win_w rounded up to 16 (keeps interiors 16-byte aligned)
    int w[ZS_MAX_LEVELS], h[ZS_MAX_LEVELS];
    int pitch[ZS_MAX_LEVELS];    // bytes per padded image row (= elements per padded derivative row)
    size_t slot_stride[ZS_MAX_LEVELS];   // elements per slot plane (pitch * padded height)
    uint8_t* img[ZS_MAX_LEVELS];
    short2* der[ZS_MAX_LEVELS];
    // blurred level-0 image for ORB (un-padded, pitch blur_pitch)
    uint8_t* blur; int blur_pitch; size_t blur_slot;
    // TMA descriptors in device memory: [2*l] = image plane of level l as a (pitch, padded rows, slots) u8 tensor with a
    // 48x33x1 box; [2*l+1] = derivative plane as u32 elements with a 36x33x1 box (TMA box origins must be 16-byte
    // aligned): one box = the patch of one 32x32 window tile.  Coordinates are padded-plane coordinates.
    const void* tmaps;
    // [3] TMA descriptors of the level-0 image plane as u32 elements for the grid-FAST strips (zs_fast.cu): boxes of
    // (16 cells x 16 px + 16) x 16 rows, (4 x 32 + 16) x 32 rows and (64 + 16) x 64 rows; null when they could not be built
    const void* fast_maps;
};

// A/B switches: environment variables read ONCE when a context is created -- never on a launch
// path -- and again only when zs_context_reload_switches is called (the tests / benches flip them between runs).
struct zs_switches {
    bool fe_no_graph, klt_no_tma, klt_no_share, lk_no_cache, fast_v1, l2_no_tensor, l2_one_tile, fast_pretest, subpix_v1, klt_no_persist, fast_no_tma, klt63_four_warps, klt63_unpacked, l2_chains, hamming_no_tensor;
    int pyr_force;               // 0 = by batch size, 1 = ZS_PYR_SPLIT, 2 = ZS_PYR_FUSED
    int hamming_splits, hamming_variant, l2_splits, l2_epi_groups;   // 0 = default
    long long hamming_tensor_min; // ZS_HAMMING_TENSOR_MIN: distances per call from which Hamming matching takes the tensor-core path (0 = default 12 M)
    int klt31_packed;            // ZS_KLT31_PACKED: 24 / 28 = packed-template form of the 31x31 kernel at that many CTAs per SM (experiment)
    int klt63_packed;            // ZS_KLT63_PACKED: 6 / 7 = the packed-template two-tile 63x63 kernel at that many CTAs per SM (default 8)
    int klt_persist_min;         // ZS_KLT_PERSIST_MIN: launches with at least this many (job, point) items take the persistent form (0 = default)
};
void zs_read_switches(zs_switches* s);

struct zs_context {
    int device;
    zs_switches sw;
    cudaStream_t stream;
    bool own_stream;
    int sm_count;
    uint64_t launches;
    // growable device scratch (single-stream use)
    void* scratch; size_t scratch_bytes;
    // pinned host staging for the _host entry points
    void* pinned; size_t pinned_bytes;
    // cached pyramids for the single-image host mirrors: [0] LK (2 slots), [1] detection (1 slot)
    zs_pyramid* host_pyr[2];
    // cached multi-scale ORB detector of zs_detect_keypoints_orb_host and the parameters it was built for
    zs_orb_detector* host_orb; int host_orb_key[8]; float host_orb_sf;
    // which frames the slots of host_pyr[0] hold (zs_host.cu: the reference's eight LK calls per stereo frame see only
    // four distinct images, two of them already seen by the previous frame's calls)
    uint64_t lk_hash[ZS_LK_CACHE_SLOTS], lk_stamp[ZS_LK_CACHE_SLOTS], lk_clock;
    uint64_t lk_hits, lk_misses;
    int* d_async_err;         // device flags raised by kernels of stream-asynchronous entries ([0]: L2 descriptors not integers in 0..255)
    int* d_klt_work;          // work counter of the persistent KLT launch (same allocation; reset in stream order before each launch)
    // two pinned frame buffers of the single-image host entries (zs_host.cu: pageable frame -> pinned, row chunk by row chunk,
    // each chunk's DMA running behind the host copy of the next one)
    uint8_t* frame_pin[2]; size_t frame_pin_bytes[2]; int frame_pin_next;
    uint8_t* lk_copy[ZS_LK_CACHE_SLOTS];   // sampled rows of the frame each slot holds: a hash hit is confirmed against them (zs_host.cu)
    size_t lk_copy_bytes[ZS_LK_CACHE_SLOTS];
};

struct zs_pyramid {
    zs_context* ctx;
    int width, height, slots, win_w, win_h, max_level;
    zs_pyr_view v;
    void* block;        // one allocation backing every plane
    size_t block_bytes;
    void* tmaps_dev;    // device copy of the CUtensorMap array (2 per level, then the 3 grid-FAST maps)
};

void zs_set_error(const char* fmt, ...);
zs_status zs_cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define ZS_CUDA(call)                                                        \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) return zs_cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define ZS_LAUNCH_CHECK(ctx)                                                 \
    do {                                                                     \
        (ctx)->launches++;                                                   \
        cudaError_t e__ = cudaGetLastError();                                \
        if (e__ != cudaSuccess) return zs_cuda_fail(e__, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define ZS_REQUIRE(cond, msg)                                                \
    do {                                                                     \
        if (!(cond)) { zs_set_error("%s:%d: %s", __FILE__, __LINE__, msg); return ZS_ERR_INVALID; } \
    } while (0)

// stream-ordered temporary of one host-mirror call: freed on every exit path (the error returns of ZS_CUDA included)
struct zs_async_buffer {
    uint8_t* p = nullptr;
    cudaStream_t stream;
    explicit zs_async_buffer(cudaStream_t s) : stream(s) {}
    ~zs_async_buffer() { if (p) cudaFreeAsync(p, stream); }
    zs_async_buffer(const zs_async_buffer&) = delete;
    zs_async_buffer& operator=(const zs_async_buffer&) = delete;
};

zs_status zs_scratch(zs_context* ctx, size_t bytes, void** out);
zs_status zs_pinned(zs_context* ctx, size_t bytes, void** out);

static inline int zs_div_up(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ int zs_reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}

__device__ __forceinline__ int zs_slot(int first, int i, int slots) { return (first + i) % slots; }

// cuTensorMapEncodeTiled through the runtime's driver entry point (zs_context.cu); null when the driver lacks it
typedef CUresult (*zs_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
zs_encode_tiled_fn zs_get_encode_tiled();

// internal launchers (defined in the per-stage .cu files)
zs_status zs_lm_radius_count(zs_context* ctx, const double* d_xyz, const int* d_n, int cap, const double* d_center, double radius,
                             int sequences, int* d_nt);
bool zs_klt_tiled_window(const zs_context* ctx, const zs_pyramid* p, int win_w, int win_h);
zs_status zs_launch_orb_blur(zs_context* ctx, const zs_pyramid* p, int first, int count);
zs_status zs_klt_launch(zs_context* ctx, const zs_pyramid* p, const int* d_prev_slot, const int* d_next_slot,
                        const float* d_prev_pts, float* d_next_pts, const int* d_count, const int* d_pts_row, int jobs, int cap,
                        const zs_lk_params* prm, uint8_t* d_status, float* d_err, int fb, double fb_thr, uint8_t* d_keep,
                        const int* d_next_slot2 = nullptr, const int* d_out_job2 = nullptr, const int* d_job_list = nullptr,
                        int n_list = 0);
