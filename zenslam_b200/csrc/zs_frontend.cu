// zs_frontend.cu -- the batched stereo front-end: the per-frame call pattern of keypoint_tracker::track
// (zenslam_core/source/tracking/keypoint_tracker.cpp:41-105) + processor's pyramid builds (processor.cpp:37,53)
// for B consecutive stereo frames per launch sequence.
//
// Slot layout of the single zs_pyramid (2B + 2 slots):   0 = carried left frame, 1 = carried right frame,
//   2 .. B+1 = left frames of the batch, B+2 .. 2B+1 = right frames.  Keypoint arrays use the same rows.
// Per batch (all on one stream, 2B images / 4B KLT jobs per launch):
//   upload -> pyramid build -> grid FAST -> ORB (blur, filter, rBRIEF) -> stereo Hamming kNN + ratio
//   -> fused forward+backward KLT with FB gate for {temporal L, temporal R, stereo L->R, stereo R->L}
//   -> carry the last frame (pyramid planes + keypoints) into slots 0/1.
#include <stdlib.h>

#include "zs_common.cuh"
#include <algorithm>

#define ZS_FE_STAGES 6          // pyramid, fast, orb, match, klt, carry
#define ZS_FE_TIMING_RING 64

struct zs_frontend {
    zs_context* ctx;
    zs_frontend_options opt;
    int B, cap, gw, gh, slots;
    zs_pyramid* pyr;
    uint8_t* dev; size_t dev_bytes;
    // device arrays (rows = slots unless noted)
    float* raw_xy; float* raw_resp; int* raw_n;          // [2B][cap]: grid candidates before ORB's border filter
    float* xy; float* resp; int* n; uint8_t* desc;       // [slots][cap]
    int* m_idx; float* m_dist; uint8_t* m_pass;          // [B][cap]
    int* job_prev; int* job_next; int* job_row;          // [4B]
    int* job_next2; int* job_out2; int* job_prev_all;    // [4B] template sharing (see zs_frontend_create)
    int* job_list; int n_job_list;                       // jobs that still own work once temporal jobs are folded
    float* t_pts; uint8_t* t_status; float* t_err; uint8_t* t_keep;   // [4B][cap]
    int* t_n;                                            // [4B] points tracked per job
    bool have_carry; bool share;
    // CUDA graph of one zs_frontend_run (the launch sequence is static once a previous frame is carried): captured on the
    // second run, replayed afterwards -- at small batches the ~25 launches per run are launch-bound
    bool graph_ok; cudaGraphExec_t gexec; void* g_scratch; uint64_t g_launches;
    // optional per-stage device timing: a ring of event sets, one set per zs_frontend_run
    int timing; int t_runs;
    cudaEvent_t ev[ZS_FE_TIMING_RING][ZS_FE_STAGES + 1];
    // pinned host staging for process_host
    uint8_t* pin; size_t pin_bytes;
    // pipelined host path (zs_frontend_submit_host / zs_frontend_wait): two in-flight batches.
    //   copy-in stream:  H2D of batch k+1 into stage[(k+1)&1]      } all three overlap; the compute stream
    //   compute stream:  unpack + hot path of batch k, snapshot    } is the context's stream
    //   copy-out stream: D2H of batch k-1 from outbox[(k-1)&1]     }
    cudaStream_t s_in, s_out;
    uint8_t* stage[2]; size_t stage_bytes;     // raw host layout (pitch, stride) of 2B images
    uint8_t* outbox[2]; size_t outbox_bytes;   // snapshot of every result array
    cudaEvent_t ev_in[2], ev_unpacked[2], ev_run[2], ev_out[2];
    bool busy[2]; uint64_t submitted, waited;
    // optional pre-processing of the raw camera frames (processor::process, processor.cpp:25-55): BGR -> gray,
    // CLAHE, remap with per-camera maps, written straight into level 0 of the pyramid slots
    int pp_enabled, pp_channels, pp_clahe; double pp_clip;
    float* pp_maps;                 // device: [4][H][W] = map_x_left, map_y_left, map_x_right, map_y_right (null: no remap)
    uint8_t* pp_tmp[2]; size_t pp_pitch;   // two gray planes [2B][H][pp_pitch]
};

#define ZS_FE_MARK(i) do { if (fe->timing) ZS_CUDA(cudaEventRecord(fe->ev[fe->t_runs % ZS_FE_TIMING_RING][i], ctx->stream)); } while (0)

static inline size_t al256(size_t v) { return (v + 255) / 256 * 256; }

// t_n[j] = n[row[j]]
// the carried frame: up to 2 x (2 planes per level + keypoints + count) device-to-device segments in one launch; every
// segment is 4-byte granular and 16-byte aligned unless it is shorter than that (the keypoint count)
#define FE_CARRY_BLOCKS 16
struct fe_carry_seg { const void* src; void* dst; size_t bytes; };
struct fe_carry_args { fe_carry_seg seg[2 * (2 * ZS_MAX_LEVELS + 2)]; int n; };

__global__ void __launch_bounds__(256) k_fe_carry(const fe_carry_args a)
{
    const fe_carry_seg s = a.seg[blockIdx.y];
    const bool wide = (((uintptr_t)s.src | (uintptr_t)s.dst) & 15) == 0;
    const size_t n16 = wide ? s.bytes >> 4 : 0;
    const size_t step = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint4* s16 = (const uint4*)s.src;
    uint4* d16 = (uint4*)s.dst;
    for (size_t i = t0; i < n16; i += step) d16[i] = s16[i];
    const uint32_t* s4 = (const uint32_t*)s.src;
    uint32_t* d4 = (uint32_t*)s.dst;
    for (size_t i = n16 * 4 + t0; i < (s.bytes >> 2); i += step) d4[i] = s4[i];
}

__global__ void k_gather_counts(const int* __restrict__ n, const int* __restrict__ row, int jobs, int* __restrict__ out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < jobs) out[j] = n[row[j]];
}

extern "C" zs_status zs_frontend_create(zs_context* ctx, const zs_frontend_options* opt, zs_frontend** out)
{
    ZS_REQUIRE(ctx && opt && out, "null argument");
    ZS_REQUIRE(opt->width > 0 && opt->height > 0 && opt->batch > 0, "bad geometry");
    ZS_REQUIRE(opt->cell_w >= 7 && opt->cell_h >= 7, "cell size must be at least 7");
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
    ZS_CUDA(cudaSetDevice(ctx->device));
    zs_frontend* fe = (zs_frontend*)calloc(1, sizeof(zs_frontend));
    fe->ctx = ctx; fe->opt = *opt; fe->B = opt->batch;
    fe->gw = opt->width / opt->cell_w; fe->gh = opt->height / opt->cell_h;
    fe->cap = fe->gw * fe->gh;
    if (fe->cap < 1) { free(fe); zs_set_error("image smaller than one cell"); return ZS_ERR_INVALID; }
    fe->slots = 2 * fe->B + 2;
    zs_status st = zs_pyramid_create(ctx, opt->width, opt->height, fe->slots, opt->klt_win_w, opt->klt_win_h,
                                     opt->klt_max_level, &fe->pyr);
    if (st != ZS_OK) { free(fe); return st; }
    const size_t B = fe->B, cap = fe->cap, S = fe->slots, J = 4 * B;
    size_t off = 0;
#define CARVE(field, type, count) const size_t o_##field = off; off += al256(sizeof(type) * (count));
    CARVE(raw_xy, float, 2 * (2 * B) * cap) CARVE(raw_resp, float, 2 * B * cap) CARVE(raw_n, int, 2 * B)
    CARVE(xy, float, 2 * S * cap) CARVE(resp, float, S * cap) CARVE(n, int, S) CARVE(desc, uint8_t, S * cap * 32)
    CARVE(m_idx, int, 2 * B * cap) CARVE(m_dist, float, 2 * B * cap) CARVE(m_pass, uint8_t, B * cap)
    CARVE(job_prev, int, J) CARVE(job_next, int, J) CARVE(job_row, int, J)
    CARVE(job_next2, int, J) CARVE(job_out2, int, J) CARVE(job_prev_all, int, J) CARVE(job_list, int, J)
    CARVE(t_pts, float, 2 * J * cap) CARVE(t_status, uint8_t, J * cap) CARVE(t_err, float, J * cap)
    CARVE(t_keep, uint8_t, J * cap) CARVE(t_n, int, J)
#undef CARVE
    cudaError_t e = cudaMalloc((void**)&fe->dev, off);
    if (e != cudaSuccess) { zs_pyramid_destroy(fe->pyr); free(fe); return zs_cuda_fail(e, "cudaMalloc(frontend)", __FILE__, __LINE__); }
    fe->dev_bytes = off;
    cudaMemsetAsync(fe->dev, 0, off, ctx->stream);
#define BIND(field, type) fe->field = (type*)(fe->dev + o_##field);
    BIND(raw_xy, float) BIND(raw_resp, float) BIND(raw_n, int) BIND(xy, float) BIND(resp, float) BIND(n, int)
    BIND(desc, uint8_t) BIND(m_idx, int) BIND(m_dist, float) BIND(m_pass, uint8_t) BIND(job_prev, int) BIND(job_next, int)
    BIND(job_row, int) BIND(job_next2, int) BIND(job_out2, int) BIND(job_prev_all, int) BIND(job_list, int) BIND(t_pts, float) BIND(t_status, uint8_t) BIND(t_err, float) BIND(t_keep, uint8_t) BIND(t_n, int)
#undef BIND
    // job tables: kind-major [4][B].  The stereo job of frame k (L_k -> R_k from the keypoints of L_k) and the temporal
    // job of frame k+1 (L_k -> L_{k+1} from the same keypoints) share their forward template: the temporal job is
    // folded into the stereo job as its second target (job_next2 / job_out2) and is itself skipped (prev slot -1).
    // Only the temporal jobs of frame 0, whose source is the carried frame, run on their own.
    int* h = (int*)malloc(sizeof(int) * 6 * J);
    int *hp = h, *hn = h + J, *hr = h + 2 * J, *hn2 = h + 3 * J, *ho2 = h + 4 * J, *hpa = h + 5 * J;
    for (size_t k = 0; k < B; ++k) {
        const int L = 2 + (int)k, R = (int)B + 2 + (int)k;
        const int Lp = k == 0 ? 0 : L - 1, Rp = k == 0 ? 1 : R - 1;
        hp[0 * B + k] = Lp; hn[0 * B + k] = L; hr[0 * B + k] = Lp;      // temporal left: previous frame's keypoints
        hp[1 * B + k] = Rp; hn[1 * B + k] = R; hr[1 * B + k] = Rp;      // temporal right
        hp[2 * B + k] = L;  hn[2 * B + k] = R; hr[2 * B + k] = L;       // stereo L -> R
        hp[3 * B + k] = R;  hn[3 * B + k] = L; hr[3 * B + k] = R;       // stereo R -> L
        for (int kind = 0; kind < 4; ++kind) { hn2[kind * B + k] = -1; ho2[kind * B + k] = -1; hpa[kind * B + k] = hp[kind * B + k]; }
    }
    for (size_t k = 0; k + 1 < B; ++k) {
        hn2[2 * B + k] = 2 + (int)k + 1;          ho2[2 * B + k] = (int)(0 * B + k + 1);   hp[0 * B + k + 1] = -1;
        hn2[3 * B + k] = (int)B + 2 + (int)k + 1; ho2[3 * B + k] = (int)(1 * B + k + 1);   hp[1 * B + k + 1] = -1;
    }
    const bool share = !ctx->sw.klt_no_share && zs_klt_tiled_window(ctx, fe->pyr, opt->klt_win_w, opt->klt_win_h);
    e = cudaMemcpyAsync(fe->job_prev, share ? hp : hpa, sizeof(int) * J, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(fe->job_next, hn, sizeof(int) * J, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(fe->job_row, hr, sizeof(int) * J, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(fe->job_next2, hn2, sizeof(int) * J, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(fe->job_out2, ho2, sizeof(int) * J, cudaMemcpyHostToDevice, ctx->stream);
    // the jobs that still launch blocks once the temporal jobs are folded: 2B + 2 of the 4B
    fe->n_job_list = 0;
    if (share) {
        int* hl = hpa;                                     // hpa is no longer needed when sharing
        for (size_t j = 0; j < J; ++j) if (hp[j] >= 0) hl[fe->n_job_list++] = (int)j;
        if (e == cudaSuccess) e = cudaMemcpyAsync(fe->job_list, hl, sizeof(int) * fe->n_job_list, cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    free(h);
    if (e != cudaSuccess) { zs_frontend_destroy(fe); return zs_cuda_fail(e, "frontend job tables", __FILE__, __LINE__); }
    fe->share = share;
    fe->graph_ok = !ctx->sw.fe_no_graph;
    *out = fe;
    return ZS_OK;
}

extern "C" void zs_frontend_destroy(zs_frontend* fe)
{
    if (!fe) return;
    cudaSetDevice(fe->ctx->device);
    cudaStreamSynchronize(fe->ctx->stream);
    if (fe->gexec) cudaGraphExecDestroy(fe->gexec);
    if (fe->pyr) zs_pyramid_destroy(fe->pyr);
    if (fe->dev) cudaFree(fe->dev);
    if (fe->pin) cudaFreeHost(fe->pin);
    if (fe->pp_maps) cudaFree(fe->pp_maps);
    if (fe->pp_tmp[0]) cudaFree(fe->pp_tmp[0]);
    if (fe->pp_tmp[1]) cudaFree(fe->pp_tmp[1]);
    if (fe->s_in) {
        cudaStreamSynchronize(fe->s_in); cudaStreamSynchronize(fe->s_out);
        for (int b = 0; b < 2; ++b) {
            if (fe->stage[b]) cudaFree(fe->stage[b]);
            if (fe->outbox[b]) cudaFree(fe->outbox[b]);
            cudaEventDestroy(fe->ev_in[b]); cudaEventDestroy(fe->ev_unpacked[b]);
            cudaEventDestroy(fe->ev_run[b]); cudaEventDestroy(fe->ev_out[b]);
        }
        cudaStreamDestroy(fe->s_in); cudaStreamDestroy(fe->s_out);
    }
    if (fe->ev[0][0])
        for (int r = 0; r < ZS_FE_TIMING_RING; ++r)
            for (int i = 0; i <= ZS_FE_STAGES; ++i) cudaEventDestroy(fe->ev[r][i]);
    free(fe);
}

extern "C" int zs_frontend_capacity(const zs_frontend* fe) { return fe ? fe->cap : 0; }

extern "C" size_t zs_frontend_h2d_bytes(const zs_frontend* fe)
{
    return fe ? (size_t)2 * fe->B * fe->opt.width * fe->opt.height * (fe->pp_enabled ? fe->pp_channels : 1) : 0;
}

extern "C" size_t zs_frontend_d2h_bytes(const zs_frontend* fe)
{
    if (!fe) return 0;
    const size_t B = fe->B, cap = fe->cap;
    // counts + keypoints + responses + descriptors (both cameras) + match idx/dist/pass + tracks pts/keep/n
    return 2 * B * 4 + 2 * B * cap * (8 + 4 + 32) + B * cap * (8 + 8 + 1) + 4 * B * cap * (8 + 1) + 4 * B * 4;
}

extern "C" zs_status zs_frontend_upload(zs_frontend* fe, const uint8_t* left, const uint8_t* right, size_t pitch, size_t stride,
                                        int src_is_host)
{
    ZS_REQUIRE(fe && left && right, "null argument");
    zs_status st = zs_pyramid_upload(fe->ctx, fe->pyr, left, pitch, stride, 2, fe->B, src_is_host);
    if (st != ZS_OK) return st;
    return zs_pyramid_upload(fe->ctx, fe->pyr, right, pitch, stride, fe->B + 2, fe->B, src_is_host);
}

static zs_status frontend_run_body(zs_frontend* fe);

extern "C" zs_status zs_frontend_run(zs_frontend* fe)
{
    ZS_REQUIRE(fe, "null argument");
    zs_context* ctx = fe->ctx;
    ZS_CUDA(cudaSetDevice(ctx->device));
    // eager: first run of a sequence (no carried frame yet; it also sizes the context scratch), per-stage timing on,
    // or capture found unusable (e.g. the legacy default stream)
    if (!fe->graph_ok || fe->timing || !fe->have_carry) return frontend_run_body(fe);
    if (fe->gexec && fe->g_scratch != ctx->scratch) {           // another call re-grew the scratch the kernels point into
        cudaGraphExecDestroy(fe->gexec);
        fe->gexec = nullptr;
    }
    if (!fe->gexec) {
        void* scratch_before = ctx->scratch;
        const uint64_t l0 = ctx->launches;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            fe->graph_ok = false;
            return frontend_run_body(fe);
        }
        const zs_status st = frontend_run_body(fe);
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
        const bool good = st == ZS_OK && e == cudaSuccess && g && ctx->scratch == scratch_before &&
                          cudaGraphInstantiate(&fe->gexec, g, 0) == cudaSuccess;
        if (g) cudaGraphDestroy(g);
        if (!good) {
            cudaGetLastError();
            fe->gexec = nullptr; fe->graph_ok = false;
            ctx->launches = l0;
            // nothing ran during the capture: run it eagerly (a call that is merely illegal while capturing, e.g. an
            // attribute change for a large-window kernel, succeeds now; a genuine error shows up again)
            return frontend_run_body(fe);
        }
        fe->g_launches = ctx->launches - l0;                    // kernels per replay (counted at launch below)
        ctx->launches = l0;
        fe->g_scratch = ctx->scratch;
    }
    ZS_CUDA(cudaGraphLaunch(fe->gexec, ctx->stream));
    ctx->launches += fe->g_launches;
    return ZS_OK;
}

static zs_status frontend_run_body(zs_frontend* fe)
{
    zs_context* ctx = fe->ctx;
    const int B = fe->B, cap = fe->cap;
    const zs_frontend_options& o = fe->opt;
    zs_status st;
    ZS_FE_MARK(0);
    // 1. pyramids of the 2B new images (utils::pyramid, processor.cpp:37,53)
    if ((st = zs_pyramid_build(ctx, fe->pyr, 2, 2 * B)) != ZS_OK) return st;
    if (!fe->have_carry) {
        // very first batch of a sequence: there is no previous frame; slots 0/1 stay empty (n = 0), so the
        // temporal jobs of frame 0 track zero points
        ZS_CUDA(cudaMemsetAsync(fe->n, 0, sizeof(int) * 2, ctx->stream));
    }
    ZS_FE_MARK(1);
    // 2. detection (keypoint_tracker.cpp:53,69 -> keypoint_detector_grid.cpp:39-150), no occupancy: every cell is searched
    if ((st = zs_fast_grid_detect(ctx, fe->pyr, 2, 2 * B, o.cell_w, o.cell_h, o.fast_threshold, nullptr, fe->raw_xy, fe->raw_resp,
                                  fe->raw_n, cap)) != ZS_OK) return st;
    // PARALLEL_GRID (what tumvi.yaml:43 ships): cv::cornerSubPix on every selected corner (keypoint_detector_parallel.cpp:160-170)
    if (o.parallel_grid && (st = zs_corner_subpix(ctx, fe->pyr, 2, 2 * B, fe->raw_xy, fe->raw_n, cap, 5, 5, 30, 0.01)) != ZS_OK) return st;
    ZS_FE_MARK(2);
    // 3. ORB::compute (keypoint_detector_grid.cpp:138)
    if ((st = zs_orb_compute(ctx, fe->pyr, 2, 2 * B, fe->raw_xy, fe->raw_resp, nullptr, fe->raw_n, cap, fe->xy + (size_t)2 * cap * 2,
                             fe->resp + (size_t)2 * cap, nullptr, fe->n + 2, fe->desc + (size_t)2 * cap * 32)) != ZS_OK) return st;
    ZS_FE_MARK(3);
    // 4. stereo kNN + ratio (matcher.cpp:60-75), left = query, right = train
    if ((st = zs_match_hamming_knn2(ctx, fe->desc + (size_t)2 * cap * 32, fe->n + 2, (size_t)cap * 32,
                                    fe->desc + (size_t)(B + 2) * cap * 32, fe->n + B + 2, (size_t)cap * 32, B, cap, cap,
                                    o.matcher_ratio, fe->m_idx, fe->m_dist, fe->m_pass)) != ZS_OK) return st;
    ZS_FE_MARK(4);
    // 5. four forward+backward KLT pairs per frame (keypoint_tracker.cpp:47,50,60-67,76-83)
    zs_lk_params prm;
    prm.win_w = o.klt_win_w; prm.win_h = o.klt_win_h; prm.max_level = o.klt_max_level; prm.max_iters = o.max_iters;
    prm.epsilon = o.epsilon; prm.flags = ZS_LK_GET_MIN_EIGENVALS; prm.min_eig_threshold = o.min_eig_threshold;
    if ((st = zs_klt_launch(ctx, fe->pyr, fe->job_prev, fe->job_next, fe->xy, fe->t_pts, fe->n, fe->job_row, 4 * B, cap, &prm,
                            fe->t_status, fe->t_err, 1, o.klt_threshold, fe->t_keep, fe->share ? fe->job_next2 : nullptr,
                            fe->share ? fe->job_out2 : nullptr, fe->share ? fe->job_list : nullptr, fe->n_job_list)) != ZS_OK) return st;
    k_gather_counts<<<zs_div_up(4 * B, 256), 256, 0, ctx->stream>>>(fe->n, fe->job_row, 4 * B, fe->t_n);
    ZS_LAUNCH_CHECK(ctx);
    ZS_FE_MARK(5);
    // 6. carry the last stereo frame into slots 0/1 (pyramid planes and keypoints): one kernel over all the segments (twenty
    //    separate copy nodes cost 65 us per run -- a quarter of a batch-1 step)
    const zs_pyr_view& v = fe->pyr->v;
    fe_carry_args ca;
    ca.n = 0;
    for (int cam = 0; cam < 2; ++cam) {
        const int src = cam == 0 ? B + 1 : 2 * B + 1, dst = cam;
        for (int l = 0; l < v.levels; ++l) {
            ca.seg[ca.n++] = { v.img[l] + (size_t)src * v.slot_stride[l], v.img[l] + (size_t)dst * v.slot_stride[l], v.slot_stride[l] };
            ca.seg[ca.n++] = { v.der[l] + (size_t)src * v.slot_stride[l], v.der[l] + (size_t)dst * v.slot_stride[l],
                               v.slot_stride[l] * sizeof(short2) };
        }
        ca.seg[ca.n++] = { fe->xy + (size_t)src * cap * 2, fe->xy + (size_t)dst * cap * 2, sizeof(float) * 2 * (size_t)cap };
        ca.seg[ca.n++] = { fe->n + src, fe->n + dst, sizeof(int) };
    }
    // blocks per segment: 64 KB per block of the largest plane, between 16 (752x480: 20 x 16 blocks) and two per SM
    const int cb = (int)std::min<size_t>(std::max<size_t>(FE_CARRY_BLOCKS, v.slot_stride[0] * sizeof(short2) / 65536), 296);
    k_fe_carry<<<dim3(cb, ca.n), 256, 0, ctx->stream>>>(ca);
    ZS_LAUNCH_CHECK(ctx);
    ZS_FE_MARK(6);
    if (fe->timing) fe->t_runs++;
    fe->have_carry = true;
    return ZS_OK;
}

// Per-stage device timing (CUDA events on the context's stream).  enable(1) resets the run counter;
// collect() synchronises and returns the summed milliseconds per stage over the (at most 64 most recent) runs.
extern "C" zs_status zs_frontend_timing_enable(zs_frontend* fe, int on)
{
    ZS_REQUIRE(fe, "null argument");
    ZS_CUDA(cudaSetDevice(fe->ctx->device));
    if (on && !fe->ev[0][0])
        for (int r = 0; r < ZS_FE_TIMING_RING; ++r)
            for (int i = 0; i <= ZS_FE_STAGES; ++i) ZS_CUDA(cudaEventCreate(&fe->ev[r][i]));
    fe->timing = on; fe->t_runs = 0;
    return ZS_OK;
}

extern "C" zs_status zs_frontend_timing_collect(zs_frontend* fe, float* stage_ms_sum, int* runs)
{
    ZS_REQUIRE(fe && stage_ms_sum && runs, "null argument");
    ZS_CUDA(cudaStreamSynchronize(fe->ctx->stream));
    const int n = fe->t_runs < ZS_FE_TIMING_RING ? fe->t_runs : ZS_FE_TIMING_RING;
    for (int i = 0; i < ZS_FE_STAGES; ++i) stage_ms_sum[i] = 0.f;
    for (int r = 0; r < n; ++r)
        for (int i = 0; i < ZS_FE_STAGES; ++i) {
            float ms = 0.f;
            ZS_CUDA(cudaEventElapsedTime(&ms, fe->ev[r][i], fe->ev[r][i + 1]));
            stage_ms_sum[i] += ms;
        }
    *runs = n;
    return ZS_OK;
}

extern "C" zs_status zs_frontend_download(zs_frontend* fe, const zs_frontend_results* r)
{
    ZS_REQUIRE(fe && r, "null argument");
    ZS_REQUIRE(r->cap == fe->cap, "results.cap must equal zs_frontend_capacity()");
    zs_context* ctx = fe->ctx;
    const size_t B = fe->B, cap = fe->cap;
    const cudaMemcpyKind k = cudaMemcpyDeviceToHost;
#define DL(dst, src, bytes) if (dst) ZS_CUDA(cudaMemcpyAsync(dst, src, bytes, k, ctx->stream));
    DL(r->n_left, fe->n + 2, sizeof(int) * B) DL(r->n_right, fe->n + B + 2, sizeof(int) * B)
    DL(r->kp_left, fe->xy + 2 * cap * 2, sizeof(float) * 2 * B * cap) DL(r->kp_right, fe->xy + (B + 2) * cap * 2, sizeof(float) * 2 * B * cap)
    DL(r->resp_left, fe->resp + 2 * cap, sizeof(float) * B * cap) DL(r->resp_right, fe->resp + (B + 2) * cap, sizeof(float) * B * cap)
    DL(r->desc_left, fe->desc + 2 * cap * 32, B * cap * 32) DL(r->desc_right, fe->desc + (B + 2) * cap * 32, B * cap * 32)
    DL(r->match_idx, fe->m_idx, sizeof(int) * 2 * B * cap) DL(r->match_dist, fe->m_dist, sizeof(float) * 2 * B * cap)
    DL(r->match_pass, fe->m_pass, B * cap)
    DL(r->track_pts, fe->t_pts, sizeof(float) * 2 * 4 * B * cap) DL(r->track_keep, fe->t_keep, 4 * B * cap)
    DL(r->track_n, fe->t_n, sizeof(int) * 4 * B)
#undef DL
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZS_OK;
}

extern "C" zs_status zs_frontend_submit_host(zs_frontend* fe, const uint8_t* left, const uint8_t* right, size_t pitch,
                                             size_t stride, const zs_frontend_results* r);
extern "C" zs_status zs_frontend_wait(zs_frontend* fe);

// synchronous end-to-end call = one submission through the staged path, waited for at once
extern "C" zs_status zs_frontend_process_host(zs_frontend* fe, const uint8_t* left, const uint8_t* right, size_t pitch,
                                              size_t stride, const zs_frontend_results* res)
{
    ZS_REQUIRE(fe, "null argument");
    zs_status st;
    while (fe->submitted > fe->waited)
        if ((st = zs_frontend_wait(fe)) != ZS_OK) return st;
    if ((st = zs_frontend_submit_host(fe, left, right, pitch, stride, res)) != ZS_OK) return st;
    return zs_frontend_wait(fe);
}

// ------------------------------------------------------------------------------------------------------
// Pipelined host path.  The synchronous zs_frontend_process_host serialises H2D -> compute -> D2H; at C2 the
// copies cost as much as the kernels.  submit/wait keeps two batches in flight on three streams so that the
// PCIe transfers of the neighbouring batches hide behind the kernels of the current one.
// ------------------------------------------------------------------------------------------------------

// staging (raw host layout) -> level-0 interiors of the padded planes; 16 bytes per thread when everything
// is 16-byte aligned (C2: width 752, pitch 752), bytes otherwise.  Items = (row, 16-byte group) flattened, so that a
// block is full whatever the row length (a block per row: 47 of 128 threads busy at 752 pixels, 123 k blocks per batch).
// grid: (ceil(h * ceil(w/16) / 256), 1, images)
__global__ void __launch_bounds__(256) k_unpack_level0(const uint8_t* __restrict__ src, size_t pitch, size_t stride,
                                                       zs_pyr_view v, int first, int vec)
{
    const int w = v.w[0], h = v.h[0];
    const int wg = (w + 15) >> 4;
    const int item = blockIdx.x * 256 + threadIdx.x;
    const int y = item / wg, x = (item - y * wg) * 16, img = blockIdx.z;
    if (y >= h) return;
    const uint8_t* s = src + (size_t)img * stride + (size_t)y * pitch;
    uint8_t* d = v.img[0] + (size_t)zs_slot(first, img, v.slots) * v.slot_stride[0] + (size_t)(v.pad_y + y) * v.pitch[0] + v.pad_x;
    if (vec && x + 16 <= w) {
        *(uint4*)(d + x) = __ldcs((const uint4*)(s + x));
    } else {
        for (int k = 0; k < 16 && x + k < w; ++k) d[x + k] = s[x + k];
    }
}

static zs_status pipeline_init(zs_frontend* fe, size_t stage_bytes)
{
    if (!fe->s_in) {
        ZS_CUDA(cudaStreamCreateWithFlags(&fe->s_in, cudaStreamNonBlocking));
        ZS_CUDA(cudaStreamCreateWithFlags(&fe->s_out, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            ZS_CUDA(cudaEventCreateWithFlags(&fe->ev_in[b], cudaEventDisableTiming));
            ZS_CUDA(cudaEventCreateWithFlags(&fe->ev_unpacked[b], cudaEventDisableTiming));
            ZS_CUDA(cudaEventCreateWithFlags(&fe->ev_run[b], cudaEventDisableTiming));
            ZS_CUDA(cudaEventCreateWithFlags(&fe->ev_out[b], cudaEventDisableTiming));
        }
        fe->outbox_bytes = al256(zs_frontend_d2h_bytes(fe)) + 16 * 256;
        for (int b = 0; b < 2; ++b) ZS_CUDA(cudaMalloc((void**)&fe->outbox[b], fe->outbox_bytes));
    }
    if (stage_bytes > fe->stage_bytes) {
        ZS_CUDA(cudaStreamSynchronize(fe->s_in));
        ZS_CUDA(cudaStreamSynchronize(fe->ctx->stream));
        for (int b = 0; b < 2; ++b) {
            if (fe->stage[b]) ZS_CUDA(cudaFree(fe->stage[b]));
            fe->stage[b] = nullptr;
            ZS_CUDA(cudaMalloc((void**)&fe->stage[b], stage_bytes));
        }
        fe->stage_bytes = stage_bytes;
    }
    return ZS_OK;
}

extern "C" zs_status zs_frontend_wait(zs_frontend* fe)
{
    ZS_REQUIRE(fe, "null argument");
    ZS_REQUIRE(fe->waited < fe->submitted, "zs_frontend_wait: nothing in flight");
    const int b = (int)(fe->waited & 1);
    ZS_CUDA(cudaEventSynchronize(fe->ev_out[b]));
    fe->busy[b] = false;
    fe->waited++;
    return ZS_OK;
}

extern "C" int zs_frontend_in_flight(const zs_frontend* fe) { return fe ? (int)(fe->submitted - fe->waited) : 0; }

extern "C" zs_status zs_frontend_submit_host(zs_frontend* fe, const uint8_t* left, const uint8_t* right, size_t pitch,
                                             size_t stride, const zs_frontend_results* r)
{
    ZS_REQUIRE(fe && left && right && r, "null argument");
    ZS_REQUIRE(r->cap == fe->cap, "results.cap must equal zs_frontend_capacity()");
    const size_t ch = fe->pp_enabled ? (size_t)fe->pp_channels : 1;
    ZS_REQUIRE(pitch >= ch * (size_t)fe->opt.width && stride >= pitch * (size_t)fe->opt.height, "bad pitch/stride");
    ZS_REQUIRE(fe->submitted - fe->waited < 2, "two batches already in flight: call zs_frontend_wait first");
    zs_context* ctx = fe->ctx;
    ZS_CUDA(cudaSetDevice(ctx->device));
    const size_t B = fe->B, cap = fe->cap;
    const size_t half = B * stride;
    zs_status st = pipeline_init(fe, 2 * half);
    if (st != ZS_OK) return st;
    const int b = (int)(fe->submitted & 1);

    // copy-in stream: the staging buffer is free once the unpack of the batch that last used it has run
    ZS_CUDA(cudaStreamWaitEvent(fe->s_in, fe->ev_unpacked[b], 0));
    ZS_CUDA(cudaMemcpyAsync(fe->stage[b], left, half, cudaMemcpyHostToDevice, fe->s_in));
    ZS_CUDA(cudaMemcpyAsync(fe->stage[b] + half, right, half, cudaMemcpyHostToDevice, fe->s_in));
    ZS_CUDA(cudaEventRecord(fe->ev_in[b], fe->s_in));

    // compute stream
    ZS_CUDA(cudaStreamWaitEvent(ctx->stream, fe->ev_in[b], 0));
    const zs_pyr_view& v = fe->pyr->v;
    const int W = fe->opt.width, H = fe->opt.height;
    if (!fe->pp_enabled) {
        const int vec = (pitch % 16 == 0 && stride % 16 == 0 && ((uintptr_t)fe->stage[b] % 16) == 0 && W % 16 == 0) ? 1 : 0;
        const dim3 grid(zs_div_up(zs_div_up(W, 16) * H, 256), 1, (unsigned)(2 * B));
        k_unpack_level0<<<grid, 256, 0, ctx->stream>>>(fe->stage[b], pitch, stride, v, 2, vec);   // slots 2..2B+1 = L then R
        ZS_LAUNCH_CHECK(ctx);
    } else {
        // raw frames -> gray -> (CLAHE) -> (remap) -> level 0 of slots 2..2B+1; every stage handles all 2B images at once
        const uint8_t* cur = fe->stage[b]; size_t cp = pitch, cs = stride;
        const size_t gp = fe->pp_pitch, gs = gp * (size_t)H;
        uint8_t* l0; size_t l0p, l0s;
        if ((st = zs_pyramid_level0(fe->pyr, 2, &l0, &l0p, &l0s)) != ZS_OK) return st;
        int next_tmp = 0;
        const bool remap = fe->pp_maps != nullptr;
        if (fe->pp_channels == 3) {
            const bool last = !fe->pp_clahe && !remap;
            uint8_t* d = last ? l0 : fe->pp_tmp[next_tmp]; const size_t dp = last ? l0p : gp, ds = last ? l0s : gs;
            if ((st = zs_cvt_bgr2gray(ctx, cur, cp, cs, W, H, (int)(2 * B), d, dp, ds)) != ZS_OK) return st;
            cur = d; cp = dp; cs = ds; next_tmp ^= 1;
        }
        if (fe->pp_clahe) {
            const bool last = !remap;
            uint8_t* d = last ? l0 : fe->pp_tmp[next_tmp]; const size_t dp = last ? l0p : gp, ds = last ? l0s : gs;
            if ((st = zs_clahe(ctx, cur, cp, cs, W, H, (int)(2 * B), fe->pp_clip, 8, 8, d, dp, ds)) != ZS_OK) return st;
            cur = d; cp = dp; cs = ds; next_tmp ^= 1;
        }
        if (remap) {
            const size_t px = (size_t)W * H;
            for (int cam = 0; cam < 2; ++cam)
                if ((st = zs_remap_linear(ctx, cur + (size_t)cam * B * cs, cp, cs, W, H, (int)B, fe->pp_maps + (size_t)(2 * cam) * px,
                                          fe->pp_maps + (size_t)(2 * cam + 1) * px, (size_t)W, 0, W, H, l0 + (size_t)cam * B * l0s, l0p, l0s)) != ZS_OK)
                    return st;
        } else if (cur == fe->stage[b]) {
            // gray input, no CLAHE, no remap: plain unpack
            const int vec = (pitch % 16 == 0 && stride % 16 == 0 && ((uintptr_t)fe->stage[b] % 16) == 0 && W % 16 == 0) ? 1 : 0;
            const dim3 grid(zs_div_up(zs_div_up(W, 16) * H, 256), 1, (unsigned)(2 * B));
            k_unpack_level0<<<grid, 256, 0, ctx->stream>>>(fe->stage[b], pitch, stride, v, 2, vec);
            ZS_LAUNCH_CHECK(ctx);
        }
    }
    ZS_CUDA(cudaEventRecord(fe->ev_unpacked[b], ctx->stream));
    if ((st = zs_frontend_run(fe)) != ZS_OK) return st;
    // the outbox is free once the D2H of the batch that last used it has finished
    ZS_CUDA(cudaStreamWaitEvent(ctx->stream, fe->ev_out[b], 0));
    struct part { const void* src; void* dst; size_t bytes; };
    const part parts[14] = {
        { fe->n + 2, r->n_left, sizeof(int) * B }, { fe->n + B + 2, r->n_right, sizeof(int) * B },
        { fe->xy + 2 * cap * 2, r->kp_left, sizeof(float) * 2 * B * cap }, { fe->xy + (B + 2) * cap * 2, r->kp_right, sizeof(float) * 2 * B * cap },
        { fe->resp + 2 * cap, r->resp_left, sizeof(float) * B * cap }, { fe->resp + (B + 2) * cap, r->resp_right, sizeof(float) * B * cap },
        { fe->desc + 2 * cap * 32, r->desc_left, B * cap * 32 }, { fe->desc + (B + 2) * cap * 32, r->desc_right, B * cap * 32 },
        { fe->m_idx, r->match_idx, sizeof(int) * 2 * B * cap }, { fe->m_dist, r->match_dist, sizeof(float) * 2 * B * cap },
        { fe->m_pass, r->match_pass, B * cap },
        { fe->t_pts, r->track_pts, sizeof(float) * 2 * 4 * B * cap }, { fe->t_keep, r->track_keep, 4 * B * cap },
        { fe->t_n, r->track_n, sizeof(int) * 4 * B } };
    size_t off[14], o = 0;
    for (int i = 0; i < 14; ++i) {
        off[i] = o;
        if (parts[i].dst) {
            ZS_CUDA(cudaMemcpyAsync(fe->outbox[b] + o, parts[i].src, parts[i].bytes, cudaMemcpyDeviceToDevice, ctx->stream));
            o += al256(parts[i].bytes);
        }
    }
    ZS_CUDA(cudaEventRecord(fe->ev_run[b], ctx->stream));

    // copy-out stream
    ZS_CUDA(cudaStreamWaitEvent(fe->s_out, fe->ev_run[b], 0));
    for (int i = 0; i < 14; ++i)
        if (parts[i].dst)
            ZS_CUDA(cudaMemcpyAsync(parts[i].dst, fe->outbox[b] + off[i], parts[i].bytes, cudaMemcpyDeviceToHost, fe->s_out));
    ZS_CUDA(cudaEventRecord(fe->ev_out[b], fe->s_out));
    fe->busy[b] = true;
    fe->submitted++;
    return ZS_OK;
}

// Raw camera frames in: the image path of processor::process (zenslam_core/source/processor.cpp:25-55) runs on the device
// in front of the pyramid build.  channels 3 = BGR (utils::convert_color), clahe_enabled = detection.clahe_enabled with
// cv::createCLAHE(clahe_clip_limit) (processor.h:38), maps = calibration.map_x / map_y per camera (CV_32FC1, W*H floats
// each, HOST pointers, all four or none).  Applies to zs_frontend_submit_host / zs_frontend_process_host.
extern "C" zs_status zs_frontend_set_preprocess(zs_frontend* fe, int channels, int clahe_enabled, double clahe_clip_limit,
                                                const float* map_x_left, const float* map_y_left, const float* map_x_right,
                                                const float* map_y_right)
{
    ZS_REQUIRE(fe, "null argument");
    ZS_REQUIRE(channels == 1 || channels == 3, "channels must be 1 (gray) or 3 (BGR)");
    const int nmaps = (map_x_left != nullptr) + (map_y_left != nullptr) + (map_x_right != nullptr) + (map_y_right != nullptr);
    ZS_REQUIRE(nmaps == 0 || nmaps == 4, "give all four rectification maps or none");
    ZS_REQUIRE(fe->submitted == fe->waited, "batches in flight: call zs_frontend_wait first");
    zs_context* ctx = fe->ctx;
    ZS_CUDA(cudaSetDevice(ctx->device));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    const size_t px = (size_t)fe->opt.width * fe->opt.height;
    if (fe->pp_maps) { ZS_CUDA(cudaFree(fe->pp_maps)); fe->pp_maps = nullptr; }
    if (nmaps == 4) {
        ZS_CUDA(cudaMalloc((void**)&fe->pp_maps, sizeof(float) * 4 * px));
        const float* m[4] = { map_x_left, map_y_left, map_x_right, map_y_right };
        for (int i = 0; i < 4; ++i) ZS_CUDA(cudaMemcpy(fe->pp_maps + (size_t)i * px, m[i], sizeof(float) * px, cudaMemcpyHostToDevice));
    }
    fe->pp_pitch = ((size_t)fe->opt.width + 15) / 16 * 16;
    for (int i = 0; i < 2; ++i)
        if (!fe->pp_tmp[i]) ZS_CUDA(cudaMalloc((void**)&fe->pp_tmp[i], fe->pp_pitch * fe->opt.height * 2 * (size_t)fe->B));
    fe->pp_channels = channels; fe->pp_clahe = clahe_enabled ? 1 : 0; fe->pp_clip = clahe_clip_limit;
    fe->pp_enabled = (channels == 3 || clahe_enabled || nmaps == 4) ? 1 : 0;
    return ZS_OK;
}
