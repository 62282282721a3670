// zs_fast.cu -- FAST-9-16 + score + 3x3 NMS, per grid cell (keypoint_detector_grid) and full frame
// (keypoint_detector_simple).
//
// Reference: zenslam_core/source/detection/keypoint_detector_grid.cpp:39-120 (cells, first-max per cell,
// row-major order) over cv::FastFeatureDetector (keypoint_detector_grid.cpp:15,90); semantics in SURVEY A.1.
//
// Grid kernel: one warp per cell.  The cell ROI is staged in shared memory with coalesced loads, a cheap
// 16-bit ring-mask test finds the (few) corners, those are compacted with a ballot and scored densely
// (one lane per corner), NMS and the first-maximum selection run on the shared score tile.  Integer only.
// HBM traffic: each image byte is read once (cells never overlap) + 16 B per cell written.
#include <stdlib.h>

#include "zs_common.cuh"
#include "zs_fast_core.cuh"

#define FAST_WARPS 8

struct fast_grid_args {
    zs_pyr_view v;
    int first, count;
    int cw, ch, gw, gh, threshold;
    const uint8_t* occupied;     // [count][gh*gw] or null
    // per-cell candidates: [count][gh*gw]
    int* cand;                   // packed: (score << 20) | (y_in_cell << 10) | x_in_cell, or -1
    const void* strip_map;       // k_fast_grid_v2: TMA descriptor of this variant's strip box, or null (vector loads)
};

// grid: (ceil(cells / FAST_WARPS), count); dynamic smem = FAST_WARPS * (tile + score tile + corner list)
__global__ void __launch_bounds__(FAST_WARPS * 32) k_fast_grid(fast_grid_args a)
{
    extern __shared__ uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cells = a.gw * a.gh;
    const int cell = blockIdx.x * FAST_WARPS + warp;
    if (cell >= cells) return;
    const int img = blockIdx.y;
    int* cand = a.cand + (size_t)img * cells + cell;
    if (a.occupied && a.occupied[(size_t)img * cells + cell]) { if (lane == 0) *cand = -1; return; }

    const int tp = (a.cw + 3) & ~3;                        // tile pitch
    const int tile_bytes = tp * a.ch;
    const int list_cap = (a.cw - 6) * (a.ch - 6);
    uint8_t* tile = smem + (size_t)warp * (2 * tile_bytes + 2 * ((list_cap + 1) & ~1));
    uint8_t* score = tile + tile_bytes;
    uint16_t* list = (uint16_t*)(score + tile_bytes);

    const int gx = cell % a.gw, gy = cell / a.gw;
    const int slot = zs_slot(a.first, img, a.v.slots);
    const int pitch = a.v.pitch[0];
    const uint8_t* src = a.v.img[0] + (size_t)slot * a.v.slot_stride[0] + (size_t)(a.v.pad_y + gy * a.ch) * pitch
                         + a.v.pad_x + gx * a.cw;

    // stage the cell (the reference clips cells to the image, but with integer-division grids every cell
    // is full-size: keypoint_detector_grid.cpp:42,79-85)
    if ((a.cw & 3) == 0) {
        const int wpr = a.cw >> 2;                          // words per row; src is 4-byte aligned
        for (int i = lane; i < wpr * a.ch; i += 32) {
            const int r = i / wpr, c = i - r * wpr;
            ((uint32_t*)(tile + r * tp))[c] = *(const uint32_t*)(src + (size_t)r * pitch + 4 * c);
        }
    } else {
        for (int i = lane; i < a.cw * a.ch; i += 32) {
            const int r = i / a.cw, c = i - r * a.cw;
            tile[r * tp + c] = src[(size_t)r * pitch + c];
        }
    }
    for (int i = lane; i < tile_bytes / 4; i += 32) ((uint32_t*)score)[i] = 0;
    __syncwarp();

    // pass 1: ring-mask test over the interior, compact corners
    const int iw = a.cw - 6, ih = a.ch - 6;
    const int npx = iw * ih;
    int ncorners = 0;
    int d[16];
    for (int base = 0; base < npx; base += 32) {
        const int i = base + lane;
        bool corner = false;
        int x = 0, y = 0;
        if (i < npx) {
            y = i / iw; x = i - y * iw; x += 3; y += 3;
            corner = fast_is_corner(tile, tp, x, y, a.threshold, d);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, corner);
        if (corner) list[ncorners + __popc(m & ((1u << lane) - 1))] = (uint16_t)(y * a.cw + x);
        ncorners += __popc(m);
    }
    if (ncorners == 0) { if (lane == 0) *cand = -1; return; }
    __syncwarp();

    // pass 2: score the corners densely
    for (int base = 0; base < ncorners; base += 32) {
        const int i = base + lane;
        if (i < ncorners) {
            const int p = list[i], y = p / a.cw, x = p - y * a.cw;
            fast_is_corner(tile, tp, x, y, a.threshold, d);
            score[y * tp + x] = (uint8_t)fast_score(d);
        }
    }
    __syncwarp();

    // pass 3: NMS (strictly greater than all 8 neighbours) + first maximum in raster order
    unsigned best = 0;     // key = (score + 1) << 20 | (0xfffff - raster); 0 = none
    for (int base = 0; base < ncorners; base += 32) {
        const int i = base + lane;
        if (i < ncorners) {
            const int p = list[i], y = p / a.cw, x = p - y * a.cw;
            const int s = score[y * tp + x];
            const uint8_t* r0 = score + (y - 1) * tp + x;
            const uint8_t* r1 = r0 + tp; const uint8_t* r2 = r1 + tp;
            const bool keep = s > r0[-1] && s > r0[0] && s > r0[1] && s > r1[-1] && s > r1[1] && s > r2[-1] && s > r2[0] &&
                              s > r2[1];
            if (keep) best = max(best, ((unsigned)(s + 1) << 20) | (0xfffffu - (unsigned)p));
        }
    }
    best = __reduce_max_sync(0xffffffffu, best);
    if (lane == 0) {
        if (best == 0) *cand = -1;
        else {
            const int s = (int)(best >> 20) - 1;
            const int p = 0xfffff - (int)(best & 0xfffffu);
            const int y = p / a.cw, x = p - y * a.cw;
            *cand = (s << 20) | (y << 10) | x;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// v2 grid kernel (even cell sizes, e.g. the reference's default 16x16 and the 32x32 of the large-frame config).
// A block owns NC consecutive cells of one cell row (16 cells of 16 px, 4 cells of 32 px: 256 / 128 px strips).  The strip is staged once with 16-byte loads and
// expanded into two u16x2 "pair planes" in shared memory (E holds pixel pairs (0,1),(2,3).., O holds (1,2),(3,4)..),
// so that the two horizontally adjacent interior pixels a thread owns see every ring position as ONE aligned
// 32-bit shared load.  The FAST decision and score then need no per-pixel branching at all:
//     dark arc  : all 9 ring pixels < c - t   <=>   min over arcs of (max over the arc) < c - t
//     bright arc: all 9 ring pixels > c + t   <=>   max over arcs of (min over the arc) > c + t
//     s' = max(c - A, B - c), corner <=> s' > t, response = s' - 1                 (same integers as v1)
// and the sliding 9-window min / max over the 16-ring is 2 x 32 three-input VIMNMX3.U16x2 (two pixels per
// instruction) plus two 16 -> 1 reductions.  NMS and the per-cell first maximum run on a u8 score tile.
// ------------------------------------------------------------------------------------------------------
#define FG2_THREADS 160

template <bool MAX>
__device__ __forceinline__ unsigned mm3(unsigned a, unsigned b, unsigned c) { return MAX ? __vimax3_u16x2(a, b, c) : __vimin3_u16x2(a, b, c); }

// reduce over all 16 circular 9-windows: MAX9 = false: max_k min(window_k)  (bright side, B);  true: min_k max(window_k)  (dark side, A)
template <bool MAX9>
__device__ __forceinline__ unsigned ring_window9(const unsigned p[16])
{
    unsigned a[16], b[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = mm3<MAX9>(p[k], p[(k + 1) & 15], p[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) b[k] = mm3<MAX9>(a[k], a[(k + 3) & 15], a[(k + 6) & 15]);
    unsigned r[6];
#pragma unroll
    for (int k = 0; k < 5; ++k) r[k] = mm3<!MAX9>(b[3 * k], b[3 * k + 1], b[3 * k + 2]);
    r[5] = b[15];
    const unsigned u = mm3<!MAX9>(r[0], r[1], r[2]), w = mm3<!MAX9>(r[3], r[4], r[5]);
    return MAX9 ? __vminu2(u, w) : __vmaxu2(u, w);
}

// TMA: one elected thread asks for the block's whole strip -- CH rows of NC cells plus 16 bytes of slack, a (TW + 16) / 4 x CH x 1
// box of u32 elements of the padded level-0 plane -- and every thread waits on the mbarrier; no load instruction, address
// arithmetic or register is spent on staging.  (Columns beyond the image's last cell then hold plane bytes instead of zeros;
// they only feed cells this block does not own.)
__device__ __forceinline__ uint32_t fg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fg_mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

template <int CW, int CH, int NC>
__global__ void __launch_bounds__(FG2_THREADS) k_fast_grid_v2(fast_grid_args a)
{
    constexpr int TW = NC * CW;            // strip width in pixels
    constexpr int PW = TW / 2 + 4;                // words per pair-plane row (+ slack for the x+4 reads at the right edge)
    constexpr int IW = CW - 6, IH = CH - 6;       // interior (tested) pixels per cell
    constexpr int PPR = IW / 2;                   // pixel pairs per cell row (IW is even)
    constexpr int ROWP = NC * PPR;         // pairs per strip row
    constexpr int TOTAL = ROWP * IH;
    static_assert(CW % 4 == 0 && CW >= 8 && CH >= 7, "cell shape");
    constexpr int RP = TW + 16;
    static_assert((CH * RP) % 128 == 0, "raw strip keeps the arrays behind it aligned");
    extern __shared__ __align__(128) uint8_t fsm[];
    __shared__ __align__(8) uint64_t s_bar;
    uint8_t* sRaw = fsm;                                            // raw strip [CH][TW + 16]: the TMA box (128-byte aligned)
    uint32_t* sE = (uint32_t*)(sRaw + CH * RP);                     // [CH][PW]
    uint32_t* sO = sE + CH * PW;                                    // [CH][PW]
    uint8_t* sS = (uint8_t*)(sO + CH * PW);                         // score tile [CH][TW]
    int* sBest = (int*)(sS + CH * TW);                              // [NC]

    const int img = blockIdx.z, gy = blockIdx.y, cell0 = blockIdx.x * NC;
    const int ncell = min(NC, a.gw - cell0);
    const int tid = threadIdx.x;
    const int slot = zs_slot(a.first, img, a.v.slots);
    const int pitch = a.v.pitch[0];
    const uint8_t* src = a.v.img[0] + (size_t)slot * a.v.slot_stride[0] + (size_t)(a.v.pad_y + gy * CH) * pitch + a.v.pad_x + cell0 * CW;
    const int valid_w = ncell * CW;               // pixels of this strip that belong to cells

    // ---- stage the raw strip: one TMA box, or 16-byte loads (interiors are 16-byte aligned: pad_x, pitch and CW*NC are
    // multiples of 16)
    const bool tma = a.strip_map != nullptr;
    if (tma) {
        const uint32_t bar = fg_smem_u32(&s_bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(CH * RP) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(fg_smem_u32(sRaw)), "l"(a.strip_map), "r"((a.v.pad_x + cell0 * CW) >> 2), "r"(a.v.pad_y + gy * CH),
                           "r"(slot), "r"(bar) : "memory");
        }
    } else {
        for (int i = tid; i < CH * (RP / 16); i += FG2_THREADS) {
            const int r = i / (RP / 16), c = (i - r * (RP / 16)) * 16;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (c < valid_w) v = *(const uint4*)(src + (size_t)r * pitch + c);
            *(uint4*)(sRaw + r * RP + c) = v;
        }
    }
    for (int i = tid; i < CH * TW / 4; i += FG2_THREADS) ((uint32_t*)sS)[i] = 0;
    if (tid < NC) sBest[tid] = 0;
    __syncthreads();                                               // (also orders the barrier's initialisation before the waits)
    if (tma) fg_mbar_wait(fg_smem_u32(&s_bar), 0);
    // ---- expand into the pair planes: word j of E = (p[2j], p[2j+1]), of O = (p[2j+1], p[2j+2]) as u16x2
    for (int i = tid; i < CH * (TW / 4 + 1); i += FG2_THREADS) {
        const int r = i / (TW / 4 + 1), q = i - r * (TW / 4 + 1);
        const uint32_t w = *(const uint32_t*)(sRaw + r * RP + 4 * q);
        const uint32_t nb = sRaw[r * RP + 4 * q + 4];
        uint32_t* e = sE + r * PW + 2 * q; uint32_t* o = sO + r * PW + 2 * q;
        e[0] = __byte_perm(w, 0, 0x4140); e[1] = __byte_perm(w, 0, 0x4342);
        o[0] = __byte_perm(w, 0, 0x4241); o[1] = (w >> 24) | (nb << 16);
    }
    __syncthreads();

    // ---- scores: one thread per pair of horizontally adjacent interior pixels
    const unsigned tt = (unsigned)a.threshold * 0x10001u;
    for (int idx = tid; idx < TOTAL; idx += FG2_THREADS) {
        const int ry = idx / ROWP, rem = idx - ry * ROWP;
        const int cell = rem / PPR, pr = rem - cell * PPR;
        if (cell >= ncell) continue;
        const int y = ry + 3, x = cell * CW + 3 + 2 * pr;          // x is odd: (x+dx) odd -> plane O, even -> plane E
        const uint32_t* eb = sE + y * PW + (x >> 1);               // word holding pair (x-1, x)
        const uint32_t* ob = sO + y * PW + (x >> 1);               // word holding pair (x, x+1)
        unsigned p[16];
        // ring order of SURVEY A.1: (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
        p[0] = ob[3 * PW];      p[1] = eb[3 * PW + 1];  p[2] = ob[2 * PW + 1];  p[3] = eb[1 * PW + 2];
        p[4] = eb[2];           p[5] = eb[-1 * PW + 2]; p[6] = ob[-2 * PW + 1]; p[7] = eb[-3 * PW + 1];
        p[8] = ob[-3 * PW];     p[9] = eb[-3 * PW];     p[10] = ob[-2 * PW - 1]; p[11] = eb[-1 * PW - 1];
        p[12] = eb[-1];         p[13] = eb[1 * PW - 1]; p[14] = ob[2 * PW - 1];  p[15] = eb[3 * PW];
        const unsigned c = ob[0];
        const unsigned A = ring_window9<true>(p);                  // min over arcs of the arc maximum
        const unsigned B = ring_window9<false>(p);                 // max over arcs of the arc minimum
        // s' = max(c - A, B - c) per half, saturating at 0 (a negative side can never win against the threshold)
        const unsigned sd = __vsubus2(c, A), sb = __vsubus2(B, c);
        const unsigned sp = __vmaxu2(sd, sb);
        const unsigned corner = __vcmpgtu2(sp, tt);                // 0xffff per half where s' > t
        const unsigned sc = __vadd2(sp, 0xffffffffu) & corner;      // response s' - 1 per half (s' >= 1 wherever corner is set)
        sS[y * TW + x] = (uint8_t)(sc & 0xff);
        sS[y * TW + x + 1] = (uint8_t)(sc >> 16);
    }
    __syncthreads();

    // ---- NMS (strictly greater than all 8 neighbours; untested border pixels score 0) + first maximum per cell
    for (int idx = tid; idx < TOTAL; idx += FG2_THREADS) {
        const int ry = idx / ROWP, rem = idx - ry * ROWP;
        const int cell = rem / PPR, pr = rem - cell * PPR;
        if (cell >= ncell) continue;
        const int y = ry + 3, xc = 3 + 2 * pr, x = cell * CW + xc;
        const uint8_t* r1 = sS + y * TW + x;
        const unsigned m0 = *(const uint16_t*)(r1 - 1), m1 = *(const uint16_t*)(r1 + 1);        // (x-1,x) (x+1,x+2)
        const int s0 = m0 >> 8, s1 = m1 & 0xff;
        if ((s0 | s1) == 0) continue;
        const unsigned t0 = *(const uint16_t*)(r1 - TW - 1), t1 = *(const uint16_t*)(r1 - TW + 1);
        const unsigned b0 = *(const uint16_t*)(r1 + TW - 1), b1 = *(const uint16_t*)(r1 + TW + 1);
        // neighbours of pixel x: columns x-1, x, x+1 of the rows above/below, x-1 and x+1 of its own row
        const int n0 = max(max(max((int)(t0 & 0xff), (int)(t0 >> 8)), max((int)(t1 & 0xff), (int)(b0 & 0xff))),
                           max(max((int)(b0 >> 8), (int)(b1 & 0xff)), max((int)(m0 & 0xff), s1)));
        const int n1 = max(max(max((int)(t0 >> 8), (int)(t1 & 0xff)), max((int)(t1 >> 8), (int)(b0 >> 8))),
                           max(max((int)(b1 & 0xff), (int)(b1 >> 8)), max(s0, (int)(m1 >> 8))));
        // the pair's second pixel is later in raster order, so on equal scores the first one must win: keys carry 0xfffff - raster
        if (s0 > n0) atomicMax(&sBest[cell], ((s0 + 1) << 20) | (0xfffff - (y * CW + xc)));
        if (s1 > n1) atomicMax(&sBest[cell], ((s1 + 1) << 20) | (0xfffff - (y * CW + xc + 1)));
    }
    __syncthreads();
    if (tid < ncell) {
        const int cellg = gy * a.gw + cell0 + tid;
        int out = -1;
        const int best = sBest[tid];
        if (best != 0 && !(a.occupied && a.occupied[(size_t)img * a.gw * a.gh + cellg])) {
            const int s = (best >> 20) - 1, pz = 0xfffff - (best & 0xfffff);
            const int yy = pz / CW, xx = pz - yy * CW;
            out = (s << 20) | (yy << 10) | xx;
        }
        a.cand[(size_t)img * a.gw * a.gh + cellg] = out;
    }
}

template <int CW, int CH, int NC>
static size_t fast_grid_v2_smem()
{
    constexpr int TW = NC * CW, PW = TW / 2 + 4;
    return (size_t)2 * CH * PW * 4 + (size_t)CH * TW + ((NC + 3) & ~3) * 4 + (size_t)CH * (TW + 16);
}

// Compaction in cell row-major order: one block per image, block-wide exclusive scan over the cells.
__global__ void __launch_bounds__(1024) k_grid_compact(const int* __restrict__ cand, int cells, int gw, int cw, int ch,
                                                       float2* __restrict__ oxy,
                                                       float* __restrict__ oresp, int* __restrict__ ocount, int cap)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int img = blockIdx.x;
    const int* c = cand + (size_t)img * cells;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < cells; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = (i < cells) ? c[i] : -1;
        const int f = v >= 0;
        const uint32_t m = __ballot_sync(0xffffffffu, f);
        const int wpre = __popc(m & ((1u << lane) - 1));
        if (lane == 0) warp_sums[warp] = __popc(m);
        __syncthreads();
        int woff = 0, total = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { const int s = warp_sums[k]; if (k < warp) woff += s; total += s; }
        const int pos = carry + woff + wpre;
        if (f && pos < cap) {
            const int gx = i % gw, gy = i / gw;
            oxy[(size_t)img * cap + pos] = make_float2((float)(gx * cw + (v & 1023)), (float)(gy * ch + ((v >> 10) & 1023)));
            oresp[(size_t)img * cap + pos] = (float)(v >> 20);
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) ocount[img] = min(carry, cap);
}

extern "C" zs_status zs_fast_grid_detect(zs_context* ctx, const zs_pyramid* p, int first, int count, int cell_w,
                                         int cell_h, int threshold, const uint8_t* d_occupied, float* d_xy,
                                         float* d_response, int* d_count, int cap)
{
    ZS_REQUIRE(ctx && p && d_xy && d_response && d_count, "null argument");
    ZS_REQUIRE(count >= 0 && count <= p->slots && first >= 0, "bad slot range");
    ZS_REQUIRE(cell_w >= 7 && cell_h >= 7 && cell_w <= 256 && cell_h <= 256, "cell size must be within 7..256");
    const int gw = p->width / cell_w, gh = p->height / cell_h;
    const int cells = gw * gh;
    ZS_REQUIRE(cap >= cells || cells == 0, "cap < number of cells");
    if (count == 0) return ZS_OK;
    if (cells == 0) { ZS_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int) * count, ctx->stream)); return ZS_OK; }
    threshold = threshold < 0 ? 0 : threshold > 255 ? 255 : threshold;
    void* scratch;
    zs_status st = zs_scratch(ctx, sizeof(int) * (size_t)cells * count, &scratch);
    if (st != ZS_OK) return st;
    fast_grid_args a;
    a.v = p->v; a.first = first; a.count = count; a.cw = cell_w; a.ch = cell_h; a.gw = gw; a.gh = gh;
    a.threshold = threshold; a.occupied = d_occupied; a.cand = (int*)scratch; a.strip_map = nullptr;
    if (!ctx->sw.fast_v1 && ((cell_w == 16 && cell_h == 16) || (cell_w == 32 && cell_h == 32) || (cell_w == 64 && cell_h == 64))) {
        const char* maps = (ctx->sw.fast_no_tma || !p->v.fast_maps) ? nullptr : (const char*)p->v.fast_maps;
        if (cell_w == 16) {
            a.strip_map = maps;
            const size_t smem = fast_grid_v2_smem<16, 16, 16>();
            k_fast_grid_v2<16, 16, 16><<<dim3(zs_div_up(gw, 16), gh, count), FG2_THREADS, smem, ctx->stream>>>(a);
        } else if (cell_w == 64) {
            // the shipped configuration (tumvi.yaml:38: cell_size [64, 64]): one 64x64 cell per block, 28 KB of planes
            a.strip_map = maps ? maps + 2 * 128 : nullptr;
            const size_t smem = fast_grid_v2_smem<64, 64, 1>();
            k_fast_grid_v2<64, 64, 1><<<dim3(gw, gh, count), FG2_THREADS, smem, ctx->stream>>>(a);
        } else {
            // 4 cells (128 px) per block keep the pair planes at 26 KB, so 8 blocks fit an SM instead of 2
            a.strip_map = maps ? maps + 1 * 128 : nullptr;
            const size_t smem = fast_grid_v2_smem<32, 32, 4>();
            k_fast_grid_v2<32, 32, 4><<<dim3(zs_div_up(gw, 4), gh, count), FG2_THREADS, smem, ctx->stream>>>(a);
        }
        ZS_LAUNCH_CHECK(ctx);
        k_grid_compact<<<count, 1024, 0, ctx->stream>>>(a.cand, cells, gw, cell_w, cell_h, (float2*)d_xy, d_response, d_count, cap);
        ZS_LAUNCH_CHECK(ctx);
        return ZS_OK;
    }
    const int tp = (cell_w + 3) & ~3;
    const int list_cap = (cell_w - 6) * (cell_h - 6);
    const size_t smem = (size_t)FAST_WARPS * (2 * tp * cell_h + 2 * ((list_cap + 1) & ~1));
    ZS_REQUIRE(smem <= 220 * 1024, "cell too large for the shared-memory tile");
    if (smem > 48 * 1024)
        ZS_CUDA(cudaFuncSetAttribute(k_fast_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_fast_grid<<<dim3(zs_div_up(cells, FAST_WARPS), count), FAST_WARPS * 32, smem, ctx->stream>>>(a);
    ZS_LAUNCH_CHECK(ctx);
    k_grid_compact<<<count, 1024, 0, ctx->stream>>>(a.cand, cells, gw, cell_w, cell_h, (float2*)d_xy, d_response, d_count, cap);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

// ------------------------------------------------------------------------------------------------------
// Full-frame FAST for keypoint_detector_simple (zenslam_core/source/detection/keypoint_detector_simple.cpp:38-63):
// cv::FAST(img, threshold, true) + mask, raster order.
//   k_fast_score_map   score (s-1, 0 = not a corner) for every pixel of the frame
//   k_fast_nms_rows    per row: NMS + mask -> (count pass) row counts / (write pass) compacted output
//   k_row_scan         exclusive scan of the row counts, one block per image
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fast_score_map(zs_pyr_view v, int first, int threshold, uint8_t* __restrict__ score)
{
    const int w = v.w[0], h = v.h[0], pitch = v.pitch[0];
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const int slot = zs_slot(first, blockIdx.z, v.slots);
    uint8_t out = 0;
    if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) {
        const uint8_t* img = v.img[0] + (size_t)slot * v.slot_stride[0] + (size_t)v.pad_y * pitch + v.pad_x;
        int d[16];
        if (fast_is_corner(img, pitch, x, y, threshold, d)) out = (uint8_t)fast_score(d);
    }
    score[((size_t)blockIdx.z * h + y) * w + x] = out;
}

// one warp per row; pass 0 counts, pass 1 writes
__global__ void __launch_bounds__(256) k_fast_nms_rows(const uint8_t* __restrict__ score, const uint8_t* __restrict__ mask,
                                                       int w, int h, int write, int* __restrict__ row_count,
                                                       const int* __restrict__ row_off, float2* __restrict__ oxy,
                                                       float* __restrict__ oresp, int cap)
{
    const int lane = threadIdx.x & 31;
    const int y = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int img = blockIdx.y;
    if (y >= h) return;
    int n = 0;
    if (y >= 3 && y < h - 3) {
        const uint8_t* r1 = score + ((size_t)img * h + y) * w;
        const uint8_t* r0 = r1 - w; const uint8_t* r2 = r1 + w;
        const int base_out = write ? row_off[(size_t)img * h + y] : 0;
        for (int x0 = 0; x0 < w; x0 += 32) {
            const int x = x0 + lane;
            bool keep = false;
            int s = 0;
            if (x >= 3 && x < w - 3) {
                s = r1[x];
                keep = s > 0 && s > r0[x - 1] && s > r0[x] && s > r0[x + 1] && s > r1[x - 1] && s > r1[x + 1] && s > r2[x - 1] &&
                       s > r2[x] && s > r2[x + 1];
                if (keep && mask) keep = mask[((size_t)img * h + y) * w + x] != 0;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (write && keep) {
                const int pos = base_out + n + __popc(m & ((1u << lane) - 1));
                if (pos < cap) {
                    oxy[(size_t)img * cap + pos] = make_float2((float)x, (float)y);
                    oresp[(size_t)img * cap + pos] = (float)s;
                }
            }
            n += __popc(m);
        }
    }
    if (!write && lane == 0) row_count[(size_t)img * h + y] = n;
}

__global__ void __launch_bounds__(1024) k_row_scan(const int* __restrict__ row_count, int h, int* __restrict__ row_off,
                                                   int* __restrict__ total)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int img = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < h; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < h ? row_count[(size_t)img * h + i] : 0;
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
        if (lane == 31) warp_sums[warp] = s;
        __syncthreads();
        int woff = 0, tot = 0;
        for (int k = 0; k < 32; ++k) { const int ws = warp_sums[k]; if (k < warp) woff += ws; tot += ws; }
        if (i < h) row_off[(size_t)img * h + i] = carry + woff + s - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) total[img] = carry;
}

extern "C" zs_status zs_fast_detect(zs_context* ctx, const zs_pyramid* p, int first, int count, int threshold,
                                    const uint8_t* d_mask, float* d_xy, float* d_response, int* d_count, int cap)
{
    ZS_REQUIRE(ctx && p && d_xy && d_response && d_count, "null argument");
    ZS_REQUIRE(count >= 0 && count <= p->slots && first >= 0 && cap > 0, "bad range");
    if (count == 0) return ZS_OK;
    threshold = threshold < 0 ? 0 : threshold > 255 ? 255 : threshold;
    const int w = p->width, h = p->height;
    const size_t plane = ((size_t)count * w * h + 255) / 256 * 256;
    void* s;
    zs_status st = zs_scratch(ctx, plane + sizeof(int) * 2 * (size_t)count * h, &s);
    if (st != ZS_OK) return st;
    uint8_t* score = (uint8_t*)s;
    int* row_count = (int*)(score + plane);
    int* row_off = row_count + (size_t)count * h;
    k_fast_score_map<<<dim3(zs_div_up(w, 32), zs_div_up(h, 8), count), 256, 0, ctx->stream>>>(p->v, first, threshold, score);
    ZS_LAUNCH_CHECK(ctx);
    k_fast_nms_rows<<<dim3(zs_div_up(h, 8), count), 256, 0, ctx->stream>>>(score, d_mask, w, h, 0, row_count, nullptr, nullptr,
                                                                          nullptr, cap);
    ZS_LAUNCH_CHECK(ctx);
    k_row_scan<<<count, 1024, 0, ctx->stream>>>(row_count, h, row_off, d_count);
    ZS_LAUNCH_CHECK(ctx);
    k_fast_nms_rows<<<dim3(zs_div_up(h, 8), count), 256, 0, ctx->stream>>>(score, d_mask, w, h, 1, nullptr, row_off, (float2*)d_xy,
                                                                          d_response, cap);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}
