// zs_pyramid.cu -- cv::buildOpticalFlowPyramid(img, pyr, win, maxLevel, withDerivatives=true) on the device
// (reference: utils::pyramid, zenslam_core/source/utils/utils_opencv.cpp:525-530).
//
// Per level: (1) fill the REFLECT_101 padding of the image plane, (2) pyrDown into the next level's
// interior, (3) Scharr (dx,dy) into the derivative plane interior (its padding stays zero).
// All three are streaming u8 stencils: 4 pixels per thread, 32-bit loads from the padded plane so no
// kernel needs border logic of its own.  HBM-bound; algorithmic bytes per image = sum over levels of
// (read w*h) + (write w*h image + 4*w*h derivative).
#include "zs_common.cuh"

// ---- (1) reflect padding ---------------------------------------------------------------------------
// The padding of a plane as flat 32-bit word items, one per thread (pad_x and the pitch are multiples of 16):
//   region A: the 2 pad_y rows above / below the image, whole padded rows (copies of the reflected interior row);
//   region B: the left and right pads of the h interior rows.
// A word whose four source pixels are consecutive and ascending is one aligned load from the interior row, otherwise
// four byte reads through REFLECT_101 -- the interior is all a pad item ever reads.  (One warp per four rows with a
// lane loop per row, the first version, was latency-bound: 39 us for the 21 MB of level-0 padding of 256 images.)
static int pad_items_host(const zs_pyr_view& v, int level)
{
    const int w = v.w[level], h = v.h[level];
    const int wtotal = (w + 2 * v.pad_x + 3) >> 2;
    return 2 * v.pad_y * wtotal + h * (wtotal - (w >> 2));
}

__device__ __forceinline__ void pad_item(const zs_pyr_view& v, int level, uint8_t* __restrict__ plane, int item)
{
    const int w = v.w[level], h = v.h[level], pitch = v.pitch[level];
    const int wpad = v.pad_x >> 2;                    // words per side pad
    const int wright0 = w >> 2;                       // first interior-relative word that contains right-pad pixels
    const int wtotal = (w + 2 * v.pad_x + 3) >> 2;
    const int na = 2 * v.pad_y * wtotal, nside = wtotal - wright0;
    int r, q;                                         // padded row, word of the padded row (columns 4q .. 4q+3)
    if (item < na) {
        const int ra = item / wtotal;
        q = item - ra * wtotal;
        r = ra < v.pad_y ? ra : h + ra;
    } else {
        const int it = item - na, y = it / nside, j = it - y * nside;
        if (y >= h) return;
        r = v.pad_y + y;
        q = j >= wpad ? j + wright0 : j;              // a partial last interior word keeps its interior bytes
    }
    const int sy = zs_reflect101(r - v.pad_y, h);
    const uint8_t* src = plane + (size_t)(sy + v.pad_y) * pitch + v.pad_x;   // interior row sy
    const int px0 = 4 * q - v.pad_x;                  // image column of the word's first pixel
    uint32_t val;
    if (px0 >= 0 && px0 + 4 <= w) {
        val = *(const uint32_t*)(src + px0);          // straight copy of four interior pixels of the reflected row
    } else {
        val = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) val |= (uint32_t)src[zs_reflect101(px0 + k, w)] << (8 * k);
    }
    *(uint32_t*)(plane + (size_t)r * pitch + 4 * q) = val;
}

// grid: (ceil(items / 256), 1, count)
__global__ void __launch_bounds__(256) k_pad_reflect(zs_pyr_view v, int level, int first)
{
    const int slot = zs_slot(first, blockIdx.z, v.slots);
    pad_item(v, level, v.img[level] + (size_t)slot * v.slot_stride[level], blockIdx.x * 256 + threadIdx.x);
}

// ---- (2) pyrDown -------------------------------------------------------------------------------------
// dst(x,y) = (sum_{i,j} k_i k_j src(2x+i-2, 2y+j-2) + 128) >> 8, k = [1 4 6 4 1]; src is the padded plane.
__device__ __forceinline__ void pd_row(const uint8_t* __restrict__ row, int h4[4])
{
    // row points at source column 2*x0 (multiple of 8 => 4-byte aligned loads at -4, 0, 4, 8)
    const uint32_t w0 = *(const uint32_t*)(row - 4), w1 = *(const uint32_t*)(row);
    const uint32_t w2 = *(const uint32_t*)(row + 4), w3 = *(const uint32_t*)(row + 8);
    int s[11];                                   // source columns 2*x0-2 .. 2*x0+8
    s[0] = (w0 >> 16) & 255; s[1] = w0 >> 24;
    s[2] = w1 & 255; s[3] = (w1 >> 8) & 255; s[4] = (w1 >> 16) & 255; s[5] = w1 >> 24;
    s[6] = w2 & 255; s[7] = (w2 >> 8) & 255; s[8] = (w2 >> 16) & 255; s[9] = w2 >> 24;
    s[10] = w3 & 255;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        h4[k] = s[2 * k] + s[2 * k + 4] + 4 * (s[2 * k + 1] + s[2 * k + 3]) + 6 * s[2 * k + 2];
}

__global__ void __launch_bounds__(256) k_pyr_down(zs_pyr_view v, int level, int first)
{
    // work items = 4-pixel groups, flattened over (row, group) so that every thread of a block has one even when a row
    // holds fewer than 256 groups (a block per row left 63 % of the threads idle at level 0 -> 1 and more above)
    const int dw = v.w[level + 1], dh = v.h[level + 1];
    const int wq = (dw + 3) >> 2;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = item / wq, x0 = (item - y * wq) * 4;
    if (y >= dh) return;
    const int slot = zs_slot(first, blockIdx.z, v.slots);
    const int sp = v.pitch[level];
    const uint8_t* src = v.img[level] + (size_t)slot * v.slot_stride[level] + (size_t)v.pad_y * sp + v.pad_x;
    const uint8_t* r = src + (ptrdiff_t)(2 * y - 2) * sp + 2 * x0;
    int acc[4] = { 0, 0, 0, 0 }, t[4];
    pd_row(r, t);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = t[k];
    pd_row(r + sp, t);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += 4 * t[k];
    pd_row(r + 2 * sp, t);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += 6 * t[k];
    pd_row(r + 3 * sp, t);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += 4 * t[k];
    pd_row(r + 4 * sp, t);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = (acc[k] + t[k] + 128) >> 8;
    const int dp = v.pitch[level + 1];
    uint8_t* dst = v.img[level + 1] + (size_t)slot * v.slot_stride[level + 1] + (size_t)(v.pad_y + y) * dp + v.pad_x + x0;
    if (x0 + 4 <= dw) {
        *(uint32_t*)dst = (uint32_t)acc[0] | ((uint32_t)acc[1] << 8) | ((uint32_t)acc[2] << 16) | ((uint32_t)acc[3] << 24);
    } else {
        for (int k = 0; x0 + k < dw; ++k) dst[k] = (uint8_t)acc[k];
    }
}

// ---- (3) Scharr ---------------------------------------------------------------------------------------
// dx = 3*(r0[x+1]-r0[x-1]) + 10*(r1[x+1]-r1[x-1]) + 3*(r2[x+1]-r2[x-1])
// dy = 3*(r2[x-1]-r0[x-1]) + 10*(r2[x]-r0[x]) + 3*(r2[x+1]-r0[x+1]);  REFLECT_101 comes from the padding.
__device__ __forceinline__ void sc_row(const uint8_t* __restrict__ row, int s[6])
{
    // row points at column x0 (multiple of 4): columns x0-1 .. x0+4
    const uint32_t w0 = *(const uint32_t*)(row - 4), w1 = *(const uint32_t*)(row), w2 = *(const uint32_t*)(row + 4);
    s[0] = w0 >> 24;
    s[1] = w1 & 255; s[2] = (w1 >> 8) & 255; s[3] = (w1 >> 16) & 255; s[4] = w1 >> 24;
    s[5] = w2 & 255;
}

__global__ void __launch_bounds__(256) k_scharr(zs_pyr_view v, int level, int first)
{
    const int w = v.w[level];
    const int wq = (w + 3) >> 2;                         // flattened (row, 4-pixel group) items, see k_pyr_down
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = item / wq, x0 = (item - y * wq) * 4;
    if (y >= v.h[level]) return;
    const int slot = zs_slot(first, blockIdx.z, v.slots);
    const int pitch = v.pitch[level];
    const size_t org = (size_t)slot * v.slot_stride[level] + (size_t)(v.pad_y + y) * pitch + v.pad_x + x0;
    const uint8_t* r1 = v.img[level] + org;
    int a[6], b[6], c[6];
    sc_row(r1 - pitch, a); sc_row(r1, b); sc_row(r1 + pitch, c);
    short2 out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int dx = 3 * (a[k + 2] - a[k]) + 10 * (b[k + 2] - b[k]) + 3 * (c[k + 2] - c[k]);
        const int dy = 3 * (c[k] - a[k]) + 10 * (c[k + 1] - a[k + 1]) + 3 * (c[k + 2] - a[k + 2]);
        out[k] = make_short2((short)dx, (short)dy);
    }
    short2* dst = v.der[level] + org;
    if (x0 + 4 <= w) {
        *(uint4*)dst = *(const uint4*)out;          // 16-byte aligned: pad_x, pitch and x0 are multiples of 4
    } else {
        for (int k = 0; x0 + k < w; ++k) dst[k] = out[k];
    }
}

// ---- fused level kernel ---------------------------------------------------------------------------------
// One launch per level instead of three: pad, pyrDown and Scharr of level l all read only level l's INTERIOR, so they
// can run side by side as block roles of one grid (Scharr blocks first: they carry the most work) -- provided the
// stencils do not use the padding that the pad blocks are writing at the same time.  Rows above / below the image are
// addressed through REFLECT_101 row arithmetic; the first and the last item of a row patch the one or two window
// columns that fall outside the image from the columns already in registers (col -1 = col 1, col w = col w-2, ...).
// Their aligned word loads may still cover pad bytes, but those bytes are never used.  Levels narrower than 16 pixels
// take a scalar path.
__device__ __forceinline__ void pd_row_tiny(const uint8_t* __restrict__ row0, int cx, int w, int h4[4])
{
    int s[11];                                   // source columns cx-2 .. cx+8 of the interior row starting at row0
#pragma unroll
    for (int k = 0; k < 11; ++k) s[k] = row0[zs_reflect101(cx - 2 + k, w)];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        h4[k] = s[2 * k] + s[2 * k + 4] + 4 * (s[2 * k + 1] + s[2 * k + 3]) + 6 * s[2 * k + 2];
}

// e = w - cx: number of image columns from cx on (only looked at when `last`)
__device__ __forceinline__ void pd_row_fix(const uint8_t* __restrict__ row, bool first, bool last, int e, int h4[4])
{
    const uint32_t w0 = first ? 0u : *(const uint32_t*)(row - 4), w1 = *(const uint32_t*)(row);
    const uint32_t w2 = *(const uint32_t*)(row + 4), w3 = *(const uint32_t*)(row + 8);
    int s[11];
    s[0] = (w0 >> 16) & 255; s[1] = w0 >> 24;
    s[2] = w1 & 255; s[3] = (w1 >> 8) & 255; s[4] = (w1 >> 16) & 255; s[5] = w1 >> 24;
    s[6] = w2 & 255; s[7] = (w2 >> 8) & 255; s[8] = (w2 >> 16) & 255; s[9] = w2 >> 24;
    s[10] = w3 & 255;
    if (first) { s[0] = s[4]; s[1] = s[3]; }                 // columns -2, -1 = columns 2, 1
    if (last) {                                              // columns w, w+1 = columns w-2, w-3
#pragma unroll
        for (int k = 1; k <= 8; ++k)
            if (e == k) { s[k + 2] = s[k]; if (k + 3 <= 10) s[k + 3] = s[k - 1]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        h4[k] = s[2 * k] + s[2 * k + 4] + 4 * (s[2 * k + 1] + s[2 * k + 3]) + 6 * s[2 * k + 2];
}

__device__ __forceinline__ void sc_row_tiny(const uint8_t* __restrict__ row0, int x0, int w, int s[6])
{
#pragma unroll
    for (int k = 0; k < 6; ++k) s[k] = row0[zs_reflect101(x0 - 1 + k, w)];
}

// r = w - x0: number of image columns from x0 on (only looked at when `last`)
__device__ __forceinline__ void sc_row_fix(const uint8_t* __restrict__ row, bool first, bool last, int r, int s[6])
{
    const uint32_t w0 = first ? 0u : *(const uint32_t*)(row - 4), w1 = *(const uint32_t*)(row);
    const uint32_t w2 = (last && r <= 4) ? 0u : *(const uint32_t*)(row + 4);
    s[0] = w0 >> 24;
    s[1] = w1 & 255; s[2] = (w1 >> 8) & 255; s[3] = (w1 >> 16) & 255; s[4] = w1 >> 24;
    s[5] = w2 & 255;
    if (first) s[0] = s[2];                                  // column -1 = column 1
    if (last) {                                              // column w = column w-2
#pragma unroll
        for (int k = 1; k <= 4; ++k) if (r == k) s[k + 1] = s[k - 1];
    }
}

__global__ void __launch_bounds__(256, 8) k_pyr_level(zs_pyr_view v, int level, int first, int nb_scharr, int nb_down)
{
    const int w = v.w[level], h = v.h[level], pitch = v.pitch[level];
    const int slot = zs_slot(first, blockIdx.z, v.slots);
    uint8_t* plane = v.img[level] + (size_t)slot * v.slot_stride[level];
    const uint8_t* src = plane + (size_t)v.pad_y * pitch + v.pad_x;             // interior origin
    const bool tiny = w < 16;
    int b = blockIdx.x;
    if (b < nb_scharr) {
        const int wq = (w + 3) >> 2;
        const int item = b * 256 + threadIdx.x;
        const int y = item / wq, q = item - y * wq, x0 = q * 4;
        if (y >= h) return;
        const uint8_t* ra = src + (size_t)zs_reflect101(y - 1, h) * pitch;
        const uint8_t* rb = src + (size_t)y * pitch;
        const uint8_t* rc = src + (size_t)zs_reflect101(y + 1, h) * pitch;
        int a[6], bb[6], c[6];
        if (!tiny) {
            const bool fi = q == 0, la = q == wq - 1;
            const int r = w - x0;
            sc_row_fix(ra + x0, fi, la, r, a); sc_row_fix(rb + x0, fi, la, r, bb); sc_row_fix(rc + x0, fi, la, r, c);
        } else { sc_row_tiny(ra, x0, w, a); sc_row_tiny(rb, x0, w, bb); sc_row_tiny(rc, x0, w, c); }
        short2 out[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int dx = 3 * (a[k + 2] - a[k]) + 10 * (bb[k + 2] - bb[k]) + 3 * (c[k + 2] - c[k]);
            const int dy = 3 * (c[k] - a[k]) + 10 * (c[k + 1] - a[k + 1]) + 3 * (c[k + 2] - a[k + 2]);
            out[k] = make_short2((short)dx, (short)dy);
        }
        short2* dst = v.der[level] + (size_t)slot * v.slot_stride[level] + (size_t)(v.pad_y + y) * pitch + v.pad_x + x0;
        if (x0 + 4 <= w) *(uint4*)dst = *(const uint4*)out;
        else {
#pragma unroll
            for (int k = 0; k < 3; ++k) if (x0 + k < w) dst[k] = out[k];
        }
        return;
    }
    b -= nb_scharr;
    if (b < nb_down) {
        const int dw = v.w[level + 1], dh = v.h[level + 1];
        const int wq = (dw + 3) >> 2;
        const int item = b * 256 + threadIdx.x;
        const int y = item / wq, q = item - y * wq, x0 = q * 4;
        if (y >= dh) return;
        const int cx = 2 * x0;
        const bool fi = q == 0, la = q == wq - 1;
        const int kw[5] = { 1, 4, 6, 4, 1 };
        int acc[4] = { 0, 0, 0, 0 }, t[4];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const uint8_t* r0 = src + (size_t)zs_reflect101(2 * y - 2 + j, h) * pitch;
            if (!tiny) pd_row_fix(r0 + cx, fi, la, w - cx, t); else pd_row_tiny(r0, cx, w, t);
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] += kw[j] * t[k];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = (acc[k] + 128) >> 8;
        const int dp = v.pitch[level + 1];
        uint8_t* dst = v.img[level + 1] + (size_t)slot * v.slot_stride[level + 1] + (size_t)(v.pad_y + y) * dp + v.pad_x + x0;
        if (x0 + 4 <= dw) *(uint32_t*)dst = (uint32_t)acc[0] | ((uint32_t)acc[1] << 8) | ((uint32_t)acc[2] << 16) | ((uint32_t)acc[3] << 24);
        else {
#pragma unroll
            for (int k = 0; k < 3; ++k) if (x0 + k < dw) dst[k] = (uint8_t)acc[k];
        }
        return;
    }
    b -= nb_down;
    // pad role
    pad_item(v, level, plane, b * 256 + threadIdx.x);
}

extern "C" zs_status zs_pyramid_build(zs_context* ctx, zs_pyramid* p, int first, int count)
{
    ZS_REQUIRE(ctx && p, "null argument");
    ZS_REQUIRE(count >= 0 && count <= p->slots && first >= 0, "bad slot range");
    if (count == 0) return ZS_OK;
    const zs_pyr_view& v = p->v;
    // Measured at 752x480 (bench stage events, fused vs three passes): 2 images 37 vs 65 us and 8 images 45 vs 70 us (28 us
    // for 2 with the flat pad items), 32 images 79 vs 82 us (whole step equal), 64: 135 vs 121, 128: 247 vs 199, 256: 468 vs
    // 357 -- with many images in flight the three separate streaming passes win, with few the launch count does.
    // ZS_PYR_SPLIT / ZS_PYR_FUSED force one or the other.
    const int force = ctx->sw.pyr_force;
    const bool split = force == 1 || (force == 0 && count > 16);
    for (int l = 0; l < v.levels; ++l) {
        const int w = v.w[l], h = v.h[l];
        const bool down = l + 1 < v.levels;
        const int nb_scharr = zs_div_up(zs_div_up(w, 4) * h, 256);
        const int nb_down = down ? zs_div_up(zs_div_up(v.w[l + 1], 4) * v.h[l + 1], 256) : 0;
        const int nb_pad = zs_div_up(pad_items_host(v, l), 256);
        if (!split) {
            k_pyr_level<<<dim3(nb_scharr + nb_down + nb_pad, 1, count), 256, 0, ctx->stream>>>(v, l, first, nb_scharr, nb_down);
            ZS_LAUNCH_CHECK(ctx);
            continue;
        }
        k_pad_reflect<<<dim3(nb_pad, 1, count), 256, 0, ctx->stream>>>(v, l, first);
        ZS_LAUNCH_CHECK(ctx);
        if (down) {
            k_pyr_down<<<dim3(nb_down, 1, count), 256, 0, ctx->stream>>>(v, l, first);
            ZS_LAUNCH_CHECK(ctx);
        }
        k_scharr<<<dim3(nb_scharr, 1, count), 256, 0, ctx->stream>>>(v, l, first);
        ZS_LAUNCH_CHECK(ctx);
    }
    return ZS_OK;
}
