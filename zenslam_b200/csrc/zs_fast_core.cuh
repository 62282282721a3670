// zs_fast_core.cuh -- FAST-9-16 ring test and score, shared by the grid / full-frame detectors
// (zs_fast.cu) and the multi-scale ORB detector (zs_orb_detect.cu).
#pragma once

#include "zs_common.cuh"

// does a 16-bit circular mask contain 9 contiguous set bits?
__device__ __forceinline__ bool has_run9(uint32_t m16)
{
    uint32_t m = m16 | (m16 << 16);
    uint32_t r = m & (m >> 1);       // runs of 2
    r &= r >> 2;                     // runs of 4
    r &= r >> 4;                     // runs of 8
    r &= m >> 8;                     // runs of 9
    return (r & 0xffffu) != 0;
}

// Ring test on a shared-memory tile: returns true if pixel (x,y) passes FAST-9 at threshold t.
__device__ __forceinline__ bool fast_is_corner(const uint8_t* tile, int tp, int x, int y, int t, int d[16])
{
    const int v = tile[y * tp + x];
    uint32_t dark = 0, bright = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        // ring offsets are compile-time after unrolling
        const int dx = (k == 0 || k == 8) ? 0 : (k == 1 || k == 7) ? 1 : (k == 2 || k == 6) ? 2 : (k >= 3 && k <= 5) ? 3
                       : (k == 9 || k == 15) ? -1 : (k == 10 || k == 14) ? -2 : -3;
        const int dy = (k == 4 || k == 12) ? 0 : (k == 3 || k == 13) ? 1 : (k == 2 || k == 14) ? 2 : (k <= 1 || k == 15) ? 3
                       : (k == 5 || k == 11) ? -1 : (k == 6 || k == 10) ? -2 : -3;
        d[k] = v - (int)tile[(y + dy) * tp + (x + dx)];
        dark |= (uint32_t)(d[k] > t) << k;
        bright |= (uint32_t)(d[k] < -t) << k;
    }
    return has_run9(dark) || has_run9(bright);
}

// s = max over 16 arcs of 9 of min(d) (dark) / min(-d) (bright); score = s - 1.
// Sliding minimum / maximum over windows of 9 by doubling (2, 4, 8, +1).
// NOTE: the arc minima and maxima are reduced separately and negated once at the end.  Folding the negation
// into the loop (`best = max(best, max(mn9, -mx9))`) is miscompiled by ptxas 12.9 for sm_100a at -O1 and
// above (3-input VIMNMX3 with a negated operand returns wrong values; verified on a B200, correct at -O0).
__device__ __forceinline__ int fast_score(const int d[16])
{
    int mn2[16], mx2[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { mn2[k] = min(d[k], d[(k + 1) & 15]); mx2[k] = max(d[k], d[(k + 1) & 15]); }
    int mn4[16], mx4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { mn4[k] = min(mn2[k], mn2[(k + 2) & 15]); mx4[k] = max(mx2[k], mx2[(k + 2) & 15]); }
    int bmn = -256, bmx = 256;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        bmn = max(bmn, min(min(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]));
        bmx = min(bmx, max(max(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]));
    }
    return max(bmn, -bmx) - 1;
}

