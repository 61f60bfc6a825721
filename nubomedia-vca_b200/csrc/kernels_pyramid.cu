// kernels_pyramid.cu — K4 + K5: the cascade's internal scale pyramid (INTER_LINEAR_EXACT, always from the
// equalised base image) fused with the row pass of the integral / squared-integral images
// (A.5), then the column pass.  All levels of a frame go through ONE launch of each kernel.
//
// Integral layout (chosen for the consumer, the cascade kernels): per level an (lh+1) x (lw+1) image of
// uint32 (sum is int32 in OpenCV; both are consumed through differences, so modulo-2^32 arithmetic on
// uint32 is bit-identical).  For ystep==2 levels the columns of every row are de-interleaved into an
// even plane and an odd plane (physical col = (c % ystep) * iplane + c / ystep): the cascade only visits
// even x there, so the 32 lanes of a warp, on windows x, x+2, x+4 …, read CONSECUTIVE words for every
// feature corner (coalesced in global memory, conflict-free once staged in shared memory).
#include "internal.h"

__device__ __forceinline__ int find_level(const PlanDev *__restrict__ plan, int idx, int LevelDesc::*first)
{
    int n = plan->nlevels, l = 0;
    while (l + 1 < n && plan->lv[l + 1].*first <= idx) l++;
    return l;
}

// One warp per level row: resize (two source rows, 8.8 x 8.8 fixed point) -> equalised value -> warp-shuffle
// inclusive scan of v and v*v along the row, carried across 32-pixel chunks.
__global__ void __launch_bounds__(256)
k_pyr_rowscan(const PlanDev *__restrict__ plan, const uint8_t *__restrict__ gray, int gstride,
              const uint8_t *__restrict__ lut, const int2 *__restrict__ ptab, uint32_t *__restrict__ sum,
              uint32_t *__restrict__ sq, uint8_t *__restrict__ pyr_debug)
{
    __shared__ uint8_t s_lut[256];
    int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    s_lut[tid] = lut[tid];
    __syncthreads();

    int l = find_level(plan, blockIdx.x, &LevelDesc::rowblk0);
    const LevelDesc &L = plan->lv[l];
    int lw = L.lw, lh = L.lh, ys = L.ystep, pitch = L.ipitch, plane = L.iplane;
    int y = (blockIdx.x - L.rowblk0) * 8 + warp;
    if (y >= lh) return;
    uint32_t *srow = sum + L.iofs, *qrow = sq + L.iofs;
    if (y == 0)                                         // first integral row is all zeros
        for (int c = lane; c < pitch; c += 32) { srow[c] = 0; qrow[c] = 0; }
    srow += (size_t)(y + 1) * pitch;
    qrow += (size_t)(y + 1) * pitch;
    if (lane == 0) { srow[0] = 0; qrow[0] = 0; }        // first integral column (c = 0 -> plane 0, col 0)

    const int2 *xt = ptab + L.xtab, *yt = ptab + L.ytab;
    int2 ty = yt[y];
    const uint8_t *g0 = gray + (size_t)ty.x * gstride, *g1 = ty.y < 0 ? g0 : g0 + gstride;
    uint32_t r1 = ty.y < 0 ? 0u : (uint32_t)ty.y, r0 = 256u - r1;
    uint32_t carry_s = 0, carry_q = 0;
    for (int x0 = 0; x0 < lw; x0 += 32) {
        int x = x0 + lane;
        uint32_t v = 0;
        if (x < lw) {
            int2 tx = xt[x];
            uint32_t h0, h1;
            if (tx.y < 0) { h0 = (uint32_t)s_lut[g0[tx.x]] << 8; h1 = (uint32_t)s_lut[g1[tx.x]] << 8; }
            else {
                uint32_t c1 = (uint32_t)tx.y, c0 = 256u - c1;
                h0 = s_lut[g0[tx.x]] * c0 + s_lut[g0[tx.x + 1]] * c1;
                h1 = s_lut[g1[tx.x]] * c0 + s_lut[g1[tx.x + 1]] * c1;
            }
            v = (h0 * r0 + h1 * r1 + 32768u) >> 16;
            if (pyr_debug) pyr_debug[L.pofs + (size_t)y * lw + x] = (uint8_t)v;
        }
        uint32_t s = v, q = v * v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t ts = __shfl_up_sync(0xffffffffu, s, d), tq = __shfl_up_sync(0xffffffffu, q, d);
            if (lane >= d) { s += ts; q += tq; }
        }
        s += carry_s; q += carry_q;
        if (x < lw) {
            int c = x + 1;
            int pc = ys == 2 ? (c & 1) * plane + (c >> 1) : c;
            srow[pc] = s; qrow[pc] = q;
        }
        carry_s = __shfl_sync(0xffffffffu, s, 31);
        carry_q = __shfl_sync(0xffffffffu, q, 31);
    }
}

// Column pass, shared-memory tiled: a block owns 32 physical columns of one array of one level; its 32 warps
// split the rows into 32 bands.  Pass 1 sums each band, the band totals are exchanged through shared memory
// and prefixed, pass 2 re-reads the band (L1/L2 hit) and writes the running column sums.  Every global access
// is a 128-byte row segment.
__global__ void __launch_bounds__(1024)
k_colscan(const PlanDev *__restrict__ plan, int total_colblk, uint32_t *__restrict__ sum, uint32_t *__restrict__ sq)
{
    __shared__ uint32_t tot[32][33];
    int lane = threadIdx.x & 31, band = threadIdx.x >> 5;
    int b = blockIdx.x;
    uint32_t *arr = sum;
    if (b >= total_colblk) { b -= total_colblk; arr = sq; }
    int l = find_level(plan, b, &LevelDesc::colblk0);
    const LevelDesc &L = plan->lv[l];
    int pitch = L.ipitch, rows = L.lh + 1;
    int col = (b - L.colblk0) * 32 + lane;
    bool ok = col < pitch;
    uint32_t *p = arr + L.iofs + col;
    int R = (rows + 31) / 32, r0 = band * R, r1 = min(rows, r0 + R);
    uint32_t acc = 0;
    if (ok)
        for (int r = r0; r < r1; r++) acc += p[(size_t)r * pitch];
    tot[band][lane] = acc;
    __syncthreads();
    uint32_t run = 0;
    for (int k = 0; k < band; k++) run += tot[k][lane];
    if (ok)
        for (int r = r0; r < r1; r++) {
            run += p[(size_t)r * pitch];
            p[(size_t)r * pitch] = run;
        }
}

// Tilted integral (cv::integral's third output; only for cascades with tilted features).  tilted(X,Y) = sum of the
// level's pixels in the 45-degree triangle whose apex is pixel (X-1, Y-1); three-term recurrence over rows
//   T(X,Y) = T(X-1,Y-1) + T(X+1,Y-1) - T(X,Y-2) + img(X-1,Y-1) + img(X-1,Y-2)
// closed on the columns 0..lw by T(-1,Y) = T(0,Y-1) and T(lw+1,Y) = T(lw,Y-1) (the strips those triangles would add lie
// outside the image).  Rows depend on the two rows above, so ONE block walks a level top to bottom, its threads
// striding over the columns with the last two rows kept in shared memory; levels run in parallel blocks.  Plain row
// layout (no column de-interleave), pitch and level offsets shared with the upright integrals.
__global__ void __launch_bounds__(1024)
k_tilted(const PlanDev *__restrict__ plan, const uint8_t *__restrict__ pyr, uint32_t *__restrict__ tilt)
{
    extern __shared__ uint32_t s_rows[];                         // 3 x (lw + 1): rows Y-2, Y-1, Y (rotating)
    const LevelDesc &L = plan->lv[blockIdx.x];
    const int lw = L.lw, lh = L.lh, pitch = L.ipitch, n = lw + 1;
    const uint8_t *img = pyr + L.pofs;
    uint32_t *out = tilt + L.iofs;
    uint32_t *r0 = s_rows, *r1 = s_rows + n, *r2 = s_rows + 2 * n;
    for (int x = threadIdx.x; x < n; x += blockDim.x) { r0[x] = 0; r1[x] = 0; out[x] = 0; }
    __syncthreads();
    for (int Y = 1; Y <= lh; Y++) {
        const uint8_t *i1 = img + (size_t)(Y - 1) * lw, *i2 = Y >= 2 ? img + (size_t)(Y - 2) * lw : nullptr;
        for (int X = threadIdx.x; X < n; X += blockDim.x) {
            uint32_t left = X > 0 ? r1[X - 1] : r0[0], right = X < lw ? r1[X + 1] : r0[lw];
            uint32_t v = left + right - r0[X];
            if (X > 0) v += (uint32_t)i1[X - 1] + (i2 ? (uint32_t)i2[X - 1] : 0u);
            r2[X] = v;
            out[(size_t)Y * pitch + X] = v;
        }
        __syncthreads();
        uint32_t *t = r0; r0 = r1; r1 = r2; r2 = t;
    }
}

cudaError_t launch_tilted(const PlanDev *plan, int nlevels, int max_lw, const uint8_t *pyr, uint32_t *tilt, cudaStream_t st)
{
    size_t smem = 3 * (size_t)(max_lw + 1) * sizeof(uint32_t);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    static std::mutex mu;
    static unsigned long long attr_set = 0ull;
    if (smem > 48 * 1024) {
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(mu);
        if (!((attr_set >> (dev & 63)) & 1ull)) {
            cudaFuncSetAttribute(k_tilted, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            attr_set |= 1ull << (dev & 63);
        }
    }
    k_tilted<<<nlevels, 1024, smem, st>>>(plan, pyr, tilt);
    return cudaGetLastError();
}

cudaError_t launch_pyr_rowscan(const PlanDev *plan, int total_rowblk, const uint8_t *gray, int gstride, const uint8_t *lut,
                               const int *ptab, uint32_t *sum, uint32_t *sq, uint8_t *pyr_debug, cudaStream_t st)
{
    k_pyr_rowscan<<<total_rowblk, 256, 0, st>>>(plan, gray, gstride, lut, (const int2 *)ptab, sum, sq, pyr_debug);
    return cudaGetLastError();
}

cudaError_t launch_colscan(const PlanDev *plan, int total_colblk, uint32_t *sum, uint32_t *sq, cudaStream_t st)
{
    k_colscan<<<2 * total_colblk, 1024, 0, st>>>(plan, total_colblk, sum, sq);
    return cudaGetLastError();
}
