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

// (A variant with eight integral columns per lane, in-lane running sums, one shuffle scan per 256 columns and 16-byte
// stores needs a third of the instructions but 48 registers and fewer, longer warps: 57 us against 46 us alone, and
// 2540 against 2600 frames/s in bench.py, where this kernel has to fit beside the resident blocks of another stream's
// cascade kernel — profiles/r1_v5_summary.md.  Kept: this one.)
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

// Column pass, shared-memory tiled: a block owns NV_COLBLK = 128 physical columns of one array of one level (four per
// lane, moved as 16-byte vectors); its 16 warps split the rows into 16 bands.  Pass 1 sums each band, the band totals
// are exchanged through shared memory and prefixed, pass 2 re-reads the band (L1/L2 hit) and writes the running
// column sums.  Every global access is a 512-byte row segment.
__device__ __forceinline__ uint4 add4(uint4 a, uint4 b) { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// 16 bands = 512 threads x 40 registers, 8 KB of shared memory: small enough to share an SM with resident blocks of the
// cascade kernel of another stream.  With 32 bands (1024 threads, a whole SM's worth of registers) the column scan had
// to wait for an SM to drain: bench.py 2418 -> 2536 frames/s from this change alone (profiles/r1_v5_summary.md).
#ifndef NV_COLBANDS
#define NV_COLBANDS 16
#endif
__global__ void __launch_bounds__(32 * NV_COLBANDS)
k_colscan(const PlanDev *__restrict__ plan, int total_colblk, uint32_t *__restrict__ sum, uint32_t *__restrict__ sq)
{
    __shared__ uint4 tot[NV_COLBANDS][33];
    int lane = threadIdx.x & 31, band = threadIdx.x >> 5;
    int b = blockIdx.x;
    uint32_t *arr = sum;
    if (b >= total_colblk) { b -= total_colblk; arr = sq; }
    int l = find_level(plan, b, &LevelDesc::colblk0);
    const LevelDesc &L = plan->lv[l];
    const int pitch = L.ipitch, rows = L.lh + 1;
    int col = (b - L.colblk0) * NV_COLBLK + lane * 4;
    bool ok = col < pitch;                              // pitch is a multiple of 4: a lane's four columns exist together
    uint4 *p = reinterpret_cast<uint4 *>(arr + L.iofs + col);
    const size_t rs = (size_t)(pitch >> 2);             // row stride in uint4
    int R = (rows + NV_COLBANDS - 1) / NV_COLBANDS, r0 = band * R, r1 = min(rows, r0 + R);
    uint4 acc = make_uint4(0, 0, 0, 0);
    if (ok) {
        const uint4 *q = p + (size_t)r0 * rs;
        int r = r0;
        for (; r + 4 <= r1; r += 4, q += 4 * rs) {
            uint4 a0 = q[0], a1 = q[rs], a2 = q[2 * rs], a3 = q[3 * rs];
            acc = add4(acc, add4(add4(a0, a1), add4(a2, a3)));
        }
        for (; r < r1; r++, q += rs) acc = add4(acc, q[0]);
    }
    tot[band][lane] = acc;
    __syncthreads();
    uint4 run = make_uint4(0, 0, 0, 0);
    for (int k = 0; k < band; k++) run = add4(run, tot[k][lane]);
    if (ok) {
        uint4 *q = p + (size_t)r0 * rs;
        int r = r0;
        for (; r + 4 <= r1; r += 4, q += 4 * rs) {
            uint4 a0 = q[0], a1 = q[rs], a2 = q[2 * rs], a3 = q[3 * rs];
            a0 = add4(a0, run); a1 = add4(a1, a0); a2 = add4(a2, a1); a3 = add4(a3, a2);
            q[0] = a0; q[rs] = a1; q[2 * rs] = a2; q[3 * rs] = a3;
            run = a3;
        }
        for (; r < r1; r++, q += rs) { run = add4(run, q[0]); q[0] = run; }
    }
}

// Tilted integral (cv::integral's third output; only for cascades with tilted features).  tilted(X,Y) = sum of the
// level's pixels in the 45-degree triangle whose apex is pixel (X-1, Y-1).  Row y < Y of the triangle is the pixel run
// [X-1-k, X-1+k], k = Y-1-y, clipped to the image, i.e. R(y, cl(X+k)) - R(y, cl(X-1-k)) with the row prefix
// R(y, x) = sum[y+1][x] - sum[y][x] taken from the finished upright integral and cl() clamping to [0, lw].  So
//   tilted(X,Y) = A(X,Y) - B(X,Y),   A = sum_y R(y, cl(X+Y-1-y)),   B = sum_y R(y, cl(X-Y+y)),
// and A only depends on the anti-diagonal d = X+Y, B on the diagonal e = X-Y: ONE THREAD PER DIAGONAL walks down its
// diagonal keeping a running sum (a thread's leading rows, where the run is clipped to a whole row or to nothing,
// collapse to sum[y0][lw] or 0), no synchronisation, consecutive threads on consecutive addresses.  Pass A writes,
// pass B subtracts.  Plain row layout, pitch and level offsets shared with the upright integrals.
struct TiltView {
    const uint32_t *S; uint32_t *T; int lw, lh, pitch, plane, ys;
    __device__ __forceinline__ uint32_t s(int y, int x) const { return __ldg(S + (size_t)y * pitch + (ys == 2 ? (x & 1) * plane + (x >> 1) : x)); }
};

__device__ __forceinline__ bool tilt_setup(const PlanDev *__restrict__ plan, const uint32_t *sum, uint32_t *tilt, TiltView &v, int &i)
{
    int l = find_level(plan, blockIdx.x, &LevelDesc::dblk0);
    const LevelDesc &L = plan->lv[l];
    v = TiltView{sum + L.iofs, tilt + L.iofs, L.lw, L.lh, L.ipitch, L.iplane, L.ystep};
    i = (blockIdx.x - L.dblk0) * 256 + threadIdx.x;              // diagonal index, 0 .. lw + lh - 1
    return i < L.lw + L.lh;
}

__global__ void __launch_bounds__(256)
k_tilt_a(const PlanDev *__restrict__ plan, const uint32_t *__restrict__ sum, uint32_t *__restrict__ tilt)
{
    TiltView v; int i;
    if (!tilt_setup(plan, sum, tilt, v, i)) return;
    const int d = i + 1;                                         // X + Y, 1 .. lw + lh
    if (d - 1 <= v.lw) v.T[d - 1] = 0u;                          // row Y = 0
    int y0 = max(0, d - 1 - v.lw);                               // rows above y0: the run is clipped to the whole row
    uint32_t acc = v.s(y0, v.lw);
    for (int y = y0; y < v.lh && y < d; y++) {
        int X = d - 1 - y;                                       // 0 <= X <= lw here
        acc += v.s(y + 1, X) - v.s(y, X);
        v.T[(size_t)(y + 1) * v.pitch + X] = acc;
    }
}

__global__ void __launch_bounds__(256)
k_tilt_b(const PlanDev *__restrict__ plan, const uint32_t *__restrict__ sum, uint32_t *__restrict__ tilt)
{
    TiltView v; int i;
    if (!tilt_setup(plan, sum, tilt, v, i)) return;
    const int e = i - v.lh;                                      // X - Y, -lh .. lw - 1
    uint32_t acc = 0u;
    for (int y = max(0, -e); y < v.lh; y++) {                    // rows above: the run starts left of the image
        int X = e + y + 1;
        if (X > v.lw) break;
        acc += v.s(y + 1, X - 1) - v.s(y, X - 1);                // R(y, X - 1), 0 <= X - 1 < lw
        v.T[(size_t)(y + 1) * v.pitch + X] -= acc;
    }
}

cudaError_t launch_tilted(const PlanDev *plan, int total_dblk, const uint32_t *sum, uint32_t *tilt, cudaStream_t st)
{
    k_tilt_a<<<total_dblk, 256, 0, st>>>(plan, sum, tilt);
    k_tilt_b<<<total_dblk, 256, 0, st>>>(plan, sum, tilt);
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
    k_colscan<<<2 * total_colblk, 32 * NV_COLBANDS, 0, st>>>(plan, total_colblk, sum, sq);
    return cudaGetLastError();
}
