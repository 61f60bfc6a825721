// kernels_tracker.cu — K11 + K12: the nubotracker per-frame path, gstnubotracker.cpp:356-377:
//   cvtColor(BGRA->gray) :356, absdiff(gray, img_prev) :361, threshold(diff, thr, 255, BINARY) :364,
//   updateMotionHistory(mask, mhi, ts, MHI_DURATION=0.2) :368, segmentMotion(mhi, ts, 32) :376,
//   img_prev = gray :419.  calcMotionGradient (:372) writes two outputs nobody reads and is skipped.
// K11 is one fused point-op kernel (reads BGRA + prev + mhi, writes prev + mhi + label seeds).
// K12 is segmentMotion as connected-component labelling: 4-connected pixels are joined when their MHI
// values differ by at most 32 (OpenCV's floating-range flood fill with lo = up = segThresh), zeros are
// never joined to anything that matters (OpenCV swaps them for FLT_MAX*0.1), a component is reported iff
// it contains a pixel with mhi == ts, in raster order of its first such pixel, with the bounding box of
// ALL its pixels.  Union-find with atomicMin links (smaller index wins), then bbox/min-seed reductions.
#include <limits.h>

#include "internal.h"

#define TRK_MAX_COMPONENTS 16384

// FMT 0: BGRA (the reference's caps, gstnubotracker.cpp:57-61); 1 / 2 / 3: I420 / NV12 / NV21 planes — the 4:2:0 ingest
// extension: gray = BGR2GRAY(cvtColor(COLOR_YUV2BGR_*)) per pixel, the BGRA frame is never built.
template <int FMT>
__global__ void __launch_bounds__(256)
k_trk_point(SrcPlanes src, int w, int h, int first, float ts, float del, int thr,
            uint8_t *__restrict__ prev, float *__restrict__ mhi, int *__restrict__ label, int4 *__restrict__ box,
            int *__restrict__ seed, uint8_t *__restrict__ mask_out)
{
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    int p = y * w + x;
    int g;
    if (FMT == 0) {
        uchar4 px = *reinterpret_cast<const uchar4 *>(src.p0 + (size_t)y * src.s0 + 4 * x);
        g = (px.x * 3735 + px.y * 19235 + px.z * 9798 + 16384) >> 15;
    } else {
        int c3[3];
        yuv_pixel(src, yuv_chroma<FMT == 0 ? 1 : FMT>(src, x, y), x, y, c3);
        g = (c3[0] * 3735 + c3[1] * 19235 + c3[2] * 9798 + 16384) >> 15;
    }
    if (!first) {
        int d = abs(g - (int)prev[p]);
        bool silh = d > thr;
        float m = mhi[p];
        m = silh ? ts : (m < del ? 0.f : m);
        mhi[p] = m;
        label[p] = m != 0.f ? p : -1;
        box[p] = make_int4(INT_MAX, INT_MAX, -1, -1);
        seed[p] = INT_MAX;
        if (mask_out) mask_out[p] = silh ? 255 : 0;
    }
    prev[p] = (uint8_t)g;
}

__device__ __forceinline__ int trk_find(volatile int *L, int a)
{
    int p;
    while ((p = L[a]) != a) a = p;
    return a;
}

__device__ __forceinline__ void trk_union(int *L, int a, int b)
{
    bool done;
    do {
        a = trk_find(L, a);
        b = trk_find(L, b);
        if (a < b) { int old = atomicMin(&L[b], a); done = old == b; b = old; }
        else if (b < a) { int old = atomicMin(&L[a], b); done = old == a; a = old; }
        else done = true;
    } while (!done);
}

__global__ void __launch_bounds__(256)
k_trk_merge(const float *__restrict__ mhi, int w, int h, float seg, int *__restrict__ label)
{
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    int p = y * w + x;
    float m = mhi[p];
    if (m == 0.f) return;
    if (x + 1 < w) {
        float q = mhi[p + 1];
        float d = q - m;
        if (q != 0.f && d >= -seg && d <= seg) trk_union(label, p, p + 1);
    }
    if (y + 1 < h) {
        float q = mhi[p + w];
        float d = q - m;
        if (q != 0.f && d >= -seg && d <= seg) trk_union(label, p, p + w);
    }
}

__global__ void __launch_bounds__(256)
k_trk_reduce(const float *__restrict__ mhi, int w, int h, float ts, int *__restrict__ label, int4 *__restrict__ box,
             int *__restrict__ seed)
{
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    int p = y * w + x;
    if (label[p] < 0) return;
    int r = trk_find(label, p);
    int *b = reinterpret_cast<int *>(box + r);
    atomicMin(b + 0, x); atomicMin(b + 1, y); atomicMax(b + 2, x); atomicMax(b + 3, y);
    if (mhi[p] == ts) atomicMin(seed + r, p);
}

// roots that own a seed pixel -> (first seed index, bbox) list
__global__ void __launch_bounds__(256)
k_trk_collect(int w, int h, const int *__restrict__ label, const int4 *__restrict__ box, const int *__restrict__ seed,
              int *__restrict__ misc, int *__restrict__ keys, int4 *__restrict__ rects)
{
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    int p = y * w + x;
    if (label[p] != p || seed[p] == INT_MAX) return;
    int pos = atomicAdd(&misc[0], 1);
    if (pos < TRK_MAX_COMPONENTS) {
        int4 b = box[p];
        keys[pos] = seed[p];
        rects[pos] = make_int4(b.x, b.y, b.z - b.x + 1, b.w - b.y + 1);
    }
}

// rank sort by first seed pixel; result block = [count, pad, pad, pad][rects...]
__global__ void __launch_bounds__(256)
k_trk_sort(int *__restrict__ misc, const int *__restrict__ keys, const int4 *__restrict__ rects, int4 *__restrict__ out)
{
    int n = min(misc[0], TRK_MAX_COMPONENTS);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int key = keys[i], rank = 0;
        for (int j = 0; j < n; j++) rank += __ldg(keys + j) < key;
        out[1 + rank] = rects[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = make_int4(misc[0], 0, 0, 0);
}

cudaError_t launch_tracker(nv_ctx *ctx, int fmt, const SrcPlanes &src, int w, int h, int first, float ts, float del,
                           int thr, int *nlaunch)
{
    dim3 grid((w + 31) / 32, (h + 7) / 8), block(32, 8);
    cudaStream_t st = ctx->stream;
    int n = w * h;
    // scratch carved from d_trk_labels: label[n] | seed[n] | keys[MAX] ; boxes: box[n] | rects[MAX] | out[1+MAX]
    int *label = ctx->d_trk_labels, *seed = label + n, *keys = seed + n;
    int4 *box = ctx->d_trk_boxes, *rects = box + n, *out = rects + TRK_MAX_COMPONENTS;
    uint8_t *mask = ctx->debug ? ctx->d_trk_mask : nullptr;
#define TRK_POINT(F) k_trk_point<F><<<grid, block, 0, st>>>(src, w, h, first, ts, del, thr, ctx->d_trk_prev, ctx->d_trk_mhi, label, box, seed, mask)
    if (fmt == NV_FMT_BGR) TRK_POINT(0);             // interleaved: BGRA here
    else if (fmt == NV_FMT_I420) TRK_POINT(1);
    else if (fmt == NV_FMT_NV12) TRK_POINT(2);
    else if (fmt == NV_FMT_NV21) TRK_POINT(3);
    else return cudaErrorInvalidValue;
#undef TRK_POINT
    (*nlaunch)++;
    if (!first) {
        cudaError_t e = cudaMemsetAsync(ctx->d_trk_misc, 0, 4 * sizeof(int), st);
        if (e != cudaSuccess) return e;
        k_trk_merge<<<grid, block, 0, st>>>(ctx->d_trk_mhi, w, h, 32.f, label);
        k_trk_reduce<<<grid, block, 0, st>>>(ctx->d_trk_mhi, w, h, ts, label, box, seed);
        k_trk_collect<<<grid, block, 0, st>>>(w, h, label, box, seed, ctx->d_trk_misc, keys, rects);
        k_trk_sort<<<64, 256, 0, st>>>(ctx->d_trk_misc, keys, rects, out);
        (*nlaunch) += 4;
    }
    return cudaGetLastError();
}
