// kernels_prep.cu — K1/K2/K3: colour conversion, cv::resize(INTER_LINEAR), equalizeHist pieces, flip.
// Arithmetic follows OpenCV 4.13 bit-exactly (SURVEY.md A.1–A.3; oracle/nubo_oracle.c is the checker).
// Reference call sites: kmsfacedetect.cpp:805-807, kmseyedetect.cpp:949-964, kmsmouthdetect.cpp:836-853,
// kmsnosedetect.cpp:834-851, kmseardetect.cpp:786-800, gstnubotracker.cpp:356.
#include <math.h>

#include "internal.h"

// ------------------------------------------------------------------------------------------------
// host: coefficient tables of cv::resize(INTER_LINEAR) for u8 (11-bit fixed point)
// ------------------------------------------------------------------------------------------------
static inline int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

void build_resize_tables(int sw, int sh, int dw, int dh, std::vector<int> &tab)
{
    tab.clear();
    if (sw == dw && sh == dh) { tab.push_back(RT_COPY); return; }
    if (sw == 2 * dw && sh == 2 * dh) { tab.push_back(RT_BOX2); return; }   // INTER_AREA fast path
    tab.resize(1 + 2 * (size_t)dw + 3 * (size_t)dh);
    tab[0] = RT_LINEAR;
    int *xofs = &tab[1], *xa = xofs + dw, *y0 = xa + dw, *y1 = y0 + dh, *yb = y1 + dh;
    double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        int a0 = sat_short((int)lrintf((1.f - fx) * 2048)), a1 = sat_short((int)lrintf(fx * 2048));
        xofs[dx] = sx;
        xa[dx] = (a0 & 0xFFFF) | (a1 << 16);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        // vertical coefficients keep their fraction; only the row indices are clamped
        int b0 = sat_short((int)lrintf((1.f - fy) * 2048)), b1 = sat_short((int)lrintf(fy * 2048));
        y0[dy] = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
        y1[dy] = sy + 1 < 0 ? 0 : (sy + 1 > sh - 1 ? sh - 1 : sy + 1);
        yb[dy] = (b0 & 0xFFFF) | (b1 << 16);
    }
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int gray_of(int b, int g, int r)
{
    return (b * 3735 + g * 19235 + r * 9798 + 16384) >> 15;     // A.1, 15-bit coefficients
}

__device__ __forceinline__ int lin_tap(const uint8_t *__restrict__ r0, const uint8_t *__restrict__ r1, int i0, int i1,
                                       int a0, int a1, int b0, int b1)
{
    int h0 = r0[i0] * a0 + r0[i1] * a1;
    int h1 = r1[i0] * a0 + r1[i1] * a1;
    return (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;       // A.2 vertical pass
}

// K1+K2 fused for the face element: resize the BGR(A) frame (3 channels computed), convert to gray,
// accumulate the histogram equalizeHist needs.  One thread per output pixel.
__global__ void __launch_bounds__(256)
k_face_prep(const uint8_t *__restrict__ src, int sw, int sh, int sstride, int cn, uint8_t *__restrict__ gray, int dw,
            int dh, const int *__restrict__ rtab, int *__restrict__ hist)
{
    __shared__ int sh_hist[256];
    int tid = threadIdx.y * 32 + threadIdx.x;
    sh_hist[tid] = 0;
    __syncthreads();
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x < dw && y < dh) {
        int mode = rtab[0], g;
        if (mode == RT_COPY) {
            const uint8_t *p = src + (size_t)y * sstride + x * cn;
            g = gray_of(p[0], p[1], p[2]);
        } else if (mode == RT_BOX2) {
            const uint8_t *p0 = src + (size_t)(2 * y) * sstride + 2 * x * cn, *p1 = p0 + sstride;
            int c3[3];
#pragma unroll
            for (int c = 0; c < 3; c++) c3[c] = (p0[c] + p0[c + cn] + p1[c] + p1[c + cn] + 2) >> 2;
            g = gray_of(c3[0], c3[1], c3[2]);
        } else {
            const int *xofs = rtab + 1, *xa = xofs + dw, *y0t = xa + dw, *y1t = y0t + dh, *ybt = y1t + dh;
            int sx = xofs[x], sx1 = min(sx + 1, sw - 1), xav = xa[x], ybv = ybt[y];
            int a0 = (short)(xav & 0xFFFF), a1 = xav >> 16, b0 = (short)(ybv & 0xFFFF), b1 = ybv >> 16;
            const uint8_t *r0 = src + (size_t)y0t[y] * sstride, *r1 = src + (size_t)y1t[y] * sstride;
            int c3[3];
#pragma unroll
            for (int c = 0; c < 3; c++) c3[c] = lin_tap(r0, r1, sx * cn + c, sx1 * cn + c, a0, a1, b0, b1);
            g = gray_of(c3[0], c3[1], c3[2]);
        }
        gray[(size_t)y * dw + x] = (uint8_t)g;
        atomicAdd(&sh_hist[g], 1);
    }
    __syncthreads();
    if (sh_hist[tid]) atomicAdd(&hist[tid], sh_hist[tid]);
}

// ---- 4:2:0 ingest: the same block fed by I420 / NV12 / NV21 planes.  Every source pixel the
// resize touches is converted with cvtColor(COLOR_YUV2BGR_*)'s arithmetic (BT.601, 20-bit fixed point, saturated to
// u8 per pixel; oracle: ora_yuv420_to_bgr), so the result equals the reference block on the converted BGR frame.
struct YuvTerms { int b, g, r; };                 // chroma contributions incl. the rounding half

template <int FMT>   // 1: I420 (three planes), 2: NV12 (UV interleaved), 3: NV21 (VU interleaved)
__device__ __forceinline__ YuvTerms yuv_chroma(const SrcPlanes &s, int x, int y)
{
    int u, v;
    if (FMT == 1) {
        u = s.p1[(size_t)(y >> 1) * s.s1 + (x >> 1)];
        v = s.p2[(size_t)(y >> 1) * s.s2 + (x >> 1)];
    } else {
        const uint8_t *uv = s.p1 + (size_t)(y >> 1) * s.s1 + (x & ~1);
        u = uv[FMT == 2 ? 0 : 1];
        v = uv[FMT == 2 ? 1 : 0];
    }
    u -= 128; v -= 128;
    YuvTerms t;
    t.b = (1 << 19) + 2116026 * u;
    t.g = (1 << 19) - 852492 * v - 409993 * u;
    t.r = (1 << 19) + 1673527 * v;
    return t;
}

__device__ __forceinline__ int sat_u8(int v) { return min(max(v, 0), 255); }

__device__ __forceinline__ void yuv_pixel(const SrcPlanes &s, const YuvTerms &t, int x, int y, int c3[3])
{
    int yy = max(0, (int)s.p0[(size_t)y * s.s0 + x] - 16) * 1220542;
    c3[0] = sat_u8((yy + t.b) >> 20);
    c3[1] = sat_u8((yy + t.g) >> 20);
    c3[2] = sat_u8((yy + t.r) >> 20);
}

template <int FMT>
__global__ void __launch_bounds__(256)
k_face_prep_yuv(SrcPlanes s, int sw, int sh, uint8_t *__restrict__ gray, int dw, int dh, const int *__restrict__ rtab,
                int *__restrict__ hist)
{
    __shared__ int sh_hist[256];
    int tid = threadIdx.y * 32 + threadIdx.x;
    sh_hist[tid] = 0;
    __syncthreads();
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x < dw && y < dh) {
        int mode = rtab[0], g, c3[3];
        if (mode == RT_COPY) {
            yuv_pixel(s, yuv_chroma<FMT>(s, x, y), x, y, c3);
        } else if (mode == RT_BOX2) {                       // the four source pixels share one chroma sample
            YuvTerms t = yuv_chroma<FMT>(s, 2 * x, 2 * y);
            int a[3], b[3], c[3], d[3];
            yuv_pixel(s, t, 2 * x, 2 * y, a);     yuv_pixel(s, t, 2 * x + 1, 2 * y, b);
            yuv_pixel(s, t, 2 * x, 2 * y + 1, c); yuv_pixel(s, t, 2 * x + 1, 2 * y + 1, d);
#pragma unroll
            for (int k = 0; k < 3; k++) c3[k] = (a[k] + b[k] + c[k] + d[k] + 2) >> 2;
        } else {
            const int *xofs = rtab + 1, *xa = xofs + dw, *y0t = xa + dw, *y1t = y0t + dh, *ybt = y1t + dh;
            int sx = xofs[x], sx1 = min(sx + 1, sw - 1), xav = xa[x], ybv = ybt[y], sy0 = y0t[y], sy1 = y1t[y];
            int a0 = (short)(xav & 0xFFFF), a1 = xav >> 16, b0 = (short)(ybv & 0xFFFF), b1 = ybv >> 16;
            int p00[3], p01[3], p10[3], p11[3];
            yuv_pixel(s, yuv_chroma<FMT>(s, sx, sy0), sx, sy0, p00);   yuv_pixel(s, yuv_chroma<FMT>(s, sx1, sy0), sx1, sy0, p01);
            yuv_pixel(s, yuv_chroma<FMT>(s, sx, sy1), sx, sy1, p10);   yuv_pixel(s, yuv_chroma<FMT>(s, sx1, sy1), sx1, sy1, p11);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                int h0 = p00[k] * a0 + p01[k] * a1, h1 = p10[k] * a0 + p11[k] * a1;
                c3[k] = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;       // A.2 vertical pass
            }
        }
        g = gray_of(c3[0], c3[1], c3[2]);
        gray[(size_t)y * dw + x] = (uint8_t)g;
        atomicAdd(&sh_hist[g], 1);
    }
    __syncthreads();
    if (sh_hist[tid]) atomicAdd(&hist[tid], sh_hist[tid]);
}

// cvtColor(COLOR_YUV2BGR_*) alone (parity tap of the ingest path); one thread per pixel
template <int FMT>
__global__ void __launch_bounds__(256)
k_yuv2bgr(SrcPlanes s, int w, int h, uint8_t *__restrict__ dst, int dstride)
{
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    int c3[3];
    yuv_pixel(s, yuv_chroma<FMT>(s, x, y), x, y, c3);
    uint8_t *d = dst + (size_t)y * dstride + 3 * x;
    d[0] = (uint8_t)c3[0]; d[1] = (uint8_t)c3[1]; d[2] = (uint8_t)c3[2];
}

__global__ void __launch_bounds__(256)
k_bgr2gray(const uint8_t *__restrict__ src, int w, int h, int sstride, int cn, uint8_t *__restrict__ dst, int dstride)
{
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t *p = src + (size_t)y * sstride + x * cn;
    dst[(size_t)y * dstride + x] = (uint8_t)gray_of(p[0], p[1], p[2]);
}

// Generic cv::resize(INTER_LINEAR) for cn interleaved channels; one thread per output pixel.
__global__ void __launch_bounds__(256)
k_resize_linear(const uint8_t *__restrict__ src, int sw, int sh, int sstride, int cn, uint8_t *__restrict__ dst, int dw,
                int dh, int dstride, const int *__restrict__ rtab)
{
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= dw || y >= dh) return;
    int mode = rtab[0];
    uint8_t *d = dst + (size_t)y * dstride + x * cn;
    if (mode == RT_COPY) {
        const uint8_t *p = src + (size_t)y * sstride + x * cn;
        for (int c = 0; c < cn; c++) d[c] = p[c];
    } else if (mode == RT_BOX2) {
        const uint8_t *p0 = src + (size_t)(2 * y) * sstride + 2 * x * cn, *p1 = p0 + sstride;
        for (int c = 0; c < cn; c++) d[c] = (uint8_t)((p0[c] + p0[c + cn] + p1[c] + p1[c + cn] + 2) >> 2);
    } else {
        const int *xofs = rtab + 1, *xa = xofs + dw, *y0t = xa + dw, *y1t = y0t + dh, *ybt = y1t + dh;
        int sx = xofs[x], sx1 = min(sx + 1, sw - 1), xav = xa[x], ybv = ybt[y];
        int a0 = (short)(xav & 0xFFFF), a1 = xav >> 16, b0 = (short)(ybv & 0xFFFF), b1 = ybv >> 16;
        const uint8_t *r0 = src + (size_t)y0t[y] * sstride, *r1 = src + (size_t)y1t[y] * sstride;
        for (int c = 0; c < cn; c++) d[c] = (uint8_t)lin_tap(r0, r1, sx * cn + c, sx1 * cn + c, a0, a1, b0, b1);
    }
}

__global__ void __launch_bounds__(256)
k_hist(const uint8_t *__restrict__ src, int w, int h, int stride, int *__restrict__ hist)
{
    __shared__ int sh_hist[256];
    int tid = threadIdx.y * 32 + threadIdx.x;
    sh_hist[tid] = 0;
    __syncthreads();
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x < w && y < h) atomicAdd(&sh_hist[src[(size_t)y * stride + x]], 1);
    __syncthreads();
    if (sh_hist[tid]) atomicAdd(&hist[tid], sh_hist[tid]);
}

// K3: histogram -> equalizeHist LUT (A.3).  One block of 256 threads; clears the histogram for the
// next frame.  lut[k] = sat_u8(rint((float)(cum[k] - cum[i0]) * (255.f / (total - hist[i0])))).
__global__ void __launch_bounds__(256) k_lut(int *__restrict__ hist, int total, uint8_t *__restrict__ lut)
{
    __shared__ int cum[256];
    __shared__ int s_i0;
    int t = threadIdx.x, hv = hist[t];
    if (t == 0) s_i0 = 256;
    cum[t] = hv;
    __syncthreads();
    if (hv) atomicMin(&s_i0, t);
    for (int d = 1; d < 256; d <<= 1) {         // Hillis-Steele inclusive scan
        int v = t >= d ? cum[t - d] : 0;
        __syncthreads();
        cum[t] += v;
        __syncthreads();
    }
    int i0 = s_i0;
    uint8_t out;
    if (i0 >= 256) out = (uint8_t)t;                                  // empty image: identity
    else {
        int h0 = cum[i0] - (i0 ? cum[i0 - 1] : 0);
        if (h0 == total) out = (uint8_t)i0;                            // constant image
        else if (t <= i0) out = 0;
        else {
            float scale = __fdiv_rn(255.f, __int2float_rn(total - h0));
            int v = __float2int_rn(__fmul_rn(__int2float_rn(cum[t] - cum[i0]), scale));
            out = (uint8_t)min(max(v, 0), 255);
        }
    }
    lut[t] = out;
    hist[t] = 0;
}

__global__ void __launch_bounds__(256)
k_apply_lut(const uint8_t *__restrict__ src, int w, int h, int sstride, const uint8_t *__restrict__ lut,
            uint8_t *__restrict__ dst, int dstride)
{
    __shared__ uint8_t s_lut[256];
    s_lut[threadIdx.y * 32 + threadIdx.x] = lut[threadIdx.y * 32 + threadIdx.x];
    __syncthreads();
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x < w && y < h) dst[(size_t)y * dstride + x] = s_lut[src[(size_t)y * sstride + x]];
}

__global__ void __launch_bounds__(256)
k_flip(const uint8_t *__restrict__ src, int w, int h, int sstride, uint8_t *__restrict__ dst, int dstride)
{
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x < w && y < h) dst[(size_t)y * dstride + x] = src[(size_t)y * sstride + (w - 1 - x)];
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static inline dim3 grid2d(int w, int h) { return dim3((w + 31) / 32, (h + 7) / 8); }

cudaError_t launch_face_prep(const uint8_t *src, int sw, int sh, int sstride, int cn, uint8_t *gray, int dw, int dh,
                             const int *rtab, int *hist, cudaStream_t st)
{
    k_face_prep<<<grid2d(dw, dh), dim3(32, 8), 0, st>>>(src, sw, sh, sstride, cn, gray, dw, dh, rtab, hist);
    return cudaGetLastError();
}
cudaError_t launch_face_prep_yuv(int fmt, const SrcPlanes &s, int sw, int sh, uint8_t *gray, int dw, int dh, const int *rtab,
                                 int *hist, cudaStream_t st)
{
    dim3 g = grid2d(dw, dh), b(32, 8);
    if (fmt == NV_FMT_I420) k_face_prep_yuv<1><<<g, b, 0, st>>>(s, sw, sh, gray, dw, dh, rtab, hist);
    else if (fmt == NV_FMT_NV12) k_face_prep_yuv<2><<<g, b, 0, st>>>(s, sw, sh, gray, dw, dh, rtab, hist);
    else if (fmt == NV_FMT_NV21) k_face_prep_yuv<3><<<g, b, 0, st>>>(s, sw, sh, gray, dw, dh, rtab, hist);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}
cudaError_t launch_yuv2bgr(int fmt, const SrcPlanes &s, int w, int h, uint8_t *dst, int dstride, cudaStream_t st)
{
    dim3 g = grid2d(w, h), b(32, 8);
    if (fmt == NV_FMT_I420) k_yuv2bgr<1><<<g, b, 0, st>>>(s, w, h, dst, dstride);
    else if (fmt == NV_FMT_NV12) k_yuv2bgr<2><<<g, b, 0, st>>>(s, w, h, dst, dstride);
    else if (fmt == NV_FMT_NV21) k_yuv2bgr<3><<<g, b, 0, st>>>(s, w, h, dst, dstride);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}
cudaError_t launch_bgr2gray(const uint8_t *src, int w, int h, int sstride, int cn, uint8_t *dst, int dstride,
                            cudaStream_t st)
{
    k_bgr2gray<<<grid2d(w, h), dim3(32, 8), 0, st>>>(src, w, h, sstride, cn, dst, dstride);
    return cudaGetLastError();
}
cudaError_t launch_resize_linear(const uint8_t *src, int sw, int sh, int sstride, int cn, uint8_t *dst, int dw, int dh,
                                 int dstride, const int *rtab, cudaStream_t st)
{
    k_resize_linear<<<grid2d(dw, dh), dim3(32, 8), 0, st>>>(src, sw, sh, sstride, cn, dst, dw, dh, dstride, rtab);
    return cudaGetLastError();
}
cudaError_t launch_hist(const uint8_t *src, int w, int h, int stride, int *hist, cudaStream_t st)
{
    k_hist<<<grid2d(w, h), dim3(32, 8), 0, st>>>(src, w, h, stride, hist);
    return cudaGetLastError();
}
cudaError_t launch_lut(int *hist, int total, uint8_t *lut, cudaStream_t st)
{
    k_lut<<<1, 256, 0, st>>>(hist, total, lut);
    return cudaGetLastError();
}
cudaError_t launch_apply_lut(const uint8_t *src, int w, int h, int sstride, const uint8_t *lut, uint8_t *dst, int dstride,
                             cudaStream_t st)
{
    k_apply_lut<<<grid2d(w, h), dim3(32, 8), 0, st>>>(src, w, h, sstride, lut, dst, dstride);
    return cudaGetLastError();
}
cudaError_t launch_flip(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride, cudaStream_t st)
{
    k_flip<<<grid2d(w, h), dim3(32, 8), 0, st>>>(src, w, h, sstride, dst, dstride);
    return cudaGetLastError();
}
