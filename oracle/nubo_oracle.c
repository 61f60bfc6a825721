/*
 * nubo_oracle.c — CPU restatement of the arithmetic behind NUBOMEDIA-VCA's per-frame
 * detection hot path.  TEST INFRASTRUCTURE ONLY: nothing under nubomedia-vca_b200/ may
 * link, import or execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * The reference (modules/nubo_face/.../kmsfacedetect.cpp:805-811 and the analogous blocks
 * in kmseyedetect.cpp:949-1005, kmsmouthdetect.cpp:836-873, kmsnosedetect.cpp:834-873,
 * kmseardetect.cpp:786-803,656-715, gstnubotracker.cpp:356-377) holds no arithmetic of its
 * own: each step is a call into OpenCV, an un-vendored, un-pinned dependency
 * (CMakeLists.txt:35,46 `opencv>=2.0.0`).  The published algorithm restated here is
 * OpenCV 4.13.0's (the only runnable implementation in this image), following
 * SURVEY.md Appendix A.  Parity pin: tests/test_oracle_vs_cv2.py checks every function
 * below against cv2 4.13 live and against fixtures under tests/golden/ that
 * tests/golden/make_golden.py generated from cv2.  The reference itself ships no tests
 * or golden vectors, so relative to the reference parity is "unpinned";
 * relative to the library the reference calls it is pinned bit-exactly.
 *
 * Plain C99, single-threaded, no FMA contraction (build with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORA_API __attribute__((visibility("default")))

/* cvRound: round half to even (default FP environment). */
static inline int ora_round(double v) { return (int)lrint(v); }
static inline int ora_roundf(float v) { return (int)lrintf(v); }
static inline int ora_min(int a, int b) { return a < b ? a : b; }
static inline int ora_max(int a, int b) { return a > b ? a : b; }
static inline short ora_sat_short(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }
static inline uint8_t ora_sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

/* ---------------------------------------------------------------------------------------
 * A.1  cvtColor(BGR2GRAY) on 3- or 4-channel u8 (kmsfacedetect.cpp:806, gstnubotracker.cpp:356)
 * ------------------------------------------------------------------------------------- */
ORA_API void ora_bgr2gray(const uint8_t *src, int w, int h, int sstride, int cn,
                          uint8_t *dst, int dstride)
{
    for (int y = 0; y < h; y++) {
        const uint8_t *s = src + (size_t)y * sstride;
        uint8_t *d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++, s += cn)
            d[x] = (uint8_t)((s[0] * 3735 + s[1] * 19235 + s[2] * 9798 + 16384) >> 15);
    }
}

/* ---------------------------------------------------------------------------------------
 * A.1b cvtColor(COLOR_YUV2BGR_I420 / _NV12 / _NV21): the 4:2:0 ingest extension.
 *      The reference elements only accept BGR caps (kmsfacedetect.cpp:1025-1031), so in a pipeline a
 *      videoconvert sits in front of them; this is OpenCV 4.13's conversion (ITU-R BT.601, 20-bit fixed
 *      point, modules/imgproc color_yuv), pinned against cv2 in tests/test_oracle_vs_cv2.py.
 *      fmt 0: three planes (I420; YV12 = the caller swaps u and v), 1: NV12 (u = interleaved UV plane),
 *      2: NV21 (u = interleaved VU plane).  w and h even.
 * ------------------------------------------------------------------------------------- */
ORA_API void ora_yuv420_to_bgr(int fmt, const uint8_t *yp, int ystride, const uint8_t *up, int ustride,
                               const uint8_t *vp, int vstride, int w, int h, uint8_t *dst, int dstride)
{
    for (int y = 0; y < h; y++) {
        uint8_t *d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            int u, v;
            if (fmt == 0) { u = up[(size_t)(y / 2) * ustride + x / 2]; v = vp[(size_t)(y / 2) * vstride + x / 2]; }
            else {
                const uint8_t *uv = up + (size_t)(y / 2) * ustride + (x / 2) * 2;
                u = uv[fmt == 1 ? 0 : 1]; v = uv[fmt == 1 ? 1 : 0];
            }
            u -= 128; v -= 128;
            int yy = ora_max(0, (int)yp[(size_t)y * ystride + x] - 16) * 1220542;
            d[3 * x + 0] = ora_sat_u8((yy + (1 << 19) + 2116026 * u) >> 20);
            d[3 * x + 1] = ora_sat_u8((yy + (1 << 19) - 852492 * v - 409993 * u) >> 20);
            d[3 * x + 2] = ora_sat_u8((yy + (1 << 19) + 1673527 * v) >> 20);
        }
    }
}

/* ---------------------------------------------------------------------------------------
 * A.2  cv::resize(INTER_LINEAR), u8, cn interleaved channels (kmsfacedetect.cpp:805,
 *      kmseyedetect.cpp:956,963 ...)
 * ------------------------------------------------------------------------------------- */
ORA_API void ora_resize_linear(const uint8_t *src, int sw, int sh, int sstride, int cn,
                               uint8_t *dst, int dw, int dh, int dstride)
{
    if (sw == dw && sh == dh) {                      /* same size: plain copy */
        for (int y = 0; y < sh; y++)
            memcpy(dst + (size_t)y * dstride, src + (size_t)y * sstride, (size_t)sw * cn);
        return;
    }
    if (sw == 2 * dw && sh == 2 * dh) {              /* exact 2x: INTER_AREA fast path */
        for (int y = 0; y < dh; y++) {
            const uint8_t *s0 = src + (size_t)(2 * y) * sstride;
            const uint8_t *s1 = s0 + sstride;
            uint8_t *d = dst + (size_t)y * dstride;
            for (int x = 0; x < dw; x++)
                for (int c = 0; c < cn; c++) {
                    int i = 2 * x * cn + c;
                    d[x * cn + c] = (uint8_t)((s0[i] + s0[i + cn] + s1[i] + s1[i + cn] + 2) >> 2);
                }
        }
        return;
    }
    double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    int *xofs = (int *)malloc(sizeof(int) * dw);
    short *xa = (short *)malloc(sizeof(short) * 2 * dw);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        xa[2 * dx] = ora_sat_short(ora_roundf((1.f - fx) * 2048));
        xa[2 * dx + 1] = ora_sat_short(ora_roundf(fx * 2048));
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        /* vertical: coefficients keep their fraction, row indices are clamped */
        short b0 = ora_sat_short(ora_roundf((1.f - fy) * 2048));
        short b1 = ora_sat_short(ora_roundf(fy * 2048));
        int y0 = ora_min(ora_max(sy, 0), sh - 1), y1 = ora_min(ora_max(sy + 1, 0), sh - 1);
        const uint8_t *s0 = src + (size_t)y0 * sstride, *s1 = src + (size_t)y1 * sstride;
        uint8_t *d = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx], sx1 = ora_min(sx + 1, sw - 1);
            int a0 = xa[2 * dx], a1 = xa[2 * dx + 1];
            for (int c = 0; c < cn; c++) {
                int h0 = s0[sx * cn + c] * a0 + s0[sx1 * cn + c] * a1;
                int h1 = s1[sx * cn + c] * a0 + s1[sx1 * cn + c] * a1;
                d[dx * cn + c] = (uint8_t)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
            }
        }
    }
    free(xofs); free(xa);
}

/* ---------------------------------------------------------------------------------------
 * A.3  equalizeHist (kmsfacedetect.cpp:807, kmseyedetect.cpp:950,964 ...)
 * ------------------------------------------------------------------------------------- */
ORA_API void ora_hist_lut(const int *hist, int total, uint8_t *lut)
{
    int i = 0;
    while (i < 256 && !hist[i]) ++i;
    if (i == 256) { for (int k = 0; k < 256; k++) lut[k] = (uint8_t)k; return; }
    if (hist[i] == total) { for (int k = 0; k < 256; k++) lut[k] = (uint8_t)i; return; }
    float scale = 255.f / (float)(total - hist[i]);
    int sum = 0;
    for (int k = 0; k <= i; k++) lut[k] = 0;
    for (++i; i < 256; ++i) {
        sum += hist[i];
        lut[i] = ora_sat_u8(ora_roundf((float)sum * scale));
    }
}

ORA_API void ora_equalize_hist(const uint8_t *src, int w, int h, int sstride,
                               uint8_t *dst, int dstride)
{
    int hist[256] = {0};
    uint8_t lut[256];
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) hist[src[(size_t)y * sstride + x]]++;
    ora_hist_lut(hist, w * h, lut);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) dst[(size_t)y * dstride + x] = lut[src[(size_t)y * sstride + x]];
}

/* ---------------------------------------------------------------------------------------
 * A.4  INTER_LINEAR_EXACT resize, u8 single channel (the cascade's internal pyramid)
 * ------------------------------------------------------------------------------------- */
static void ora_exact_coefs(int ssize, int dsize, int *ofs, int *c1)
{
    double scale = 1.0 / ((double)dsize / ssize);
    for (int d = 0; d < dsize; d++) {
        double f = scale * (d + 0.5) - 0.5;
        int i = (int)floor(f);
        if (i >= 0 && i < ssize - 1) { ofs[d] = i; c1[d] = ora_round((f - i) * 256.0); }
        else if (i < 0) { ofs[d] = 0; c1[d] = -1; }          /* replicate first sample */
        else { ofs[d] = ssize - 1; c1[d] = -1; }              /* replicate last sample  */
    }
}

ORA_API void ora_resize_linear_exact(const uint8_t *src, int sw, int sh, int sstride,
                                     uint8_t *dst, int dw, int dh, int dstride)
{
    int *xo = (int *)malloc(sizeof(int) * dw), *xc = (int *)malloc(sizeof(int) * dw);
    int *yo = (int *)malloc(sizeof(int) * dh), *yc = (int *)malloc(sizeof(int) * dh);
    ora_exact_coefs(sw, dw, xo, xc);
    ora_exact_coefs(sh, dh, yo, yc);
    for (int dy = 0; dy < dh; dy++) {
        const uint8_t *s0 = src + (size_t)yo[dy] * sstride;
        const uint8_t *s1 = yc[dy] < 0 ? s0 : s0 + sstride;
        uint32_t r1 = yc[dy] < 0 ? 0 : (uint32_t)yc[dy], r0 = 256 - r1;
        for (int dx = 0; dx < dw; dx++) {
            uint32_t h0, h1;
            if (xc[dx] < 0) { h0 = (uint32_t)s0[xo[dx]] * 256; h1 = (uint32_t)s1[xo[dx]] * 256; }
            else {
                uint32_t c1 = (uint32_t)xc[dx], c0 = 256 - c1;
                h0 = s0[xo[dx]] * c0 + s0[xo[dx] + 1] * c1;
                h1 = s1[xo[dx]] * c0 + s1[xo[dx] + 1] * c1;
            }
            dst[(size_t)dy * dstride + dx] = (uint8_t)((h0 * r0 + h1 * r1 + 32768) >> 16);
        }
    }
    free(xo); free(xc); free(yo); free(yc);
}

/* ---------------------------------------------------------------------------------------
 * A.5  integral + squared integral, (h+1) x (w+1), int32 / u32 modulo 2^32
 * ------------------------------------------------------------------------------------- */
ORA_API void ora_integral(const uint8_t *src, int w, int h, int sstride,
                          int32_t *sum, uint32_t *sqsum)
{
    int p = w + 1;
    memset(sum, 0, sizeof(int32_t) * p);
    memset(sqsum, 0, sizeof(uint32_t) * p);
    for (int y = 0; y < h; y++) {
        int32_t rs = 0; uint32_t rq = 0;
        sum[(size_t)(y + 1) * p] = 0; sqsum[(size_t)(y + 1) * p] = 0;
        for (int x = 0; x < w; x++) {
            uint32_t v = src[(size_t)y * sstride + x];
            rs += (int32_t)v; rq += v * v;
            sum[(size_t)(y + 1) * p + x + 1] = sum[(size_t)y * p + x + 1] + rs;
            sqsum[(size_t)(y + 1) * p + x + 1] = sqsum[(size_t)y * p + x + 1] + rq;
        }
    }
}

/* cv::integral's third output (tilted sums), CV_32S: tilted(X,Y) = sum of img(x,y) over y < Y, |x - X + 1| <= Y - y - 1,
 * i.e. the 45-degree triangle whose apex is pixel (X-1, Y-1).  Restated through the three-term recurrence
 *   T(X,Y) = T(X-1,Y-1) + T(X+1,Y-1) - T(X,Y-2) + img(X-1,Y-1) + img(X-1,Y-2)
 * evaluated on a domain widened by h columns on both sides, where the triangles beyond the border are empty and the
 * zero boundary is exact.  Pinned against cv2.integral3 in tests/test_oracle_vs_cv2.py.  tilted: [(h+1)*(w+1)]. */
ORA_API void ora_integral_tilted(const uint8_t *img, int w, int h, int stride, int32_t *tilted)
{
    int pad = h + 1, ew = w + 1 + 2 * pad;                     /* extended X = -pad .. w + pad */
    int32_t *r0 = (int32_t *)calloc((size_t)ew, sizeof(int32_t)), *r1 = (int32_t *)calloc((size_t)ew, sizeof(int32_t)),
            *r2 = (int32_t *)calloc((size_t)ew, sizeof(int32_t));   /* rows Y-2, Y-1, Y */
    for (int x = 0; x <= w; x++) tilted[x] = 0;
    for (int Y = 1; Y <= h; Y++) {
        for (int e = 1; e < ew - 1; e++) {
            int X = e - pad, px = X - 1;
            int32_t a = (px >= 0 && px < w) ? img[(size_t)(Y - 1) * stride + px] : 0;
            int32_t b = (px >= 0 && px < w && Y >= 2) ? img[(size_t)(Y - 2) * stride + px] : 0;
            r2[e] = r1[e - 1] + r1[e + 1] - r0[e] + a + b;
        }
        for (int x = 0; x <= w; x++) tilted[(size_t)Y * (w + 1) + x] = r2[x + pad];
        int32_t *t = r0; r0 = r1; r1 = r2; r2 = t;
    }
    free(r0); free(r1); free(r2);
}

/* ---------------------------------------------------------------------------------------
 * Cascade model (flat arrays; filled from the XML by oracle/oracle.py)
 * ------------------------------------------------------------------------------------- */
typedef struct {
    int win_w, win_h;
    int nstages, nstumps;
    const int *stage_ntrees;     /* [nstages] */
    const float *stage_thr;      /* [nstages] raw XML thresholds */
    const int *stump_feat;       /* [nstumps] feature index */
    const float *stump_thr;      /* [nstumps] */
    const float *stump_left;     /* [nstumps] */
    const float *stump_right;    /* [nstumps] */
    int nfeatures;
    const int *feat_rect;        /* [nfeatures*3*4] x,y,w,h (w==0: unused rect) */
    const float *feat_weight;    /* [nfeatures*3] */
    /* general model: weak classifiers that are trees of more than one node and/or tilted features (OpenCV's
     * predictOrdered path).  general == 0: the stump arrays above describe the whole cascade. */
    int general;
    const int *tree_nnodes;      /* [nstumps] internal nodes per weak classifier */
    const int *node_feat;        /* [sum nnodes] */
    const float *node_thr;
    const int *node_left;        /* > 0: next node of the same tree; <= 0: leaf -idx of the same tree */
    const int *node_right;
    const float *leaves;         /* [sum (nnodes + 1)] */
    const unsigned char *feat_tilted;   /* [nfeatures] */
    /* LBP model (OpenCV featureType LBP, predictCategoricalStump / predictCategorical): a feature is ONE cell rect
     * (feat_rect[12 f .. 12 f + 3]) spanning a 3 x 3 grid of such cells; a node holds a 256-bit subset instead of a
     * threshold.  No variance normalisation. */
    int lbp;
    const int *node_subset;      /* [sum nnodes * 8] */
} ora_cascade;

#define ORA_DEPTH_VARREJ  (-100)    /* OpenCV result -1 from the variance test */
#define ORA_DEPTH_SKIPPED (-32768)  /* window never evaluated (stage-0 skip rule) */

/* A.4 scale list */
ORA_API int ora_scales(int W, int H, int win_w, int win_h, double scale_factor,
                       int min_w, int min_h, int max_w, int max_h, float *scales, int cap)
{
    int n = 0;
    if (max_w == 0 || max_h == 0) { max_w = W; max_h = H; }
    for (double f = 1;; f *= scale_factor) {
        int ww = ora_round(win_w * f), wh = ora_round(win_h * f);
        if (ww > max_w || wh > max_h || ww > W || wh > H) break;
        if (ww < min_w || wh < min_h) continue;
        if (n < cap) scales[n] = (float)f;
        n++;
        if (scale_factor <= 1.0) break;   /* guard: OpenCV would loop forever */
    }
    return n;
}

ORA_API void ora_level_size(int W, int H, float sc, int *lw, int *lh)
{
    *lw = ora_roundf((float)W / sc);
    *lh = ora_roundf((float)H / sc);
}

/* A.6: evaluate one window at (x,y) of a level given its integrals.  Returns the depth code. */
static int ora_run_at(const ora_cascade *c, const int32_t *sum, const uint32_t *sq, int p,
                      int x, int y)
{
    int nw = c->win_w - 2, nh = c->win_h - 2;
    const int32_t *s = sum + (size_t)(y + 1) * p + (x + 1);
    const uint32_t *q = sq + (size_t)(y + 1) * p + (x + 1);
    int valsum = s[0] - s[nw] - s[(size_t)nh * p] + s[(size_t)nh * p + nw];
    uint32_t valsq = q[0] - q[nw] - q[(size_t)nh * p] + q[(size_t)nh * p + nw];
    double area = (double)nw * nh;
    double nf = area * valsq - (double)valsum * valsum;
    float vnf;
    if (nf > 0.) {
        nf = sqrt(nf);
        vnf = (float)(1. / nf);
        if (!(area * vnf < 1e-1)) return ORA_DEPTH_VARREJ;
    } else
        return ORA_DEPTH_VARREJ;

    const int32_t *w0 = sum + (size_t)y * p + x;
    int si = 0;
    for (int st = 0; st < c->nstages; st++) {
        double tmp = 0.;   /* Haar stump path accumulates leaves in double (probed: tests/test_oracle_vs_cv2.py) */
        for (int i = 0; i < c->stage_ntrees[st]; i++, si++) {
            int f = c->stump_feat[si];
            const int *r = c->feat_rect + (size_t)f * 12;
            const float *wt = c->feat_weight + (size_t)f * 3;
            float v = 0.f;
            for (int k = 0; k < 3; k++) {
                if (k == 2 && wt[2] == 0.f) break;
                const int32_t *a = w0 + (size_t)r[4 * k + 1] * p + r[4 * k];
                int rs = a[0] - a[r[4 * k + 2]] - a[(size_t)r[4 * k + 3] * p]
                       + a[(size_t)r[4 * k + 3] * p + r[4 * k + 2]];
                float t = wt[k] * (float)rs;
                v = (k == 0) ? t : v + t;
            }
            v *= vnf;
            tmp += (double)((v < c->stump_thr[si]) ? c->stump_left[si] : c->stump_right[si]);
        }
        float thr = c->stage_thr[st] - 1e-5f;
        if (tmp < (double)thr) return -st;
    }
    return 1;
}

/* Feature value on the upright or the tilted integral.  Tilted rect (x, y, w, h): corners (x, y), (x - h, y + h),
 * (x + w, y + w), (x + w - h, y + w + h) of the tilted integral, combined p0 - p1 - p2 + p3 (OpenCV CV_TILTED_OFS). */
static float ora_feature(const ora_cascade *c, const int32_t *w0, const int32_t *t0, int p, int f)
{
    const int *r = c->feat_rect + (size_t)f * 12;
    const float *wt = c->feat_weight + (size_t)f * 3;
    int tilted = c->feat_tilted && c->feat_tilted[f];
    float v = 0.f;
    for (int k = 0; k < 3; k++) {
        if (k == 2 && wt[2] == 0.f) break;
        int x = r[4 * k], y = r[4 * k + 1], w = r[4 * k + 2], h = r[4 * k + 3], rs;
        if (tilted)
            rs = t0[(size_t)y * p + x] - t0[(size_t)(y + h) * p + x - h] - t0[(size_t)(y + w) * p + x + w]
               + t0[(size_t)(y + w + h) * p + x + w - h];
        else {
            const int32_t *a = w0 + (size_t)y * p + x;
            rs = a[0] - a[w] - a[(size_t)h * p] + a[(size_t)h * p + w];
        }
        float t = wt[k] * (float)rs;
        v = (k == 0) ? t : v + t;
    }
    return v;
}

/* OpenCV predictOrdered<HaarEvaluator>: every weak classifier is walked from its root; the float feature value times
 * the variance factor is compared with the node threshold; leaves accumulate in double. */
static int ora_run_at_general(const ora_cascade *c, const int32_t *sum, const uint32_t *sq, const int32_t *tilt, int p,
                              int x, int y)
{
    int nw = c->win_w - 2, nh = c->win_h - 2;
    const int32_t *s = sum + (size_t)(y + 1) * p + (x + 1);
    const uint32_t *q = sq + (size_t)(y + 1) * p + (x + 1);
    int valsum = s[0] - s[nw] - s[(size_t)nh * p] + s[(size_t)nh * p + nw];
    uint32_t valsq = q[0] - q[nw] - q[(size_t)nh * p] + q[(size_t)nh * p + nw];
    double area = (double)nw * nh;
    double nf = area * valsq - (double)valsum * valsum;
    float vnf;
    if (nf > 0.) {
        nf = sqrt(nf);
        vnf = (float)(1. / nf);
        if (!(area * vnf < 1e-1)) return ORA_DEPTH_VARREJ;
    } else
        return ORA_DEPTH_VARREJ;
    const int32_t *w0 = sum + (size_t)y * p + x, *t0 = tilt ? tilt + (size_t)y * p + x : NULL;
    int ti = 0, node0 = 0, leaf0 = 0;
    for (int st = 0; st < c->nstages; st++) {
        double tmp = 0.;
        for (int i = 0; i < c->stage_ntrees[st]; i++, ti++) {
            int idx = 0;
            do {
                int n = node0 + idx;
                float v = ora_feature(c, w0, t0, p, c->node_feat[n]) * vnf;
                idx = v < c->node_thr[n] ? c->node_left[n] : c->node_right[n];
            } while (idx > 0);
            tmp += (double)c->leaves[leaf0 - idx];
            node0 += c->tree_nnodes[ti]; leaf0 += c->tree_nnodes[ti] + 1;
        }
        float thr = c->stage_thr[st] - 1e-5f;
        if (tmp < (double)thr) return -st;
    }
    return 1;
}

/* OpenCV LBPEvaluator::OptFeature::calc: the 8 neighbour cells of the 3 x 3 grid compared (>=) with the centre cell,
 * clockwise from the top-left, most significant bit first. */
static int ora_lbp_code(const int32_t *w0, int p, const int *r)
{
    int x = r[0], y = r[1], w = r[2], h = r[3];
    int c[3][3];
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) {
            const int32_t *a = w0 + (size_t)(y + j * h) * p + x + i * w;
            c[j][i] = a[0] - a[w] - a[(size_t)h * p] + a[(size_t)h * p + w];
        }
    int cv = c[1][1];
    return (c[0][0] >= cv ? 128 : 0) | (c[0][1] >= cv ? 64 : 0) | (c[0][2] >= cv ? 32 : 0) | (c[1][2] >= cv ? 16 : 0) |
           (c[2][2] >= cv ? 8 : 0) | (c[2][1] >= cv ? 4 : 0) | (c[2][0] >= cv ? 2 : 0) | (c[1][0] >= cv ? 1 : 0);
}

ORA_API int ora_lbp_code_at(const ora_cascade *c, const int32_t *sum, int p, int x, int y, int f)
{
    return ora_lbp_code(sum + (size_t)y * p + x, p, c->feat_rect + (size_t)f * 12);
}

/* OpenCV predictCategorical<LBPEvaluator> (and its stump form, which walks one-node trees the same way): the code of
 * the node's feature picks the left child when its bit is set in the node's subset; leaves accumulate in double
 * (probed against cv2: tests/test_oracle_vs_cv2.py); no variance test — setWindow only checks the window bounds. */
static int ora_run_at_lbp(const ora_cascade *c, const int32_t *sum, int p, int x, int y)
{
    const int32_t *w0 = sum + (size_t)y * p + x;
    int ti = 0, node0 = 0, leaf0 = 0;
    for (int st = 0; st < c->nstages; st++) {
        double tmp = 0.;
        for (int i = 0; i < c->stage_ntrees[st]; i++, ti++) {
            int idx = 0;
            do {
                int n = node0 + idx;
                int code = ora_lbp_code(w0, p, c->feat_rect + (size_t)c->node_feat[n] * 12);
                const int *subset = c->node_subset + (size_t)n * 8;
                idx = (subset[code >> 5] & (1 << (code & 31))) ? c->node_left[n] : c->node_right[n];
            } while (idx > 0);
            tmp += (double)c->leaves[leaf0 - idx];
            node0 += c->tree_nnodes[ti]; leaf0 += c->tree_nnodes[ti] + 1;
        }
        float thr = c->stage_thr[st] - 1e-5f;
        if (tmp < (double)thr) return -st;
    }
    return 1;
}

/* Debug tap for the pin tests: normalised value of feature `f` at window (x,y); returns 0 and
 * leaves *out untouched when the variance test rejects the window. */
ORA_API int ora_feature_value(const ora_cascade *c, const int32_t *sum, const uint32_t *sq, int p,
                              int x, int y, int f, float *out)
{
    int nw = c->win_w - 2, nh = c->win_h - 2;
    const int32_t *s = sum + (size_t)(y + 1) * p + (x + 1);
    const uint32_t *q = sq + (size_t)(y + 1) * p + (x + 1);
    int valsum = s[0] - s[nw] - s[(size_t)nh * p] + s[(size_t)nh * p + nw];
    uint32_t valsq = q[0] - q[nw] - q[(size_t)nh * p] + q[(size_t)nh * p + nw];
    double area = (double)nw * nh, nf = area * valsq - (double)valsum * valsum;
    if (!(nf > 0.)) return 0;
    float vnf = (float)(1. / sqrt(nf));
    if (!(area * vnf < 1e-1)) return 0;
    const int32_t *w0 = sum + (size_t)y * p + x;
    const int *r = c->feat_rect + (size_t)f * 12;
    const float *wt = c->feat_weight + (size_t)f * 3;
    float v = 0.f;
    for (int k = 0; k < 3; k++) {
        if (k == 2 && wt[2] == 0.f) break;
        const int32_t *a = w0 + (size_t)r[4 * k + 1] * p + r[4 * k];
        int rs = a[0] - a[r[4 * k + 2]] - a[(size_t)r[4 * k + 3] * p] + a[(size_t)r[4 * k + 3] * p + r[4 * k + 2]];
        float t = wt[k] * (float)rs;
        v = (k == 0) ? t : v + t;
    }
    *out = v * vnf;
    return 1;
}

/* A.6 window loop over one level.  depth: [ny*nx] int16 (may be NULL), nx/ny in ystep units.
 * cand: [cap*4] candidate rects in level order (may be NULL).  Returns #passes. */
/* Rows OpenCV actually visits: detectMultiScale cuts every level into `nstripes` horizontal
 * stripes of max(ceil((ry/ystep)/nstripes),1)*ystep rows, nstripes = ceil(rx_of_first_scale/32),
 * and clamps the last one to ry — so with ystep 2 and odd ry the final row is visited only when
 * nstripes does not divide ry/2 (probed with an always-pass cascade). */
ORA_API int ora_row_limit(int ry, int ystep, int nstripes)
{
    if (nstripes < 1) nstripes = 1;
    int stripe = ora_max((ry / ystep + nstripes - 1) / nstripes, 1) * ystep;
    long long lim = (long long)stripe * nstripes;
    return lim < ry ? (int)lim : ry;
}

ORA_API int ora_has_tilted(const ora_cascade *c)
{
    if (!c->feat_tilted) return 0;
    for (int f = 0; f < c->nfeatures; f++)
        if (c->feat_tilted[f]) return 1;
    return 0;
}

ORA_API int ora_eval_level(const ora_cascade *c, const int32_t *sum, const uint32_t *sq, const int32_t *tilt,
                           int lw, int lh, int ystep, float sc, int nstripes, int16_t *depth,
                           int *cand, int cap, int *ncand_io)
{
    int p = lw + 1;
    int rx = lw + 1 - c->win_w, ry = lh + 1 - c->win_h;
    if (rx <= 0 || ry <= 0) return 0;
    ry = ora_row_limit(ry, ystep, nstripes);
    int nx = (rx + ystep - 1) / ystep;
    int ww = ora_roundf(c->win_w * sc), wh = ora_roundf(c->win_h * sc);
    int npass = 0, iy = 0;
    for (int y = 0; y < ry; y += ystep, iy++) {
        if (depth) for (int i = 0; i < nx; i++) depth[(size_t)iy * nx + i] = ORA_DEPTH_SKIPPED;
        for (int x = 0; x < rx; x += ystep) {
            int r = c->lbp ? ora_run_at_lbp(c, sum, p, x, y)
                  : c->general ? ora_run_at_general(c, sum, sq, tilt, p, x, y) : ora_run_at(c, sum, sq, p, x, y);
            if (depth) depth[(size_t)iy * nx + x / ystep] = (int16_t)r;
            if (r > 0) {
                npass++;
                if (cand && *ncand_io < cap) {
                    int *o = cand + 4 * (size_t)(*ncand_io);
                    o[0] = ora_roundf(x * sc); o[1] = ora_roundf(y * sc); o[2] = ww; o[3] = wh;
                }
                if (ncand_io) (*ncand_io)++;
            }
            if (r == 0) x += ystep;
        }
    }
    return npass;
}

/* ---------------------------------------------------------------------------------------
 * A.7  groupRectangles(list, thr, eps); in/out rects [n*4]; returns new count; weights opt.
 * ------------------------------------------------------------------------------------- */
static int ora_similar(const int *a, const int *b, double eps)
{
    double delta = eps * (ora_min(a[2], b[2]) + ora_min(a[3], b[3])) * 0.5;
    return abs(a[0] - b[0]) <= delta && abs(a[1] - b[1]) <= delta &&
           abs(a[0] + a[2] - b[0] - b[2]) <= delta && abs(a[1] + a[3] - b[1] - b[3]) <= delta;
}

ORA_API int ora_group_rectangles(int *rects, int n, int thr, double eps, int *weights)
{
    if (thr <= 0 || n == 0) {
        if (weights) for (int i = 0; i < n; i++) weights[i] = 1;
        return n;
    }
    /* cv::partition: disjoint-set forest over all ordered pairs, classes numbered by first member */
    int *parent = (int *)malloc(sizeof(int) * n), *rank = (int *)calloc(n, sizeof(int));
    int *labels = (int *)malloc(sizeof(int) * n);
    for (int i = 0; i < n; i++) parent[i] = -1;
    for (int i = 0; i < n; i++) {
        int root = i;
        while (parent[root] >= 0) root = parent[root];
        for (int j = 0; j < n; j++) {
            if (i == j || !ora_similar(rects + 4 * i, rects + 4 * j, eps)) continue;
            int root2 = j;
            while (parent[root2] >= 0) root2 = parent[root2];
            if (root2 != root) {
                if (rank[root] > rank[root2]) parent[root2] = root;
                else { parent[root] = root2; rank[root2] += rank[root] == rank[root2]; root = root2; }
                int k = j, pp;
                while ((pp = parent[k]) >= 0) { parent[k] = root; k = pp; }
                k = i;
                while ((pp = parent[k]) >= 0) { parent[k] = root; k = pp; }
            }
        }
    }
    int ncls = 0;
    for (int i = 0; i < n; i++) {
        int root = i;
        while (parent[root] >= 0) root = parent[root];
        if (rank[root] >= 0) rank[root] = ~ncls++;
        labels[i] = ~rank[root];
    }
    int *acc = (int *)calloc((size_t)ncls * 4, sizeof(int)), *cnt = (int *)calloc(ncls, sizeof(int));
    for (int i = 0; i < n; i++) {
        int cl = labels[i];
        for (int k = 0; k < 4; k++) acc[4 * cl + k] += rects[4 * i + k];
        cnt[cl]++;
    }
    for (int i = 0; i < ncls; i++) {
        float s = 1.f / cnt[i];
        for (int k = 0; k < 4; k++) acc[4 * i + k] = ora_roundf(acc[4 * i + k] * s);
    }
    int m = 0;
    for (int i = 0; i < ncls; i++) {
        const int *r1 = acc + 4 * i; int n1 = cnt[i], j;
        if (n1 <= thr) continue;
        for (j = 0; j < ncls; j++) {
            int n2 = cnt[j];
            if (j == i || n2 <= thr) continue;
            const int *r2 = acc + 4 * j;
            int dx = ora_round(r2[2] * eps), dy = ora_round(r2[3] * eps);
            if (r1[0] >= r2[0] - dx && r1[1] >= r2[1] - dy &&
                r1[0] + r1[2] <= r2[0] + r2[2] + dx && r1[1] + r1[3] <= r2[1] + r2[3] + dy &&
                (n2 > ora_max(3, n1) || n1 < 3))
                break;
        }
        if (j == ncls) {
            memcpy(rects + 4 * m, r1, sizeof(int) * 4);
            if (weights) weights[m] = n1;
            m++;
        }
    }
    free(parent); free(rank); free(labels); free(acc); free(cnt);
    return m;
}

/* ---------------------------------------------------------------------------------------
 * CascadeClassifier::detectMultiScale on a gray image (kmsfacedetect.cpp:809-811).
 * out: [cap*4] rects; depth_maps / integrals are exposed through ora_eval_level for tests.
 * min_neighbors==0 returns the raw candidates in scale -> y -> x order (A.9), clipped (A.8).
 * ------------------------------------------------------------------------------------- */
ORA_API int ora_detect_multiscale(const ora_cascade *c, const uint8_t *gray, int W, int H,
                                  int stride, double scale_factor, int min_neighbors,
                                  int min_w, int min_h, int max_w, int max_h,
                                  int *out, int cap, int *weights, long long *nwindows)
{
    float scales[256];
    int ns = ora_scales(W, H, c->win_w, c->win_h, scale_factor, min_w, min_h, max_w, max_h, scales, 256);
    if (ns > 256) ns = 256;
    int ncand = 0, capc = 1 << 20;
    int *cand = (int *)malloc(sizeof(int) * 4 * (size_t)capc);
    long long nwin = 0;
    int nstripes = 1;
    if (ns > 0) {
        int lw0, lh0;
        ora_level_size(W, H, scales[0], &lw0, &lh0);
        nstripes = (ora_max(lw0 + 1 - c->win_w, 0) + 31) / 32;
    }
    for (int k = 0; k < ns; k++) {
        int lw, lh;
        ora_level_size(W, H, scales[k], &lw, &lh);
        if (lw + 1 - c->win_w <= 0 || lh + 1 - c->win_h <= 0) continue;
        uint8_t *lvl = (uint8_t *)malloc((size_t)lw * lh);
        int32_t *sum = (int32_t *)malloc(sizeof(int32_t) * (size_t)(lw + 1) * (lh + 1));
        uint32_t *sq = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(lw + 1) * (lh + 1));
        if (lw == W && lh == H)
            for (int y = 0; y < H; y++) memcpy(lvl + (size_t)y * lw, gray + (size_t)y * stride, W);
        else
            ora_resize_linear_exact(gray, W, H, stride, lvl, lw, lh, lw);
        ora_integral(lvl, lw, lh, lw, sum, sq);
        int32_t *tilt = NULL;
        if (c->general && ora_has_tilted(c)) {
            tilt = (int32_t *)malloc(sizeof(int32_t) * (size_t)(lw + 1) * (lh + 1));
            ora_integral_tilted(lvl, lw, lh, lw, tilt);
        }
        int ystep = scales[k] >= 2.f ? 1 : 2;
        ora_eval_level(c, sum, sq, tilt, lw, lh, ystep, scales[k], nstripes, NULL, cand, capc, &ncand);
        free(tilt);
        nwin += (long long)((lw + 1 - c->win_w + ystep - 1) / ystep) *
                ((ora_row_limit(lh + 1 - c->win_h, ystep, nstripes) + ystep - 1) / ystep);
        free(lvl); free(sum); free(sq);
    }
    if (nwindows) *nwindows = nwin;
    if (ncand > capc) ncand = capc;
    int n = ora_group_rectangles(cand, ncand, min_neighbors, 0.2, weights && ncand <= cap ? weights : NULL);
    /* A.8 clip to the image; empty intersections are dropped (clipObjects) */
    int m = 0;
    for (int i = 0; i < n; i++) {
        int x0 = ora_max(cand[4 * i], 0), y0 = ora_max(cand[4 * i + 1], 0);
        int x1 = ora_min(cand[4 * i] + cand[4 * i + 2], W), y1 = ora_min(cand[4 * i + 1] + cand[4 * i + 3], H);
        if (x1 <= x0 || y1 <= y0) continue;
        if (m < cap) {
            out[4 * m] = x0; out[4 * m + 1] = y0; out[4 * m + 2] = x1 - x0; out[4 * m + 3] = y1 - y0;
            if (weights && m != i) weights[m] = weights[i];
        }
        m++;
    }
    free(cand);
    return m;
}

/* The face element's whole hot block, kmsfacedetect.cpp:770-811: integer scale, resize the
 * BGR frame, gray, equalizeHist, detectMultiScale(sf, 3, 0, (cols/20, rows/20)). */
ORA_API int ora_face_process(const ora_cascade *c, const uint8_t *bgr, int W, int H, int stride,
                             int width_to_process, double scale_factor, int min_neighbors,
                             int min_w, int min_h, int *out, int cap, uint8_t *gray_eq_out)
{
    int iscale = W / width_to_process;               /* kmsfacedetect.cpp:304 integer division */
    double scale = iscale;
    int rows = H, cols = W;
    if (iscale > 0 && ora_round(H / scale) > 0) rows = ora_round(H / scale); else scale = 1;
    if (iscale > 0 && ora_round(W / scale) > 0) cols = ora_round(W / scale); else scale = 1;
    uint8_t *aux = (uint8_t *)malloc((size_t)rows * cols * 3);
    uint8_t *gray = (uint8_t *)malloc((size_t)rows * cols);
    ora_resize_linear(bgr, W, H, stride, 3, aux, cols, rows, cols * 3);
    ora_bgr2gray(aux, cols, rows, cols * 3, 3, gray, cols);
    ora_equalize_hist(gray, cols, rows, cols, gray, cols);
    if (gray_eq_out) memcpy(gray_eq_out, gray, (size_t)rows * cols);
    if (min_w < 0) { min_w = cols / 20; min_h = rows / 20; }   /* kmsfacedetect.cpp:811 */
    int n = ora_detect_multiscale(c, gray, cols, rows, cols, scale_factor, min_neighbors,
                                  min_w, min_h, 0, 0, out, cap, NULL, NULL);
    free(aux); free(gray);
    return n;
}

/* ---------------------------------------------------------------------------------------
 * Tracker (gstnubotracker.cpp:356-377): gray, absdiff, threshold, updateMotionHistory,
 * segmentMotion.  OpenCV 2.4 motempl.cpp semantics (A.9); cv2.motempl is absent from this
 * image, so this part is checked against cv2.floodFill / connectedComponentsWithStats via
 * the equivalence argument of SURVEY.md §3.4 (tests/test_oracle_vs_cv2.py).
 * ------------------------------------------------------------------------------------- */
ORA_API void ora_absdiff_threshold(const uint8_t *a, const uint8_t *b, int n, int thr, uint8_t *mask)
{
    for (int i = 0; i < n; i++) {
        int d = abs((int)a[i] - (int)b[i]);
        mask[i] = d > thr ? 255 : 0;
    }
}

ORA_API void ora_update_mhi(const uint8_t *silh, float *mhi, int n, double ts, double duration)
{
    float fts = (float)ts, del = (float)(ts - duration);
    for (int i = 0; i < n; i++) {
        if (silh[i]) mhi[i] = fts;
        else if (mhi[i] < del) mhi[i] = 0.f;
    }
}

/* segmentMotion: seeds where mhi==ts in raster order, 4-connected floating-range flood fill
 * (|neighbour - current| <= seg_thresh), zeros replaced by FLT_MAX*0.1.  labels: int32 [w*h]
 * (0 = none, k = k-th component).  rects: [cap*4] x,y,w,h.  Returns #components. */
ORA_API int ora_segment_motion(const float *mhi, int w, int h, double ts, double seg_thresh,
                               int32_t *labels, int *rects, int cap)
{
    float fts = (float)ts, thr = (float)seg_thresh, stub = FLT_MAX * 0.1f;
    int n = w * h, ncomp = 0;
    float *m = (float *)malloc(sizeof(float) * (size_t)n);
    int *stack = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) { m[i] = mhi[i] == 0.f ? stub : mhi[i]; labels[i] = 0; }
    for (int i = 0; i < n; i++) {
        if (m[i] != fts || labels[i]) continue;
        int sp = 0, x0 = w, y0 = h, x1 = -1, y1 = -1;
        ncomp++;
        labels[i] = ncomp; stack[sp++] = i;
        while (sp) {
            int pidx = stack[--sp], px = pidx % w, py = pidx / w;
            if (px < x0) x0 = px; if (px > x1) x1 = px;
            if (py < y0) y0 = py; if (py > y1) y1 = py;
            float v = m[pidx];
            int nb[4] = { px > 0 ? pidx - 1 : -1, px < w - 1 ? pidx + 1 : -1,
                          py > 0 ? pidx - w : -1, py < h - 1 ? pidx + w : -1 };
            for (int k = 0; k < 4; k++) {
                int q = nb[k];
                if (q < 0 || labels[q]) continue;
                float d = m[q] - v;
                if (d >= -thr && d <= thr) { labels[q] = ncomp; stack[sp++] = q; }
            }
        }
        if (ncomp <= cap) {
            int *r = rects + 4 * (size_t)(ncomp - 1);
            r[0] = x0; r[1] = y0; r[2] = x1 - x0 + 1; r[3] = y1 - y0 + 1;
        }
    }
    free(m); free(stack);
    return ncomp;
}

/* gstnubotracker.cpp:119-200: calc_dist / __merge / __join_objects (in place; returns count) */
static float ora_calc_dist(const int *a, const int *b)
{
    int c1x = a[0] + a[2] / 2, c1y = a[1] + a[3] / 2, c2x = b[0] + b[2] / 2, c2y = b[1] + b[3] / 2;
    return (float)sqrt((double)((c1x - c2x) * (c1x - c2x) + (c1y - c2y) * (c1y - c2y)));
}
static int ora_pt_inside(int px, int py, const int *r)
{
    return r[0] <= px && px < r[0] + r[2] && r[1] <= py && py < r[1] + r[3];
}
static void ora_merge(const int *r1, const int *r2, int *out)
{
    int b1x = r1[0] + r1[2], b1y = r1[1] + r1[3], b2x = r2[0] + r2[2], b2y = r2[1] + r2[3];
    if (ora_pt_inside(r2[0], r2[1], r1) && ora_pt_inside(b2x, b2y, r1)) { memcpy(out, r1, 16); return; }
    if (ora_pt_inside(r1[0], r1[1], r2) && ora_pt_inside(b1x, b1y, r2)) { memcpy(out, r2, 16); return; }
    int tx = ora_min(r1[0], r2[0]), ty = ora_min(r1[1], r2[1]);
    int bx = ora_max(b1x, b2x), by = ora_max(b1y, b2y);
    out[0] = tx; out[1] = ty; out[2] = bx - tx; out[3] = by - ty;
}
ORA_API int ora_join_objects(int *r, int n, int min_area, long max_area, int distance)
{
#define AREA_OK(i) ((long)r[4*(i)+2] * r[4*(i)+3] > min_area && (long)r[4*(i)+2] * r[4*(i)+3] < max_area)
    for (int a = n - 1; a >= 0; a--) {
        if (AREA_OK(a)) {
            for (int b = a - 1; b >= 0; b--) {
                if (AREA_OK(b) && (float)distance > ora_calc_dist(r + 4 * a, r + 4 * b)) {
                    int m[4];
                    ora_merge(r + 4 * a, r + 4 * b, m);
                    memcpy(r + 4 * b, m, 16);
                    memmove(r + 4 * a, r + 4 * (a + 1), sizeof(int) * 4 * (size_t)(n - a - 1));
                    n--;
                    break;
                }
            }
        } else {
            memmove(r + 4 * a, r + 4 * (a + 1), sizeof(int) * 4 * (size_t)(n - a - 1));
            n--;
        }
    }
#undef AREA_OK
    return n;
}

/* One tracker frame (gstnubotracker.cpp:356-380).  prev_gray/mhi are state, updated in place.
 * first_frame!=0 reproduces num_frames==0 (only the gray copy happens). */
ORA_API int ora_tracker_process(const uint8_t *bgra, int w, int h, int stride, int first_frame,
                                uint8_t *prev_gray, float *mhi, double ts, int threshold,
                                int min_area, long max_area, int distance,
                                int *rects, int cap, int *nraw, uint8_t *mask_out)
{
    int n = w * h, nr = 0;
    uint8_t *gray = (uint8_t *)malloc(n), *mask = (uint8_t *)malloc(n);
    ora_bgr2gray(bgra, w, h, stride, 4, gray, w);
    if (!first_frame) {
        int32_t *labels = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
        ora_absdiff_threshold(gray, prev_gray, n, threshold, mask);
        ora_update_mhi(mask, mhi, n, ts, 0.2);          /* MHI_DURATION, gstnubotracker.cpp:28 */
        nr = ora_segment_motion(mhi, w, h, ts, 32, labels, rects, cap);   /* SEGMENTATION 32, :31 */
        if (nraw) *nraw = nr;
        if (nr > cap) nr = cap;
        nr = ora_join_objects(rects, nr, min_area, max_area, distance);
        if (mask_out) memcpy(mask_out, mask, n);
        free(labels);
    } else if (nraw) *nraw = 0;
    memcpy(prev_gray, gray, n);
    free(gray); free(mask);
    return nr;
}
