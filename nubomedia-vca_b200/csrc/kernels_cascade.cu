// kernels_cascade.cu — K6 + K7: Haar cascade stage evaluation over every level of the pyramid.
// Replaces the inside of cv::CascadeClassifier::detectMultiScale (kmsfacedetect.cpp:809-811 and the
// analogous calls of the eye/mouth/nose/ear elements).  Arithmetic per SURVEY.md A.6 plus the three
// facts the oracle probes added (oracle/nubo_oracle.c): leaves accumulate in DOUBLE, the stage
// threshold is (float)xml - 1e-5f, a window is valid iff area * (float)(1/sqrt(nf)) < 0.1.
// No FMA contraction anywhere on the float path: explicit __fmul_rn / __fadd_rn.
//
// Pass structure (all levels per launch):
//   k_stage0        one lane per window, 32 consecutive windows of a row per warp: variance
//                   normalisation + stage 0, results as two ballot bit-words per 32 windows.
//   k_skip_compact  one warp per window row: resolves OpenCV's "skip the next window after a stage-0
//                   reject" rule (a serial automaton along x) on the bit-words, and compacts the
//                   windows still alive into a global queue with warp-aggregated atomics.
//   k_queue_stages  drains the queue: remaining stages with early exit; passes become candidates.
#include "internal.h"

__device__ __forceinline__ int find_level_c(const PlanDev *__restrict__ plan, int idx, int LevelDesc::*first)
{
    int n = plan->nlevels, l = 0;
    while (l + 1 < n && plan->lv[l + 1].*first <= idx) l++;
    return l;
}

struct LevelView {
    const uint32_t *sum;
    int pitch, plane, ys;
};

__device__ __forceinline__ int corner(const LevelView &v, int dx, int dy)
{
    return dy * v.pitch + (v.ys == 2 ? (dx & 1) * v.plane + (dx >> 1) : dx);
}

__device__ __forceinline__ int rect_sum(const uint32_t *__restrict__ wb, const LevelView &v, uint32_t packed)
{
    int x = packed & 255, y = (packed >> 8) & 255, w = (packed >> 16) & 255, h = packed >> 24;
    uint32_t a = __ldg(wb + corner(v, x, y)), b = __ldg(wb + corner(v, x + w, y));
    uint32_t c = __ldg(wb + corner(v, x, y + h)), d = __ldg(wb + corner(v, x + w, y + h));
    return (int)(a - b - c + d);
}

struct StumpRegs {
    uint32_t r0, r1, r2;
    float w0, w1, w2, thr, left, right;
};

__device__ __forceinline__ StumpRegs load_stump(const DevStump *__restrict__ s)
{
    const uint4 *p = reinterpret_cast<const uint4 *>(s);
    uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    StumpRegs o;
    o.r0 = a.x; o.r1 = a.y; o.r2 = a.z;
    o.w0 = __uint_as_float(a.w); o.w1 = __uint_as_float(b.x); o.w2 = __uint_as_float(b.y);
    o.thr = __uint_as_float(b.z); o.left = __uint_as_float(b.w); o.right = __uint_as_float(c.x);
    return o;
}

// value of one weak classifier at the window whose top-left integral element is wb
__device__ __forceinline__ float stump_leaf(const uint32_t *__restrict__ wb, const LevelView &v, const StumpRegs &s,
                                            float vnf)
{
    float f = __fmul_rn(s.w0, __int2float_rn(rect_sum(wb, v, s.r0)));
    f = __fadd_rn(f, __fmul_rn(s.w1, __int2float_rn(rect_sum(wb, v, s.r1))));
    if (s.w2 != 0.f) f = __fadd_rn(f, __fmul_rn(s.w2, __int2float_rn(rect_sum(wb, v, s.r2))));
    f = __fmul_rn(f, vnf);
    return f < s.thr ? s.left : s.right;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_stage0(const PlanDev *__restrict__ plan, int total_chunks, const DevCascade *__restrict__ meta,
         const DevStump *__restrict__ stumps, const uint32_t *__restrict__ sum, const uint32_t *__restrict__ sq,
         float *__restrict__ vnf_out, uint32_t *__restrict__ bits_fail, uint32_t *__restrict__ bits_ok)
{
    int lane = threadIdx.x & 31;
    int chunk = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (chunk >= total_chunks) return;
    int l = find_level_c(plan, chunk, &LevelDesc::chunk0);
    const LevelDesc &L = plan->lv[l];
    int rel = chunk - L.chunk0, iy = rel / L.nxw, cx = rel - iy * L.nxw;
    int ix = cx * 32 + lane;
    bool valid = ix < L.nx;
    int ixc = valid ? ix : L.nx - 1;
    LevelView v{sum + L.iofs, L.ipitch, L.iplane, L.ystep};
    size_t base = (size_t)iy * L.ystep * L.ipitch + ixc;
    const uint32_t *wb = v.sum + base, *qb = sq + L.iofs + base;

    int ww = plan->win_w, wh = plan->win_h;
    int c00 = corner(v, 1, 1), c10 = corner(v, ww - 1, 1), c01 = corner(v, 1, wh - 1), c11 = corner(v, ww - 1, wh - 1);
    int valsum = (int)(__ldg(wb + c00) - __ldg(wb + c10) - __ldg(wb + c01) + __ldg(wb + c11));
    uint32_t valsq = __ldg(qb + c00) - __ldg(qb + c10) - __ldg(qb + c01) + __ldg(qb + c11);
    double area = (double)((ww - 2) * (wh - 2));
    double nf = __dsub_rn(__dmul_rn(area, (double)valsq), __dmul_rn((double)valsum, (double)valsum));
    float vnf = 0.f;
    bool ok = false;
    if (nf > 0.) {
        vnf = __double2float_rn(__ddiv_rn(1.0, __dsqrt_rn(nf)));
        ok = __dmul_rn(area, (double)vnf) < 1e-1;
    }
    bool fail = false;
    if (ok) {
        double tmp = 0.;
        int n0 = meta->stage_first[1];
        for (int i = 0; i < n0; i++) {
            StumpRegs s = load_stump(stumps + i);
            tmp = __dadd_rn(tmp, (double)stump_leaf(wb, v, s, vnf));
        }
        fail = tmp < (double)meta->stage_thr[0];
    }
    ok = ok && valid;
    fail = fail && ok;
    uint32_t m_ok = __ballot_sync(0xffffffffu, ok), m_fail = __ballot_sync(0xffffffffu, fail);
    if (lane == 0) {
        bits_ok[L.bofs + iy * L.nxw + cx] = m_ok;
        bits_fail[L.bofs + iy * L.nxw + cx] = m_fail;
    }
    if (valid) vnf_out[L.wofs + iy * L.nx + ix] = ok ? vnf : 0.f;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_skip_compact(const PlanDev *__restrict__ plan, int total_rows, const float *__restrict__ vnf,
               const uint32_t *__restrict__ bits_fail, const uint32_t *__restrict__ bits_ok, uint2 *__restrict__ queue,
               int *__restrict__ counters, int queue_cap, int16_t *__restrict__ depth)
{
    int lane = threadIdx.x & 31;
    int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= total_rows) return;
    int l = find_level_c(plan, row, &LevelDesc::row0);
    const LevelDesc &L = plan->lv[l];
    int iy = row - L.row0;
    const uint32_t *bf = bits_fail + L.bofs + iy * L.nxw, *bo = bits_ok + L.bofs + iy * L.nxw;
    bool e = true;                                  // is the next window visited?  (x = 0 always is)
    for (int cx = 0; cx < L.nxw; cx++) {
        uint32_t f = bf[cx], o = bo[cx];
        // e[i+1] = !(e[i] && stage0_failed[i]); every lane runs the same 32-step automaton
        uint32_t em = 0;
#pragma unroll
        for (int i = 0; i < 32; i++) {
            em |= (uint32_t)e << i;
            e = !(e && ((f >> i) & 1u));
        }
        int ix = cx * 32 + lane;
        bool valid = ix < L.nx;
        bool visited = (em >> lane) & 1u, okl = (o >> lane) & 1u, fl = (f >> lane) & 1u;
        bool alive = valid && visited && okl && !fl;
        uint32_t am = __ballot_sync(0xffffffffu, alive);
        int base = 0;
        if (lane == 0 && am) base = atomicAdd(&counters[0], __popc(am));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (alive) {
            int pos = base + __popc(am & ((1u << lane) - 1u));
            if (pos < queue_cap)
                queue[pos] = make_uint2(((uint32_t)l << 26) | ((uint32_t)iy << 13) | (uint32_t)ix,
                                        __float_as_uint(vnf[L.wofs + iy * L.nx + ix]));
            else
                counters[2] = 1;
        }
        if (depth && valid && !alive)
            depth[L.wofs + iy * L.nx + ix] =
                (int16_t)(!visited ? NV_DEPTH_SKIPPED : (!okl ? NV_DEPTH_VARREJ : 0));
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_queue_stages(const PlanDev *__restrict__ plan, const DevCascade *__restrict__ meta, const DevStump *__restrict__ stumps,
               const uint32_t *__restrict__ sum, const uint2 *__restrict__ queue, int *__restrict__ counters,
               uint32_t *__restrict__ cand, int cand_cap, int16_t *__restrict__ depth)
{
    int n = counters[0];
    int nstages = meta->nstages;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint2 q = queue[i];
        int l = q.x >> 26, iy = (q.x >> 13) & 8191, ix = q.x & 8191;
        float vnf = __uint_as_float(q.y);
        const LevelDesc &L = plan->lv[l];
        LevelView v{sum + L.iofs, L.ipitch, L.iplane, L.ystep};
        const uint32_t *wb = v.sum + (size_t)iy * L.ystep * L.ipitch + ix;
        int code = NV_DEPTH_PASS;
        for (int st = 1; st < nstages; st++) {
            double tmp = 0.;
            int s0 = meta->stage_first[st], s1 = meta->stage_first[st + 1];
            for (int k = s0; k < s1; k++) {
                StumpRegs s = load_stump(stumps + k);
                tmp = __dadd_rn(tmp, (double)stump_leaf(wb, v, s, vnf));
            }
            if (tmp < (double)meta->stage_thr[st]) { code = -st; break; }
        }
        if (depth) depth[L.wofs + iy * L.nx + ix] = (int16_t)code;
        if (code == NV_DEPTH_PASS) {
            int pos = atomicAdd(&counters[1], 1);
            if (pos < cand_cap) cand[pos] = q.x;
            else counters[2] = 1;
        }
    }
}

// ------------------------------------------------------------------------------------------------
cudaError_t launch_stage0(const PlanDev *plan, int total_chunks, const DevCascade *meta, const DevStump *stumps,
                          const uint32_t *sum, const uint32_t *sq, float *vnf, uint32_t *bits_fail, uint32_t *bits_ok,
                          cudaStream_t st)
{
    k_stage0<<<(total_chunks + 7) / 8, 256, 0, st>>>(plan, total_chunks, meta, stumps, sum, sq, vnf, bits_fail, bits_ok);
    return cudaGetLastError();
}

cudaError_t launch_skip_compact(const PlanDev *plan, int total_rows, const float *vnf, const uint32_t *bits_fail,
                                const uint32_t *bits_ok, uint2 *queue, int *counters, int queue_cap, int16_t *depth,
                                cudaStream_t st)
{
    k_skip_compact<<<(total_rows + 7) / 8, 256, 0, st>>>(plan, total_rows, vnf, bits_fail, bits_ok, queue, counters,
                                                        queue_cap, depth);
    return cudaGetLastError();
}

cudaError_t launch_queue_stages(const PlanDev *plan, const DevCascade *meta, const DevStump *stumps, const uint32_t *sum,
                                const uint2 *queue, int *counters, uint32_t *cand, int cand_cap, int16_t *depth,
                                int nblocks, cudaStream_t st)
{
    k_queue_stages<<<nblocks, 256, 0, st>>>(plan, meta, stumps, sum, queue, counters, cand, cand_cap, depth);
    return cudaGetLastError();
}
