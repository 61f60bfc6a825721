// streams_bench — BASELINE config 5 driven the way a media server drives it: many per-stream element contexts on one
// GPU, fed from several native host threads through the C ABI (include/nubovca.h) with page-locked host frames.
// Every stream keeps one frame in flight: a thread walks its streams, collects a stream's previous frame and
// submits its next one (the steady state of a GStreamer streaming thread per element, kmsfacedetect.cpp:857-898,
// folded onto fewer threads).  Prints one JSON line: frames/s over all streams, streams@30fps, and the number of
// rectangles found (bench.py checks it against the Python path on the same frames).
//
//   streams_bench --frames-file F --nframes K --fmt bgr|nv12|i420 --width 1280 --height 720 --width-to-process 640
//                 --xml cascade.xml [--gpu 0] [--streams 32] [--threads 4] [--iters 200] [--warmup 20]
//                 [--scale-factor 1.25] [--min-size -1] [--memory pinned|pageable|registered] [--sync 0|1]
//
// F holds K frames back to back (bgr: 3wh bytes each; nv12 / i420: 3wh/2 bytes each).
// --memory pageable: the frames live in malloc() memory, as the buffers GStreamer hands an element do (the library
//   stages them through its pinned buffer); registered: malloc() memory page-locked in place with cudaHostRegister through
//   nv_host_register (what a shell can do once per upstream buffer-pool block); pinned: nv_host_alloc.
// --sync 1: the element's own call shape — one synchronous nv_face_detect per buffer (kmsfacedetect.cpp:857-898), every
//   host thread walking its streams with nothing in flight between calls.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "nubovca.h"

struct Args {
    std::string file, fmt = "bgr", xml;
    int nframes = 1, w = 1280, h = 720, w2p = 640, gpu = 0, streams = 32, threads = 4, iters = 200, warmup = 20;
    double sf = 1.25;
    int min_size = -1, sync = 0;
    std::string memory = "pinned";
};

static void die(const char *what, int rc) { fprintf(stderr, "streams_bench: %s failed (%d): %s\n", what, rc, nv_last_error()); exit(2); }

int main(int argc, char **argv)
{
    Args a;
    for (int i = 1; i + 1 < argc; i += 2) {
        std::string k = argv[i], v = argv[i + 1];
        if (k == "--frames-file") a.file = v; else if (k == "--fmt") a.fmt = v; else if (k == "--xml") a.xml = v;
        else if (k == "--nframes") a.nframes = atoi(v.c_str()); else if (k == "--width") a.w = atoi(v.c_str());
        else if (k == "--height") a.h = atoi(v.c_str()); else if (k == "--width-to-process") a.w2p = atoi(v.c_str());
        else if (k == "--gpu") a.gpu = atoi(v.c_str()); else if (k == "--streams") a.streams = atoi(v.c_str());
        else if (k == "--threads") a.threads = atoi(v.c_str()); else if (k == "--iters") a.iters = atoi(v.c_str());
        else if (k == "--warmup") a.warmup = atoi(v.c_str()); else if (k == "--scale-factor") a.sf = atof(v.c_str());
        else if (k == "--min-size") a.min_size = atoi(v.c_str()); else if (k == "--sync") a.sync = atoi(v.c_str());
        else if (k == "--memory") a.memory = v;
        else { fprintf(stderr, "unknown option %s\n", k.c_str()); return 2; }
    }
    const bool yuv = a.fmt != "bgr";
    const int fmt = a.fmt == "nv12" ? NV_FMT_NV12 : a.fmt == "i420" ? NV_FMT_I420 : NV_FMT_BGR;
    const size_t fbytes = yuv ? (size_t)a.w * a.h * 3 / 2 : (size_t)a.w * a.h * 3;
    if (a.file.empty() || a.xml.empty() || a.nframes < 1 || a.threads < 1 || a.streams < a.threads) { fprintf(stderr, "bad arguments\n"); return 2; }

    int rc;
    nv_cascade *casc = nullptr;
    if ((rc = nv_cascade_load(a.xml.c_str(), &casc)) != NV_OK) die("nv_cascade_load", rc);
    uint8_t *frames = nullptr;
    if (a.memory == "pinned") {
        if ((rc = nv_host_alloc(fbytes * a.nframes, (void **)&frames)) != NV_OK) die("nv_host_alloc", rc);
    } else {
        if (posix_memalign((void **)&frames, 4096, fbytes * a.nframes) != 0) { fprintf(stderr, "out of memory\n"); return 2; }
        if (a.memory == "registered" && (rc = nv_host_register(frames, fbytes * a.nframes)) != NV_OK) die("nv_host_register", rc);
    }
    FILE *f = fopen(a.file.c_str(), "rb");
    if (!f || fread(frames, fbytes, a.nframes, f) != (size_t)a.nframes) { fprintf(stderr, "cannot read %d frames from %s\n", a.nframes, a.file.c_str()); return 2; }
    fclose(f);

    std::vector<nv_ctx *> ctx(a.streams);
    for (auto &c : ctx) if ((rc = nv_ctx_create(a.gpu, a.w, a.h, &c)) != NV_OK) die("nv_ctx_create", rc);
    nv_face_params fp = {a.w2p, a.sf, 3, a.min_size, a.min_size};

    auto submit = [&](int s, long j) {
        const uint8_t *p = frames + fbytes * ((s + j) % a.nframes);
        if (!yuv) return nv_face_submit(ctx[s], casc, p, a.w, a.h, 3 * a.w, &fp);
        nv_yuv_frame y = {};
        y.format = fmt; y.width = a.w; y.height = a.h;
        y.plane[0] = p; y.stride[0] = a.w;
        y.plane[1] = p + (size_t)a.w * a.h;
        if (fmt == NV_FMT_I420) { y.stride[1] = y.stride[2] = a.w / 2; y.plane[2] = y.plane[1] + (size_t)a.w * a.h / 4; }
        else y.stride[1] = a.w;
        return nv_face_submit_yuv(ctx[s], casc, &y, &fp);
    };

    std::vector<long long> nrects(a.threads, 0);
    auto detect_sync = [&](int s, long j, nv_rect *out, int cap, int *n) {
        const uint8_t *p = frames + fbytes * ((s + j) % a.nframes);
        if (!yuv) return nv_face_detect(ctx[s], casc, p, a.w, a.h, 3 * a.w, &fp, out, cap, n);
        nv_yuv_frame y = {};
        y.format = fmt; y.width = a.w; y.height = a.h;
        y.plane[0] = p; y.stride[0] = a.w;
        y.plane[1] = p + (size_t)a.w * a.h;
        if (fmt == NV_FMT_I420) { y.stride[1] = y.stride[2] = a.w / 2; y.plane[2] = y.plane[1] + (size_t)a.w * a.h / 4; }
        else y.stride[1] = a.w;
        return nv_face_detect_yuv(ctx[s], casc, &y, &fp, out, cap, n);
    };
    auto run = [&](int t, int iters, bool count) {
        if (a.sync) {                                   // one blocking call per buffer, as transform_frame_ip makes it
            nv_rect out[256];
            int n;
            for (long j = 0; j < iters; j++)
                for (int s = t; s < a.streams; s += a.threads) {
                    int r = detect_sync(s, j, out, 256, &n);
                    if (r != NV_OK) die("nv_face_detect", r);
                    if (count) nrects[t] += n;
                }
            return;
        }
        // streams t, t + threads, ...: collect the previous frame of a stream right before submitting its next one
        nv_rect out[256];
        int n;
        for (long j = 0; j <= iters; j++)
            for (int s = t; s < a.streams; s += a.threads) {
                if (j > 0) {
                    int r = nv_face_collect(ctx[s], out, 256, &n);
                    if (r != NV_OK) die("nv_face_collect", r);
                    if (count) nrects[t] += n;
                }
                if (j < iters) { int r = submit(s, j); if (r != NV_OK) die("nv_face_submit", r); }
            }
    };
    auto run_all = [&](int iters, bool count) {
        std::vector<std::thread> th;
        for (int t = 0; t < a.threads; t++) th.emplace_back(run, t, iters, count);
        for (auto &x : th) x.join();
    };
    run_all(a.warmup, false);
    auto t0 = std::chrono::steady_clock::now();
    run_all(a.iters, true);
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long long total = 0;
    for (auto v : nrects) total += v;
    double fps = (double)a.streams * a.iters / dt;
    printf("{\"frames_per_s\": %.1f, \"streams_at_30fps\": %.1f, \"streams\": %d, \"host_threads\": %d, \"frames\": %lld, "
           "\"rects\": %lld, \"fmt\": \"%s\", \"h2d_bytes_per_frame\": %zu, \"seconds\": %.3f, \"memory\": \"%s\", \"sync\": %d}\n",
           fps, fps / 30.0, a.streams, a.threads, (long long)a.streams * a.iters, total, a.fmt.c_str(), fbytes, dt, a.memory.c_str(), a.sync);
    for (auto c : ctx) nv_ctx_destroy(c);
    if (a.memory == "pinned") nv_host_free(frames);
    else { if (a.memory == "registered") nv_host_unregister(frames); free(frames); }
    nv_cascade_free(casc);
    return 0;
}
