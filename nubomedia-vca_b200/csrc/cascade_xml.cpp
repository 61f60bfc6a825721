// cascade_xml.cpp — loads OpenCV "opencv-cascade-classifier" XML (new format), the model files the
// reference elements pass to cv::CascadeClassifier::load (kmsfacedetect.cpp:40,163-177;
// kmseyedetect.cpp:27-29,171-183; kmsmouthdetect.cpp:37-38; kmsnosedetect.cpp:31-32;
// kmseardetect.cpp:29-31).  Host-only; no OpenCV, no libxml: the grammar is small enough for a
// purpose-built tokenizer.  Supports HAAR features, upright rectangles, depth-1 trees (stumps).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <memory>

#include "internal.h"

static thread_local char g_err[512] = "";

void nv_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" const char *nv_last_error(void) { return g_err; }

namespace {

struct Node {
    std::string name, text;
    std::vector<std::unique_ptr<Node>> kids;
    const Node *child(const char *n) const
    {
        for (auto &k : kids)
            if (k->name == n) return k.get();
        return nullptr;
    }
    std::string child_text(const char *n) const
    {
        const Node *c = child(n);
        return c ? c->text : std::string();
    }
};

// Recursive-descent over <tag attr..>text<child/>..</tag>; comments, <?..?> and attributes skipped.
struct Parser {
    const char *p, *end;
    bool fail = false;

    void skip_misc()
    {
        for (;;) {
            while (p < end && isspace((unsigned char)*p)) p++;
            if (end - p >= 4 && !memcmp(p, "<!--", 4)) {
                const char *q = (const char *)memmem(p + 4, end - p - 4, "-->", 3);
                if (!q) { fail = true; p = end; return; }
                p = q + 3;
            } else if (end - p >= 2 && p[0] == '<' && p[1] == '?') {
                const char *q = (const char *)memmem(p, end - p, "?>", 2);
                if (!q) { fail = true; p = end; return; }
                p = q + 2;
            } else
                return;
        }
    }

    std::unique_ptr<Node> element()
    {
        skip_misc();
        if (p >= end || *p != '<') { fail = true; return nullptr; }
        p++;
        const char *s = p;
        while (p < end && !isspace((unsigned char)*p) && *p != '>' && *p != '/') p++;
        auto n = std::make_unique<Node>();
        n->name.assign(s, p);
        while (p < end && *p != '>') p++;          // attributes are not needed
        if (p >= end) { fail = true; return nullptr; }
        bool selfclose = p[-1] == '/';
        p++;
        if (selfclose) return n;
        for (;;) {
            const char *t = p;
            while (p < end && *p != '<') p++;
            n->text.append(t, p);
            if (p >= end) { fail = true; return nullptr; }
            if (end - p >= 2 && p[1] == '/') {      // closing tag
                while (p < end && *p != '>') p++;
                if (p < end) p++;
                return n;
            }
            if (end - p >= 4 && !memcmp(p, "<!--", 4)) { skip_misc(); continue; }
            auto k = element();
            if (fail || !k) return nullptr;
            n->kids.push_back(std::move(k));
        }
    }
};

bool parse_floats(const std::string &s, std::vector<double> &out)
{
    out.clear();
    const char *c = s.c_str();
    char *e;
    for (;;) {
        while (*c && isspace((unsigned char)*c)) c++;
        if (!*c) return true;
        double v = strtod(c, &e);
        if (e == c) return false;
        out.push_back(v);
        c = e;
    }
}

static int finish_cascade(HostCascade *hc)
{
    // Exactness certificate for parallel stage sums: OpenCV adds the float leaves one by one into a
    // double.  If, for every stage, (sum of |leaf|) / (smallest unit-in-last-place of any leaf) fits in
    // 2^52, no addition can round, so any summation order gives the same double.
    hc->order_free = 1;
    size_t si = 0;
    for (int nt : hc->stage_ntrees) {
        double mag = 0, min_ulp = INFINITY;
        for (int i = 0; i < nt; i++, si++)
            for (float leaf : {hc->stump_left[si], hc->stump_right[si]}) {
                if (leaf == 0.f) continue;
                if (!isfinite(leaf)) { hc->order_free = 0; continue; }
                int e;
                frexp((double)leaf, &e);                 // |leaf| in [2^(e-1), 2^e)
                mag += fabs((double)leaf);
                min_ulp = fmin(min_ulp, ldexp(1.0, e - 24));
            }
        if (mag > 0 && mag / min_ulp >= 4503599627370496.0) hc->order_free = 0;
    }
    return NV_OK;
}

}  // namespace

int nv_parse_cascade_xml(const char *path, HostCascade *hc)
{
    FILE *f = fopen(path, "rb");
    if (!f) { nv_set_error("cannot open cascade file %s", path); return NV_ERR_IO; }
    std::string buf;
    char tmp[65536];
    size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.append(tmp, n);
    fclose(f);

    Parser ps{buf.data(), buf.data() + buf.size()};
    auto root = ps.element();
    if (ps.fail || !root || root->name != "opencv_storage") {
        nv_set_error("%s: not an opencv_storage XML document", path);
        return NV_ERR_FORMAT;
    }
    const Node *c = root->child("cascade");
    std::vector<double> v;
    if (!c) {
        // OpenCV 1.x/2.x "opencv-haar-classifier" layout — what /usr/share/opencv/haarcascades held on the OpenCV 2.4
        // systems the reference was deployed on.  OpenCV >= 3 converts it to the new layout on load and evaluates
        // it identically (checked against cv2 4.13 in tests/test_oracle_vs_cv2.py), so it maps onto the same model:
        // one feature per tree node, <left_val>/<right_val> leaves.
        const Node *o = nullptr;
        for (auto &k : root->kids)
            if (k->child("stages") && k->child("size")) { o = k.get(); break; }
        if (!o) { nv_set_error("%s: neither a new-format <cascade> nor an old-format haar classifier", path); return NV_ERR_FORMAT; }
        if (!parse_floats(o->child_text("size"), v) || v.size() != 2) { nv_set_error("%s: malformed <size>", path); return NV_ERR_FORMAT; }
        hc->win_w = (int)v[0]; hc->win_h = (int)v[1];
        if (hc->win_w < 3 || hc->win_h < 3 || hc->win_w > 255 || hc->win_h > 255) { nv_set_error("%s: window out of range", path); return NV_ERR_FORMAT; }
        for (auto &st : o->child("stages")->kids) {
            const Node *trees = st->child("trees");
            if (!trees || !parse_floats(st->child_text("stage_threshold"), v) || v.size() != 1) { nv_set_error("%s: malformed stage", path); return NV_ERR_FORMAT; }
            hc->stage_thr.push_back((float)v[0]);
            int nt = 0;
            for (auto &tree : trees->kids) {
                if (tree->kids.size() != 1) { nv_set_error("%s: tree weak classifiers (depth > 1) are not supported yet", path); return NV_ERR_UNSUPPORTED; }
                const Node *node = tree->kids[0].get();
                const Node *ft = node->child("feature");
                if (!ft || !ft->child("rects") || !node->child("left_val") || !node->child("right_val")) {
                    nv_set_error("%s: malformed or non-stump tree node", path);
                    return node->child("left_node") || node->child("right_node") ? NV_ERR_UNSUPPORTED : NV_ERR_FORMAT;
                }
                if (atoi(ft->child_text("tilted").c_str()) != 0) { nv_set_error("%s: tilted features are not supported yet", path); return NV_ERR_UNSUPPORTED; }
                int r[12] = {0};
                float w[3] = {0, 0, 0};
                int k = 0;
                for (auto &rc : ft->child("rects")->kids) {
                    if (k >= 3 || !parse_floats(rc->text, v) || v.size() != 5) { nv_set_error("%s: malformed feature rectangle", path); return NV_ERR_FORMAT; }
                    for (int i = 0; i < 4; i++) r[4 * k + i] = (int)v[i];
                    w[k] = (float)v[4];
                    if (r[4 * k] < 0 || r[4 * k + 1] < 0 || r[4 * k + 2] <= 0 || r[4 * k + 3] <= 0 ||
                        r[4 * k] + r[4 * k + 2] > hc->win_w || r[4 * k + 1] + r[4 * k + 3] > hc->win_h) {
                        nv_set_error("%s: feature rectangle outside the window", path);
                        return NV_ERR_FORMAT;
                    }
                    k++;
                }
                if (k < 2) { nv_set_error("%s: feature with fewer than two rects", path); return NV_ERR_FORMAT; }
                std::vector<double> t, l, rr;
                if (!parse_floats(node->child_text("threshold"), t) || t.size() != 1 || !parse_floats(node->child_text("left_val"), l) ||
                    l.size() != 1 || !parse_floats(node->child_text("right_val"), rr) || rr.size() != 1) {
                    nv_set_error("%s: malformed tree node", path);
                    return NV_ERR_FORMAT;
                }
                if (w[2] != 0.f) hc->n3rect++;
                hc->stump_feat.push_back((int)hc->feat_weight.size() / 3);
                hc->feat_rect.insert(hc->feat_rect.end(), r, r + 12);
                hc->feat_weight.insert(hc->feat_weight.end(), w, w + 3);
                hc->stump_thr.push_back((float)t[0]); hc->stump_left.push_back((float)l[0]); hc->stump_right.push_back((float)rr[0]);
                nt++;
            }
            if (nt == 0) { nv_set_error("%s: empty stage", path); return NV_ERR_FORMAT; }
            hc->stage_ntrees.push_back(nt);
        }
        if (hc->stage_ntrees.empty() || hc->stage_ntrees.size() > NV_MAX_STAGES) {
            nv_set_error("%s: %zu stages (supported: 1..%d)", path, hc->stage_ntrees.size(), NV_MAX_STAGES);
            return hc->stage_ntrees.empty() ? NV_ERR_FORMAT : NV_ERR_UNSUPPORTED;
        }
        return finish_cascade(hc);
    }
    if (!c->child("stages") || !c->child("features")) {
        nv_set_error("%s: incomplete cascade (no <stages>/<features>)", path);
        return NV_ERR_FORMAT;
    }
    auto trimmed = [](std::string s) {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    };
    if (trimmed(c->child_text("featureType")) != "HAAR" || trimmed(c->child_text("stageType")) != "BOOST") {
        nv_set_error("%s: only BOOST/HAAR cascades are supported (LBP/HOG are not)", path);
        return NV_ERR_UNSUPPORTED;
    }
    hc->win_w = atoi(c->child_text("width").c_str());
    hc->win_h = atoi(c->child_text("height").c_str());
    if (hc->win_w < 3 || hc->win_h < 3 || hc->win_w > 255 || hc->win_h > 255) {
        nv_set_error("%s: window %dx%d out of range", path, hc->win_w, hc->win_h);
        return NV_ERR_FORMAT;
    }
    for (auto &st : c->child("stages")->kids) {
        const Node *weak = st->child("weakClassifiers");
        if (!weak || !parse_floats(st->child_text("stageThreshold"), v) || v.size() != 1) {
            nv_set_error("%s: malformed stage", path);
            return NV_ERR_FORMAT;
        }
        hc->stage_thr.push_back((float)v[0]);
        int nt = 0;
        for (auto &wc : weak->kids) {
            std::vector<double> nodes, leaves;
            if (!parse_floats(wc->child_text("internalNodes"), nodes) || !parse_floats(wc->child_text("leafValues"), leaves)) {
                nv_set_error("%s: malformed weak classifier", path);
                return NV_ERR_FORMAT;
            }
            if (nodes.size() != 4 || leaves.size() != 2 || nodes[0] != 0 || nodes[1] != -1) {
                nv_set_error("%s: tree weak classifiers (depth > 1) are not supported yet", path);
                return NV_ERR_UNSUPPORTED;
            }
            hc->stump_feat.push_back((int)nodes[2]);
            hc->stump_thr.push_back((float)nodes[3]);
            hc->stump_left.push_back((float)leaves[0]);
            hc->stump_right.push_back((float)leaves[1]);
            nt++;
        }
        if (nt == 0) { nv_set_error("%s: empty stage", path); return NV_ERR_FORMAT; }
        hc->stage_ntrees.push_back(nt);
    }
    if (hc->stage_ntrees.empty() || hc->stage_ntrees.size() > NV_MAX_STAGES) {
        nv_set_error("%s: %zu stages (supported: 1..%d)", path, hc->stage_ntrees.size(), NV_MAX_STAGES);
        return hc->stage_ntrees.empty() ? NV_ERR_FORMAT : NV_ERR_UNSUPPORTED;
    }
    for (auto &ft : c->child("features")->kids) {
        const Node *rects = ft->child("rects");
        if (!rects) { nv_set_error("%s: feature without rects", path); return NV_ERR_FORMAT; }
        if (atoi(ft->child_text("tilted").c_str()) != 0) {
            nv_set_error("%s: tilted features are not supported yet", path);
            return NV_ERR_UNSUPPORTED;
        }
        int r[12] = {0};
        float w[3] = {0, 0, 0};
        int k = 0;
        for (auto &rc : rects->kids) {
            if (k >= 3 || !parse_floats(rc->text, v) || v.size() != 5) {
                nv_set_error("%s: malformed feature rectangle", path);
                return NV_ERR_FORMAT;
            }
            for (int i = 0; i < 4; i++) r[4 * k + i] = (int)v[i];
            w[k] = (float)v[4];
            if (r[4 * k] < 0 || r[4 * k + 1] < 0 || r[4 * k + 2] <= 0 || r[4 * k + 3] <= 0 ||
                r[4 * k] + r[4 * k + 2] > hc->win_w || r[4 * k + 1] + r[4 * k + 3] > hc->win_h) {
                nv_set_error("%s: feature rectangle outside the window", path);
                return NV_ERR_FORMAT;
            }
            k++;
        }
        if (k < 2) { nv_set_error("%s: feature with fewer than two rects", path); return NV_ERR_FORMAT; }
        if (w[2] != 0.f) hc->n3rect++;
        hc->feat_rect.insert(hc->feat_rect.end(), r, r + 12);
        hc->feat_weight.insert(hc->feat_weight.end(), w, w + 3);
    }
    int nfeat = (int)hc->feat_weight.size() / 3;
    for (int fi : hc->stump_feat)
        if (fi < 0 || fi >= nfeat) { nv_set_error("%s: feature index out of range", path); return NV_ERR_FORMAT; }

    return finish_cascade(hc);
}
