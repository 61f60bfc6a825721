// cascade_xml.cpp — loads OpenCV "opencv-cascade-classifier" XML (new format), the model files the
// reference elements pass to cv::CascadeClassifier::load (kmsfacedetect.cpp:40,163-177;
// kmseyedetect.cpp:27-29,171-183; kmsmouthdetect.cpp:37-38; kmsnosedetect.cpp:31-32;
// kmseardetect.cpp:29-31).  Host-only; no OpenCV, no libxml: the grammar is small enough for a
// purpose-built tokenizer.  Supports BOOST/HAAR cascades (stumps or trees, upright or tilted rectangles) and BOOST/LBP
// cascades (categorical stumps or trees over 256 LBP codes).
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

// One weak classifier as parsed: internal nodes in file order and nnodes + 1 leaves.
struct ParsedTree {
    struct N { int feat; float thr; int left, right; int subset[8]; };
    std::vector<N> nodes;
    std::vector<float> leaves;
};

static int parse_rects(const Node *rects, bool tilted, int win_w, int win_h, const char *path, int r[12], float w[3])
{
    std::vector<double> v;
    int k = 0;
    for (auto &rc : rects->kids) {
        if (k >= 3 || !parse_floats(rc->text, v) || v.size() != 5) { nv_set_error("%s: malformed feature rectangle", path); return NV_ERR_FORMAT; }
        for (int i = 0; i < 4; i++) r[4 * k + i] = (int)v[i];
        w[k] = (float)v[4];
        int x = r[4 * k], y = r[4 * k + 1], ww = r[4 * k + 2], hh = r[4 * k + 3];
        // upright: the rect lies in the window.  tilted (x, y, w, h): the corners (x - h, y + h), (x + w, y + w) and
        // (x + w - h, y + w + h) of the rotated rect must be elements of the window's tilted integral.
        bool ok = x >= 0 && y >= 0 && ww > 0 && hh > 0 &&
                  (tilted ? (x - hh >= 0 && x + ww <= win_w && y + ww + hh <= win_h) : (x + ww <= win_w && y + hh <= win_h));
        if (!ok) { nv_set_error("%s: feature rectangle outside the window", path); return NV_ERR_FORMAT; }
        k++;
    }
    if (k < 2) { nv_set_error("%s: feature with fewer than two rects", path); return NV_ERR_FORMAT; }
    return NV_OK;
}

static int finish_cascade(HostCascade *hc, const std::vector<ParsedTree> &trees, const char *path)
{
    int nfeat = (int)hc->feat_weight.size() / 3;
    hc->general = hc->lbp ? 1 : 0;
    for (uint8_t t : hc->feat_tilted) if (t) { hc->general = 1; hc->has_tilted = 1; }
    for (const ParsedTree &t : trees) {
        int nn = (int)t.nodes.size();
        if (nn < 1 || (int)t.leaves.size() != nn + 1) { nv_set_error("%s: malformed weak classifier", path); return NV_ERR_FORMAT; }
        if (nn != 1) hc->general = 1;
        for (const auto &n : t.nodes) {
            if (n.feat < 0 || n.feat >= nfeat) { nv_set_error("%s: feature index out of range", path); return NV_ERR_FORMAT; }
            for (int c : {n.left, n.right})
                if (c >= nn || -c > nn) { nv_set_error("%s: tree child index out of range", path); return NV_ERR_FORMAT; }
        }
        // a child index must point forward, or the walk could loop for ever
        for (int i = 0; i < nn; i++)
            for (int c : {t.nodes[i].left, t.nodes[i].right})
                if (c > 0 && c <= i) { nv_set_error("%s: tree child index points backwards", path); return NV_ERR_FORMAT; }
        if (nn == 1 && (t.nodes[0].left != 0 || t.nodes[0].right != -1)) hc->general = 1;
    }
    for (const ParsedTree &t : trees) {
        hc->tree_nnodes.push_back((int)t.nodes.size());
        for (const auto &n : t.nodes) {
            hc->node_feat.push_back(n.feat); hc->node_thr.push_back(n.thr);
            hc->node_left.push_back(n.left); hc->node_right.push_back(n.right);
            if (hc->lbp) hc->node_subset.insert(hc->node_subset.end(), n.subset, n.subset + 8);
        }
        hc->leaves.insert(hc->leaves.end(), t.leaves.begin(), t.leaves.end());
        // the stump arrays stay index-compatible with the weak classifiers; they are only meaningful when !general
        hc->stump_feat.push_back(t.nodes[0].feat);
        hc->stump_thr.push_back(t.nodes[0].thr);
        hc->stump_left.push_back(t.leaves[0]);
        hc->stump_right.push_back(t.leaves[1]);
    }
    for (size_t f = 0; f < hc->feat_weight.size() / 3; f++)
        if (hc->feat_weight[3 * f + 2] != 0.f) hc->n3rect++;

    // Exactness certificate for parallel stage sums: OpenCV adds one float leaf per weak classifier, one by one, into
    // a double.  Every partial sum in ANY order is a sum of one leaf from each of some trees: a multiple of the smallest
    // unit-in-last-place of any leaf of the stage, bounded by the sum over trees of the largest |leaf|.  If that bound
    // stays below 2^52 such units, no addition can round and every order gives the same double.
    hc->order_free = 1;
    size_t ti = 0, li = 0;
    for (int nt : hc->stage_ntrees) {
        double mag = 0, min_ulp = INFINITY;
        for (int i = 0; i < nt; i++, ti++) {
            double big = 0;
            for (int k = 0; k <= hc->tree_nnodes[ti]; k++, li++) {
                float leaf = hc->leaves[li];
                if (leaf == 0.f) continue;
                if (!isfinite(leaf)) { hc->order_free = 0; continue; }
                int e;
                frexp((double)leaf, &e);                 // |leaf| in [2^(e-1), 2^e)
                big = fmax(big, fabs((double)leaf));
                min_ulp = fmin(min_ulp, ldexp(1.0, e - 24));
            }
            mag += big;
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
    std::vector<ParsedTree> trees;
    if (!c) {
        // OpenCV 1.x/2.x "opencv-haar-classifier" layout — what /usr/share/opencv/haarcascades held on the OpenCV 2.4
        // systems the reference was deployed on.  OpenCV >= 3 converts it to the new layout on load and evaluates
        // it identically (checked against cv2 4.13 in tests/test_oracle_vs_cv2.py), so it maps onto the same model:
        // one feature per tree node in file order; a <left_val>/<right_val> becomes the next leaf of its tree, a
        // <left_node>/<right_node> the index of the child node.
        const Node *o = nullptr;
        for (auto &k : root->kids)
            if (k->child("stages") && k->child("size")) { o = k.get(); break; }
        if (!o) { nv_set_error("%s: neither a new-format <cascade> nor an old-format haar classifier", path); return NV_ERR_FORMAT; }
        if (!parse_floats(o->child_text("size"), v) || v.size() != 2) { nv_set_error("%s: malformed <size>", path); return NV_ERR_FORMAT; }
        hc->win_w = (int)v[0]; hc->win_h = (int)v[1];
        if (hc->win_w < 3 || hc->win_h < 3 || hc->win_w > 255 || hc->win_h > 255) { nv_set_error("%s: window out of range", path); return NV_ERR_FORMAT; }
        for (auto &st : o->child("stages")->kids) {
            const Node *tr = st->child("trees");
            if (!tr || !parse_floats(st->child_text("stage_threshold"), v) || v.size() != 1) { nv_set_error("%s: malformed stage", path); return NV_ERR_FORMAT; }
            hc->stage_thr.push_back((float)v[0]);
            int nt = 0;
            for (auto &tree : tr->kids) {
                ParsedTree pt;
                for (auto &nodep : tree->kids) {
                    const Node *node = nodep.get();
                    const Node *ft = node->child("feature");
                    if (!ft || !ft->child("rects")) { nv_set_error("%s: malformed tree node", path); return NV_ERR_FORMAT; }
                    bool tilted = atoi(ft->child_text("tilted").c_str()) != 0;
                    int r[12] = {0};
                    float w[3] = {0, 0, 0};
                    int rc = parse_rects(ft->child("rects"), tilted, hc->win_w, hc->win_h, path, r, w);
                    if (rc != NV_OK) return rc;
                    std::vector<double> t;
                    if (!parse_floats(node->child_text("threshold"), t) || t.size() != 1) { nv_set_error("%s: malformed tree node", path); return NV_ERR_FORMAT; }
                    int child[2];
                    const char *val[2] = {"left_val", "right_val"}, *nod[2] = {"left_node", "right_node"};
                    for (int sd = 0; sd < 2; sd++) {
                        std::vector<double> cv;
                        if (node->child(val[sd])) {
                            if (!parse_floats(node->child_text(val[sd]), cv) || cv.size() != 1) { nv_set_error("%s: malformed leaf", path); return NV_ERR_FORMAT; }
                            child[sd] = -(int)pt.leaves.size();
                            pt.leaves.push_back((float)cv[0]);
                        } else if (node->child(nod[sd])) {
                            if (!parse_floats(node->child_text(nod[sd]), cv) || cv.size() != 1) { nv_set_error("%s: malformed child index", path); return NV_ERR_FORMAT; }
                            child[sd] = (int)cv[0];
                        } else { nv_set_error("%s: tree node without children", path); return NV_ERR_FORMAT; }
                    }
                    pt.nodes.push_back({(int)hc->feat_weight.size() / 3, (float)t[0], child[0], child[1], {0}});
                    hc->feat_rect.insert(hc->feat_rect.end(), r, r + 12);
                    hc->feat_weight.insert(hc->feat_weight.end(), w, w + 3);
                    hc->feat_tilted.push_back(tilted ? 1 : 0);
                }
                trees.push_back(std::move(pt));
                nt++;
            }
            if (nt == 0) { nv_set_error("%s: empty stage", path); return NV_ERR_FORMAT; }
            hc->stage_ntrees.push_back(nt);
        }
        if (hc->stage_ntrees.empty() || hc->stage_ntrees.size() > NV_MAX_STAGES) {
            nv_set_error("%s: %zu stages (supported: 1..%d)", path, hc->stage_ntrees.size(), NV_MAX_STAGES);
            return hc->stage_ntrees.empty() ? NV_ERR_FORMAT : NV_ERR_UNSUPPORTED;
        }
        return finish_cascade(hc, trees, path);
    }
    if (!c->child("stages") || !c->child("features")) {
        nv_set_error("%s: incomplete cascade (no <stages>/<features>)", path);
        return NV_ERR_FORMAT;
    }
    auto trimmed = [](std::string s) {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    };
    const std::string ftype = trimmed(c->child_text("featureType"));
    if ((ftype != "HAAR" && ftype != "LBP") || trimmed(c->child_text("stageType")) != "BOOST") {
        nv_set_error("%s: only BOOST/HAAR and BOOST/LBP cascades are supported (HOG is not)", path);
        return NV_ERR_UNSUPPORTED;
    }
    hc->lbp = ftype == "LBP";
    if (hc->lbp) {
        // OpenCV's subset size is (maxCatCount + 31) / 32 words per node; LBP codes are 8 bits, the trainer writes 256
        const Node *fp = c->child("featureParams");
        if (!fp || atoi(fp->child_text("maxCatCount").c_str()) != 256) {
            nv_set_error("%s: LBP cascade with maxCatCount != 256", path);
            return NV_ERR_UNSUPPORTED;
        }
    }
    const size_t per_node = hc->lbp ? 11 : 4;                    // left right featureIdx, then the threshold or 8 subset words
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
            if (!parse_floats(wc->child_text("internalNodes"), nodes) || !parse_floats(wc->child_text("leafValues"), leaves) ||
                nodes.empty() || nodes.size() % per_node != 0 || leaves.size() != nodes.size() / per_node + 1) {
                nv_set_error("%s: malformed weak classifier", path);
                return NV_ERR_FORMAT;
            }
            ParsedTree pt;
            for (size_t i = 0; i < nodes.size(); i += per_node) {
                ParsedTree::N n = {(int)nodes[i + 2], hc->lbp ? 0.f : (float)nodes[i + 3], (int)nodes[i], (int)nodes[i + 1], {0}};
                if (hc->lbp)
                    for (int k = 0; k < 8; k++) n.subset[k] = (int)(long long)nodes[i + 3 + k];   // written as signed 32-bit ints
                pt.nodes.push_back(n);
            }
            for (double l : leaves) pt.leaves.push_back((float)l);
            trees.push_back(std::move(pt));
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
        if (hc->lbp) {                                            // <rect>x y w h</rect>: one cell of the 3 x 3 grid
            if (!parse_floats(ft->child_text("rect"), v) || v.size() != 4) { nv_set_error("%s: malformed LBP feature", path); return NV_ERR_FORMAT; }
            int r[12] = {(int)v[0], (int)v[1], (int)v[2], (int)v[3], 0, 0, 0, 0, 0, 0, 0, 0};
            if (r[0] < 0 || r[1] < 0 || r[2] <= 0 || r[3] <= 0 || r[0] + 3 * r[2] > hc->win_w || r[1] + 3 * r[3] > hc->win_h) {
                nv_set_error("%s: LBP feature outside the window", path);
                return NV_ERR_FORMAT;
            }
            const float w[3] = {0, 0, 0};
            hc->feat_rect.insert(hc->feat_rect.end(), r, r + 12);
            hc->feat_weight.insert(hc->feat_weight.end(), w, w + 3);
            hc->feat_tilted.push_back(0);
            continue;
        }
        const Node *rects = ft->child("rects");
        if (!rects) { nv_set_error("%s: feature without rects", path); return NV_ERR_FORMAT; }
        bool tilted = atoi(ft->child_text("tilted").c_str()) != 0;
        int r[12] = {0};
        float w[3] = {0, 0, 0};
        int rc = parse_rects(rects, tilted, hc->win_w, hc->win_h, path, r, w);
        if (rc != NV_OK) return rc;
        hc->feat_rect.insert(hc->feat_rect.end(), r, r + 12);
        hc->feat_weight.insert(hc->feat_weight.end(), w, w + 3);
        hc->feat_tilted.push_back(tilted ? 1 : 0);
    }
    return finish_cascade(hc, trees, path);
}
