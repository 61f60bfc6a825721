// opencv2/opencv.hpp — stand-in for the OpenCV 2.x headers, just wide enough for the reference's own element sources
// (/root/reference/modules/*/*/src/gst-plugins/*.cpp, Faces.cpp, BaseFace.cpp) to compile UNMODIFIED here.
//
// TEST INFRASTRUCTURE ONLY (part of oracle/): nothing under nubomedia-vca_b200/ may include or link this.
//
// Every pixel operation is forwarded to the CPU oracle (oracle/nubo_oracle.c, pinned to cv2 4.13 by
// tests/test_oracle_vs_cv2.py), so a reference element built against this header is "the reference's glue — gating,
// ROI arithmetic, temporal smoothing, events — on top of the oracle's arithmetic".  Drawing calls (cvRectangle,
// cv::circle) are not rasterised here: they are RECORDED (refcv::draw_log) and the tests replay them with the real
// cv2.rectangle / cv2.circle, so the pixel check does not rest on a second restatement by the same author.
// cv::Mat::operator()(Rect) throws cv::Exception for a rectangle outside the image, as OpenCV does.
#ifndef REFBUILD_OPENCV_HPP
#define REFBUILD_OPENCV_HPP

#include <emmintrin.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <exception>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

// ---- the oracle's C entry points (oracle/nubo_oracle.c) -------------------------------------------------------------
extern "C" {
struct ora_cascade;
void ora_bgr2gray(const uint8_t *src, int w, int h, int sstride, int cn, uint8_t *dst, int dstride);
void ora_resize_linear(const uint8_t *src, int sw, int sh, int sstride, int cn, uint8_t *dst, int dw, int dh, int dstride);
void ora_equalize_hist(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride);
int ora_detect_multiscale(const ora_cascade *c, const uint8_t *gray, int W, int H, int stride, double scale_factor,
                          int min_neighbors, int min_w, int min_h, int max_w, int max_h, int *out, int cap, int *weights,
                          long long *nwindows);
void ora_update_mhi(const uint8_t *silh, float *mhi, int n, double ts, double duration);
int ora_segment_motion(const float *mhi, int w, int h, double ts, double seg_thresh, int32_t *labels, int *rects, int cap);
}

typedef unsigned char uchar;

namespace cv {

using std::vector;      // OpenCV 2.x core.hpp does this; the elements write vector<Rect> / string under `using namespace cv`
using std::string;

class Exception : public std::exception {
public:
    std::string msg;
    Exception() {}
    explicit Exception(const std::string &m) : msg(m) {}
    virtual ~Exception() throw() {}
    virtual const char *what() const throw() { return msg.c_str(); }
};

template <typename T> struct Rect_;
template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
    bool inside(const Rect_<T> &r) const;
};
template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T a, T b, T w, T h) : x(a), y(b), width(w), height(h) {}
    Rect_(const Point_<T> &p1, const Point_<T> &p2)
    {
        x = std::min(p1.x, p2.x); y = std::min(p1.y, p2.y);
        width = std::max(p1.x, p2.x) - x; height = std::max(p1.y, p2.y) - y;
    }
    Point_<T> tl() const { return Point_<T>(x, y); }
    Point_<T> br() const { return Point_<T>(x + width, y + height); }
    Size_<T> size() const { return Size_<T>(width, height); }
    T area() const { return width * height; }
    bool contains(const Point_<T> &p) const { return x <= p.x && p.x < x + width && y <= p.y && p.y < y + height; }
};
template <typename T> inline bool Point_<T>::inside(const Rect_<T> &r) const { return r.contains(*this); }

typedef Point_<int> Point;
typedef Size_<int> Size;
typedef Rect_<int> Rect;

struct Scalar {
    double val[4];
    Scalar() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar(double v0) { val[0] = v0; val[1] = val[2] = val[3] = 0; }
    Scalar(double v0, double v1, double v2 = 0, double v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    double operator[](int i) const { return val[i]; }
};

}  // namespace cv

// ---- the C API pieces the elements use ------------------------------------------------------------------------------
typedef void CvArr;
typedef cv::Point CvPoint;
typedef cv::Scalar CvScalar;
typedef cv::Size CvSize;
typedef cv::Rect CvRect;
struct CvMemStorage { int dummy; };
struct CvSeq { int dummy; };

#define IPL_DEPTH_8U 8
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_8UC4 24
#define CV_32FC1 5
#define CV_BGR2GRAY 6
#define CV_INTER_LINEAR 1
#define CV_LOAD_IMAGE_UNCHANGED (-1)
#define CV_HAAR_DO_CANNY_PRUNING 1
#define CV_HAAR_SCALE_IMAGE 2
#define CV_HAAR_FIND_BIGGEST_OBJECT 4
#define CV_HAAR_DO_ROUGH_SEARCH 8
#define CV_RGB(r, g, b) cv::Scalar((b), (g), (r), 0)

struct IplImage {
    int nChannels, depth, width, height, widthStep;
    char *imageData;
    int owns_data;
};

inline int cvRound(double v) { return _mm_cvtsd_si32(_mm_set_sd(v)); }      // OpenCV's own SSE2 form: INT_MIN for inf / NaN
inline CvPoint cvPoint(int x, int y) { return CvPoint(x, y); }
inline CvSize cvSize(int w, int h) { return CvSize(w, h); }

inline IplImage *cvCreateImageHeader(CvSize size, int depth, int channels)
{
    IplImage *im = (IplImage *)calloc(1, sizeof(IplImage));
    im->nChannels = channels; im->depth = depth; im->width = size.width; im->height = size.height;
    im->widthStep = (size.width * channels * (depth / 8) + 3) & ~3;          // IplImage rows are 4-byte aligned
    return im;
}
inline IplImage *cvCreateImage(CvSize size, int depth, int channels)
{
    IplImage *im = cvCreateImageHeader(size, depth, channels);
    im->imageData = (char *)calloc((size_t)im->widthStep * (size_t)std::max(im->height, 1), 1);
    im->owns_data = 1;
    return im;
}
inline void cvReleaseImage(IplImage **im)
{
    if (!im || !*im) return;
    if ((*im)->owns_data) free((*im)->imageData);
    free(*im);
    *im = NULL;
}
inline IplImage *cvLoadImage(const char *, int) { return NULL; }                       // no overlay images here
inline void cvResize(const IplImage *src, IplImage *dst, int)
{
    ora_resize_linear((const uint8_t *)src->imageData, src->width, src->height, src->widthStep, src->nChannels,
                      (uint8_t *)dst->imageData, dst->width, dst->height, dst->widthStep);
}
inline CvMemStorage *cvCreateMemStorage(int) { return (CvMemStorage *)calloc(1, sizeof(CvMemStorage)); }
inline CvSeq *cvCreateSeq(int, size_t, size_t, CvMemStorage *) { return (CvSeq *)calloc(1, sizeof(CvSeq)); }
inline void cvClearMemStorage(CvMemStorage *) {}
inline void cvClearSeq(CvSeq *) {}
inline void cvReleaseMemStorage(CvMemStorage **s) { if (s && *s) { free(*s); *s = NULL; } }

// ---- recorded drawing and hooks (read by oracle/refbuild/ref_harness.cpp) -------------------------------------------
namespace refcv {
struct DrawCall {
    int kind;                      // 0: cvRectangle(p1, p2), 1: cv::circle(centre = p1, radius = p2.x)
    const void *target;            // imageData the call drew into
    int x0, y0, x1, y1;
    double color[4];
    int thickness, line_type, shift;
};
std::vector<DrawCall> &draw_log();
// CascadeClassifier::load(path): the harness maps the reference's hard-coded /usr/share/opencv/haarcascades/<file> to a
// model registered by the test (parsed by oracle/oracle.py); NULL = file missing (load fails, as on a bare machine)
const ora_cascade *find_cascade(const std::string &path);
}  // namespace refcv

inline void cvRectangle(CvArr *img, CvPoint p1, CvPoint p2, CvScalar color, int thickness = 1, int line_type = 8, int shift = 0)
{
    refcv::DrawCall d;
    d.kind = 0; d.target = ((IplImage *)img)->imageData;
    d.x0 = p1.x; d.y0 = p1.y; d.x1 = p2.x; d.y1 = p2.y;
    for (int i = 0; i < 4; i++) d.color[i] = color.val[i];
    d.thickness = thickness; d.line_type = line_type; d.shift = shift;
    refcv::draw_log().push_back(d);
}

namespace cv {

enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { COLOR_BGR2GRAY = 6 };
enum { CASCADE_DO_CANNY_PRUNING = 1, CASCADE_SCALE_IMAGE = 2, CASCADE_FIND_BIGGEST_OBJECT = 4, CASCADE_DO_ROUGH_SEARCH = 8 };
enum { THRESH_BINARY = 0 };

class Mat {
public:
    int rows, cols;
    uchar *data;
    size_t step;                                        // bytes per row
    Mat() : rows(0), cols(0), data(NULL), step(0), type_(0) {}
    Mat(int r, int c, int type) : rows(0), cols(0), data(NULL), step(0), type_(0) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar &s) : rows(0), cols(0), data(NULL), step(0), type_(0) { create(r, c, type); *this = s; }
    Mat(const IplImage *im, bool copyData = false) : rows(0), cols(0), data(NULL), step(0), type_(0)        // OpenCV 2.x only
    {
        if (!im) return;
        rows = im->height; cols = im->width; step = (size_t)im->widthStep; data = (uchar *)im->imageData;
        type_ = im->nChannels == 1 ? CV_8UC1 : im->nChannels == 3 ? CV_8UC3 : CV_8UC4;
        if (copyData) *this = clone();
    }
    static int elem_size(int type) { return type == CV_8UC1 ? 1 : type == CV_8UC3 ? 3 : 4; }    // CV_8UC4 and CV_32FC1: 4 bytes
    int type() const { return type_; }
    int channels() const { return type_ == CV_8UC3 ? 3 : type_ == CV_8UC4 ? 4 : 1; }
    size_t elemSize() const { return (size_t)elem_size(type_); }
    bool empty() const { return data == NULL || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    void create(int r, int c, int type)
    {
        if (r < 0 || c < 0) throw Exception("Mat::create: negative size");
        if (data && r == rows && c == cols && type == type_ && buf_) return;
        rows = r; cols = c; type_ = type; step = (size_t)c * elem_size(type);
        buf_.reset(new std::vector<uchar>(std::max<size_t>(step * (size_t)r, 1), 0));
        data = buf_->data();
    }
    void release() { buf_.reset(); data = NULL; rows = cols = 0; step = 0; }
    Mat clone() const
    {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
        return m;
    }
    Mat &operator=(const Scalar &s)
    {
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++) {
                if (type_ == CV_32FC1) ((float *)(data + (size_t)y * step))[x] = (float)s.val[0];
                else for (int c = 0; c < channels(); c++) data[(size_t)y * step + (size_t)x * channels() + c] = (uchar)s.val[c];
            }
        return *this;
    }
    Mat operator()(const Rect &roi) const
    {
        if (!(0 <= roi.x && 0 <= roi.width && roi.x + roi.width <= cols && 0 <= roi.y && 0 <= roi.height && roi.y + roi.height <= rows))
            throw Exception("Mat::operator()(Rect): roi outside the image");
        Mat m;
        m.rows = roi.height; m.cols = roi.width; m.type_ = type_; m.step = step; m.buf_ = buf_;
        m.data = data + (size_t)roi.y * step + (size_t)roi.x * elemSize();
        return m;
    }
    template <typename T> T *ptr(int y) { return (T *)(data + (size_t)y * step); }
    template <typename T> const T *ptr(int y) const { return (const T *)(data + (size_t)y * step); }

private:
    int type_;
    std::shared_ptr<std::vector<uchar> > buf_;
};

// dst is (re)allocated like OpenCV's OutputArray::create; src == dst works through a temporary
inline void cvtColor(const Mat &src, Mat &dst, int code)
{
    if (code != COLOR_BGR2GRAY || src.channels() < 3) throw Exception("cvtColor: only BGR(A)2GRAY");
    Mat out(src.rows, src.cols, CV_8UC1);
    ora_bgr2gray(src.data, src.cols, src.rows, (int)src.step, src.channels(), out.data, (int)out.step);
    dst = out;
}
inline void resize(const Mat &src, Mat &dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR)
{
    if (interpolation != INTER_LINEAR) throw Exception("resize: only INTER_LINEAR");
    if (dsize.width == 0 || dsize.height == 0) { dsize.width = cvRound(src.cols * fx); dsize.height = cvRound(src.rows * fy); }
    if (dsize.width <= 0 || dsize.height <= 0) throw Exception("resize: empty destination");
    Mat out(dsize.height, dsize.width, src.type());
    ora_resize_linear(src.data, src.cols, src.rows, (int)src.step, src.channels(), out.data, out.cols, out.rows, (int)out.step);
    dst = out;
}
inline void equalizeHist(const Mat &src, Mat &dst)
{
    if (src.type() != CV_8UC1) throw Exception("equalizeHist: 8UC1 only");
    Mat out(src.rows, src.cols, CV_8UC1);
    ora_equalize_hist(src.data, src.cols, src.rows, (int)src.step, out.data, (int)out.step);
    dst = out;
}
inline void flip(const Mat &src, Mat &dst, int flipCode)
{
    if (flipCode != 1 || src.type() != CV_8UC1) throw Exception("flip: only 8UC1 around the y axis");
    Mat out(src.rows, src.cols, CV_8UC1);
    for (int y = 0; y < src.rows; y++)
        for (int x = 0; x < src.cols; x++) out.data[(size_t)y * out.step + x] = src.data[(size_t)y * src.step + (src.cols - 1 - x)];
    dst = out;
}
inline void absdiff(const Mat &a, const Mat &b, Mat &dst)
{
    if (a.rows != b.rows || a.cols != b.cols || a.type() != CV_8UC1 || b.type() != CV_8UC1) throw Exception("absdiff: size/type mismatch");
    Mat out(a.rows, a.cols, CV_8UC1);
    for (int y = 0; y < a.rows; y++)
        for (int x = 0; x < a.cols; x++) out.data[(size_t)y * out.step + x] = (uchar)abs((int)a.data[(size_t)y * a.step + x] - (int)b.data[(size_t)y * b.step + x]);
    dst = out;
}
inline double threshold(const Mat &src, Mat &dst, double thresh, double maxval, int type)
{
    if (type != THRESH_BINARY || src.type() != CV_8UC1) throw Exception("threshold: only THRESH_BINARY on 8UC1");
    Mat out(src.rows, src.cols, CV_8UC1);
    const int ithr = (int)floor(thresh);                                      // 8U: the threshold is floored first
    const uchar mv = (uchar)std::min(std::max(cvRound(maxval), 0), 255);
    for (int y = 0; y < src.rows; y++)
        for (int x = 0; x < src.cols; x++) out.data[(size_t)y * out.step + x] = src.data[(size_t)y * src.step + x] > ithr ? mv : 0;
    dst = out;
    return thresh;
}
// motion templates (OpenCV 2.4 video/motempl.cpp semantics): the MHI must be a whole,
// continuous 32FC1 image, as the tracker's is
inline void updateMotionHistory(const Mat &silhouette, Mat &mhi, double timestamp, double duration)
{
    if (silhouette.type() != CV_8UC1 || mhi.type() != CV_32FC1 || silhouette.rows != mhi.rows || silhouette.cols != mhi.cols ||
        silhouette.step != (size_t)silhouette.cols || mhi.step != (size_t)mhi.cols * 4)
        throw Exception("updateMotionHistory: layout");
    ora_update_mhi(silhouette.data, (float *)mhi.data, mhi.rows * mhi.cols, timestamp, duration);
}
inline void calcMotionGradient(const Mat &, Mat &, Mat &, double, double, int = 3) {}       // outputs never read (TRK:368-370)
inline void segmentMotion(const Mat &mhi, Mat &segmask, std::vector<Rect> &boundingRects, double timestamp, double segThresh)
{
    if (mhi.type() != CV_32FC1 || mhi.step != (size_t)mhi.cols * 4) throw Exception("segmentMotion: layout");
    std::vector<int32_t> labels((size_t)mhi.rows * mhi.cols);
    int cap = 1 << 16;
    std::vector<int> rects((size_t)cap * 4);
    int n = ora_segment_motion((const float *)mhi.data, mhi.cols, mhi.rows, timestamp, segThresh, labels.data(), rects.data(), cap);
    (void)segmask;
    boundingRects.clear();
    for (int i = 0; i < std::min(n, cap); i++) boundingRects.push_back(Rect(rects[4 * i], rects[4 * i + 1], rects[4 * i + 2], rects[4 * i + 3]));
}
inline void circle(Mat &img, Point center, int radius, const Scalar &color, int thickness = 1, int lineType = 8, int shift = 0)
{
    refcv::DrawCall d;
    d.kind = 1; d.target = img.data;
    d.x0 = center.x; d.y0 = center.y; d.x1 = radius; d.y1 = 0;
    for (int i = 0; i < 4; i++) d.color[i] = color.val[i];
    d.thickness = thickness; d.line_type = lineType; d.shift = shift;
    refcv::draw_log().push_back(d);
}

class CascadeClassifier {
public:
    CascadeClassifier() : c_(NULL) {}
    bool load(const std::string &filename) { c_ = refcv::find_cascade(filename); return c_ != NULL; }
    bool empty() const { return c_ == NULL; }
    // OpenCV >= 3 semantics for new-format models (the oracle's): `flags` is ignored
    void detectMultiScale(const Mat &image, std::vector<Rect> &objects, double scaleFactor = 1.1, int minNeighbors = 3, int flags = 0,
                          Size minSize = Size(), Size maxSize = Size())
    {
        (void)flags;
        objects.clear();
        if (getenv("REFCV_TRACE"))
            fprintf(stderr, "detectMultiScale %dx%d step %zu sf %.17g mn %d min %dx%d max %dx%d cascade %p win %dx%d stages %d\n", image.cols, image.rows, image.step,
                    scaleFactor, minNeighbors, minSize.width, minSize.height, maxSize.width, maxSize.height, (const void *)c_, c_ ? ((const int *)c_)[0] : -1,
                    c_ ? ((const int *)c_)[1] : -1, c_ ? ((const int *)c_)[2] : -1);
        if (!c_ || image.empty()) return;
        if (image.type() != CV_8UC1) throw Exception("detectMultiScale: 8UC1 only");
        int cap = 1 << 14;
        std::vector<int> out((size_t)cap * 4);
        int n = ora_detect_multiscale(c_, image.data, image.cols, image.rows, (int)image.step, scaleFactor, minNeighbors, minSize.width,
                                      minSize.height, maxSize.width, maxSize.height, out.data(), cap, NULL, NULL);
        for (int i = 0; i < std::min(n, cap); i++) objects.push_back(Rect(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]));
    }

private:
    const ora_cascade *c_;
};

}  // namespace cv

#endif
