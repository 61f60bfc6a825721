// harness.cpp — C driver around minigst for ctypes: create an element by factory name, set / get properties, send sink
// events, run transform_frame_ip on a frame, and read back what the element pushed, emitted and (reference build only)
// drew.  Linked into BOTH mock builds (the reference's elements and this repo's shells), so the tests drive the two
// through the very same calls.  TEST INFRASTRUCTURE ONLY.
#include <math.h>
#include <sys/time.h>
#include <time.h>

#include <string>

#include "minigst.h"

#ifdef MH_HAVE_REFCV
#include <opencv2/opencv.hpp>
namespace refcv {
std::vector<DrawCall> &draw_log() { static std::vector<DrawCall> v; return v; }
static std::map<std::string, const ora_cascade *> &cascades() { static std::map<std::string, const ora_cascade *> m; return m; }
const ora_cascade *find_cascade(const std::string &path)
{
    size_t s = path.find_last_of('/');
    auto it = cascades().find(s == std::string::npos ? path : path.substr(s + 1));
    return it == cascades().end() ? NULL : it->second;
}
}  // namespace refcv
#endif

#define MH_API extern "C" __attribute__((visibility("default")))

// ---- injected time (see prelude.h): clock() for the tracker's MHI timestamp, gettimeofday() for the events-ms limit ----
static double g_clock_ms = 0, g_wall_ms = 0;
extern "C" clock_t mh_hook_clock(void) { return (clock_t)llround(g_clock_ms * (CLOCKS_PER_SEC / 1000.0)); }
extern "C" int mh_hook_gettimeofday(struct timeval *tv, void *)
{
    long long us = llround(g_wall_ms * 1000.0);
    tv->tv_sec = (time_t)(us / 1000000); tv->tv_usec = (suseconds_t)(us % 1000000);
    return 0;
}
MH_API void mh_set_time(double clock_ms, double wall_ms) { g_clock_ms = clock_ms; g_wall_ms = wall_ms; }

typedef struct { char field[16]; char name[16]; char type[16]; unsigned x, y, width, height; } mh_rect;
typedef struct { int kind; int x0, y0, x1, y1; double color[4]; int thickness, line_type, shift; long long offset; } mh_draw;

#ifdef MH_HAVE_REFCV
MH_API void mh_register_cascade(const char *basename, const void *ora_cascade_ptr)
{
    if (ora_cascade_ptr) refcv::cascades()[basename] = (const ora_cascade *)ora_cascade_ptr;
    else refcv::cascades().erase(basename);
}
// drawing calls recorded since the last mh_clear_draws; offset = target pointer - frame (0 for the frame itself)
MH_API int mh_get_draws(const void *frame, mh_draw *out, int cap)
{
    int n = 0;
    for (const refcv::DrawCall &d : refcv::draw_log()) {
        if (n < cap) {
            mh_draw &o = out[n];
            o.kind = d.kind; o.x0 = d.x0; o.y0 = d.y0; o.x1 = d.x1; o.y1 = d.y1;
            for (int i = 0; i < 4; i++) o.color[i] = d.color[i];
            o.thickness = d.thickness; o.line_type = d.line_type; o.shift = d.shift;
            o.offset = (long long)((const char *)d.target - (const char *)frame);
        }
        n++;
    }
    return n;
}
MH_API void mh_clear_draws(void) { refcv::draw_log().clear(); }
#endif

MH_API int mh_warning_count(void) { return minigst_warning_count(); }

MH_API void *mh_element_new(const char *factory)
{
    GType t = minigst_element_factory_type(factory);
    return t ? g_object_new(t, NULL) : NULL;
}
MH_API void mh_element_free(void *e) { g_object_unref(e); }

MH_API int mh_set_long(void *e, const char *prop, long v)
{
    GParamSpec *p = g_object_class_find_property(G_OBJECT_GET_CLASS(e), prop);
    if (!p) return -1;
    GValue val; memset(&val, 0, sizeof val); val.g_type = p->value_type;
    if (p->value_type == G_TYPE_LONG) val.data.v_long = v;
    else if (p->value_type == G_TYPE_INT || p->value_type == G_TYPE_BOOLEAN) val.data.v_int = (int)v;
    else return -2;
    return minigst_object_set_value(e, prop, &val) ? 0 : -3;
}
MH_API int mh_get_long(void *e, const char *prop, long *v)
{
    GParamSpec *p = g_object_class_find_property(G_OBJECT_GET_CLASS(e), prop);
    if (!p) return -1;
    if (p->value_type == G_TYPE_LONG) { glong x = 0; g_object_get(e, prop, &x, NULL); *v = x; }
    else if (p->value_type == G_TYPE_INT || p->value_type == G_TYPE_BOOLEAN) { gint x = 0; g_object_get(e, prop, &x, NULL); *v = x; }
    else return -2;
    return 0;
}
// boxed GstStructure property (image-to-overlay): `st` stays the caller's
MH_API int mh_set_structure(void *e, const char *prop, void *st)
{
    GParamSpec *p = g_object_class_find_property(G_OBJECT_GET_CLASS(e), prop);
    if (!p || p->value_type != GST_TYPE_STRUCTURE) return -1;
    g_object_set(e, prop, st, NULL);
    return 0;
}
MH_API void *mh_get_structure(void *e, const char *prop)               // the caller frees with mh_st_free
{
    GParamSpec *p = g_object_class_find_property(G_OBJECT_GET_CLASS(e), prop);
    if (!p || p->value_type != GST_TYPE_STRUCTURE) return NULL;
    GstStructure *s = NULL;
    g_object_get(e, prop, &s, NULL);
    return s;
}

static void append(std::string &o, const char *fmt, ...)
{
    char b[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof b, fmt, ap); va_end(ap);
    o += b;
}
// everything class_init declared, as text: factory, rank, type chain, metadata, pad templates, properties, signals
MH_API int mh_describe(const char *factory, char *buf, int cap)
{
    GType t = minigst_element_factory_type(factory);
    if (!t) return -1;
    GstElementClass *ec = (GstElementClass *)g_type_class_ref(t);
    std::string o;
    append(o, "factory|%s|rank=%u\n", factory, minigst_element_factory_rank(factory));
    append(o, "type|%s|parent=%s\n", g_type_name(t), g_type_name(g_type_parent(t)));
    append(o, "metadata|%s|%s|%s|%s\n", ec->longname ? ec->longname : "", ec->classification ? ec->classification : "",
           ec->description ? ec->description : "", ec->author ? ec->author : "");
    for (GstPadTemplate *pt : *ec->padtemplates)
        append(o, "pad|%s|%s|%s|%s\n", pt->name.c_str(), pt->direction == GST_PAD_SRC ? "src" : "sink",
               pt->presence == GST_PAD_ALWAYS ? "always" : "other", pt->caps->str.c_str());
    guint n = 0;
    GParamSpec **ps = g_object_class_list_properties((GObjectClass *)ec, &n);
    for (guint i = 0; i < n; i++) {
        const char *tn = ps[i]->value_type == G_TYPE_INT ? "int" : ps[i]->value_type == G_TYPE_LONG ? "long"
                         : ps[i]->value_type == G_TYPE_BOOLEAN ? "boolean" : ps[i]->value_type == GST_TYPE_STRUCTURE ? "GstStructure" : "?";
        append(o, "property|%s|%s|min=%ld|max=%ld|default=%ld|flags=%u|id=%u|nick=%s\n", ps[i]->name, tn, ps[i]->minimum, ps[i]->maximum,
               ps[i]->default_value, ps[i]->flags & 3u, ps[i]->param_id, ps[i]->nick ? ps[i]->nick : "");
    }
    g_free(ps);
    for (const MiniSignalInfo &si : minigst_signals())
        if (si.itype == t) {
            append(o, "signal|%s|return=%lu|params=", si.name.c_str(), (unsigned long)si.return_type);
            for (GType p : si.params) append(o, "%lu,", (unsigned long)p);
            o += "\n";
        }
    GstVideoFilterClass *vc = (GstVideoFilterClass *)ec;
    append(o, "vfuncs|transform_frame_ip=%d|transform_frame=%d|sink_event=%d\n", vc->transform_frame_ip != NULL, vc->transform_frame != NULL,
           vc->parent_class.sink_event != NULL);
    snprintf(buf, (size_t)cap, "%s", o.c_str());
    return (int)o.size();
}

// ---- GstStructure builders for sink events ----------------------------------------------------------------------------
MH_API void *mh_st_new(const char *name) { return gst_structure_new_empty(name); }
MH_API void mh_st_free(void *s) { gst_structure_free((GstStructure *)s); }
MH_API void mh_st_set_uint(void *s, const char *f, unsigned v) { gst_structure_set((GstStructure *)s, f, G_TYPE_UINT, v, NULL); }
MH_API void mh_st_set_int(void *s, const char *f, int v) { gst_structure_set((GstStructure *)s, f, G_TYPE_INT, v, NULL); }
MH_API void mh_st_set_uint64(void *s, const char *f, unsigned long long v) { gst_structure_set((GstStructure *)s, f, G_TYPE_UINT64, (guint64)v, NULL); }
MH_API void mh_st_set_double(void *s, const char *f, double v) { gst_structure_set((GstStructure *)s, f, G_TYPE_DOUBLE, v, NULL); }
MH_API void mh_st_set_string(void *s, const char *f, const char *v) { gst_structure_set((GstStructure *)s, f, G_TYPE_STRING, v, NULL); }
MH_API void mh_st_set_struct(void *s, const char *f, void *child) { gst_structure_set((GstStructure *)s, f, GST_TYPE_STRUCTURE, (GstStructure *)child, NULL); }
MH_API int mh_st_to_string(void *s, char *buf, int cap)
{
    gchar *t = gst_structure_to_string((GstStructure *)s);
    int n = snprintf(buf, (size_t)cap, "%s", t);
    g_free(t);
    return n;
}

// GstBaseTransformClass::sink_event with a custom event (downstream != 0) or an EOS-like other event; takes `st`
MH_API int mh_send_event(void *e, void *st, int downstream_custom)
{
    GstBaseTransformClass *bc = GST_BASE_TRANSFORM_GET_CLASS(e);
    if (!bc->sink_event) return -1;
    GstEvent *ev = gst_event_new_custom(downstream_custom ? GST_EVENT_CUSTOM_DOWNSTREAM : GST_EVENT_EOS, (GstStructure *)st);
    bc->sink_event((GstBaseTransform *)e, ev);        // the reference's face element returns an uninitialised value: ignored
    return 0;
}

// ---- one buffer -------------------------------------------------------------------------------------------------------
// format: 0 BGR, 1 BGRA (one plane); 2 I420, 3 NV12 (planes laid out back to back in `data`, strides = width / width/2)
// returns 0, or 1 when the element let an exception escape (the streaming thread would have died)
MH_API int mh_transform_frame(void *e, unsigned char *data, int width, int height, int stride, unsigned long long pts_ns, int format)
{
    GstVideoFilterClass *vc = (GstVideoFilterClass *)(((GTypeInstance *)e)->g_class);
    if (!vc->transform_frame_ip) return -1;
    static const GstVideoFormatInfo finfo[4] = {{GST_VIDEO_FORMAT_BGR, "BGR"}, {GST_VIDEO_FORMAT_BGRA, "BGRA"},
                                                {GST_VIDEO_FORMAT_I420, "I420"}, {GST_VIDEO_FORMAT_NV12, "NV12"}};
    GstBuffer buf; memset(&buf, 0, sizeof buf);
    buf.data = data; buf.pts = pts_ns; buf.dts = pts_ns;
    GstVideoFrame fr; memset(&fr, 0, sizeof fr);
    fr.info.finfo = &finfo[format]; fr.info.width = width; fr.info.height = height;
    fr.buffer = &buf;
    if (format <= 1) {
        buf.size = (gsize)stride * height;
        fr.info.stride[0] = stride; fr.data[0] = data;
    } else {
        gsize ysz = (gsize)stride * height;
        fr.info.stride[0] = stride; fr.data[0] = data; fr.info.offset[0] = 0;
        if (format == 2) {
            fr.info.stride[1] = fr.info.stride[2] = stride / 2;
            fr.info.offset[1] = ysz; fr.info.offset[2] = ysz + ysz / 4;
            fr.data[1] = data + ysz; fr.data[2] = data + ysz + ysz / 4;
        } else {
            fr.info.stride[1] = stride; fr.info.offset[1] = ysz; fr.data[1] = data + ysz;
        }
        buf.size = ysz * 3 / 2;
    }
    fr.info.size = buf.size;
    GstVideoFilter *vf = (GstVideoFilter *)e;
    vf->in_info = fr.info; vf->out_info = fr.info; vf->negotiated = TRUE;
    try {
        vc->transform_frame_ip(vf, &fr);
    } catch (...) {
        return 1;
    }
    return 0;
}

// ---- what came out ----------------------------------------------------------------------------------------------------
MH_API int mh_pushed_count(void *e) { return (int)((GstBaseTransform *)e)->srcpad->pushed->size(); }
MH_API void mh_clear_pushed(void *e)
{
    std::vector<GstEvent *> *v = ((GstBaseTransform *)e)->srcpad->pushed;
    for (GstEvent *ev : *v) gst_event_unref(ev);
    v->clear();
}
MH_API int mh_pushed_to_string(void *e, int index, char *buf, int cap)
{
    std::vector<GstEvent *> *v = ((GstBaseTransform *)e)->srcpad->pushed;
    if (index < 0 || index >= (int)v->size()) return -1;
    gchar *t = gst_structure_to_string((*v)[index]->structure);
    int n = snprintf(buf, (size_t)cap, "%d:%s", (int)(*v)[index]->type, t);
    g_free(t);
    return n;
}
// the rectangle sub-structures of pushed event `index` in field order; *pts = its timestamp.pts (or ~0)
MH_API int mh_pushed_rects(void *e, int index, mh_rect *out, int cap, unsigned long long *pts, char *name16)
{
    std::vector<GstEvent *> *v = ((GstBaseTransform *)e)->srcpad->pushed;
    if (index < 0 || index >= (int)v->size()) return -1;
    const GstStructure *m = (*v)[index]->structure;
    if (name16) snprintf(name16, 16, "%s", m->name.c_str());
    if (pts) *pts = ~0ull;
    int n = 0;
    for (const MiniField &f : m->fields) {
        if (f.type != GST_TYPE_STRUCTURE || !f.st) continue;
        if (f.name == "timestamp") {
            for (const MiniField &g : f.st->fields) if (g.name == "pts" && pts) *pts = g.u;
            continue;
        }
        if (n < cap) {
            mh_rect &r = out[n];
            memset(&r, 0, sizeof r);
            snprintf(r.field, sizeof r.field, "%s", f.name.c_str());
            snprintf(r.name, sizeof r.name, "%s", f.st->name.c_str());
            for (const MiniField &g : f.st->fields) {
                if (g.name == "type") snprintf(r.type, sizeof r.type, "%s", g.s.c_str());
                else if (g.name == "x") r.x = (unsigned)g.u;
                else if (g.name == "y") r.y = (unsigned)g.u;
                else if (g.name == "width") r.width = (unsigned)g.u;
                else if (g.name == "height") r.height = (unsigned)g.u;
            }
        }
        n++;
    }
    return n;
}
// signal emissions since the last clear: "signal-name\targument\n" per emission
MH_API int mh_emissions(void *e, char *buf, int cap)
{
    std::string o;
    for (auto &p : minigst_emissions(e)) o += p.first + "\t" + p.second + "\n";
    snprintf(buf, (size_t)cap, "%s", o.c_str());
    return (int)minigst_emissions(e).size();
}
MH_API void mh_clear_emissions(void *e) { minigst_emissions(e).clear(); }

// debugging aid: MH_BACKTRACE_AFTER=<seconds> prints a native backtrace of the calling thread when a harness call runs that long
#include <execinfo.h>
#include <signal.h>
static void mh_alarm_handler(int)
{
    void *bt[64];
    int n = backtrace(bt, 64);
    backtrace_symbols_fd(bt, n, 2);
    _exit(97);
}
MH_API void mh_arm_backtrace(int seconds)
{
    signal(SIGALRM, mh_alarm_handler);
    alarm((unsigned)seconds);
}
