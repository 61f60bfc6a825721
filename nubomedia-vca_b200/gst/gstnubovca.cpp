// gstnubovca.cpp — GStreamer 1.x shells of the six NUBOMEDIA-VCA elements over libnubovca.so.
//
// Same factory names, rank, pad templates, GObject properties (names, ranges, pspec defaults, the boxed image-to-overlay),
// signals, downstream "message" events and sink-event handling as the reference's elements:
//   nubofacedetector  kmsfacedetect.cpp   (class_init :1015-1118, transform_frame_ip :857-898, send_event :179-249, sink_event :251-280)
//   nuboeyedetector   kmseyedetect.cpp    (:1244-1343, :1107-1141, :220-308, :192-218)
//   nubomouthdetector kmsmouthdetect.cpp  (:1048-1138, :912-946, :200-262, :173-198)
//   nubonosedetector  kmsnosedetect.cpp   (:1060-1150, :915-955, :209-268, :180-207)
//   nuboeardetector   kmseardetect.cpp    (:965-1055, :830-866, :192-290; no sink_event handler)
//   nubotracker       gstnubotracker.cpp  (:470-560, :423-445; no sink_event handler, no downstream event)
// Everything per frame happens behind nv_element_transform_frame_ip (include/nubovca.h); this file only translates between
// GStreamer objects and that C ABI.  One class implementation serves the six GTypes (a descriptor per factory).
//
// Builds against real GStreamer (CMakeLists.txt beside this file, pkg-config gstreamer-video-1.0) and against the mock
// under tests/mock_gst/ (this image has no GStreamer): `make -C nubomedia-vca_b200` produces lib/libnubovca_gst_mock.so,
// which tests/test_gst_shells*.py drive next to the reference's own elements.
//
// Not carried over (out of scope): downloading and blending the image-to-overlay picture.  The
// property is kept (type, name, get/set round trip) so that applications setting it keep working.
// GPU placement: NUBOVCA_GPU=<index> pins every instance of the process; NUBOVCA_GPU=auto spreads instances round-robin
// over the visible devices (streams are independent); default 0.
#include <gst/gst.h>
#include <gst/video/gstvideofilter.h>
#include <gst/video/video.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "nubovca.h"

#ifndef PACKAGE
#define PACKAGE "nubovca"
#endif
#ifndef VERSION
#define VERSION "0.2.0"
#endif

GST_DEBUG_CATEGORY_STATIC(gst_nubovca_debug);
#define GST_CAT_DEFAULT gst_nubovca_debug

typedef struct {
    const char *factory, *type_name, *signal, *caps_format, *longname, *description;
    gboolean overlay_prop;          // the five detectors carry image-to-overlay; the tracker does not
    gboolean own_sink_event;        // face / eye / mouth / nose queue custom downstream events
    gboolean forwards_custom;       // eye / mouth / nose hand every event on (gst_pad_event_default); the face element keeps custom ones
} NuboDesc;

static const NuboDesc DESCS[6] = {
    {"nubofacedetector", "GstNuboVcaFaceDetect", "face-event", "BGR", "face detection filter element", "Face detector", TRUE, TRUE, FALSE},
    {"nuboeyedetector", "GstNuboVcaEyeDetect", "eye-event", "BGR", "eye detection filter element", "Eye detector", TRUE, TRUE, TRUE},
    {"nubomouthdetector", "GstNuboVcaMouthDetect", "mouth-event", "BGR", "mouth detection filter element", "Mouth detector", TRUE, TRUE, TRUE},
    {"nubonosedetector", "GstNuboVcaNoseDetect", "nose-event", "BGR", "nose detection filter element", "Nose detector", TRUE, TRUE, TRUE},
    {"nuboeardetector", "GstNuboVcaEarDetect", "ear-event", "BGR", "ear detection filter element", "Ear detector", TRUE, FALSE, FALSE},
    {"nubotracker", "GstNuboVcaTracker", "tracker-event", "BGRA", "Motion tracker filter element", "Motion Tracker", FALSE, FALSE, FALSE},
};

typedef struct _GstNuboVca {
    GstVideoFilter base;
    const NuboDesc *desc;
    nv_element *el;
    GRecMutex mutex;                 // the reference's per-element GRecMutex (kmsfacedetect.cpp:44-48,873-885)
    GstStructure *image_to_overlay;
} GstNuboVca;

typedef struct _GstNuboVcaClass {
    GstVideoFilterClass parent_class;
    const NuboDesc *desc;
    guint signal_id;
    guint n_props;                   // nv_element properties are ids 1 .. n_props, image-to-overlay is n_props + 1
    gpointer parent;                 // parent class, for chaining up
} GstNuboVcaClass;

#define NUBO(obj) ((GstNuboVca *)(obj))
#define NUBO_GET_CLASS(obj) ((GstNuboVcaClass *)(((GTypeInstance *)(obj))->g_class))

static int pick_gpu(void)
{
    static int next = 0;
    const char *s = getenv("NUBOVCA_GPU");
    if (!s) return 0;
    if (!strcmp(s, "auto")) { int n = nv_device_count(); return n > 0 ? next++ % n : 0; }
    return atoi(s);
}

static void gst_nubovca_set_property(GObject *object, guint property_id, const GValue *value, GParamSpec *pspec)
{
    GstNuboVca *self = NUBO(object);
    GstNuboVcaClass *klass = NUBO_GET_CLASS(object);
    g_rec_mutex_lock(&self->mutex);
    if (klass->desc->overlay_prop && property_id == klass->n_props + 1) {          // kmsfacedetect.cpp:568-574
        if (self->image_to_overlay) gst_structure_free(self->image_to_overlay);
        self->image_to_overlay = (GstStructure *)g_value_dup_boxed(value);
    } else if (property_id >= 1 && property_id <= klass->n_props && self->el) {
        long v = pspec->value_type == G_TYPE_LONG ? g_value_get_long(value) : g_value_get_int(value);
        if (nv_element_set_property(self->el, pspec->name, v) != NV_OK) GST_WARNING_OBJECT(self, "%s: %s", pspec->name, nv_last_error());
    } else
        G_OBJECT_WARN_INVALID_PROPERTY_ID(object, property_id, pspec);
    g_rec_mutex_unlock(&self->mutex);
}

static void gst_nubovca_get_property(GObject *object, guint property_id, GValue *value, GParamSpec *pspec)
{
    GstNuboVca *self = NUBO(object);
    GstNuboVcaClass *klass = NUBO_GET_CLASS(object);
    g_rec_mutex_lock(&self->mutex);
    if (klass->desc->overlay_prop && property_id == klass->n_props + 1) {          // kmsfacedetect.cpp:641-647
        if (!self->image_to_overlay) self->image_to_overlay = gst_structure_new_empty("image_to_overlay");
        g_value_set_boxed(value, self->image_to_overlay);
    } else if (property_id >= 1 && property_id <= klass->n_props && self->el) {
        long v = 0;
        nv_element_get_property(self->el, pspec->name, &v);
        if (pspec->value_type == G_TYPE_LONG) g_value_set_long(value, v);
        else g_value_set_int(value, (gint)v);
    } else
        G_OBJECT_WARN_INVALID_PROPERTY_ID(object, property_id, pspec);
    g_rec_mutex_unlock(&self->mutex);
}

// one custom downstream event -> its fields, as __get_timestamp / __get_event_message would read them (nv_event_field)
static void queue_message(GstNuboVca *self, const GstStructure *message)
{
    enum { MAXF = 256 };
    nv_event_field fields[MAXF] = {};
    gchar *types[MAXF] = {};
    int n = 0;
    gint len = gst_structure_n_fields(message);
    for (gint i = 0; i < len && n < MAXF; i++) {
        const gchar *name = gst_structure_nth_field_name(message, i);
        nv_event_field f;
        memset(&f, 0, sizeof f);
        f.name = name;
        types[n] = NULL;
        GstStructure *data = NULL;
        if (gst_structure_get(message, name, GST_TYPE_STRUCTURE, &data, NULL) && data) {
            guint x = 0, y = 0, w = 0, h = 0;
            f.is_structure = 1;
            if (gst_structure_get(data, "type", G_TYPE_STRING, &types[n], NULL)) f.type = types[n];
            gst_structure_get(data, "x", G_TYPE_UINT, &x, NULL);
            gst_structure_get(data, "y", G_TYPE_UINT, &y, NULL);
            gst_structure_get(data, "width", G_TYPE_UINT, &w, NULL);
            gst_structure_get(data, "height", G_TYPE_UINT, &h, NULL);
            f.rect.x = (int)x; f.rect.y = (int)y; f.rect.width = (int)w; f.rect.height = (int)h;
            gst_structure_free(data);
        }
        fields[n++] = f;
    }
    GST_OBJECT_LOCK(self);
    nv_element_push_message(self->el, fields, n);
    GST_OBJECT_UNLOCK(self);
    for (int i = 0; i < n; i++) g_free(types[i]);
}

static gboolean gst_nubovca_sink_event(GstBaseTransform *trans, GstEvent *event)
{
    GstNuboVca *self = NUBO(trans);
    GstNuboVcaClass *klass = NUBO_GET_CLASS(trans);
    if (GST_EVENT_TYPE(event) == GST_EVENT_CUSTOM_DOWNSTREAM && self->el) {
        const GstStructure *st = gst_event_get_structure(event);
        if (st) queue_message(self, st);
        if (!klass->desc->forwards_custom) {           // the reference's face element keeps custom events to itself (:258-267)
            gst_event_unref(event);
            return TRUE;
        }
    }
    return GST_BASE_TRANSFORM_CLASS(klass->parent)->sink_event(trans, event);
}

static double wall_clock_ms(void)
{
    struct timeval t;
    gettimeofday(&t, NULL);
    return t.tv_sec * 1000.0 + t.tv_usec / 1000.0;
}

static GstFlowReturn gst_nubovca_transform_frame_ip(GstVideoFilter *filter, GstVideoFrame *frame)
{
    GstNuboVca *self = NUBO(filter);
    GstNuboVcaClass *klass = NUBO_GET_CLASS(filter);
    if (!self->el) return GST_FLOW_OK;
    const guint64 pts = GST_BUFFER_PTS(frame->buffer);
    const int W = GST_VIDEO_FRAME_WIDTH(frame), H = GST_VIDEO_FRAME_HEIGHT(frame);
    int rc;
    g_rec_mutex_lock(&self->mutex);
    switch (GST_VIDEO_FRAME_FORMAT(frame)) {
    case GST_VIDEO_FORMAT_I420: case GST_VIDEO_FORMAT_YV12: case GST_VIDEO_FORMAT_NV12: case GST_VIDEO_FORMAT_NV21: {
        // 4:2:0 ingest (only offered when the caps were widened, see NUBOVCA_GST_YUV_CAPS): the decoder's planes as they are
        nv_yuv_frame f;
        memset(&f, 0, sizeof f);
        GstVideoFormat fmt = GST_VIDEO_FRAME_FORMAT(frame);
        f.format = fmt == GST_VIDEO_FORMAT_NV12 ? NV_FMT_NV12 : fmt == GST_VIDEO_FORMAT_NV21 ? NV_FMT_NV21 : NV_FMT_I420;
        f.width = W; f.height = H;
        const int np = (f.format == NV_FMT_I420) ? 3 : 2;
        for (int p = 0; p < np; p++) {
            int src = (fmt == GST_VIDEO_FORMAT_YV12 && p > 0) ? 3 - p : p;          // YV12 = I420 with the chroma planes swapped
            f.plane[p] = (const uint8_t *)GST_VIDEO_FRAME_PLANE_DATA(frame, src);
            f.stride[p] = GST_VIDEO_FRAME_PLANE_STRIDE(frame, src);
        }
        rc = nv_element_transform_frame_yuv(self->el, &f, pts, wall_clock_ms());
        break;
    }
    default:
        rc = nv_element_transform_frame_ip(self->el, (uint8_t *)GST_VIDEO_FRAME_PLANE_DATA(frame, 0), W, H,
                                           GST_VIDEO_FRAME_PLANE_STRIDE(frame, 0), pts, wall_clock_ms());
        break;
    }
    if (rc != NV_OK) GST_ERROR_OBJECT(self, "frame passed through untouched: %s", nv_last_error());       // the reference logs and carries on

    // kms_*_send_event: the downstream custom event ...
    enum { MAXR = 1024 };
    static __thread nv_meta_rect rects[MAXR];
    int n = 0, pushed = 0, has_ts = 1;
    char top[16] = "message";
    nv_element_get_message(self->el, rects, MAXR, &n, &pushed);
    nv_element_get_message_info(self->el, top, &has_ts);
    if (pushed) {
        GstStructure *message = gst_structure_new_empty(top);
        if (has_ts) {
            GstStructure *ts = gst_structure_new("time", "pts", G_TYPE_UINT64, pts, NULL);
            gst_structure_set(message, "timestamp", GST_TYPE_STRUCTURE, ts, NULL);
            gst_structure_free(ts);
        }
        for (int i = 0; i < n; i++) {
            char id[16];
            GstStructure *s = gst_structure_new(rects[i].name, "type", G_TYPE_STRING, rects[i].type, "x", G_TYPE_UINT, rects[i].x, "y",
                                                G_TYPE_UINT, rects[i].y, "width", G_TYPE_UINT, rects[i].width, "height", G_TYPE_UINT,
                                                rects[i].height, NULL);
            g_snprintf(id, sizeof id, "%d", i);
            gst_structure_set(message, id, GST_TYPE_STRUCTURE, s, NULL);
            gst_structure_free(s);
        }
        gst_pad_push_event(GST_BASE_TRANSFORM_SRC_PAD(filter), gst_event_new_custom(GST_EVENT_CUSTOM_DOWNSTREAM, message));
    }
    // ... and the rate-limited application signal
    static __thread char payload[1 << 16];
    int emitted = 0;
    nv_element_get_signal(self->el, payload, sizeof payload, &emitted);
    g_rec_mutex_unlock(&self->mutex);
    if (emitted) g_signal_emit(G_OBJECT(self), klass->signal_id, 0, payload);
    return GST_FLOW_OK;                                   // always, like the reference
}

static void gst_nubovca_finalize(GObject *object)
{
    GstNuboVca *self = NUBO(object);
    GstNuboVcaClass *klass = NUBO_GET_CLASS(object);
    nv_element_destroy(self->el);
    self->el = NULL;
    if (self->image_to_overlay) gst_structure_free(self->image_to_overlay);
    g_rec_mutex_clear(&self->mutex);
    G_OBJECT_CLASS(klass->parent)->finalize(object);
}

static void gst_nubovca_init(GTypeInstance *instance, gpointer g_class)
{
    GstNuboVca *self = NUBO(instance);
    self->desc = ((GstNuboVcaClass *)g_class)->desc;
    g_rec_mutex_init(&self->mutex);
    // cascade directory: NUBOVCA_CASCADE_DIR, else /usr/share/opencv/haarcascades as the reference hard-codes (kmsfacedetect.cpp:40)
    if (nv_element_create(self->desc->factory, pick_gpu(), NULL, &self->el) != NV_OK) {
        GST_ERROR_OBJECT(self, "nv_element_create: %s", nv_last_error());
        self->el = NULL;
    } else if (nv_last_error()[0])
        GST_WARNING_OBJECT(self, "%s", nv_last_error());          // a cascade file is missing: the element runs without it
}

static const NuboDesc *desc_of_type(GType t)
{
    const gchar *name = g_type_name(t);
    for (int i = 0; i < 6; i++) if (!strcmp(DESCS[i].type_name, name)) return &DESCS[i];
    return NULL;
}

static void gst_nubovca_class_init(gpointer g_class, gpointer class_data)
{
    (void)class_data;
    GstNuboVcaClass *klass = (GstNuboVcaClass *)g_class;
    GObjectClass *gobject_class = G_OBJECT_CLASS(g_class);
    GstElementClass *element_class = GST_ELEMENT_CLASS(g_class);
    const NuboDesc *d = klass->desc = desc_of_type(G_TYPE_FROM_CLASS(g_class));
    klass->parent = g_type_class_peek_parent(g_class);

    gchar *caps;
#ifdef NUBOVCA_GST_YUV_CAPS        // opt-in: also accept the decoder's 4:2:0 output (the detectors' result is defined on OpenCV's YUV->BGR)
    caps = g_strconcat("video/x-raw, format = (string) { ", d->caps_format, ", I420, YV12, NV12, NV21 }, width = (int) [ 1, max ], "
                       "height = (int) [ 1, max ], framerate = (fraction) [ 0, max ]", NULL);
#else
    caps = g_strconcat("video/x-raw, format = (string) { ", d->caps_format, " }, width = (int) [ 1, max ], height = (int) [ 1, max ], "
                       "framerate = (fraction) [ 0, max ]", NULL);                        // == GST_VIDEO_CAPS_MAKE("{ BGR }")
#endif
    gst_element_class_add_pad_template(element_class, gst_pad_template_new("src", GST_PAD_SRC, GST_PAD_ALWAYS, gst_caps_from_string(caps)));
    gst_element_class_add_pad_template(element_class, gst_pad_template_new("sink", GST_PAD_SINK, GST_PAD_ALWAYS, gst_caps_from_string(caps)));
    g_free(caps);
    gst_element_class_set_static_metadata(element_class, d->longname, "Video/Filter", d->description, "nubovca-b200");

    gobject_class->set_property = gst_nubovca_set_property;
    gobject_class->get_property = gst_nubovca_get_property;
    gobject_class->finalize = gst_nubovca_finalize;

    // the property table comes from the library (names and ranges of the reference's class_init); like the reference, every
    // pspec default is 0 and the live values are those of *_init()
    nv_element *probe = NULL;
    klass->n_props = 0;
    if (nv_element_create(d->factory, 0, NULL, &probe) == NV_OK) {
        const char *name; long lo, hi, def;
        for (int i = 0; nv_element_property_info(probe, i, &name, &lo, &hi, &def) == NV_OK; i++) {
            gchar *nick = g_strdup(name);
            for (gchar *c = nick; *c; c++) if (*c == '-') *c = ' ';
            GParamSpec *ps = strcmp(name, "set_max_area") == 0                            // gstnubotracker.cpp:523: the one glong property
                                 ? g_param_spec_long(g_strdup(name), nick, nick, lo, hi, 0, (GParamFlags)G_PARAM_READWRITE)
                                 : g_param_spec_int(g_strdup(name), nick, nick, (gint)lo, (gint)hi, 0, (GParamFlags)G_PARAM_READWRITE);
            g_object_class_install_property(gobject_class, ++klass->n_props, ps);
        }
        nv_element_destroy(probe);
    }
    if (d->overlay_prop)
        g_object_class_install_property(gobject_class, klass->n_props + 1,
                                        g_param_spec_boxed("image-to-overlay", "image to overlay", "set the url of the image to overlay the faces",
                                                           GST_TYPE_STRUCTURE, (GParamFlags)(G_PARAM_READWRITE | G_PARAM_STATIC_STRINGS)));

    klass->signal_id = g_signal_new(d->signal, G_TYPE_FROM_CLASS(g_class), G_SIGNAL_RUN_LAST, 0, NULL, NULL, NULL, G_TYPE_NONE, 1, G_TYPE_STRING);

    GST_VIDEO_FILTER_CLASS(g_class)->transform_frame_ip = GST_DEBUG_FUNCPTR(gst_nubovca_transform_frame_ip);
    if (d->own_sink_event) GST_BASE_TRANSFORM_CLASS(g_class)->sink_event = GST_DEBUG_FUNCPTR(gst_nubovca_sink_event);
}

static GType gst_nubovca_get_type(const NuboDesc *d)
{
    GType t = g_type_from_name(d->type_name);
    if (!t)
        t = g_type_register_static_simple(GST_TYPE_VIDEO_FILTER, d->type_name, sizeof(GstNuboVcaClass), gst_nubovca_class_init,
                                          sizeof(GstNuboVca), gst_nubovca_init, (GTypeFlags)0);
    return t;
}

static gboolean plugin_init(GstPlugin *plugin)
{
    GST_DEBUG_CATEGORY_INIT(gst_nubovca_debug, "nubovca", 0, "NUBOMEDIA-VCA elements on libnubovca");
    gboolean ok = TRUE;
    for (int i = 0; i < 6; i++) ok = gst_element_register(plugin, DESCS[i].factory, GST_RANK_NONE, gst_nubovca_get_type(&DESCS[i])) && ok;
    return ok;
}

GST_PLUGIN_DEFINE(GST_VERSION_MAJOR, GST_VERSION_MINOR, nubovca, "NUBOMEDIA-VCA detection elements on libnubovca (B200)", plugin_init, VERSION,
                  "LGPL", PACKAGE, "https://github.com/nubomedia/NUBOMEDIA-VCA")

#ifdef MINIGST_H        // mock build: there is no plugin loader, register at load time
namespace { struct NuboMockRegistrar { NuboMockRegistrar() { plugin_init(NULL); } } nubo_mock_registrar; }
#endif
