// minigst.h — a small FUNCTIONAL stand-in for the slice of GLib / GObject / GStreamer 1.x that a GstVideoFilter-based
// element touches: type registration (G_DEFINE_TYPE_WITH_CODE), class / instance / private structs, GParamSpec-backed
// properties with GLib's range validation, signals, GstStructure with typed and nested fields, custom events, pads that
// record what is pushed through them, GQueue, GstBuffer / GstVideoFrame.  TEST INFRASTRUCTURE ONLY.
//
// Two things are compiled against it, each into its own shared object (each gets a private copy of this state):
//   * the reference's own element sources (/root/reference/modules/*/*/src/gst-plugins/kms*detect.cpp, gstnubotracker.cpp),
//     UNMODIFIED -> oracle/_ref/libnubo_ref_elements.so, the CPU reference of the complete element behaviour;
//   * this repo's GStreamer shells (nubomedia-vca_b200/gst/) -> lib/libnubovca_gst_mock.so, so that the shells are built
//     and driven in tests on a machine without GStreamer (this image has none).
// The same driver (tests/mock_gst/harness.cpp) is linked into both; tests compare what the two emit field by field.
//
// Semantics kept from the real libraries because the elements depend on them: g_object_set rejects out-of-range values
// (GLib warns and leaves the property untouched); gst_structure_set replaces a field of the same name and keeps the
// insertion order of the others; gst_structure_get fails when the field is missing or has another type; a
// GST_TYPE_STRUCTURE field is copied in and out; instance and private memory are zero-filled.
#ifndef MINIGST_H
#define MINIGST_H

#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <deque>
#include <map>
#include <string>
#include <vector>

// ---------------------------------------------------------------------------------------------------------------------
// GLib basics
// ---------------------------------------------------------------------------------------------------------------------
typedef char gchar;
typedef int gint;
typedef unsigned int guint;
typedef long glong;
typedef unsigned long gulong;
typedef int gboolean;
typedef double gdouble;
typedef float gfloat;
typedef void *gpointer;
typedef const void *gconstpointer;
typedef unsigned char guint8;
typedef uint32_t guint32;
typedef uint64_t guint64;
typedef int64_t gint64;
typedef size_t gsize;
typedef guint32 GQuark;
typedef gsize GType;

#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#define G_BEGIN_DECLS
#define G_END_DECLS
#define G_GNUC_UNUSED __attribute__((unused))

inline gchar *g_strdup(const gchar *s) { return s ? strdup(s) : NULL; }
inline void g_free(gpointer p) { free(p); }
inline int g_strcmp0(const char *a, const char *b) { return !a ? -(a != b) : !b ? 1 : strcmp(a, b); }
gchar *g_strconcat(const gchar *first, ...);
#define g_snprintf snprintf
inline gchar *g_mkdtemp(gchar *tmpl) { return mkdtemp(tmpl); }
inline int g_remove(const gchar *path) { return remove(path); }
#define g_return_val_if_fail(expr, val) do { if (!(expr)) return (val); } while (0)
#define g_once_init_enter(loc) (*(loc) == 0)
#define g_once_init_leave(loc, v) (*(loc) = (v))

typedef struct { int dummy; } GRegex;
typedef int GRegexCompileFlags;
typedef int GRegexMatchFlags;
#define G_REGEX_MATCH_ANCHORED 16
inline GRegex *g_regex_new(const gchar *, GRegexCompileFlags, GRegexMatchFlags, gpointer) { return (GRegex *)calloc(1, sizeof(GRegex)); }
inline gboolean g_regex_match(const GRegex *, const gchar *, int, gpointer) { return FALSE; }      // no network here: never a URL
inline void g_regex_unref(GRegex *r) { free(r); }

typedef struct { int depth; } GRecMutex;
typedef struct { int depth; } GMutex;
inline void g_rec_mutex_init(GRecMutex *m) { m->depth = 0; }
inline void g_rec_mutex_lock(GRecMutex *m) { m->depth++; }
inline void g_rec_mutex_unlock(GRecMutex *m) { m->depth--; }
inline void g_rec_mutex_clear(GRecMutex *) {}
inline void g_mutex_clear(GMutex *) {}

typedef struct _GQueue { std::deque<gpointer> *q; } GQueue;
inline GQueue *g_queue_new(void) { GQueue *g = (GQueue *)calloc(1, sizeof(GQueue)); g->q = new std::deque<gpointer>(); return g; }
inline void g_queue_push_tail(GQueue *g, gpointer d) { g->q->push_back(d); }
inline gpointer g_queue_pop_head(GQueue *g) { if (g->q->empty()) return NULL; gpointer d = g->q->front(); g->q->pop_front(); return d; }
inline guint g_queue_get_length(GQueue *g) { return (guint)g->q->size(); }
inline void g_queue_free(GQueue *g) { if (g) { delete g->q; free(g); } }

// ---------------------------------------------------------------------------------------------------------------------
// GType / GValue / GParamSpec
// ---------------------------------------------------------------------------------------------------------------------
// fundamental type ids are GType-typed: they travel through varargs (gst_structure_new, g_signal_new)
#define G_TYPE_INVALID ((GType)0)
#define G_TYPE_NONE ((GType)1)
#define G_TYPE_BOOLEAN ((GType)5)
#define G_TYPE_INT ((GType)6)
#define G_TYPE_UINT ((GType)7)
#define G_TYPE_LONG ((GType)8)
#define G_TYPE_ULONG ((GType)9)
#define G_TYPE_INT64 ((GType)10)
#define G_TYPE_UINT64 ((GType)11)
#define G_TYPE_FLOAT ((GType)14)
#define G_TYPE_DOUBLE ((GType)15)
#define G_TYPE_STRING ((GType)16)
#define G_TYPE_POINTER ((GType)17)
#define G_TYPE_BOXED ((GType)18)
#define G_TYPE_OBJECT ((GType)20)
enum { MINIGST_TYPE_STRUCTURE = 40, MINIGST_TYPE_OBJECT = 41, MINIGST_TYPE_ELEMENT = 42, MINIGST_TYPE_BASE_TRANSFORM = 43,
       MINIGST_TYPE_VIDEO_FILTER = 44, MINIGST_TYPE_FIRST_DYNAMIC = 100 };
#define GST_TYPE_STRUCTURE ((GType)MINIGST_TYPE_STRUCTURE)
#define GST_TYPE_OBJECT ((GType)MINIGST_TYPE_OBJECT)
#define GST_TYPE_ELEMENT ((GType)MINIGST_TYPE_ELEMENT)
#define GST_TYPE_BASE_TRANSFORM ((GType)MINIGST_TYPE_BASE_TRANSFORM)
#define GST_TYPE_VIDEO_FILTER ((GType)MINIGST_TYPE_VIDEO_FILTER)

typedef struct _GstStructure GstStructure;

typedef struct _GValue {
    GType g_type;
    union { gint v_int; guint v_uint; glong v_long; guint64 v_uint64; gdouble v_double; gpointer v_pointer; } data;
} GValue;
inline void g_value_set_int(GValue *v, gint i) { v->data.v_int = i; }
inline gint g_value_get_int(const GValue *v) { return v->data.v_int; }
inline void g_value_set_long(GValue *v, glong i) { v->data.v_long = i; }
inline glong g_value_get_long(const GValue *v) { return v->data.v_long; }
void g_value_set_boxed(GValue *v, gconstpointer boxed);        // copies (GstStructure only)
gpointer g_value_dup_boxed(const GValue *v);

typedef enum {
    G_PARAM_READABLE = 1, G_PARAM_WRITABLE = 2, G_PARAM_READWRITE = 3, G_PARAM_CONSTRUCT = 4, G_PARAM_STATIC_NAME = 32,
    G_PARAM_STATIC_NICK = 64, G_PARAM_STATIC_BLURB = 128, G_PARAM_STATIC_STRINGS = 224
} GParamFlags;

typedef struct _GParamSpec {
    const gchar *name, *nick, *blurb;
    GType value_type, owner_type;
    guint flags, param_id;
    glong minimum, maximum, default_value;      // int and long specs
} GParamSpec;
GParamSpec *g_param_spec_int(const gchar *name, const gchar *nick, const gchar *blurb, gint minimum, gint maximum, gint default_value, GParamFlags flags);
GParamSpec *g_param_spec_long(const gchar *name, const gchar *nick, const gchar *blurb, glong minimum, glong maximum, glong default_value, GParamFlags flags);
GParamSpec *g_param_spec_boolean(const gchar *name, const gchar *nick, const gchar *blurb, gboolean default_value, GParamFlags flags);
GParamSpec *g_param_spec_boxed(const gchar *name, const gchar *nick, const gchar *blurb, GType boxed_type, GParamFlags flags);

// ---------------------------------------------------------------------------------------------------------------------
// GObject
// ---------------------------------------------------------------------------------------------------------------------
typedef struct _GTypeClass { GType g_type; } GTypeClass;
typedef struct _GTypeInstance { GTypeClass *g_class; } GTypeInstance;
typedef struct _GObject { GTypeInstance g_type_instance; guint ref_count; gpointer qdata; } GObject;
typedef struct _GObjectClass {
    GTypeClass g_type_class;
    void (*set_property)(GObject *object, guint property_id, const GValue *value, GParamSpec *pspec);
    void (*get_property)(GObject *object, guint property_id, GValue *value, GParamSpec *pspec);
    void (*dispose)(GObject *object);
    void (*finalize)(GObject *object);
    void (*constructed)(GObject *object);
} GObjectClass;
typedef void (*GClassInitFunc)(gpointer g_class, gpointer class_data);
typedef void (*GInstanceInitFunc)(GTypeInstance *instance, gpointer g_class);
typedef int GTypeFlags;

GType g_type_register_static_simple(GType parent, const gchar *name, guint class_size, GClassInitFunc class_init, guint instance_size,
                                    GInstanceInitFunc instance_init, GTypeFlags flags);
gpointer g_type_class_peek_parent(gpointer g_class);
gpointer g_type_class_ref(GType type);                       // creates the class on first use (runs the class_init chain)
const gchar *g_type_name(GType type);
GType g_type_from_name(const gchar *name);
GType g_type_parent(GType type);
gboolean g_type_is_a(GType type, GType is_a);
void g_type_class_add_private(gpointer g_class, gsize private_size);
gpointer g_type_instance_get_private(GTypeInstance *instance, GType private_type);
gpointer minigst_check_instance_cast(gpointer instance, GType type);

#define G_TYPE_FROM_CLASS(k) (((GTypeClass *)(k))->g_type)
#define G_TYPE_FROM_INSTANCE(i) (((GTypeInstance *)(i))->g_class->g_type)
#define G_OBJECT_TYPE(o) G_TYPE_FROM_INSTANCE(o)
#define G_TYPE_CHECK_INSTANCE_CAST(obj, type, T) ((T *)minigst_check_instance_cast((gpointer)(obj), (type)))
#define G_TYPE_CHECK_CLASS_CAST(k, type, T) ((T *)(k))
#define G_TYPE_CHECK_INSTANCE_TYPE(obj, type) ((obj) != NULL && g_type_is_a(G_TYPE_FROM_INSTANCE(obj), (type)))
#define G_TYPE_CHECK_CLASS_TYPE(k, type) ((k) != NULL && g_type_is_a(G_TYPE_FROM_CLASS(k), (type)))
#define G_TYPE_INSTANCE_GET_PRIVATE(obj, type, T) ((T *)g_type_instance_get_private((GTypeInstance *)(obj), (type)))
#define G_TYPE_INSTANCE_GET_CLASS(obj, type, T) ((T *)(((GTypeInstance *)(obj))->g_class))
#define G_OBJECT(o) ((GObject *)(o))
#define G_OBJECT_CLASS(k) ((GObjectClass *)(k))
#define G_OBJECT_GET_CLASS(o) ((GObjectClass *)(((GTypeInstance *)(o))->g_class))
#define G_OBJECT_WARN_INVALID_PROPERTY_ID(object, property_id, pspec) minigst_warn("invalid property id %u", (guint)(property_id))

void minigst_warn(const char *fmt, ...);
int minigst_warning_count(void);

#define G_DEFINE_TYPE_WITH_CODE(TN, t_n, T_P, _C_)                                                                              \
    static void t_n##_init(TN *self);                                                                                           \
    static void t_n##_class_init(TN##Class *klass);                                                                             \
    static gpointer t_n##_parent_class G_GNUC_UNUSED = NULL;                                                                    \
    static void t_n##_class_intern_init(gpointer klass, gpointer)                                                               \
    {                                                                                                                           \
        t_n##_parent_class = g_type_class_peek_parent(klass);                                                                   \
        t_n##_class_init((TN##Class *)klass);                                                                                   \
    }                                                                                                                           \
    GType t_n##_get_type(void)                                                                                                  \
    {                                                                                                                           \
        static GType g_define_type_id = 0;                                                                                      \
        if (g_define_type_id == 0) {                                                                                            \
            g_define_type_id = g_type_register_static_simple(T_P, #TN, sizeof(TN##Class), (GClassInitFunc)t_n##_class_intern_init, \
                                                             sizeof(TN), (GInstanceInitFunc)t_n##_init, (GTypeFlags)0);         \
            { _C_; }                                                                                                            \
        }                                                                                                                       \
        return g_define_type_id;                                                                                                \
    }
#define G_DEFINE_TYPE(TN, t_n, T_P) G_DEFINE_TYPE_WITH_CODE(TN, t_n, T_P, ;)

void g_object_class_install_property(GObjectClass *oclass, guint property_id, GParamSpec *pspec);
GParamSpec *g_object_class_find_property(GObjectClass *oclass, const gchar *name);
GParamSpec **g_object_class_list_properties(GObjectClass *oclass, guint *n_properties);      // g_free the array
gpointer g_object_new(GType type, const gchar *first_property_name, ...);
void g_object_set(gpointer object, const gchar *first_property_name, ...);
void g_object_get(gpointer object, const gchar *first_property_name, ...);
gboolean minigst_object_set_value(gpointer object, const gchar *name, const GValue *value);   // FALSE: unknown name or rejected
gpointer g_object_ref(gpointer object);
void g_object_unref(gpointer object);                                                        // NULL-tolerant here
#define g_clear_object(pp) do { if (*(pp)) { g_object_unref(*(pp)); *(pp) = NULL; } } while (0)

typedef enum { G_SIGNAL_RUN_FIRST = 1, G_SIGNAL_RUN_LAST = 2, G_SIGNAL_ACTION = 32 } GSignalFlags;
typedef void (*GCallback)(void);
#define G_CALLBACK(f) ((GCallback)(f))
guint g_signal_new(const gchar *signal_name, GType itype, GSignalFlags flags, guint class_offset, gpointer accumulator, gpointer accu_data,
                   gpointer c_marshaller, GType return_type, guint n_params, ...);
void g_signal_emit(gpointer instance, guint signal_id, GQuark detail, ...);                  // one G_TYPE_STRING argument supported
gulong g_signal_connect(gpointer instance, const gchar *detailed_signal, GCallback handler, gpointer data);   // handler(instance, const gchar*, data)
struct MiniSignalInfo { std::string name; GType itype; GType return_type; std::vector<GType> params; };
const std::vector<MiniSignalInfo> &minigst_signals(void);
std::vector<std::pair<std::string, std::string> > &minigst_emissions(gpointer instance);   // (signal name, string argument), oldest first

// ---------------------------------------------------------------------------------------------------------------------
// GstStructure / GstEvent
// ---------------------------------------------------------------------------------------------------------------------
struct MiniField {
    std::string name;
    GType type;
    guint64 u;            // BOOLEAN / INT / UINT / LONG / UINT64 (two's complement)
    double d;
    std::string s;
    GstStructure *st;     // owned copy
};
struct _GstStructure {
    std::string name;
    std::vector<MiniField> fields;
};
GstStructure *gst_structure_new_empty(const gchar *name);
GstStructure *gst_structure_new(const gchar *name, const gchar *firstfield, ...);
void gst_structure_set(GstStructure *s, const gchar *fieldname, ...);
gboolean gst_structure_get(const GstStructure *s, const gchar *first_fieldname, ...);
GstStructure *gst_structure_copy(const GstStructure *s);
void gst_structure_free(GstStructure *s);
inline gint gst_structure_n_fields(const GstStructure *s) { return (gint)s->fields.size(); }
inline const gchar *gst_structure_nth_field_name(const GstStructure *s, guint i) { return i < s->fields.size() ? s->fields[i].name.c_str() : NULL; }
inline const gchar *gst_structure_get_name(const GstStructure *s) { return s->name.c_str(); }
gboolean gst_structure_has_field(const GstStructure *s, const gchar *fieldname);
gchar *gst_structure_to_string(const GstStructure *s);                                        // g_free; close to GStreamer's serialisation

typedef enum {
    GST_EVENT_UNKNOWN = 0, GST_EVENT_FLUSH_START = 1, GST_EVENT_EOS = 2, GST_EVENT_CAPS = 3, GST_EVENT_SEGMENT = 4,
    GST_EVENT_CUSTOM_UPSTREAM = 10, GST_EVENT_CUSTOM_DOWNSTREAM = 11
} GstEventType;
typedef struct _GstEvent { GstEventType type; GstStructure *structure; int refs; } GstEvent;
#define GST_EVENT_TYPE(e) ((e)->type)
GstEvent *gst_event_new_custom(GstEventType type, GstStructure *structure);                   // takes the structure
inline const GstStructure *gst_event_get_structure(GstEvent *e) { return e->structure; }
void gst_event_unref(GstEvent *e);

// ---------------------------------------------------------------------------------------------------------------------
// GstObject / GstElement / GstPad / GstBaseTransform / GstVideoFilter
// ---------------------------------------------------------------------------------------------------------------------
typedef guint64 GstClockTime;
#define GST_CLOCK_TIME_NONE ((GstClockTime)-1)
typedef enum { GST_FLOW_OK = 0, GST_FLOW_ERROR = -5 } GstFlowReturn;
typedef enum { GST_PAD_UNKNOWN, GST_PAD_SRC, GST_PAD_SINK } GstPadDirection;
typedef enum { GST_PAD_ALWAYS, GST_PAD_SOMETIMES, GST_PAD_REQUEST } GstPadPresence;
typedef enum { GST_RANK_NONE = 0, GST_RANK_MARGINAL = 64, GST_RANK_SECONDARY = 128, GST_RANK_PRIMARY = 256 } GstRank;
typedef enum { GST_MAP_READ = 1, GST_MAP_WRITE = 2, GST_MAP_READWRITE = 3 } GstMapFlags;

typedef struct _GstCaps { std::string str; } GstCaps;
typedef struct _GstPadTemplate { std::string name; GstPadDirection direction; GstPadPresence presence; GstCaps *caps; } GstPadTemplate;
inline GstCaps *gst_caps_from_string(const gchar *s) { GstCaps *c = new GstCaps(); c->str = s ? s : ""; return c; }
inline GstPadTemplate *gst_pad_template_new(const gchar *name, GstPadDirection d, GstPadPresence p, GstCaps *caps)
{
    GstPadTemplate *t = new GstPadTemplate(); t->name = name; t->direction = d; t->presence = p; t->caps = caps; return t;
}
#define GST_VIDEO_CAPS_MAKE(format)                                                                                             \
    "video/x-raw, format = (string) " format ", width = (int) [ 1, max ], height = (int) [ 1, max ], framerate = (fraction) [ 0, max ]"

typedef struct _GstObject { GObject object; gchar *name; } GstObject;
typedef struct _GstObjectClass { GObjectClass parent_class; } GstObjectClass;
#define GST_OBJECT(o) ((GstObject *)(o))
#define GST_OBJECT_LOCK(o) ((void)(o))
#define GST_OBJECT_UNLOCK(o) ((void)(o))

typedef struct _GstPad {
    GstPadDirection direction;
    gpointer parent;
    std::vector<GstEvent *> *pushed;       // events pushed through this pad, oldest first (owned)
} GstPad;
typedef struct _GstElement { GstObject object; } GstElement;
typedef struct _GstElementClass {
    GstObjectClass parent_class;
    std::vector<GstPadTemplate *> *padtemplates;
    const gchar *longname, *classification, *description, *author;
} GstElementClass;
#define GST_ELEMENT(o) ((GstElement *)(o))
#define GST_ELEMENT_CLASS(k) ((GstElementClass *)(k))
void gst_element_class_add_pad_template(GstElementClass *klass, GstPadTemplate *templ);
void gst_element_class_set_static_metadata(GstElementClass *klass, const gchar *longname, const gchar *classification,
                                           const gchar *description, const gchar *author);
gboolean gst_pad_push_event(GstPad *pad, GstEvent *event);                 // records, takes the event
gboolean gst_pad_event_default(GstPad *pad, GstObject *parent, GstEvent *event);      // sink pad: forwards through the src pad

typedef struct _GstPlugin { int dummy; } GstPlugin;
gboolean gst_element_register(GstPlugin *plugin, const gchar *name, guint rank, GType type);
GType minigst_element_factory_type(const gchar *name);                     // 0 if not registered
guint minigst_element_factory_rank(const gchar *name);

typedef struct _GstBaseTransform { GstElement element; GstPad *sinkpad; GstPad *srcpad; } GstBaseTransform;
typedef struct _GstBaseTransformClass {
    GstElementClass parent_class;
    gboolean passthrough_on_same_caps, transform_ip_on_passthrough;
    gboolean (*sink_event)(GstBaseTransform *trans, GstEvent *event);
    gboolean (*src_event)(GstBaseTransform *trans, GstEvent *event);
    gboolean (*start)(GstBaseTransform *trans);
    gboolean (*stop)(GstBaseTransform *trans);
} GstBaseTransformClass;
#define GST_BASE_TRANSFORM(o) ((GstBaseTransform *)(o))
#define GST_BASE_TRANSFORM_CLASS(k) ((GstBaseTransformClass *)(k))
#define GST_BASE_TRANSFORM_GET_CLASS(o) ((GstBaseTransformClass *)(((GTypeInstance *)(o))->g_class))
#define GST_BASE_TRANSFORM_SINK_PAD(o) (((GstBaseTransform *)(o))->sinkpad)
#define GST_BASE_TRANSFORM_SRC_PAD(o) (((GstBaseTransform *)(o))->srcpad)

typedef struct _GstBuffer { guint8 *data; gsize size; GstClockTime pts, dts; } GstBuffer;
typedef struct _GstMapInfo { guint8 *data; gsize size, maxsize; GstMapFlags flags; } GstMapInfo;
#define GST_BUFFER_PTS(b) ((b)->pts)
#define GST_BUFFER_DTS(b) ((b)->dts)
#define GST_IS_BUFFER(b) ((b) != NULL)
inline gboolean gst_buffer_map(GstBuffer *b, GstMapInfo *info, GstMapFlags flags) { info->data = b->data; info->size = info->maxsize = b->size; info->flags = flags; return TRUE; }
inline void gst_buffer_unmap(GstBuffer *, GstMapInfo *) {}

typedef enum { GST_VIDEO_FORMAT_UNKNOWN = 0, GST_VIDEO_FORMAT_I420 = 2, GST_VIDEO_FORMAT_YV12 = 3, GST_VIDEO_FORMAT_BGRA = 12,
               GST_VIDEO_FORMAT_BGR = 16, GST_VIDEO_FORMAT_NV12 = 23, GST_VIDEO_FORMAT_NV21 = 24 } GstVideoFormat;
#define GST_VIDEO_MAX_PLANES 4
typedef struct _GstVideoFormatInfo { GstVideoFormat format; const gchar *name; } GstVideoFormatInfo;
typedef struct _GstVideoInfo {
    const GstVideoFormatInfo *finfo;
    gint width, height;
    gsize size;
    gsize offset[GST_VIDEO_MAX_PLANES];
    gint stride[GST_VIDEO_MAX_PLANES];
} GstVideoInfo;
typedef struct _GstVideoFrame {
    GstVideoInfo info;
    GstBuffer *buffer;
    gpointer data[GST_VIDEO_MAX_PLANES];
    GstMapInfo map[GST_VIDEO_MAX_PLANES];
} GstVideoFrame;
#define GST_VIDEO_INFO_FORMAT(i) ((i)->finfo ? (i)->finfo->format : GST_VIDEO_FORMAT_UNKNOWN)
#define GST_VIDEO_INFO_WIDTH(i) ((i)->width)
#define GST_VIDEO_INFO_HEIGHT(i) ((i)->height)
#define GST_VIDEO_FRAME_FORMAT(f) GST_VIDEO_INFO_FORMAT(&(f)->info)
#define GST_VIDEO_FRAME_WIDTH(f) ((f)->info.width)
#define GST_VIDEO_FRAME_HEIGHT(f) ((f)->info.height)
#define GST_VIDEO_FRAME_PLANE_DATA(f, p) ((f)->data[p])
#define GST_VIDEO_FRAME_PLANE_STRIDE(f, p) ((f)->info.stride[p])

typedef struct _GstVideoFilter { GstBaseTransform element; gboolean negotiated; GstVideoInfo in_info, out_info; } GstVideoFilter;
typedef struct _GstVideoFilterClass {
    GstBaseTransformClass parent_class;
    gboolean (*set_info)(GstVideoFilter *filter, GstCaps *incaps, GstVideoInfo *in_info, GstCaps *outcaps, GstVideoInfo *out_info);
    GstFlowReturn (*transform_frame)(GstVideoFilter *filter, GstVideoFrame *inframe, GstVideoFrame *outframe);
    GstFlowReturn (*transform_frame_ip)(GstVideoFilter *trans, GstVideoFrame *frame);
} GstVideoFilterClass;
#define GST_VIDEO_FILTER(o) ((GstVideoFilter *)(o))
#define GST_VIDEO_FILTER_CLASS(k) ((GstVideoFilterClass *)(k))

typedef struct _GstMeta { int flags; gconstpointer info; } GstMeta;
typedef struct _GstMetaInfo { GType api, type; gsize size; } GstMetaInfo;

// logging: compiled out (the elements' messages carry no behaviour)
#define GST_DEBUG_CATEGORY_STATIC(cat) static int cat G_GNUC_UNUSED = 0
#define GST_DEBUG_CATEGORY_INIT(cat, name, color, desc) ((void)0)
#define GST_DEBUG_FUNCPTR(f) (f)
inline void minigst_log_sink(const void *, ...) {}
#define GST_DEBUG(...) minigst_log_sink(0, __VA_ARGS__)
#define GST_INFO(...) minigst_log_sink(0, __VA_ARGS__)
#define GST_WARNING(...) minigst_log_sink(0, __VA_ARGS__)
#define GST_ERROR(...) minigst_log_sink(0, __VA_ARGS__)
#define GST_DEBUG_OBJECT(o, ...) minigst_log_sink((o), __VA_ARGS__)
#define GST_INFO_OBJECT(o, ...) minigst_log_sink((o), __VA_ARGS__)
#define GST_WARNING_OBJECT(o, ...) minigst_log_sink((o), __VA_ARGS__)
#define GST_ERROR_OBJECT(o, ...) minigst_log_sink((o), __VA_ARGS__)

#define GST_VERSION_MAJOR 1
#define GST_VERSION_MINOR 5
#define GST_PLUGIN_DEFINE(major, minor, name, description, init, version, license, package, origin)                            \
    extern "C" gboolean minigst_plugin_init_##name(void) { return init((GstPlugin *)NULL); }

// kms-core / libsoup names the elements mention on paths that are never taken here (commented-out metadata, overlay download)
typedef struct _KmsSerializableMeta { GstMeta meta; GstStructure *data; } KmsSerializableMeta;
inline KmsSerializableMeta *kms_buffer_get_serializable_meta(GstBuffer *) { return NULL; }
inline KmsSerializableMeta *kms_buffer_add_serializable_meta(GstBuffer *, GstStructure *) { return NULL; }
typedef struct { const char *data; gsize length; } SoupMessageBody;
typedef struct { GObject parent; SoupMessageBody *response_body; } SoupMessage;
typedef struct { GObject parent; } SoupSession;
inline SoupSession *soup_session_sync_new(void) { return NULL; }
inline SoupMessage *soup_message_new(const char *, const char *) { return NULL; }
inline guint soup_session_send_message(SoupSession *, SoupMessage *) { return 0; }

#endif /* MINIGST_H */
