// minigst.cpp — state and non-inline functions of the GLib / GObject / GStreamer stand-in (see minigst.h).
// TEST INFRASTRUCTURE ONLY.
#include "minigst.h"

#include <algorithm>

namespace {

struct TypeRec {
    std::string name;
    GType parent = 0;
    guint class_size = 0, instance_size = 0;
    GClassInitFunc class_init = nullptr;
    GInstanceInitFunc instance_init = nullptr;
    gsize private_size = 0;
    void *klass = nullptr;
    std::vector<GParamSpec *> props;
};

struct InstanceRec {
    std::map<GType, void *> priv;
    std::vector<std::pair<std::string, std::pair<GCallback, gpointer>>> handlers;
};

struct Factory { GType type; guint rank; };

struct State {
    std::map<GType, TypeRec> types;
    GType next_type = MINIGST_TYPE_FIRST_DYNAMIC;
    std::map<void *, InstanceRec> instances;
    std::vector<MiniSignalInfo> signals;
    std::map<std::string, Factory> factories;
    int warnings = 0;
};

State &S()
{
    static State *s = nullptr;
    if (!s) {
        s = new State();
        auto add = [&](GType id, const char *name, GType parent, guint csz, guint isz, GClassInitFunc ci, GInstanceInitFunc ii) {
            TypeRec t; t.name = name; t.parent = parent; t.class_size = csz; t.instance_size = isz; t.class_init = ci; t.instance_init = ii;
            s->types[id] = t;
        };
        add(G_TYPE_OBJECT, "GObject", 0, sizeof(GObjectClass), sizeof(GObject),
            [](gpointer k, gpointer) {
                GObjectClass *c = (GObjectClass *)k;
                c->dispose = [](GObject *) {}; c->finalize = [](GObject *) {}; c->constructed = [](GObject *) {};
            }, nullptr);
        add(GST_TYPE_OBJECT, "GstObject", G_TYPE_OBJECT, sizeof(GstObjectClass), sizeof(GstObject), nullptr, nullptr);
        add(GST_TYPE_ELEMENT, "GstElement", GST_TYPE_OBJECT, sizeof(GstElementClass), sizeof(GstElement), nullptr, nullptr);
        add(GST_TYPE_BASE_TRANSFORM, "GstBaseTransform", GST_TYPE_ELEMENT, sizeof(GstBaseTransformClass), sizeof(GstBaseTransform),
            [](gpointer k, gpointer) {
                GstBaseTransformClass *c = (GstBaseTransformClass *)k;
                // GstBaseTransform's default sink_event: (caps / segment bookkeeping, then) forward downstream
                c->sink_event = [](GstBaseTransform *t, GstEvent *e) -> gboolean { return gst_pad_event_default(t->sinkpad, GST_OBJECT(t), e); };
                c->src_event = [](GstBaseTransform *, GstEvent *e) -> gboolean { gst_event_unref(e); return TRUE; };
            },
            [](GTypeInstance *inst, gpointer) {
                GstBaseTransform *t = (GstBaseTransform *)inst;
                for (GstPad **pp : {&t->sinkpad, &t->srcpad}) {
                    GstPad *p = new GstPad();
                    p->direction = pp == &t->sinkpad ? GST_PAD_SINK : GST_PAD_SRC; p->parent = t; p->pushed = new std::vector<GstEvent *>();
                    *pp = p;
                }
            });
        add(GST_TYPE_VIDEO_FILTER, "GstVideoFilter", GST_TYPE_BASE_TRANSFORM, sizeof(GstVideoFilterClass), sizeof(GstVideoFilter), nullptr, nullptr);
    }
    return *s;
}

TypeRec *rec(GType t)
{
    auto it = S().types.find(t);
    return it == S().types.end() ? nullptr : &it->second;
}

GParamSpec *find_pspec(GType t, const char *name)
{
    for (; t; t = rec(t)->parent)
        for (GParamSpec *p : rec(t)->props) if (!strcmp(p->name, name)) return p;
    return nullptr;
}

GParamSpec *new_pspec(const char *name, const char *nick, const char *blurb, GType vt, long lo, long hi, long def, guint flags)
{
    GParamSpec *p = (GParamSpec *)calloc(1, sizeof(GParamSpec));
    p->name = name; p->nick = nick; p->blurb = blurb; p->value_type = vt; p->minimum = lo; p->maximum = hi; p->default_value = def; p->flags = flags;
    return p;
}

void free_field(MiniField &f) { if (f.st) gst_structure_free(f.st); f.st = nullptr; }

// one (fieldname, GType, value) triple of gst_structure_new / gst_structure_set
void set_field_va(GstStructure *s, const char *fieldname, va_list &ap)
{
    MiniField f; f.name = fieldname; f.type = va_arg(ap, GType); f.u = 0; f.d = 0; f.st = nullptr;
    switch (f.type) {
    case G_TYPE_BOOLEAN: case G_TYPE_INT: f.u = (guint64)(gint64)va_arg(ap, int); break;
    case G_TYPE_UINT: f.u = va_arg(ap, unsigned int); break;
    case G_TYPE_LONG: case G_TYPE_INT64: f.u = (guint64)va_arg(ap, gint64); break;
    case G_TYPE_ULONG: case G_TYPE_UINT64: f.u = va_arg(ap, guint64); break;
    case G_TYPE_FLOAT: case G_TYPE_DOUBLE: f.d = va_arg(ap, double); break;
    case G_TYPE_STRING: { const char *v = va_arg(ap, const char *); f.s = v ? v : ""; break; }
    case GST_TYPE_STRUCTURE: { const GstStructure *v = va_arg(ap, const GstStructure *); f.st = v ? gst_structure_copy(v) : nullptr; break; }
    default: minigst_warn("gst_structure_set: unsupported field type %lu for '%s'", (unsigned long)f.type, fieldname); (void)va_arg(ap, gpointer); break;
    }
    for (MiniField &g : s->fields)
        if (g.name == f.name) { free_field(g); g = f; return; }
    s->fields.push_back(f);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
void minigst_warn(const char *fmt, ...)
{
    S().warnings++;
    if (getenv("MINIGST_VERBOSE")) {
        va_list ap; va_start(ap, fmt);
        fprintf(stderr, "minigst WARNING: "); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n");
        va_end(ap);
    }
}
int minigst_warning_count(void) { return S().warnings; }

gchar *g_strconcat(const gchar *first, ...)
{
    std::string r = first ? first : "";
    va_list ap; va_start(ap, first);
    for (const char *p = first ? va_arg(ap, const char *) : nullptr; p; p = va_arg(ap, const char *)) r += p;
    va_end(ap);
    return strdup(r.c_str());
}

void g_value_set_boxed(GValue *v, gconstpointer boxed) { v->data.v_pointer = boxed ? gst_structure_copy((const GstStructure *)boxed) : nullptr; }
gpointer g_value_dup_boxed(const GValue *v) { return v->data.v_pointer ? gst_structure_copy((const GstStructure *)v->data.v_pointer) : nullptr; }

GParamSpec *g_param_spec_int(const gchar *n, const gchar *k, const gchar *b, gint lo, gint hi, gint def, GParamFlags f) { return new_pspec(n, k, b, G_TYPE_INT, lo, hi, def, f); }
GParamSpec *g_param_spec_long(const gchar *n, const gchar *k, const gchar *b, glong lo, glong hi, glong def, GParamFlags f) { return new_pspec(n, k, b, G_TYPE_LONG, lo, hi, def, f); }
GParamSpec *g_param_spec_boolean(const gchar *n, const gchar *k, const gchar *b, gboolean def, GParamFlags f) { return new_pspec(n, k, b, G_TYPE_BOOLEAN, 0, 1, def, f); }
GParamSpec *g_param_spec_boxed(const gchar *n, const gchar *k, const gchar *b, GType boxed, GParamFlags f) { return new_pspec(n, k, b, boxed, 0, 0, 0, f); }

// ---- types ----------------------------------------------------------------------------------------------------------
GType g_type_register_static_simple(GType parent, const gchar *name, guint class_size, GClassInitFunc class_init, guint instance_size,
                                    GInstanceInitFunc instance_init, GTypeFlags)
{
    TypeRec t; t.name = name; t.parent = parent; t.class_size = class_size; t.instance_size = instance_size;
    t.class_init = class_init; t.instance_init = instance_init;
    GType id = S().next_type++;
    S().types[id] = t;
    return id;
}
const gchar *g_type_name(GType t) { TypeRec *r = rec(t); return r ? r->name.c_str() : NULL; }
GType g_type_from_name(const gchar *name)
{
    for (auto &kv : S().types) if (kv.second.name == name) return kv.first;
    return 0;
}
GType g_type_parent(GType t) { TypeRec *r = rec(t); return r ? r->parent : 0; }
gboolean g_type_is_a(GType t, GType is_a)
{
    for (; t; t = g_type_parent(t)) if (t == is_a) return TRUE;
    return FALSE;
}
gpointer g_type_class_ref(GType type)
{
    TypeRec *r = rec(type);
    if (!r) return NULL;
    if (r->klass) return r->klass;
    void *pk = r->parent ? g_type_class_ref(r->parent) : NULL;
    TypeRec *pr = r->parent ? rec(r->parent) : NULL;
    r->klass = calloc(1, r->class_size);
    if (pk) memcpy(r->klass, pk, pr->class_size);                            // a class starts as a copy of its parent class
    ((GTypeClass *)r->klass)->g_type = type;
    if (g_type_is_a(type, GST_TYPE_ELEMENT)) {                                // GstElement's base_init: own pad template list
        GstElementClass *ec = (GstElementClass *)r->klass;
        ec->padtemplates = new std::vector<GstPadTemplate *>(ec->padtemplates ? *ec->padtemplates : std::vector<GstPadTemplate *>());
    }
    if (r->class_init) r->class_init(r->klass, NULL);
    return r->klass;
}
gpointer g_type_class_peek_parent(gpointer g_class)
{
    TypeRec *r = rec(((GTypeClass *)g_class)->g_type);
    return r && r->parent ? g_type_class_ref(r->parent) : NULL;
}
void g_type_class_add_private(gpointer g_class, gsize private_size) { rec(((GTypeClass *)g_class)->g_type)->private_size = private_size; }
gpointer g_type_instance_get_private(GTypeInstance *instance, GType private_type)
{
    auto it = S().instances.find(instance);
    if (it == S().instances.end()) return NULL;
    auto jt = it->second.priv.find(private_type);
    return jt == it->second.priv.end() ? NULL : jt->second;
}
gpointer minigst_check_instance_cast(gpointer instance, GType type)
{
    if (instance && !g_type_is_a(G_TYPE_FROM_INSTANCE(instance), type))
        minigst_warn("invalid cast from '%s' to '%s'", g_type_name(G_TYPE_FROM_INSTANCE(instance)), g_type_name(type));
    return instance;
}

// ---- objects --------------------------------------------------------------------------------------------------------
void g_object_class_install_property(GObjectClass *oclass, guint property_id, GParamSpec *pspec)
{
    GType t = G_TYPE_FROM_CLASS(oclass);
    pspec->param_id = property_id; pspec->owner_type = t;
    rec(t)->props.push_back(pspec);
}
GParamSpec *g_object_class_find_property(GObjectClass *oclass, const gchar *name) { return find_pspec(G_TYPE_FROM_CLASS(oclass), name); }
GParamSpec **g_object_class_list_properties(GObjectClass *oclass, guint *n)
{
    std::vector<GParamSpec *> all;
    std::vector<GType> chain;
    for (GType t = G_TYPE_FROM_CLASS(oclass); t; t = g_type_parent(t)) chain.push_back(t);
    for (auto it = chain.rbegin(); it != chain.rend(); ++it) for (GParamSpec *p : rec(*it)->props) all.push_back(p);
    GParamSpec **arr = (GParamSpec **)calloc(all.size() + 1, sizeof(GParamSpec *));
    std::copy(all.begin(), all.end(), arr);
    if (n) *n = (guint)all.size();
    return arr;
}

gboolean minigst_object_set_value(gpointer object, const gchar *name, const GValue *value)
{
    GParamSpec *p = find_pspec(G_OBJECT_TYPE(object), name);
    if (!p) { minigst_warn("object class '%s' has no property named '%s'", g_type_name(G_OBJECT_TYPE(object)), name); return FALSE; }
    if (!(p->flags & G_PARAM_WRITABLE)) { minigst_warn("property '%s' is not writable", name); return FALSE; }
    if (p->value_type == G_TYPE_INT || p->value_type == G_TYPE_LONG || p->value_type == G_TYPE_BOOLEAN) {
        long v = p->value_type == G_TYPE_LONG ? value->data.v_long : (long)value->data.v_int;
        if (v < p->minimum || v > p->maximum) {     // g_param_value_validate changes the value -> GLib warns and does not set it
            minigst_warn("value \"%ld\" is invalid or out of range for property '%s'", v, name);
            return FALSE;
        }
    }
    GObjectClass *oc = (GObjectClass *)g_type_class_ref(p->owner_type);
    if (!oc->set_property) return FALSE;
    oc->set_property((GObject *)object, p->param_id, value, p);
    return TRUE;
}

static void set_valist(gpointer object, const gchar *name, va_list &ap)
{
    for (; name; name = va_arg(ap, const gchar *)) {
        GParamSpec *p = find_pspec(G_OBJECT_TYPE(object), name);
        if (!p) { minigst_warn("object class '%s' has no property named '%s'", g_type_name(G_OBJECT_TYPE(object)), name); return; }
        GValue v; memset(&v, 0, sizeof v); v.g_type = p->value_type;
        gpointer boxed = NULL;
        if (p->value_type == G_TYPE_LONG) v.data.v_long = va_arg(ap, long);
        else if (p->value_type == G_TYPE_INT || p->value_type == G_TYPE_BOOLEAN) v.data.v_int = va_arg(ap, int);
        else { boxed = va_arg(ap, gpointer); g_value_set_boxed(&v, boxed); }
        minigst_object_set_value(object, name, &v);
        if (p->value_type == GST_TYPE_STRUCTURE && v.data.v_pointer) gst_structure_free((GstStructure *)v.data.v_pointer);
    }
}

gpointer g_object_new(GType type, const gchar *first_property_name, ...)
{
    TypeRec *r = rec(type);
    if (!r) return NULL;
    void *klass = g_type_class_ref(type);
    GObject *o = (GObject *)calloc(1, r->instance_size);
    o->g_type_instance.g_class = (GTypeClass *)klass;
    o->ref_count = 1;
    InstanceRec &ir = S().instances[o];
    std::vector<GType> chain;
    for (GType t = type; t; t = g_type_parent(t)) chain.push_back(t);
    for (auto it = chain.rbegin(); it != chain.rend(); ++it)
        if (rec(*it)->private_size) ir.priv[*it] = calloc(1, rec(*it)->private_size);
    for (auto it = chain.rbegin(); it != chain.rend(); ++it)
        if (rec(*it)->instance_init) rec(*it)->instance_init((GTypeInstance *)o, klass);
    if (first_property_name) { va_list ap; va_start(ap, first_property_name); set_valist(o, first_property_name, ap); va_end(ap); }
    return o;
}
void g_object_set(gpointer object, const gchar *first_property_name, ...)
{
    va_list ap; va_start(ap, first_property_name); set_valist(object, first_property_name, ap); va_end(ap);
}
void g_object_get(gpointer object, const gchar *name, ...)
{
    va_list ap; va_start(ap, name);
    for (; name; name = va_arg(ap, const gchar *)) {
        GParamSpec *p = find_pspec(G_OBJECT_TYPE(object), name);
        if (!p) { minigst_warn("object class '%s' has no property named '%s'", g_type_name(G_OBJECT_TYPE(object)), name); break; }
        GValue v; memset(&v, 0, sizeof v); v.g_type = p->value_type;
        GObjectClass *oc = (GObjectClass *)g_type_class_ref(p->owner_type);
        if (oc->get_property) oc->get_property((GObject *)object, p->param_id, &v, p);
        gpointer dst = va_arg(ap, gpointer);
        if (p->value_type == G_TYPE_LONG) *(glong *)dst = v.data.v_long;
        else if (p->value_type == G_TYPE_INT || p->value_type == G_TYPE_BOOLEAN) *(gint *)dst = v.data.v_int;
        else *(gpointer *)dst = v.data.v_pointer;                            // boxed: the caller owns the copy
    }
    va_end(ap);
}
gpointer g_object_ref(gpointer object) { if (object) ((GObject *)object)->ref_count++; return object; }
void g_object_unref(gpointer object)
{
    if (!object) return;
    GObject *o = (GObject *)object;
    auto it = S().instances.find(o);
    if (it == S().instances.end()) return;                                   // not one of ours
    if (--o->ref_count > 0) return;
    GObjectClass *oc = G_OBJECT_GET_CLASS(o);
    if (oc->dispose) oc->dispose(o);
    if (oc->finalize) oc->finalize(o);
    if (g_type_is_a(G_OBJECT_TYPE(o), GST_TYPE_BASE_TRANSFORM)) {
        GstBaseTransform *t = (GstBaseTransform *)o;
        for (GstPad *p : {t->sinkpad, t->srcpad}) {
            if (!p) continue;
            for (GstEvent *e : *p->pushed) gst_event_unref(e);
            delete p->pushed; delete p;
        }
    }
    for (auto &kv : it->second.priv) free(kv.second);
    S().instances.erase(it);
    free(o);
}

// ---- signals --------------------------------------------------------------------------------------------------------
guint g_signal_new(const gchar *signal_name, GType itype, GSignalFlags, guint, gpointer, gpointer, gpointer, GType return_type, guint n_params, ...)
{
    MiniSignalInfo si; si.name = signal_name; si.itype = itype; si.return_type = return_type;
    va_list ap; va_start(ap, n_params);
    for (guint i = 0; i < n_params; i++) si.params.push_back(va_arg(ap, GType));
    va_end(ap);
    S().signals.push_back(si);
    return (guint)S().signals.size();
}
const std::vector<MiniSignalInfo> &minigst_signals(void) { return S().signals; }
gulong g_signal_connect(gpointer instance, const gchar *detailed_signal, GCallback handler, gpointer data)
{
    InstanceRec &ir = S().instances[instance];
    ir.handlers.push_back({detailed_signal, {handler, data}});
    return (gulong)ir.handlers.size();
}
void g_signal_emit(gpointer instance, guint signal_id, GQuark detail, ...)
{
    if (signal_id == 0 || signal_id > S().signals.size()) { minigst_warn("g_signal_emit: bad signal id %u", signal_id); return; }
    const MiniSignalInfo si = S().signals[signal_id - 1];
    if (!g_type_is_a(G_OBJECT_TYPE(instance), si.itype)) { minigst_warn("signal '%s' is invalid for this instance", si.name.c_str()); return; }
    std::string arg;
    va_list ap; va_start(ap, detail);
    if (si.params.size() == 1 && si.params[0] == G_TYPE_STRING) { const char *a = va_arg(ap, const char *); arg = a ? a : ""; }
    else if (!si.params.empty()) minigst_warn("g_signal_emit: only one G_TYPE_STRING parameter is supported ('%s')", si.name.c_str());
    va_end(ap);
    minigst_emissions(instance).push_back({si.name, arg});
    auto it = S().instances.find(instance);
    if (it != S().instances.end())
        for (auto &h : it->second.handlers)
            if (h.first == si.name) ((void (*)(gpointer, const gchar *, gpointer))h.second.first)(instance, arg.c_str(), h.second.second);
}
std::vector<std::pair<std::string, std::string>> &minigst_emissions(gpointer instance)
{
    static std::map<gpointer, std::vector<std::pair<std::string, std::string>>> log;
    return log[instance];
}

// ---- structures and events ------------------------------------------------------------------------------------------
GstStructure *gst_structure_new_empty(const gchar *name) { GstStructure *s = new GstStructure(); s->name = name ? name : ""; return s; }
GstStructure *gst_structure_new(const gchar *name, const gchar *firstfield, ...)
{
    GstStructure *s = gst_structure_new_empty(name);
    va_list ap; va_start(ap, firstfield);
    for (const char *f = firstfield; f; f = va_arg(ap, const char *)) set_field_va(s, f, ap);
    va_end(ap);
    return s;
}
void gst_structure_set(GstStructure *s, const gchar *fieldname, ...)
{
    va_list ap; va_start(ap, fieldname);
    for (const char *f = fieldname; f; f = va_arg(ap, const char *)) set_field_va(s, f, ap);
    va_end(ap);
}
gboolean gst_structure_get(const GstStructure *s, const gchar *first_fieldname, ...)
{
    gboolean ok = TRUE;
    va_list ap; va_start(ap, first_fieldname);
    for (const char *fn = first_fieldname; fn && ok; fn = va_arg(ap, const char *)) {
        GType want = va_arg(ap, GType);
        gpointer dst = va_arg(ap, gpointer);
        const MiniField *f = nullptr;
        for (const MiniField &g : s->fields) if (g.name == fn) f = &g;
        if (!f || f->type != want) { ok = FALSE; break; }
        switch (want) {
        case G_TYPE_BOOLEAN: case G_TYPE_INT: *(gint *)dst = (gint)f->u; break;
        case G_TYPE_UINT: *(guint *)dst = (guint)f->u; break;
        case G_TYPE_LONG: case G_TYPE_INT64: case G_TYPE_ULONG: case G_TYPE_UINT64: *(guint64 *)dst = f->u; break;
        case G_TYPE_FLOAT: *(gfloat *)dst = (gfloat)f->d; break;
        case G_TYPE_DOUBLE: *(gdouble *)dst = f->d; break;
        case G_TYPE_STRING: *(gchar **)dst = g_strdup(f->s.c_str()); break;
        case GST_TYPE_STRUCTURE: *(GstStructure **)dst = f->st ? gst_structure_copy(f->st) : NULL; break;
        default: ok = FALSE; break;
        }
    }
    va_end(ap);
    return ok;
}
GstStructure *gst_structure_copy(const GstStructure *s)
{
    GstStructure *c = new GstStructure(*s);
    for (MiniField &f : c->fields) if (f.st) f.st = gst_structure_copy(f.st);
    return c;
}
void gst_structure_free(GstStructure *s)
{
    if (!s) return;
    for (MiniField &f : s->fields) free_field(f);
    delete s;
}
gboolean gst_structure_has_field(const GstStructure *s, const gchar *fieldname)
{
    for (const MiniField &f : s->fields) if (f.name == fieldname) return TRUE;
    return FALSE;
}
static void to_string(const GstStructure *s, std::string &out)
{
    out += s->name;
    for (const MiniField &f : s->fields) {
        char b[64];
        out += ", " + f.name + "=";
        switch (f.type) {
        case G_TYPE_BOOLEAN: out += f.u ? "(boolean)true" : "(boolean)false"; break;
        case G_TYPE_INT: snprintf(b, sizeof b, "(int)%d", (int)f.u); out += b; break;
        case G_TYPE_UINT: snprintf(b, sizeof b, "(uint)%u", (guint)f.u); out += b; break;
        case G_TYPE_LONG: case G_TYPE_INT64: snprintf(b, sizeof b, "(gint64)%lld", (long long)f.u); out += b; break;
        case G_TYPE_ULONG: case G_TYPE_UINT64: snprintf(b, sizeof b, "(guint64)%llu", (unsigned long long)f.u); out += b; break;
        case G_TYPE_FLOAT: case G_TYPE_DOUBLE: snprintf(b, sizeof b, "(double)%.17g", f.d); out += b; break;
        case G_TYPE_STRING: out += "(string)" + f.s; break;
        case GST_TYPE_STRUCTURE: out += "(structure){"; if (f.st) to_string(f.st, out); out += "}"; break;
        default: out += "(?)"; break;
        }
    }
    out += ";";
}
gchar *gst_structure_to_string(const GstStructure *s) { std::string o; to_string(s, o); return strdup(o.c_str()); }

GstEvent *gst_event_new_custom(GstEventType type, GstStructure *structure)
{
    GstEvent *e = new GstEvent(); e->type = type; e->structure = structure; e->refs = 1; return e;
}
void gst_event_unref(GstEvent *e)
{
    if (!e || --e->refs > 0) return;
    gst_structure_free(e->structure);
    delete e;
}

// ---- elements -------------------------------------------------------------------------------------------------------
void gst_element_class_add_pad_template(GstElementClass *klass, GstPadTemplate *templ) { klass->padtemplates->push_back(templ); }
void gst_element_class_set_static_metadata(GstElementClass *klass, const gchar *longname, const gchar *classification,
                                           const gchar *description, const gchar *author)
{
    klass->longname = longname; klass->classification = classification; klass->description = description; klass->author = author;
}
gboolean gst_pad_push_event(GstPad *pad, GstEvent *event)
{
    if (!pad || !event) return FALSE;
    pad->pushed->push_back(event);
    return TRUE;
}
gboolean gst_pad_event_default(GstPad *pad, GstObject *parent, GstEvent *event)
{
    // an event arriving on the sink pad of a filter goes out of its src pad
    if (pad && pad->direction == GST_PAD_SINK && parent && g_type_is_a(G_OBJECT_TYPE(parent), GST_TYPE_BASE_TRANSFORM))
        return gst_pad_push_event(((GstBaseTransform *)parent)->srcpad, event);
    gst_event_unref(event);
    return TRUE;
}
gboolean gst_element_register(GstPlugin *, const gchar *name, guint rank, GType type)
{
    if (!g_type_is_a(type, GST_TYPE_ELEMENT)) { minigst_warn("gst_element_register: '%s' is not a GstElement", name); return FALSE; }
    S().factories[name] = Factory{type, rank};
    return TRUE;
}
GType minigst_element_factory_type(const gchar *name)
{
    auto it = S().factories.find(name);
    return it == S().factories.end() ? 0 : it->second.type;
}
guint minigst_element_factory_rank(const gchar *name)
{
    auto it = S().factories.find(name);
    return it == S().factories.end() ? 0 : it->second.rank;
}
