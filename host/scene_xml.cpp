/*
 * host/scene_xml.cpp — Mitsuba scene XML -> plugin objects -> mer_render_desc.
 *
 * A small recursive-descent XML reader plus the part of SceneHandler that the volumetric scenes use
 * (src/librender/scenehandler.cpp): property tags (float, integer, boolean, string, spectrum, rgb, point,
 * vector, transform with translate/scale/rotate/lookat/matrix), object tags, <ref id=...>, <default>, and
 * `$name` substitution from `-D name=value` (:210-219; src/mitsuba/mitsuba.cpp:168-173).
 * Objects are instantiated bottom-up exactly like scenehandler.cpp:712-777: ctor(Properties) -> addChild
 * for every nested object -> configure().
 */
#include <cctype>
#include <fstream>
#include <functional>

#include "mer_host.hpp"

namespace merhost {

/* ------------------------------------------------------------------ tiny XML */
struct XmlNode {
    std::string tag;
    std::map<std::string, std::string> attr;
    std::vector<XmlNode> children;
    int line = 0;
};

class XmlReader {
public:
    explicit XmlReader(const std::string &text) : s(text) {}
    XmlNode parseDocument() {
        skipMisc();
        XmlNode root = parseElement();
        skipMisc();
        if (pos != s.size()) fail("trailing content after the root element");
        return root;
    }
private:
    const std::string &s;
    size_t pos = 0;
    int line = 1;
    [[noreturn]] void fail(const std::string &m) { logError("XML parse error (line " + std::to_string(line) + "): " + m); }
    bool starts(const char *t) const { return s.compare(pos, std::strlen(t), t) == 0; }
    void adv(size_t n = 1) { for (size_t i = 0; i < n && pos < s.size(); i++) { if (s[pos] == '\n') line++; pos++; } }
    void skipWs() { while (pos < s.size() && std::isspace((unsigned char) s[pos])) adv(); }
    void skipMisc() {
        for (;;) {
            skipWs();
            if (starts("<!--")) { size_t e = s.find("-->", pos); if (e == std::string::npos) fail("unterminated comment"); adv(e + 3 - pos); }
            else if (starts("<?")) { size_t e = s.find("?>", pos); if (e == std::string::npos) fail("unterminated declaration"); adv(e + 2 - pos); }
            else if (starts("<!")) { size_t e = s.find('>', pos); if (e == std::string::npos) fail("unterminated doctype"); adv(e + 1 - pos); }
            else return;
        }
    }
    std::string name() {
        size_t b = pos;
        while (pos < s.size() && (std::isalnum((unsigned char) s[pos]) || s[pos] == '_' || s[pos] == '-' || s[pos] == ':' || s[pos] == '.')) adv();
        if (b == pos) fail("expected a name");
        return s.substr(b, pos - b);
    }
    static std::string unescape(const std::string &v) {
        std::string o;
        for (size_t i = 0; i < v.size(); i++) {
            if (v[i] != '&') { o += v[i]; continue; }
            static const std::pair<const char *, char> ents[] = {{"&amp;", '&'}, {"&lt;", '<'}, {"&gt;", '>'}, {"&quot;", '"'}, {"&apos;", '\''}};
            bool hit = false;
            for (auto &e : ents) if (v.compare(i, std::strlen(e.first), e.first) == 0) { o += e.second; i += std::strlen(e.first) - 1; hit = true; break; }
            if (!hit) o += v[i];
        }
        return o;
    }
    XmlNode parseElement() {
        if (pos >= s.size() || s[pos] != '<') fail("expected '<'");
        adv();
        XmlNode n;
        n.line = line;
        n.tag = name();
        for (;;) {
            skipWs();
            if (starts("/>")) { adv(2); return n; }
            if (starts(">")) { adv(); break; }
            std::string k = name();
            skipWs();
            if (!starts("=")) fail("expected '=' after attribute " + k);
            adv();
            skipWs();
            char q = pos < s.size() ? s[pos] : 0;
            if (q != '"' && q != '\'') fail("expected a quoted attribute value");
            adv();
            size_t e = s.find(q, pos);
            if (e == std::string::npos) fail("unterminated attribute value");
            n.attr[k] = unescape(s.substr(pos, e - pos));
            adv(e + 1 - pos);
        }
        for (;;) {
            skipMisc();
            if (starts("</")) {
                adv(2);
                std::string close = name();
                if (close != n.tag) fail("mismatched closing tag </" + close + "> for <" + n.tag + ">");
                skipWs();
                if (!starts(">")) fail("expected '>'");
                adv();
                return n;
            }
            if (pos >= s.size()) fail("unexpected end of file inside <" + n.tag + ">");
            if (s[pos] == '<') n.children.push_back(parseElement());
            else adv(); /* character data is not used by scene files */
        }
    }
};

/* ------------------------------------------------------------------ SceneHandler subset */
namespace {

const std::set<std::string> kObjectTags = {"scene", "integrator", "sensor", "sampler", "film", "rfilter", "medium", "volume",
                                           "phase", "shape", "emitter", "bsdf", "texture", "subsurface"};

struct Loader {
    std::map<std::string, std::string> params;
    std::map<std::string, ObjectRef> byId;
    std::vector<ObjectRef> all;
    std::string sceneDir;

    std::string resolveFile(const std::string &v) const {
        if (v.empty() || v[0] == '/') return v;
        std::ifstream direct(v);
        if (direct) return v;
        return sceneDir + "/" + v;
    }

    std::string subst(const std::string &v, int line) const { /* scenehandler.cpp:210-219 */
        if (v.find('$') == std::string::npos) return v;
        std::string out = v;
        /* longest names first so that $ab is not clobbered by $a */
        std::vector<std::pair<std::string, std::string>> ps(params.begin(), params.end());
        std::sort(ps.begin(), ps.end(), [](const auto &a, const auto &b) { return a.first.size() > b.first.size(); });
        for (auto &kv : ps) {
            std::string key = "$" + kv.first;
            for (size_t p = out.find(key); p != std::string::npos; p = out.find(key, p + kv.second.size())) out.replace(p, key.size(), kv.second);
        }
        if (out.find('$') != std::string::npos)
            logError("The scene referenced an undefined parameter: \"" + out + "\" (line " + std::to_string(line) + ")");
        return out;
    }
    std::string attr(const XmlNode &n, const std::string &k, const char *def = nullptr) const {
        auto it = n.attr.find(k);
        if (it == n.attr.end()) {
            if (def) return def;
            logError("<" + n.tag + "> is missing the attribute \"" + k + "\" (line " + std::to_string(n.line) + ")");
        }
        return subst(it->second, n.line);
    }
    static double num(const std::string &s) {
        char *e = nullptr;
        double d = std::strtod(s.c_str(), &e);
        if (e == s.c_str()) logError("Could not parse \"" + s + "\" as a number");
        return d;
    }
    Vec3 xyz(const XmlNode &n, double def) const {
        if (n.attr.count("value") && !n.attr.count("x")) {
            std::vector<double> v = Properties::numbers(attr(n, "value"));
            if (v.size() == 1) return {v[0], v[0], v[0]};
            if (v.size() == 3) return {v[0], v[1], v[2]};
            logError("<" + n.tag + "> value needs 1 or 3 numbers");
        }
        return {num(attr(n, "x", std::to_string(def).c_str())), num(attr(n, "y", std::to_string(def).c_str())), num(attr(n, "z", std::to_string(def).c_str()))};
    }
    static Vec3 triple(const std::string &s, const char *what) {
        std::vector<double> v = Properties::numbers(s);
        if (v.size() != 3) logError(std::string("lookat: \"") + what + "\" needs three numbers");
        return {v[0], v[1], v[2]};
    }
    Transform transform(const XmlNode &n) const {
        Transform t;
        for (const XmlNode &c : n.children) {
            Transform e;
            if (c.tag == "translate") e = Transform::translate(xyz(c, 0));
            else if (c.tag == "scale") e = Transform::scale(xyz(c, 1));
            else if (c.tag == "rotate") e = Transform::rotate(xyz(c, 0), num(attr(c, "angle")));
            else if (c.tag == "matrix") {
                std::vector<double> v = Properties::numbers(attr(c, "value"));
                if (v.size() != 16) logError("<matrix> needs 16 values");
                for (int i = 0; i < 16; i++) e.m[i] = v[i];
            } else if (c.tag == "lookat") { /* Transform::lookAt, src/libcore/transform.cpp:191-214 */
                Vec3 o = triple(attr(c, "origin"), "origin"), tg = triple(attr(c, "target"), "target"), up = triple(attr(c, "up", "0, 1, 0"), "up");
                Vec3 d = {tg.x - o.x, tg.y - o.y, tg.z - o.z};
                double len = std::sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
                if (len == 0) logError("lookAt(): 'origin' and 'target' coincide!");
                d = {d.x / len, d.y / len, d.z / len};
                Vec3 l = {up.y * d.z - up.z * d.y, up.z * d.x - up.x * d.z, up.x * d.y - up.y * d.x};
                len = std::sqrt(l.x * l.x + l.y * l.y + l.z * l.z);
                if (len == 0) logError("lookAt(): the forward and upward direction must be linearly independent!");
                l = {l.x / len, l.y / len, l.z / len};
                Vec3 u = {d.y * l.z - d.z * l.y, d.z * l.x - d.x * l.z, d.x * l.y - d.y * l.x};
                double mm[16] = {l.x, u.x, d.x, o.x, l.y, u.y, d.y, o.y, l.z, u.z, d.z, o.z, 0, 0, 0, 1};
                std::memcpy(e.m, mm, sizeof(mm));
            } else logError("unknown transform element <" + c.tag + ">");
            t = e * t; /* each new element is applied after the previous ones */
        }
        return t;
    }

    ObjectRef create(const XmlNode &n) { /* PluginManager::createObject */
        const std::string type = attr(n, "type", "");
        ObjectRef o;
        if (n.tag == "volume" && type == "splinevolume") o = std::make_shared<SplineDataSource>();
        else if (n.tag == "volume" && type == "gridvolume") o = std::make_shared<GridDataSource>();
        else if (n.tag == "phase" && type == "hg") o = std::make_shared<HGPhaseFunction>();
        else if (n.tag == "medium" && type == "heterogeneousrefractive") o = std::make_shared<HeterogeneousRefractiveMedium>();
        else if (n.tag == "shape" && (type == "cube" || type == "sphere" || type == "obj" || type == "ply" || type == "serialized")) o = std::make_shared<Shape>();
        else if (n.tag == "bsdf" && (type == "hdielectric" || type == "null" || type == "dielectric")) o = std::make_shared<SurfaceBsdf>();
        else if (n.tag == "bsdf" || n.tag == "texture") o = std::make_shared<Ignored>();
        else if (n.tag == "integrator" || n.tag == "sensor" || n.tag == "sampler" || n.tag == "film" || n.tag == "rfilter" ||
                 n.tag == "emitter" || n.tag == "scene" || (n.tag == "shape" && type == "rectangle")) {
            auto g = std::make_shared<Generic>();
            g->tag = n.tag;
            o = g;
        } else {
            logError("Plugin \"" + type + "\" (<" + n.tag + ">) is not part of the eikonal path of libmitsubaer_b200 (line " + std::to_string(n.line) + ")");
        }
        o->props.pluginName = type;
        o->props.id = attr(n, "id", "");
        return o;
    }

    ObjectRef build(const XmlNode &n) {
        ObjectRef o = create(n);
        for (const XmlNode &c : n.children) {
            const std::string nm = attr(c, "name", "");
            if (c.tag == "default") { if (!params.count(attr(c, "name"))) params[attr(c, "name")] = attr(c, "value"); }
            else if (c.tag == "string" && nm == "filename") o->props.setRaw(nm, c.tag, resolveFile(attr(c, "value")));
            else if (c.tag == "float" || c.tag == "integer" || c.tag == "boolean" || c.tag == "string") o->props.setRaw(nm, c.tag, attr(c, "value"));
            else if (c.tag == "spectrum" || c.tag == "rgb" || c.tag == "srgb") o->props.setRaw(nm, "spectrum", attr(c, "value"));
            else if (c.tag == "point" || c.tag == "vector") { Vec3 v = xyz(c, 0); o->props.setRaw(nm, "point", std::to_string(v.x) + " " + std::to_string(v.y) + " " + std::to_string(v.z)); }
            else if (c.tag == "transform") o->props.setTransform(nm, transform(c));
            else if (c.tag == "ref") {
                auto it = byId.find(attr(c, "id"));
                if (it == byId.end()) logError("Referenced object \"" + attr(c, "id") + "\" not found (line " + std::to_string(c.line) + ")");
                o->addChild(nm, it->second);
            } else if (kObjectTags.count(c.tag)) {
                ObjectRef child = build(c);
                o->addChild(nm, child);
            } else logError("unknown tag <" + c.tag + "> (line " + std::to_string(c.line) + ")");
        }
        o->configure();
        if (!std::dynamic_pointer_cast<Generic>(o) && !std::dynamic_pointer_cast<Ignored>(o)) {
            std::vector<std::string> u = o->props.unqueried();
            if (!u.empty()) logError(std::string(o->className()) + ": unused property \"" + u[0] + "\"");
        }
        if (!o->props.id.empty()) byId[o->props.id] = o;
        all.push_back(o);
        return o;
    }
};

} /* namespace */

Scene loadScene(const std::string &xmlPath, const std::map<std::string, std::string> &params) {
    std::ifstream f(xmlPath);
    if (!f) logError("Unable to open the scene file \"" + xmlPath + "\"");
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string text = ss.str();
    XmlNode root = XmlReader(text).parseDocument();
    if (root.tag != "scene") logError("the root element must be <scene>");
    Loader L;
    L.params = params;
    /* file names are resolved relative to the scene file, like the FileResolver does (src/mitsuba/mitsuba.cpp:339-346) */
    L.sceneDir = xmlPath.find('/') == std::string::npos ? "." : xmlPath.substr(0, xmlPath.rfind('/'));
    auto sceneObj = std::dynamic_pointer_cast<Generic>(L.build(root));

    Scene S;
    std::memset(&S.render, 0, sizeof(S.render));
    mer_render_desc &R = S.render;
    std::shared_ptr<Generic> integrator = sceneObj->child("integrator"), sensor = sceneObj->child("sensor");
    if (!integrator) logError("the scene has no <integrator>");
    if (!sensor) logError("the scene has no <sensor>");
    S.integratorType = integrator->props.pluginName;
    if (S.integratorType != "ervolpath" && S.integratorType != "volpath" && S.integratorType != "volpath_simple")
        logError("integrator \"" + S.integratorType + "\": only the unidirectional volumetric path tracer (ervolpath; volpath is accepted as an alias) is on this path");
    R.max_depth = (int) integrator->props.getInteger("maxDepth", -1); /* MonteCarloIntegrator, integrator.cpp:190-225 */
    R.rr_depth = (int) integrator->props.getInteger("rrDepth", 5);
    if (R.max_depth == 0 || R.max_depth < -1) logError("maxDepth must be set to -1 (infinite) or a value greater than zero!");
    R.pool_paths = (int) integrator->props.getInteger("poolPaths", 0);
    R.steps_per_pass = (int) integrator->props.getInteger("stepsPerPass", 0);
    R.direct_connections = integrator->props.getBoolean("directConnections", false) ? 1 : 0;
    /* multiple importance sampling between the connection and phase sampling, as volpath does it (volpath.cpp:120-147, 430-433) */
    if (integrator->props.getBoolean("misConnections", false)) R.direct_connections = 2;
    R.light_tracing = integrator->props.getBoolean("lightTracing", false) ? 1 : 0; /* emitter-side walk + t = 1 sensor connections */
    const std::string connectionStart = integrator->props.getString("connectionStart", "straight");
    if (connectionStart != "straight" && connectionStart != "random") logError("connectionStart must be \"straight\" or \"random\"");

    if (sensor->props.pluginName != "perspective") logError("sensor \"" + sensor->props.pluginName + "\": only `perspective` is on this path");
    Transform toWorld = sensor->props.getTransform("toWorld", Transform());
    Vec3 o = toWorld.point({0, 0, 0}), d = toWorld.vector({0, 0, 1}), up = toWorld.vector({0, 1, 0});
    R.cam_origin[0] = (float) o.x; R.cam_origin[1] = (float) o.y; R.cam_origin[2] = (float) o.z;
    R.cam_target[0] = (float) (o.x + d.x); R.cam_target[1] = (float) (o.y + d.y); R.cam_target[2] = (float) (o.z + d.z);
    R.cam_up[0] = (float) up.x; R.cam_up[1] = (float) up.y; R.cam_up[2] = (float) up.z;
    R.fov_deg = (float) sensor->props.getFloat("fov", 50.0);
    if (sensor->props.getString("fovAxis", "x") != "x") logError("fovAxis: only \"x\" is carried by this path");
    std::shared_ptr<Generic> sampler = sensor->child("sampler"), film = sensor->child("film");
    R.spp_total = (int) (sampler ? sampler->props.getInteger("sampleCount", 4) : 4);
    R.seed = (uint64_t) (sampler ? sampler->props.getInteger("seed", 20201201) : 20201201);
    R.sample_begin = 0;
    R.sample_stride = 1;
    R.width = (int) (film ? film->props.getInteger("width", 768) : 768); /* film.cpp defaults */
    R.height = (int) (film ? film->props.getInteger("height", 576) : 576);
    if (film) { /* transient film, src/librender/film.cpp:56-78 */
        std::string dec = film->props.getString("decomposition", "none");
        for (auto &c : dec) c = (char) std::tolower((unsigned char) c);
        if (dec != "none" && dec != "transient") logError("The \"decomposition\" parameter must be equal to either \"none\" or \"transient\" on this path");
        const double lo = film->props.getFloat("minBound", 0.0), hi = film->props.getFloat("maxBound", 0.0), bw = film->props.getFloat("binWidth", 1.0);
        std::string mod = film->props.getString("modulation", "none"); /* PathLengthSampler, src/librender/pathlengthsampler.cpp:6-35 */
        for (auto &c : mod) c = (char) std::tolower((unsigned char) c);
        if (mod != "none" && mod != "sine" && mod != "square" && mod != "hamiltonian")
            logError("modulation \"" + mod + "\": none, sine, square and hamiltonian are carried by this path");
        if (mod != "none" && dec == "transient") {
            R.modulation = mod == "sine" ? MER_MODULATION_SINE : mod == "square" ? MER_MODULATION_SQUARE : MER_MODULATION_HAMILTONIAN;
            R.lambda = (float) film->props.getFloat("lambda", 1.0);
            R.phase_deg = (float) film->props.getFloat("phase", 0.0);
            if (!(R.lambda > 0)) logError("modulation: lambda must be positive");
        } else {
            film->props.getFloat("lambda", 1.0);
            film->props.getFloat("phase", 0.0);
        }
        if (dec == "transient") {
            if (!(bw > 0) || !(hi > lo)) logError("transient film: binWidth must be positive and maxBound > minBound");
            R.frames = (int) std::ceil((hi - lo) / bw);
            R.min_bound = (float) lo;
            R.bin_width = (float) bw;
            R.calibrated_transient = film->props.getBoolean("calibratedTransient", false) ? 1 : 0;
            if (R.modulation) R.frames = 0; /* a modulated film has one frame, film.cpp:76-78 */
        }
    }
    std::shared_ptr<Generic> rf = film ? film->child("rfilter") : nullptr;
    std::string rfType = rf ? rf->props.pluginName : "gaussian"; /* hdrfilm default filter */
    if (rfType == "gaussian") R.filter = MER_FILTER_GAUSSIAN;
    else if (rfType == "box") R.filter = MER_FILTER_BOX;
    else logError("rfilter \"" + rfType + "\": gaussian and box are carried by this path");

    bool haveBeam = false;
    for (auto &kv : sceneObj->children) {
        if (auto g = std::dynamic_pointer_cast<Generic>(kv.second)) {
            if (g->tag == "emitter" && g->props.pluginName == "collimated") { /* src/emitters/collimated.cpp:59-110 */
                Transform t = g->props.getTransform("toWorld", Transform());
                Vec3 o = t.point({0, 0, 0}), d = t.vector({0, 0, 1});
                Spectrum3 pw = g->props.getSpectrum("power", 1.0f);
                R.emitter_type = MER_EMITTER_COLLIMATED;
                R.beam_origin[0] = (float) o.x; R.beam_origin[1] = (float) o.y; R.beam_origin[2] = (float) o.z;
                R.beam_direction[0] = (float) d.x; R.beam_direction[1] = (float) d.y; R.beam_direction[2] = (float) d.z;
                for (int i = 0; i < 3; i++) R.beam_power[i] = pw.c[i];
                haveBeam = true;
            } else if (g->tag == "emitter") {
                if (g->props.pluginName != "constant") logError("emitter \"" + g->props.pluginName + "\": constant environment, collimated beam and rectangle area emitters are carried by this path");
                Spectrum3 L = g->props.getSpectrum("radiance", 1.0f);
                for (int i = 0; i < 3; i++) R.env_radiance[i] = L.c[i];
            } else if (g->tag == "shape") { /* rectangle with an area emitter */
                std::shared_ptr<Generic> em = g->child("emitter");
                if (!em || em->props.pluginName != "area") logError("a <shape type=\"rectangle\"> must carry an area emitter on this path");
                Transform t = g->props.getTransform("toWorld", Transform());
                Vec3 q = t.point({-1, -1, 0}), u = t.vector({2, 0, 0}), v = t.vector({0, 2, 0});
                Spectrum3 L = em->props.getSpectrum("radiance", 1.0f);
                R.has_quad = 1;
                R.quad_origin[0] = (float) q.x; R.quad_origin[1] = (float) q.y; R.quad_origin[2] = (float) q.z;
                R.quad_u[0] = (float) u.x; R.quad_u[1] = (float) u.y; R.quad_u[2] = (float) u.z;
                R.quad_v[0] = (float) v.x; R.quad_v[1] = (float) v.y; R.quad_v[2] = (float) v.z;
                for (int i = 0; i < 3; i++) R.quad_radiance[i] = L.c[i];
            }
        } else if (auto sh = std::dynamic_pointer_cast<Shape>(kv.second)) {
            auto med = std::dynamic_pointer_cast<HeterogeneousRefractiveMedium>(sh->interior);
            if (!med) logError("the container shape has no interior heterogeneousrefractive medium");
            if (S.medium) logError("only one heterogeneousrefractive medium per scene is carried by this path");
            med->attach(*sh); /* Shape::addChild("interior") -> medium->m_shape, src/librender/shape.cpp:165-178 */
            S.medium = med;
        }
    }
    if (!S.medium) logError("the scene contains no shape with an interior heterogeneousrefractive medium");
    R.connection = S.medium->connection;
    R.connection.start_mode = connectionStart == "random" ? MER_START_RANDOM : MER_START_STRAIGHT;
    if (R.direct_connections && !R.light_tracing && !R.has_quad) logError("directConnections needs an area emitter (rectangle)");
    if (haveBeam && !R.light_tracing) logError("a collimated beam has no extent: only the lightTracing mode of ervolpath can render it");
    if (R.light_tracing && !haveBeam && !R.has_quad) logError("lightTracing needs a collimated beam or a rectangle area emitter");
    S.keepAlive = L.all;
    return S;
}

void writePFM(const std::string &path, int w, int h, const float *rgb) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) logError("cannot create \"" + path + "\"");
    std::fprintf(f, "PF\n%d %d\n-1.0\n", w, h);
    for (int y = h - 1; y >= 0; y--) std::fwrite(rgb + (size_t) y * w * 3, sizeof(float), (size_t) w * 3, f); /* PFM is bottom-up */
    std::fclose(f);
}

} /* namespace merhost */
