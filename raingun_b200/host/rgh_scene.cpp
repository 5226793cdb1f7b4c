// YAML document -> flattened `rg_scene_desc` (include/raingun_b200.h): the native scene-upload
// front end.  Accepts what the reference's serde derive accepts:
//   * root struct, camelCase keys, deny_unknown_fields, defaults fov=90 / depth=10 / black
//     background ................................. raingun-lib/src/scene.rs:11-31
//   * externally tagged enums Body / Light / Coloration / Surface
//     ............................................ bodies.rs:41-47, lights.rs:22-26, material.rs:20-24,49-54
//   * vectors as {x,y,z} maps or [x,y,z] sequences (cgmath "eders") ... examples/test1.yml:5-8 vs :13
//   * unit variants written `Diffuse` or `Diffuse:` ...................... examples/test1.yml:36 vs :44-45
//   * colours "#rrggbb" only, byte/255 in f32 ............................ color.rs:114-130
//   * texture keys image / x_offset / y_offset, path relative to the CWD .. material.rs:26-47
//   * f32 fields arrive as f64 and are narrowed ......... material.rs:10,30-31,52-53; lights.rs:12,19
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "rgh_internal.h"
#include "rgh_yaml.h"

struct rgh_scene {
    rg_scene_desc desc;
    std::vector<uint8_t> body_kind, coloration_kind, surface_kind, light_kind;
    std::vector<double> body_geom, light_vec;
    std::vector<float> color, texture_offset, albedo, surface_param, light_color, light_intensity;
    std::vector<int32_t> texture_id;
    std::vector<rg_texture_desc> textures;
    std::vector<uint8_t *> texture_pixels;
    std::vector<std::string> texture_paths;
};

namespace rgh {
namespace {

struct SchemaError {
    int code;
    std::string msg;
};
[[noreturn]] void bad(const std::string &msg, int code = RGH_E_SCHEMA) { throw SchemaError{code, msg}; }

std::string describe(const YamlNode &n) {
    double r;
    long long i;
    bool b;
    switch (n.kind) {
        case YamlNode::Seq: return "sequence";
        case YamlNode::Map: return "map";
        default: break;
    }
    switch (yaml_scalar_type(n, &r, &i, &b)) {
        case ScalarType::Null: return "unit value";
        case ScalarType::Bool: return std::string("boolean `") + (b ? "true" : "false") + "`";
        case ScalarType::Int: return "integer `" + n.text + "`";
        case ScalarType::Real: return "floating point `" + n.text + "`";
        default: return "string \"" + n.text + "\"";
    }
}

double num(const YamlNode &n, const std::string &what) {
    double r = 0;
    long long i = 0;
    bool b = false;
    if (n.kind == YamlNode::Scalar) {
        switch (yaml_scalar_type(n, &r, &i, &b)) {
            case ScalarType::Int: return (double)i;
            case ScalarType::Real: return r;
            default: break;
        }
    }
    bad(what + ": invalid type: " + describe(n) + ", expected f64");
}
// serde_yaml hands f64 to the f32 visitor, which narrows with `as f32`
float f32(const YamlNode &n, const std::string &what) { return (float)num(n, what); }

const YamlNode &need(const YamlNode &m, const char *key, const std::string &what) {
    if (m.kind != YamlNode::Map) bad(what + ": invalid type: " + describe(m) + ", expected a struct");
    const YamlNode *v = m.get(key);
    if (!v) bad(what + ": missing field `" + key + "`");
    return *v;
}

void vec3(const YamlNode &n, const std::string &what, double out[3]) {
    if (n.kind == YamlNode::Map) {
        out[0] = num(need(n, "x", what), what + ".x");
        out[1] = num(need(n, "y", what), what + ".y");
        out[2] = num(need(n, "z", what), what + ".z");
        return;
    }
    if (n.kind == YamlNode::Seq && n.items.size() == 3) {
        for (int k = 0; k < 3; ++k) out[k] = num(n.items[(size_t)k], what);
        return;
    }
    bad(what + ": expected [x, y, z] or {x, y, z}");
}

// color.rs:114-130
void color(const YamlNode &n, const std::string &what, float out[3]) {
    if (n.kind != YamlNode::Scalar) bad(what + ": invalid type: " + describe(n) + ", expected a string of a simple hex color (#000000 - #ffffff)");
    double r;
    long long i;
    bool b;
    if (yaml_scalar_type(n, &r, &i, &b) != ScalarType::String)
        bad(what + ": invalid type: " + describe(n) + ", expected a string of a simple hex color (#000000 - #ffffff)");
    const std::string &s = n.text;
    bool ok = s.size() == 7 && s[0] == '#';
    size_t k = 1;
    if (ok && s[1] == '+') k = 2;  // u64::from_str_radix accepts a leading '+'
    for (size_t j = k; ok && j < 7; ++j) ok = std::isxdigit((unsigned char)s[j]) != 0;
    if (!ok) bad(what + ": " + s + " is not a valid color");
    const unsigned long v = std::strtoul(s.c_str() + k, nullptr, 16);
    out[0] = (float)((v & 0xff0000u) >> 16) / 255.0f;
    out[1] = (float)((v & 0x00ff00u) >> 8) / 255.0f;
    out[2] = (float)(v & 0x0000ffu) / 255.0f;
}

// Externally tagged enum: `Name` (unit variants only) or a single-key map {Name: payload}.
const YamlNode *variant(const YamlNode &n, const std::string &what, std::string &name) {
    static const YamlNode null_node;
    if (n.kind == YamlNode::Scalar) {
        name = n.text;
        return &null_node;
    }
    if (n.kind == YamlNode::Map && n.entries.size() == 1) {
        name = n.entries[0].first;
        return &n.entries[0].second;
    }
    bad(what + ": expected a single-key map naming the variant, got " + describe(n));
}

struct Loader {
    const char *texture_root;
    rgh_texture_cb cb;
    void *user;
    int load(const std::string &path, rgh_image *img) const {
        if (cb) return cb(path.c_str(), img, user);
        std::string full = path;
        if (texture_root && texture_root[0] && !(path.size() && path[0] == '/')) full = std::string(texture_root) + "/" + path;
        return rgh_image_open(full.c_str(), img);
    }
};

void build(const YamlNode &doc, const Loader &loader, rgh_scene &s) {
    std::memset(&s.desc, 0, sizeof s.desc);
    s.desc.abi_version = RG_ABI_VERSION;
    s.desc.max_recursion_depth = 10;  // scene.rs:22-31
    s.desc.fov = 90.0;
    static const YamlNode empty;
    const YamlNode *bodies = &empty, *lights = &empty;
    if (doc.kind == YamlNode::Map) {
        for (const auto &e : doc.entries) {
            const std::string &k = e.first;
            if (k == "fov") s.desc.fov = num(e.second, "fov");
            else if (k == "defaultColor") color(e.second, "defaultColor", s.desc.default_color);
            else if (k == "maxRecursionDepth") {
                double r;
                long long i = -1;
                bool b;
                if (e.second.kind != YamlNode::Scalar || yaml_scalar_type(e.second, &r, &i, &b) != ScalarType::Int || i < 0 || i > 0xFFFFFFFFll)
                    bad("maxRecursionDepth: invalid value: " + describe(e.second) + ", expected u32");
                s.desc.max_recursion_depth = (uint32_t)i;
            } else if (k == "bodies") bodies = &e.second;
            else if (k == "lights") lights = &e.second;
            else  // scene.rs:12 deny_unknown_fields
                bad("unknown field `" + k + "`, expected one of `fov`, `defaultColor`, `maxRecursionDepth`, `bodies`, `lights`");
        }
    } else if (doc.kind != YamlNode::Null) {
        bad("invalid type: " + describe(doc) + ", expected struct Scene");
    }
    auto as_seq = [](const YamlNode *n, const char *what) -> const std::vector<YamlNode> & {
        static const std::vector<YamlNode> none;
        if (n->kind == YamlNode::Seq) return n->items;
        if (n->kind == YamlNode::Null && n->line == 0) return none;  // key absent: #[serde(default)]
        bad(std::string(what) + ": invalid type: " + describe(*n) + ", expected a sequence");
    };
    const auto &B = as_seq(bodies, "bodies");
    const auto &L = as_seq(lights, "lights");
    const size_t n = B.size();
    s.body_kind.assign(n, 0);
    s.body_geom.assign(n * 8, 0.0);
    s.coloration_kind.assign(n, 0);
    s.color.assign(n * 3, 0.0f);
    s.texture_id.assign(n, -1);
    s.texture_offset.assign(n * 2, 0.0f);
    s.albedo.assign(n, 0.0f);
    s.surface_kind.assign(n, 0);
    s.surface_param.assign(n * 2, 0.0f);
    std::map<std::string, int> tex_index;
    for (size_t i = 0; i < n; ++i) {
        const std::string what = "bodies[" + std::to_string(i) + "]";
        std::string kind;
        const YamlNode &p = *variant(B[i], what, kind);
        double *g = &s.body_geom[i * 8];
        if (kind == "Sphere") {
            s.body_kind[i] = RG_BODY_SPHERE;
            vec3(need(p, "center", what), what + ".center", g);
            g[3] = num(need(p, "radius", what), what + ".radius");
        } else if (kind == "Plane") {
            s.body_kind[i] = RG_BODY_PLANE;
            vec3(need(p, "origin", what), what + ".origin", g);
            vec3(need(p, "normal", what), what + ".normal", g + 3);
        } else if (kind == "Disk") {
            s.body_kind[i] = RG_BODY_DISK;
            vec3(need(p, "origin", what), what + ".origin", g);
            vec3(need(p, "normal", what), what + ".normal", g + 3);
            g[6] = num(need(p, "radius", what), what + ".radius");
        } else if (kind == "AABB") {
            s.body_kind[i] = RG_BODY_AABB;
            const YamlNode &bounds = need(p, "bounds", what);
            if (bounds.kind != YamlNode::Seq || bounds.items.size() != 2) bad(what + ".bounds: expected an array of length 2");
            vec3(bounds.items[0], what + ".bounds[0]", g);
            vec3(bounds.items[1], what + ".bounds[1]", g + 3);
        } else {
            bad(what + ": unknown variant `" + kind + "`, expected one of `Sphere`, `Plane`, `Disk`, `AABB`");
        }
        const YamlNode &m = need(p, "material", what);
        std::string ckind;
        const YamlNode &cp = *variant(need(m, "coloration", what + ".material"), what + ".coloration", ckind);
        if (ckind == "Color") {
            s.coloration_kind[i] = RG_COLORATION_COLOR;
            color(cp, what + ".coloration", &s.color[i * 3]);
        } else if (ckind == "Texture") {
            s.coloration_kind[i] = RG_COLORATION_TEXTURE;
            const YamlNode &img = need(cp, "image", what + ".Texture");
            if (img.kind != YamlNode::Scalar) bad(what + ".Texture.image: invalid type: " + describe(img) + ", expected a string");
            const std::string &path = img.text;
            auto it = tex_index.find(path);
            if (it == tex_index.end()) {
                rgh_image im;
                std::memset(&im, 0, sizeof im);
                if (loader.load(path, &im) != 0 || !im.pixels) {
                    const char *why = rgh_last_error();
                    bad("Could not load texture file " + path + ": " + (why && *why ? why : "loader failed"));  // material.rs:43-46
                }
                if (im.channels == 1) {  // L8 -> RGB8: get_pixel(x, y) of a grey image is (l, l, l, 255)
                    uint8_t *rgb = (uint8_t *)rgh_alloc((size_t)im.width * im.height * 3);
                    for (size_t k = 0; k < (size_t)im.width * im.height; ++k) rgb[3 * k] = rgb[3 * k + 1] = rgb[3 * k + 2] = im.pixels[k];
                    rgh_free(im.pixels);
                    im.pixels = rgb;
                    im.channels = 3;
                }
                it = tex_index.emplace(path, (int)s.textures.size()).first;
                rg_texture_desc td;
                td.width = im.width;
                td.height = im.height;
                td.channels = im.channels;
                td.reserved = 0;
                td.pixels = im.pixels;
                s.textures.push_back(td);
                s.texture_pixels.push_back(im.pixels);
                s.texture_paths.push_back(path);
            }
            s.texture_id[i] = it->second;
            s.texture_offset[i * 2] = f32(need(cp, "x_offset", what + ".Texture"), what + ".x_offset");
            s.texture_offset[i * 2 + 1] = f32(need(cp, "y_offset", what + ".Texture"), what + ".y_offset");
        } else {
            bad(what + ".coloration: unknown variant `" + ckind + "`, expected `Color` or `Texture`");
        }
        s.albedo[i] = f32(need(m, "albedo", what + ".material"), what + ".albedo");
        const YamlNode &sn = need(m, "surface", what + ".material");
        std::string skind;
        const YamlNode &sp = *variant(sn, what + ".surface", skind);
        if (skind == "Diffuse") {
            if (sp.kind != YamlNode::Null) bad(what + ".surface: invalid type: " + describe(sp) + ", expected unit variant Surface::Diffuse");
            s.surface_kind[i] = RG_SURFACE_DIFFUSE;
        } else if (skind == "Reflecting" && sn.kind == YamlNode::Map) {
            s.surface_kind[i] = RG_SURFACE_REFLECTING;
            s.surface_param[i * 2] = f32(need(sp, "reflectivity", what + ".Reflecting"), what + ".reflectivity");
        } else if (skind == "Refractive" && sn.kind == YamlNode::Map) {
            s.surface_kind[i] = RG_SURFACE_REFRACTIVE;
            s.surface_param[i * 2] = f32(need(sp, "index", what + ".Refractive"), what + ".index");
            s.surface_param[i * 2 + 1] = f32(need(sp, "transparency", what + ".Refractive"), what + ".transparency");
        } else {
            bad(what + ".surface: unknown or non-unit variant `" + skind + "`, expected one of `Diffuse`, `Reflecting`, `Refractive`");
        }
    }
    const size_t nl = L.size();
    s.light_kind.assign(nl, 0);
    s.light_vec.assign(nl * 3, 0.0);
    s.light_color.assign(nl * 3, 0.0f);
    s.light_intensity.assign(nl, 0.0f);
    for (size_t i = 0; i < nl; ++i) {
        const std::string what = "lights[" + std::to_string(i) + "]";
        std::string kind;
        const YamlNode &p = *variant(L[i], what, kind);
        if (kind == "Directional") {
            s.light_kind[i] = RG_LIGHT_DIRECTIONAL;
            vec3(need(p, "direction", what), what + ".direction", &s.light_vec[i * 3]);
        } else if (kind == "Spherical") {
            s.light_kind[i] = RG_LIGHT_SPHERICAL;
            vec3(need(p, "position", what), what + ".position", &s.light_vec[i * 3]);
        } else {
            bad(what + ": unknown variant `" + kind + "`, expected `Directional` or `Spherical`");
        }
        color(need(p, "color", what), what + ".color", &s.light_color[i * 3]);
        s.light_intensity[i] = f32(need(p, "intensity", what), what + ".intensity");
    }
    rg_scene_desc &d = s.desc;
    d.n_bodies = (uint32_t)n;
    d.n_lights = (uint32_t)nl;
    d.n_textures = (uint32_t)s.textures.size();
    d.body_kind = s.body_kind.data();
    d.body_geom = s.body_geom.data();
    d.coloration_kind = s.coloration_kind.data();
    d.color = s.color.data();
    d.texture_id = s.texture_id.data();
    d.texture_offset = s.texture_offset.data();
    d.albedo = s.albedo.data();
    d.surface_kind = s.surface_kind.data();
    d.surface_param = s.surface_param.data();
    d.light_kind = s.light_kind.data();
    d.light_vec = s.light_vec.data();
    d.light_color = s.light_color.data();
    d.light_intensity = s.light_intensity.data();
    d.textures = s.textures.empty() ? nullptr : s.textures.data();
}

}  // namespace
}  // namespace rgh

extern "C" {

int rgh_scene_parse(const char *yaml, size_t len, const char *texture_root, rgh_texture_cb loader, void *user,
                    rgh_scene **out) {
    if (!yaml || !out) return rgh::set_error(RGH_E_INVALID, "rgh_scene_parse: null argument");
    *out = nullptr;
    rgh::YamlNode doc;
    std::string err;
    int rc;
    try {
        rc = rgh::yaml_parse(yaml, len, doc, err);
    } catch (const std::exception &e) {
        return rgh::set_error(RGH_E_FORMAT, std::string("Could not load YAML: ") + e.what());
    }
    if (rc != RGH_OK) return rgh::set_error(rc, "Could not load YAML: " + err);
    rgh_scene *s = new rgh_scene();
    try {
        rgh::build(doc, rgh::Loader{texture_root, loader, user}, *s);
    } catch (const rgh::SchemaError &e) {
        rgh_scene_destroy(s);
        return rgh::set_error(e.code, "Could not load YAML: " + e.msg);
    } catch (const std::exception &e) {   // nothing unwinds across the C ABI
        rgh_scene_destroy(s);
        return rgh::set_error(RGH_E_FORMAT, std::string("Could not load YAML: ") + e.what());
    }
    *out = s;
    return RGH_OK;
}

int rgh_scene_load(const char *path, const char *texture_root, rgh_scene **out) {
    if (!path || !out) return rgh::set_error(RGH_E_INVALID, "rgh_scene_load: null argument");
    std::vector<uint8_t> text;
    if (!rgh::read_file(path, text)) return rgh::set_error(RGH_E_IO, std::string("Could not open input file ") + path);
    return rgh_scene_parse((const char *)text.data(), text.size(), texture_root, nullptr, nullptr, out);
}

const rg_scene_desc *rgh_scene_desc(const rgh_scene *scene) { return scene ? &scene->desc : nullptr; }

void rgh_scene_limit_depth(rgh_scene *scene, uint32_t limit) {
    if (scene && limit < scene->desc.max_recursion_depth) scene->desc.max_recursion_depth = limit;
}

const char *rgh_scene_texture_path(const rgh_scene *scene, uint32_t i) {
    return scene && i < scene->texture_paths.size() ? scene->texture_paths[i].c_str() : nullptr;
}

void rgh_scene_destroy(rgh_scene *scene) {
    if (!scene) return;
    for (uint8_t *p : scene->texture_pixels) rgh_free(p);
    delete scene;
}

}  // extern "C"
