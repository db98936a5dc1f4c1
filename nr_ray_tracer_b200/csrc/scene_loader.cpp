// scene_loader.cpp — native scene loader: TOML / JSON scene file -> nrrt_graph_desc + camera.
//
// C++ stand-in for the reference's Rust loader (no Rust toolchain in the build image):
//   SceneConfig::try_load_scene (format by extension)            ray-tracer/src/scene_config.rs:475-492
//   TextureConfig / MaterialConfig / ObjectConfig builders         scene_config.rs:53-122, 163-197, 278-380
//   try_build_aux: textures -> materials -> instances -> objects   scene_config.rs:411-473
//   CameraConfig get_size / merge_with / try_update                ray-tracer/src/cli.rs:272-402
// Tolerant of the three schema generations found in scenes/ (SURVEY.md note B): [id, cfg] pair arrays (v3),
// tables keyed by id (v2) and anonymous arrays addressed by index with objects under `objects` (v1).
// Its output is checked field by field against the Python loader (tests/test_native_loader.py).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/nrrt.h"
#include "jpeg_baseline.hpp"

namespace nrrt_loader {

struct LoadError {
    std::string msg;
};
[[noreturn]] static void fail(const std::string& m) { throw LoadError{m}; }

// ---------------------------------------------------------------------------------------------- generic value
struct Value;
using ValuePtr = std::shared_ptr<Value>;
struct Value {
    enum Kind { Null, Bool, Int, Float, String, Array, Table } kind = Null;
    bool b = false;
    long long i = 0;
    double f = 0.0;
    std::string s;
    std::vector<ValuePtr> arr;
    std::vector<std::pair<std::string, ValuePtr>> tab;  // insertion order preserved
    bool defined_inline = false;                         // TOML: closed by an inline table / value

    static ValuePtr make(Kind k) {
        auto v = std::make_shared<Value>();
        v->kind = k;
        return v;
    }
    ValuePtr get(const std::string& key) const {
        for (auto& kv : tab)
            if (kv.first == key) return kv.second;
        return nullptr;
    }
    bool is_num() const { return kind == Int || kind == Float; }
    double num() const { return kind == Int ? (double)i : f; }
};

// ---------------------------------------------------------------------------------------------- JSON
// Number grammar shared by the two parsers (serde_json and the toml crate are strict about it): an optional sign
// ('-' only in JSON), an integer part without leading zeros, an optional fraction with at least one digit, an optional
// exponent with at least one digit.  TOML additionally allows '+' and '_' between digits (already removed here).
static bool strict_number(const std::string& t, bool toml) {
    size_t i = 0, n = t.size();
    if (i < n && (t[i] == '-' || (toml && t[i] == '+'))) ++i;
    if (i >= n || !isdigit((unsigned char)t[i])) return false;
    if (t[i] == '0') {
        ++i;
    } else {
        while (i < n && isdigit((unsigned char)t[i])) ++i;
    }
    if (i < n && t[i] == '.') {
        ++i;
        if (i >= n || !isdigit((unsigned char)t[i])) return false;
        while (i < n && isdigit((unsigned char)t[i])) ++i;
    }
    if (i < n && (t[i] == 'e' || t[i] == 'E')) {
        ++i;
        if (i < n && (t[i] == '-' || t[i] == '+')) ++i;
        if (i >= n || !isdigit((unsigned char)t[i])) return false;
        while (i < n && isdigit((unsigned char)t[i])) ++i;
    }
    return i == n;
}

class Json {
  public:
    explicit Json(const std::string& t) : s_(t) {}
    ValuePtr parse() {
        ws();
        ValuePtr v = value();
        ws();
        if (p_ != s_.size()) err("trailing characters");
        return v;
    }

  private:
    const std::string& s_;
    size_t p_ = 0;
    [[noreturn]] void err(const std::string& m) { fail("JSON: " + m + " at offset " + std::to_string(p_)); }
    void ws() {
        while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t' || s_[p_] == '\n' || s_[p_] == '\r')) ++p_;
    }
    ValuePtr value() {
        if (p_ >= s_.size()) err("unexpected end");
        char c = s_[p_];
        if (c == '{') return object();
        if (c == '[') return array();
        if (c == '"') {
            auto v = Value::make(Value::String);
            v->s = string();
            return v;
        }
        if (s_.compare(p_, 4, "true") == 0) {
            p_ += 4;
            auto v = Value::make(Value::Bool);
            v->b = true;
            return v;
        }
        if (s_.compare(p_, 5, "false") == 0) {
            p_ += 5;
            return Value::make(Value::Bool);
        }
        if (s_.compare(p_, 4, "null") == 0) {
            p_ += 4;
            return Value::make(Value::Null);
        }
        return number();
    }
    ValuePtr number() {
        size_t st = p_;
        if (p_ < s_.size() && (s_[p_] == '-' || s_[p_] == '+')) ++p_;
        bool is_float = false;
        while (p_ < s_.size() && (isdigit((unsigned char)s_[p_]) || s_[p_] == '.' || s_[p_] == 'e' || s_[p_] == 'E' ||
                                  s_[p_] == '-' || s_[p_] == '+')) {
            if (s_[p_] == '.' || s_[p_] == 'e' || s_[p_] == 'E') is_float = true;
            ++p_;
        }
        if (p_ == st) err("value expected");
        std::string t = s_.substr(st, p_ - st);
        if (!strict_number(t, false)) err("bad number '" + t + "'");
        auto v = Value::make(is_float ? Value::Float : Value::Int);
        char* e = nullptr;
        if (is_float) {
            v->f = std::strtod(t.c_str(), &e);
        } else {
            v->i = std::strtoll(t.c_str(), &e, 10);
        }
        if (!e || *e) err("bad number '" + t + "'");
        return v;
    }
    std::string string() {
        ++p_;
        std::string out;
        while (p_ < s_.size() && s_[p_] != '"') {
            char c = s_[p_++];
            if ((unsigned char)c < 0x20) err("control character in string");
            if (c == '\\') {
                if (p_ >= s_.size()) err("bad escape");
                char e = s_[p_++];
                switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case '"': out += '"'; break;
                    case '\\': out += '\\'; break;
                    case '/': out += '/'; break;
                    case 'u': {
                        if (p_ + 4 > s_.size()) err("bad \\u escape");
                        unsigned cp = (unsigned)std::strtoul(s_.substr(p_, 4).c_str(), nullptr, 16);
                        p_ += 4;
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: err("bad escape");
                }
            } else {
                out += c;
            }
        }
        if (p_ >= s_.size()) err("unterminated string");
        ++p_;
        return out;
    }
    ValuePtr array() {
        ++p_;
        auto v = Value::make(Value::Array);
        ws();
        if (p_ < s_.size() && s_[p_] == ']') {
            ++p_;
            return v;
        }
        for (;;) {
            ws();
            v->arr.push_back(value());
            ws();
            if (p_ < s_.size() && s_[p_] == ',') {
                ++p_;
                continue;
            }
            if (p_ < s_.size() && s_[p_] == ']') {
                ++p_;
                return v;
            }
            err("',' or ']' expected");
        }
    }
    ValuePtr object() {
        ++p_;
        auto v = Value::make(Value::Table);
        ws();
        if (p_ < s_.size() && s_[p_] == '}') {
            ++p_;
            return v;
        }
        for (;;) {
            ws();
            if (p_ >= s_.size() || s_[p_] != '"') err("key expected");
            std::string k = string();
            ws();
            if (p_ >= s_.size() || s_[p_] != ':') err("':' expected");
            ++p_;
            ws();
            v->tab.emplace_back(k, value());
            ws();
            if (p_ < s_.size() && s_[p_] == ',') {
                ++p_;
                continue;
            }
            if (p_ < s_.size() && s_[p_] == '}') {
                ++p_;
                return v;
            }
            err("',' or '}' expected");
        }
    }
};

// ---------------------------------------------------------------------------------------------- TOML (subset)
class Toml {
  public:
    explicit Toml(const std::string& t) : s_(t) {}
    ValuePtr parse() {
        root_ = Value::make(Value::Table);
        cur_ = root_;
        for (;;) {
            skip_ws_nl();
            if (p_ >= s_.size()) break;
            if (s_[p_] == '[') {
                header();
                skip_ws();
                if (p_ < s_.size() && s_[p_] == '#') skip_comment();
                if (p_ < s_.size() && s_[p_] != '\n' && s_[p_] != '\r') err("newline expected after table header");
            } else {
                keyval(cur_);
                skip_ws();
                if (p_ < s_.size() && s_[p_] == '#') skip_comment();
                if (p_ < s_.size() && s_[p_] != '\n' && s_[p_] != '\r') err("newline expected after key/value");
            }
        }
        return root_;
    }

  private:
    const std::string& s_;
    size_t p_ = 0;
    ValuePtr root_, cur_;
    [[noreturn]] void err(const std::string& m) {
        size_t line = 1;
        for (size_t i = 0; i < p_ && i < s_.size(); ++i) line += s_[i] == '\n';
        fail("TOML: " + m + " (line " + std::to_string(line) + ")");
    }
    void skip_ws() {
        while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t')) ++p_;
    }
    void skip_comment() {
        while (p_ < s_.size() && s_[p_] != '\n') ++p_;
    }
    void skip_ws_nl() {
        for (;;) {
            while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t' || s_[p_] == '\n' || s_[p_] == '\r')) ++p_;
            if (p_ < s_.size() && s_[p_] == '#') {
                skip_comment();
                continue;
            }
            break;
        }
    }
    std::string key_part() {
        skip_ws();
        if (p_ >= s_.size()) err("key expected");
        if (s_[p_] == '"') return basic_string();
        if (s_[p_] == '\'') return literal_string();
        size_t st = p_;
        while (p_ < s_.size() && (isalnum((unsigned char)s_[p_]) || s_[p_] == '_' || s_[p_] == '-')) ++p_;
        if (p_ == st) err("bad key");
        return s_.substr(st, p_ - st);
    }
    std::vector<std::string> dotted_key() {
        std::vector<std::string> parts;
        for (;;) {
            parts.push_back(key_part());
            skip_ws();
            if (p_ < s_.size() && s_[p_] == '.') {
                ++p_;
                continue;
            }
            return parts;
        }
    }
    // walk / create intermediate tables; an array-of-tables segment resolves to its last element
    ValuePtr descend(ValuePtr t, const std::string& k) {
        ValuePtr c = t->get(k);
        if (!c) {
            c = Value::make(Value::Table);
            t->tab.emplace_back(k, c);
            return c;
        }
        if (c->kind == Value::Array) {
            if (c->arr.empty() || c->arr.back()->kind != Value::Table || c->defined_inline) err("cannot extend '" + k + "'");
            return c->arr.back();
        }
        if (c->kind != Value::Table || c->defined_inline) err("key '" + k + "' is not a table");
        return c;
    }
    void header() {
        bool is_array = p_ + 1 < s_.size() && s_[p_ + 1] == '[';
        p_ += is_array ? 2 : 1;
        std::vector<std::string> parts = dotted_key();
        skip_ws();
        if (is_array) {
            if (s_.compare(p_, 2, "]]") != 0) err("']]' expected");
            p_ += 2;
        } else {
            if (p_ >= s_.size() || s_[p_] != ']') err("']' expected");
            ++p_;
        }
        ValuePtr t = root_;
        for (size_t i = 0; i + 1 < parts.size(); ++i) t = descend(t, parts[i]);
        const std::string& last = parts.back();
        if (is_array) {
            ValuePtr a = t->get(last);
            if (!a) {
                a = Value::make(Value::Array);
                t->tab.emplace_back(last, a);
            }
            if (a->kind != Value::Array || a->defined_inline) err("'" + last + "' is not an array of tables");
            a->arr.push_back(Value::make(Value::Table));
            cur_ = a->arr.back();
        } else {
            cur_ = descend(t, last);
        }
        skip_ws();
        if (p_ < s_.size() && s_[p_] == '#') skip_comment();
    }
    void keyval(ValuePtr table) {
        std::vector<std::string> parts = dotted_key();
        skip_ws();
        if (p_ >= s_.size() || s_[p_] != '=') err("'=' expected");
        ++p_;
        skip_ws();
        ValuePtr t = table;
        for (size_t i = 0; i + 1 < parts.size(); ++i) t = descend(t, parts[i]);
        if (t->get(parts.back())) err("duplicate key '" + parts.back() + "'");
        ValuePtr v = value();
        v->defined_inline = true;
        t->tab.emplace_back(parts.back(), v);
    }
    std::string basic_string() {
        ++p_;
        std::string out;
        while (p_ < s_.size() && s_[p_] != '"') {
            char c = s_[p_++];
            if ((unsigned char)c < 0x20 && c != '\t') err("control character in string");
            if (c == '\\') {
                if (p_ >= s_.size()) err("bad escape");
                char e = s_[p_++];
                switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case '"': out += '"'; break;
                    case '\\': out += '\\'; break;
                    default: err("unsupported escape");
                }
            } else {
                out += c;
            }
        }
        if (p_ >= s_.size()) err("unterminated string");
        ++p_;
        return out;
    }
    std::string literal_string() {
        ++p_;
        size_t st = p_;
        while (p_ < s_.size() && s_[p_] != '\'' && s_[p_] != '\n') ++p_;
        if (p_ >= s_.size() || s_[p_] != '\'') err("unterminated literal string");
        std::string out = s_.substr(st, p_ - st);
        ++p_;
        return out;
    }
    ValuePtr value() {
        if (p_ >= s_.size()) err("value expected");
        char c = s_[p_];
        if (c == '"') {
            if (s_.compare(p_, 3, "\"\"\"") == 0) err("multi-line strings are not supported");
            auto v = Value::make(Value::String);
            v->s = basic_string();
            return v;
        }
        if (c == '\'') {
            auto v = Value::make(Value::String);
            v->s = literal_string();
            return v;
        }
        if (c == '[') {
            ++p_;
            auto v = Value::make(Value::Array);
            for (;;) {
                skip_ws_nl();
                if (p_ < s_.size() && s_[p_] == ']') {
                    ++p_;
                    return v;
                }
                v->arr.push_back(value());
                skip_ws_nl();
                if (p_ < s_.size() && s_[p_] == ',') {
                    ++p_;
                    continue;
                }
                if (p_ < s_.size() && s_[p_] == ']') {
                    ++p_;
                    return v;
                }
                err("',' or ']' expected in array");
            }
        }
        if (c == '{') {
            ++p_;
            auto v = Value::make(Value::Table);
            skip_ws();
            if (p_ < s_.size() && s_[p_] == '}') {
                ++p_;
                return v;
            }
            for (;;) {
                skip_ws();  // TOML 1.0 (what the toml crate implements): no newline inside an inline table
                keyval(v);
                skip_ws();
                if (p_ < s_.size() && s_[p_] == ',') {
                    ++p_;
                    continue;
                }
                if (p_ < s_.size() && s_[p_] == '}') {
                    ++p_;
                    return v;
                }
                err("',' or '}' expected in inline table");
            }
        }
        if (s_.compare(p_, 4, "true") == 0) {
            p_ += 4;
            auto v = Value::make(Value::Bool);
            v->b = true;
            return v;
        }
        if (s_.compare(p_, 5, "false") == 0) {
            p_ += 5;
            return Value::make(Value::Bool);
        }
        // number
        size_t st = p_;
        while (p_ < s_.size() && (isalnum((unsigned char)s_[p_]) || s_[p_] == '.' || s_[p_] == '-' || s_[p_] == '+' || s_[p_] == '_'))
            ++p_;
        std::string t;
        for (size_t i = st; i < p_; ++i)
            if (s_[i] != '_') t += s_[i];
        if (t.empty()) err("value expected");
        std::string bare = (t[0] == '+' || t[0] == '-') ? t.substr(1) : t;
        if (bare == "inf" || bare == "nan") {
            auto v = Value::make(Value::Float);
            v->f = bare == "inf" ? (t[0] == '-' ? -INFINITY : INFINITY) : NAN;
            return v;
        }
        if (!strict_number(t, true)) err("bad value '" + t + "'");
        bool is_float = t.find_first_of(".eE") != std::string::npos;
        auto v = Value::make(is_float ? Value::Float : Value::Int);
        char* e = nullptr;
        if (is_float) v->f = std::strtod(t.c_str(), &e);
        else v->i = std::strtoll(t.c_str(), &e, 10);
        if (!e || *e) err("bad value '" + t + "'");
        return v;
    }
};

// ---------------------------------------------------------------------------------------------- files
static std::string read_file(const std::string& path, bool binary = false) {
    std::ifstream f(path, binary ? std::ios::binary : std::ios::in);
    if (!f) fail("cannot read " + path);
    return std::string((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static bool ends_with(const std::string& s, const char* suf) {
    size_t n = std::strlen(suf);
    if (s.size() < n) return false;
    for (size_t i = 0; i < n; ++i)
        if (tolower((unsigned char)s[s.size() - n + i]) != suf[i]) return false;
    return true;
}
static ValuePtr load_scene_file(const std::string& path) {  // try_load_scene, scene_config.rs:475-492
    if (ends_with(path, ".json")) return Json(read_file(path)).parse();
    if (ends_with(path, ".toml")) return Toml(read_file(path)).parse();
    fail("invalid scene file format!");
}

// ---------------------------------------------------------------------------------------------- scene building
struct Graph {
    std::vector<nrrt_object> objects;
    std::vector<uint32_t> child_ids;
    std::vector<nrrt_material> materials;
    std::vector<nrrt_texture> textures;
    std::vector<nrrt_jpeg::Image> image_store;
    std::map<std::string, uint32_t> image_by_path;
    std::vector<nrrt_image> images;
    uint32_t root = 0;
};

// ids may be strings (v2/v3) or positions (v1): keep both as strings with a type tag
static std::string id_of(const Value& v) {
    if (v.kind == Value::String) return "s:" + v.s;
    if (v.kind == Value::Int) return "i:" + std::to_string(v.i);
    fail("id must be a string or an index");
}
using IdMap = std::map<std::string, uint32_t>;

struct Builder {
    Graph& g;
    std::string base_dir;
    // scene files being loaded right now, outermost first: a `Scene { path }` object that names one of them (itself,
    // or a file that includes it) would recurse until the native stack overflows — the reference does the same
    // (scene_config.rs:335-340) and dies of it; here it is an error the caller gets back
    std::vector<std::string> include_stack;

    std::string resolve(const std::string& p) const {
        if (!p.empty() && p[0] == '/') return p;
        return base_dir.empty() ? p : base_dir + "/" + p;
    }
    static void variant(const Value& cfg, const char* what, std::string& kind, ValuePtr& body) {
        if (cfg.kind != Value::Table || cfg.tab.size() != 1) fail(std::string(what) + ": expected a single-variant table");
        kind = cfg.tab[0].first;
        body = cfg.tab[0].second;
        if (!body || body->kind == Value::Null) body = Value::make(Value::Table);
        if (body->kind != Value::Table) fail(std::string(what) + ": variant body must be a table");
    }
    static void vec3(const ValuePtr& v, const char* what, double out[3]) {
        if (!v || v->kind != Value::Array || v->arr.size() != 3) fail(std::string(what) + ": expected a 3-vector");
        for (int i = 0; i < 3; ++i) {
            if (!v->arr[i]->is_num()) fail(std::string(what) + ": expected numbers");
            out[i] = v->arr[i]->num();
        }
    }
    static double num(const ValuePtr& v, const char* what) {
        if (!v || !v->is_num()) fail(std::string(what) + ": expected a number");
        return v->num();
    }
    static bool present(const ValuePtr& v) { return v && v->kind != Value::Null; }

    uint32_t add_texture(const nrrt_texture& t) {
        g.textures.push_back(t);
        return (uint32_t)g.textures.size() - 1;
    }
    uint32_t add_solid(double r, double gg, double b) {
        nrrt_texture t;
        std::memset(&t, 0, sizeof t);
        t.kind = NRRT_TEX_SOLID;
        t.color[0] = r, t.color[1] = gg, t.color[2] = b;
        return add_texture(t);
    }
    uint32_t add_material(uint32_t kind, uint32_t tex, double param) {
        g.materials.push_back(nrrt_material{kind, tex, param});
        return (uint32_t)g.materials.size() - 1;
    }
    uint32_t add_object(uint32_t kind, uint32_t material, const std::vector<uint32_t>& children, const double* v, int nv) {
        nrrt_object o;
        std::memset(&o, 0, sizeof o);
        o.kind = kind;
        o.material = material;
        o.first_child = (uint32_t)g.child_ids.size();
        o.n_children = (uint32_t)children.size();
        for (uint32_t c : children) g.child_ids.push_back(c);
        for (int i = 0; i < nv; ++i) o.v[i] = v[i];
        g.objects.push_back(o);
        return (uint32_t)g.objects.size() - 1;
    }

    // TextureConfig::try_make_texture (scene_config.rs:53-122)
    uint32_t make_texture(const Value& cfg, const IdMap& textures) {
        std::string kind;
        ValuePtr p;
        variant(cfg, "texture", kind, p);
        nrrt_texture t;
        std::memset(&t, 0, sizeof t);
        if (kind == "SolidColor") {
            double c[3];
            vec3(p->get("color"), "SolidColor.color", c);
            return add_solid(c[0], c[1], c[2]);
        }
        if (kind == "Checker") {
            auto ref = [&](const char* name, double dflt) -> uint32_t {
                ValuePtr id = p->get(name);
                if (!present(id)) return add_solid(dflt, dflt, dflt);  // checker.rs:53-54
                auto it = textures.find(id_of(*id));
                if (it == textures.end()) fail("invalid texture index");
                return it->second;
            };
            uint32_t even = ref("even", 1.0), odd = ref("odd", 0.0);
            t.kind = NRRT_TEX_CHECKER;
            t.a = even, t.b = odd;
            t.f0 = present(p->get("scale")) ? num(p->get("scale"), "Checker.scale") : 0.5;
            return add_texture(t);
        }
        if (kind == "Image") {
            ValuePtr pv = p->get("path");
            if (!pv || pv->kind != Value::String) fail("Image.path: expected a string");
            std::string path = resolve(pv->s);
            uint32_t idx;
            auto it = g.image_by_path.find(path);
            if (it != g.image_by_path.end()) {
                idx = it->second;
            } else {
                std::string bytes;
                try {
                    bytes = read_file(path, true);
                } catch (const LoadError&) {
                    fail("cannot load image " + path);
                }
                nrrt_jpeg::Image im;
                std::string e;
                if (!nrrt_jpeg::decode((const uint8_t*)bytes.data(), bytes.size(), im, e))
                    fail("cannot decode image " + path + ": " + e);
                g.image_store.push_back(std::move(im));
                idx = (uint32_t)g.image_store.size() - 1;
                g.image_by_path[path] = idx;
            }
            t.kind = NRRT_TEX_IMAGE;
            t.a = idx;
            return add_texture(t);
        }
        const double PI = 3.14159265358979323846264338327950288;
        if (kind == "Marble") {
            t.kind = NRRT_TEX_MARBLE;
            t.seed = present(p->get("seed")) ? (uint32_t)num(p->get("seed"), "Marble.seed") : 0u;
            t.octaves = 7;
            t.f0 = present(p->get("frequency")) ? num(p->get("frequency"), "Marble.frequency") : 1.0;
            return add_texture(t);
        }
        if (kind == "Noise") {
            t.kind = NRRT_TEX_NOISE;
            t.seed = present(p->get("seed")) ? (uint32_t)num(p->get("seed"), "Noise.seed") : 0u;
            t.octaves = present(p->get("octaves")) ? (uint32_t)num(p->get("octaves"), "Noise.octaves") : 1u;
            t.f0 = present(p->get("frequency")) ? num(p->get("frequency"), "Noise.frequency") : 1.0;
            t.f1 = present(p->get("lacunarity")) ? num(p->get("lacunarity"), "Noise.lacunarity") : PI * 2.0 / 3.0;
            t.f2 = present(p->get("persistence")) ? num(p->get("persistence"), "Noise.persistence") : 0.5;
            return add_texture(t);
        }
        fail("unknown texture kind '" + kind + "'");
    }

    // MaterialConfig::try_make_material (scene_config.rs:163-197)
    uint32_t make_material(const Value& cfg, const IdMap& textures, uint32_t texture_fallback) {
        std::string kind;
        ValuePtr p;
        variant(cfg, "material", kind, p);
        auto tex = [&]() -> uint32_t {
            ValuePtr id = p->get("texture");
            if (!present(id)) return texture_fallback;
            auto it = textures.find(id_of(*id));
            if (it == textures.end()) fail("invalid texture id");
            return it->second;
        };
        if (kind == "Dielectric") return add_material(NRRT_MAT_DIELECTRIC, 0, num(p->get("refraction_index"), "refraction_index"));
        if (kind == "DiffuseLight") return add_material(NRRT_MAT_DIFFUSE_LIGHT, tex(), num(p->get("intensity"), "intensity"));
        if (kind == "Lambertian") return add_material(NRRT_MAT_LAMBERTIAN, tex(), 0.0);
        if (kind == "Metal") return add_material(NRRT_MAT_METAL, tex(), num(p->get("fuzz"), "fuzz"));
        fail("unknown material kind '" + kind + "'");
    }

    // ObjectConfig::try_make_object (scene_config.rs:278-380)
    uint32_t make_object(const Value& cfg, const IdMap& instances, const IdMap& materials, uint32_t material_fallback,
                         int depth) {
        if (depth > 64) fail("object nesting too deep");
        std::string kind;
        ValuePtr p;
        variant(cfg, "object", kind, p);
        auto mat = [&]() -> uint32_t {
            ValuePtr id = p->get("material");
            if (!present(id)) return material_fallback;
            auto it = materials.find(id_of(*id));
            if (it == materials.end()) fail("invalid material id");
            return it->second;
        };
        auto inner = [&]() -> uint32_t {
            ValuePtr o = p->get("object");
            if (!o) fail(kind + ": missing object");
            return make_object(*o, instances, materials, material_fallback, depth + 1);
        };
        double v[9] = {0};
        if (kind == "Quad" || kind == "Triangle") {
            vec3(p->get("point"), "point", v);
            vec3(p->get("u"), "u", v + 3);
            vec3(p->get("v"), "v", v + 6);
            return add_object(kind == "Quad" ? NRRT_OBJ_QUAD : NRRT_OBJ_TRIANGLE, mat(), {}, v, 9);
        }
        if (kind == "Sphere") {
            vec3(p->get("center"), "center", v);
            v[3] = num(p->get("radius"), "radius");
            return add_object(NRRT_OBJ_SPHERE, mat(), {}, v, 4);
        }
        if (kind == "Group") {
            uint32_t m = mat();
            std::vector<uint32_t> kids;
            ValuePtr objs = p->get("objects");
            if (objs && objs->kind == Value::Array)
                for (auto& o : objs->arr) kids.push_back(make_object(*o, instances, materials, m, depth + 1));
            return add_object(NRRT_OBJ_GROUP, 0, kids, v, 0);
        }
        if (kind == "Scene") {
            uint32_t m = mat();
            ValuePtr pv = p->get("path");
            if (!pv || pv->kind != Value::String) fail("Scene.path: expected a string");
            const std::string file = resolve(pv->s);
            for (const std::string& open_file : include_stack)
                if (open_file == file) fail("recursive scene include: " + file);
            if (include_stack.size() >= 16) fail("scene includes nested deeper than 16 files: " + file);
            ValuePtr sub = load_scene_file(file);
            include_stack.push_back(file);
            const uint32_t obj = build_aux(*sub, true, m);
            include_stack.pop_back();
            return obj;
        }
        if (kind == "Ref") {
            ValuePtr id = p->get("id");
            if (!present(id)) fail("Ref: missing id");
            auto it = instances.find(id_of(*id));
            if (it == instances.end()) fail("invalid object id");
            return it->second;
        }
        if (kind == "RotateX" || kind == "RotateY" || kind == "RotateZ") {
            uint32_t c = inner();
            v[0] = num(p->get("angle"), "angle");
            uint32_t k = kind == "RotateX" ? NRRT_OBJ_ROTATE_X : (kind == "RotateY" ? NRRT_OBJ_ROTATE_Y : NRRT_OBJ_ROTATE_Z);
            return add_object(k, 0, {c}, v, 1);
        }
        if (kind == "ScaleU") {
            uint32_t c = inner();
            double f = num(p->get("factor"), "factor");
            v[0] = f * 1.0, v[1] = f * 1.0, v[2] = f * 1.0;  // factor*DVec3::ONE (scale.rs:60-65)
            return add_object(NRRT_OBJ_SCALE, 0, {c}, v, 3);
        }
        if (kind == "ScaleV") {
            uint32_t c = inner();
            vec3(p->get("scale"), "scale", v);
            return add_object(NRRT_OBJ_SCALE, 0, {c}, v, 3);
        }
        if (kind == "Translate") {
            uint32_t c = inner();
            vec3(p->get("offset"), "offset", v);
            return add_object(NRRT_OBJ_TRANSLATE, 0, {c}, v, 3);
        }
        fail("unknown object kind '" + kind + "'");
    }

    // [(id, cfg)] from a v3 pair array, a v2 table or a v1 anonymous array
    static std::vector<std::pair<std::string, ValuePtr>> pairs(const ValuePtr& section, const char* what) {
        std::vector<std::pair<std::string, ValuePtr>> out;
        if (!section || section->kind == Value::Null) return out;
        if (section->kind == Value::Table) {
            for (auto& kv : section->tab) out.emplace_back("s:" + kv.first, kv.second);
            return out;
        }
        if (section->kind == Value::Array) {
            long long i = 0;
            for (auto& e : section->arr) {
                if (e->kind == Value::Array && e->arr.size() == 2 && e->arr[0]->kind == Value::String)
                    out.emplace_back("s:" + e->arr[0]->s, e->arr[1]);
                else if (e->kind == Value::Table)
                    out.emplace_back("i:" + std::to_string(i), e);
                else
                    fail(std::string(what) + ": malformed entry");
                ++i;
            }
            return out;
        }
        fail(std::string(what) + ": expected an array or a table");
    }

    // SceneConfig::try_build_aux (scene_config.rs:411-473); returns the GROUP object of scene.objects
    uint32_t build_aux(const Value& cfg, bool have_fallback, uint32_t material_fallback) {
        if (cfg.kind != Value::Table) fail("scene file: top level must be a table");
        IdMap textures;
        auto tex_pairs = pairs(cfg.get("textures"), "textures");
        if (cfg.get("textures") && cfg.get("textures")->kind == Value::Table) {
            // a table has no order: non-Checker textures first, then Checkers (they look up earlier ids)
            std::vector<std::pair<std::string, ValuePtr>> a, b;
            for (auto& kv : tex_pairs) {
                bool chk = kv.second->kind == Value::Table && kv.second->get("Checker") != nullptr;
                (chk ? b : a).push_back(kv);
            }
            a.insert(a.end(), b.begin(), b.end());
            tex_pairs.swap(a);
        }
        for (auto& kv : tex_pairs) textures[kv.first] = make_texture(*kv.second, textures);
        uint32_t texture_fallback;
        if (present(cfg.get("texture_fallback"))) texture_fallback = make_texture(*cfg.get("texture_fallback"), textures);
        else texture_fallback = add_solid(0.5 * 1.0, 0.5 * 1.0, 0.5 * 1.0);
        IdMap materials;
        for (auto& kv : pairs(cfg.get("materials"), "materials"))
            materials[kv.first] = make_material(*kv.second, textures, texture_fallback);
        if (!have_fallback) {
            if (present(cfg.get("material_fallback")))
                material_fallback = make_material(*cfg.get("material_fallback"), textures, texture_fallback);
            else
                material_fallback = add_material(NRRT_MAT_LAMBERTIAN, texture_fallback, 0.0);
        }
        IdMap instances;
        for (auto& kv : pairs(cfg.get("instances"), "instances"))
            instances[kv.first] = make_object(*kv.second, instances, materials, material_fallback, 0);
        ValuePtr list = cfg.get("scene");
        if (!list) list = cfg.get("objects");  // v1 schema
        std::vector<uint32_t> objs;
        if (list && list->kind == Value::Array)
            for (auto& o : list->arr) objs.push_back(make_object(*o, instances, materials, material_fallback, 0));
        double none[1] = {0};
        return add_object(NRRT_OBJ_GROUP, 0, objs, none, 0);
    }
};

}  // namespace nrrt_loader

// ================================================================================================ C ABI
using namespace nrrt_loader;

struct nrrt_loaded_scene {
    Graph g;
    nrrt_graph_desc desc;
    nrrt_camera_file cam;
};

static thread_local std::string g_load_error;

static void read_camera(const Value& root, nrrt_camera_file& c) {
    std::memset(&c, 0, sizeof c);
    ValuePtr cam = root.get("camera");
    if (!cam || cam->kind != Value::Table) return;
    auto num = [&](const char* k, uint32_t bit, double& dst) {
        ValuePtr v = cam->get(k);
        if (v && v->is_num()) {
            dst = v->num();
            c.present |= bit;
        }
    };
    auto vec = [&](const char* k, uint32_t bit, double* dst) {
        ValuePtr v = cam->get(k);
        if (v && v->kind == Value::Array && v->arr.size() == 3 && v->arr[0]->is_num() && v->arr[1]->is_num() && v->arr[2]->is_num()) {
            for (int i = 0; i < 3; ++i) dst[i] = v->arr[i]->num();
            c.present |= bit;
        }
    };
    double w = 0, h = 0, spp = 0, nb = 0;
    num("width", NRRT_CAM_WIDTH, w);
    num("height", NRRT_CAM_HEIGHT, h);
    num("aspect_ratio", NRRT_CAM_ASPECT_RATIO, c.aspect_ratio);
    vec("background_color", NRRT_CAM_BACKGROUND, c.background);
    vec("look_at", NRRT_CAM_LOOK_AT, c.look_at);
    vec("look_from", NRRT_CAM_LOOK_FROM, c.look_from);
    vec("view_up", NRRT_CAM_VIEW_UP, c.view_up);
    num("field_of_view", NRRT_CAM_FOV, c.field_of_view_deg);
    num("defocus_angle", NRRT_CAM_DEFOCUS, c.defocus_angle_deg);
    num("focus_distance", NRRT_CAM_FOCUS, c.focus_distance);
    num("samples_per_pixel", NRRT_CAM_SPP, spp);
    num("ray_max_bounces", NRRT_CAM_BOUNCES, nb);
    c.width = (uint32_t)w, c.height = (uint32_t)h, c.samples_per_pixel = (uint32_t)spp, c.ray_max_bounces = (uint32_t)nb;
}

extern "C" {

const char* nrrt_load_last_error(void) { return g_load_error.c_str(); }

nrrt_loaded_scene* nrrt_load_scene(const char* path, const char* base_dir) {
    g_load_error.clear();
    if (!path) {
        g_load_error = "nrrt_load_scene: null path";
        return nullptr;
    }
    try {
        auto ls = std::make_unique<nrrt_loaded_scene>();
        ValuePtr root = load_scene_file(path);
        Builder b{ls->g, base_dir ? std::string(base_dir) : std::string(), {}};
        b.include_stack.push_back(path);
        ls->g.root = b.build_aux(*root, false, 0);
        read_camera(*root, ls->cam);
        Graph& g = ls->g;
        for (auto& im : g.image_store) g.images.push_back(nrrt_image{im.width, im.height, im.rgb.data()});
        nrrt_graph_desc& d = ls->desc;
        std::memset(&d, 0, sizeof d);
        d.n_objects = (uint32_t)g.objects.size(), d.objects = g.objects.data();
        d.n_child_ids = (uint32_t)g.child_ids.size(), d.child_ids = g.child_ids.data();
        d.n_materials = (uint32_t)g.materials.size(), d.materials = g.materials.data();
        d.n_textures = (uint32_t)g.textures.size(), d.textures = g.textures.data();
        d.n_images = (uint32_t)g.images.size(), d.images = g.images.data();
        d.root = g.root;
        return ls.release();
    } catch (const LoadError& e) {
        g_load_error = e.msg;
    } catch (const std::exception& e) {
        g_load_error = e.what();
    }
    return nullptr;
}

const nrrt_graph_desc* nrrt_loaded_graph(const nrrt_loaded_scene* s) { return s ? &s->desc : nullptr; }

int nrrt_loaded_camera(const nrrt_loaded_scene* s, nrrt_camera_file* out) {
    if (!s || !out) return NRRT_ERR_INVALID;
    *out = s->cam;
    return NRRT_OK;
}

void nrrt_loaded_free(nrrt_loaded_scene* s) { delete s; }

// CameraConfig::merge_with (cli.rs:316-355): every field present in `over` replaces the one in `base`
void nrrt_camera_file_merge(nrrt_camera_file* base, const nrrt_camera_file* over) {
    if (!base || !over) return;
    const uint32_t p = over->present;
    if (p & NRRT_CAM_WIDTH) base->width = over->width;
    if (p & NRRT_CAM_HEIGHT) base->height = over->height;
    if (p & NRRT_CAM_ASPECT_RATIO) base->aspect_ratio = over->aspect_ratio;
    if (p & NRRT_CAM_BACKGROUND) std::memcpy(base->background, over->background, sizeof base->background);
    if (p & NRRT_CAM_LOOK_AT) std::memcpy(base->look_at, over->look_at, sizeof base->look_at);
    if (p & NRRT_CAM_LOOK_FROM) std::memcpy(base->look_from, over->look_from, sizeof base->look_from);
    if (p & NRRT_CAM_VIEW_UP) std::memcpy(base->view_up, over->view_up, sizeof base->view_up);
    if (p & NRRT_CAM_FOV) base->field_of_view_deg = over->field_of_view_deg;
    if (p & NRRT_CAM_DEFOCUS) base->defocus_angle_deg = over->defocus_angle_deg;
    if (p & NRRT_CAM_FOCUS) base->focus_distance = over->focus_distance;
    if (p & NRRT_CAM_SPP) base->samples_per_pixel = over->samples_per_pixel;
    if (p & NRRT_CAM_BOUNCES) base->ray_max_bounces = over->ray_max_bounces;
    base->present |= p;
}

// CameraConfig::try_update (cli.rs:357-402) applied to CameraBuilder::default() (camera.rs:162-203)
int nrrt_camera_file_to_config(const nrrt_camera_file* c, nrrt_camera_config* out) {
    if (!c || !out) return NRRT_ERR_INVALID;
    const double PI = 3.14159265358979323846264338327950288;
    std::memset(out, 0, sizeof *out);
    const bool w = c->present & NRRT_CAM_WIDTH, h = c->present & NRRT_CAM_HEIGHT, r = c->present & NRRT_CAM_ASPECT_RATIO;
    out->width = 1200, out->height = 800;  // camera.rs:163-164
    if (w && h && !r) {
        out->width = c->width, out->height = c->height;
    } else if (w && !h && r) {  // image.rs:26-32
        out->width = c->width;
        long long hh = (long long)((double)c->width / c->aspect_ratio);
        out->height = (uint32_t)(hh < 1 ? 1 : hh);
    } else if (!w && h && r) {  // image.rs:34-40
        out->height = c->height;
        long long ww = (long long)((double)c->height * c->aspect_ratio);
        out->width = (uint32_t)(ww < 1 ? 1 : ww);
    } else if (w || h || r) {
        return NRRT_ERR_INVALID;  // cli.rs:287-310: needs exactly two of width / height / aspect ratio
    }
    const double bg0[3] = {0, 0, 0}, from0[3] = {1, 1, 1}, at0[3] = {0, 0, 0}, up0[3] = {0, 1, 0};
    for (int i = 0; i < 3; ++i) {
        out->background[i] = (c->present & NRRT_CAM_BACKGROUND) ? c->background[i] : bg0[i];
        out->look_from[i] = (c->present & NRRT_CAM_LOOK_FROM) ? c->look_from[i] : from0[i];
        out->look_at[i] = (c->present & NRRT_CAM_LOOK_AT) ? c->look_at[i] : at0[i];
        out->view_up[i] = (c->present & NRRT_CAM_VIEW_UP) ? c->view_up[i] : up0[i];
    }
    out->field_of_view = (c->present & NRRT_CAM_FOV) ? (c->field_of_view_deg * PI) / 180.0 : PI / 2.;
    out->focus_dist = (c->present & NRRT_CAM_FOCUS) ? c->focus_distance : 1.0;
    out->defocus_angle = (c->present & NRRT_CAM_DEFOCUS) ? (c->defocus_angle_deg * PI) / 180.0 : 0.0;
    out->samples_per_pixel = (c->present & NRRT_CAM_SPP) ? c->samples_per_pixel : 10;
    out->ray_max_bounces = (c->present & NRRT_CAM_BOUNCES) ? c->ray_max_bounces : 10;
    return NRRT_OK;
}

}  // extern "C"
