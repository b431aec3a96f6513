// scene_loader.cpp — see scene_loader.h. glTF 2.0 (.gltf + external / embedded buffers, .glb) -> the flat scene arrays.
// Behaviour follows the reference's src/scene/scene_loader.cpp (line numbers cited per step); the parsing machinery
// (JSON, accessors, PNG) is this file's own.
#include "scene_loader.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <map>
#include <memory>

// Image decoding = the reference's own decoder, stb_image.h (scene_loader.cpp:284,290: stbi_load / stbi_load_from_memory with
// STBI_rgb_alpha), taken from the reference's vendored copy THROUGH THE INCLUDE PATH (-I /root/reference/dependencies; nothing is copied
// into this repository): PNG, JPEG (Sponza / Bistro ship both), BMP, TGA, ... Where that tree is not on the include path the library is
// built with this file's own PNG decoder only and says so (HasStbImage()).
#if defined(__has_include)
#if __has_include("stb/stb_image.h")
#define VHR_HAVE_STB_IMAGE 1
#define STB_IMAGE_IMPLEMENTATION
#define STB_IMAGE_STATIC
#define STBI_NO_STDIO
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wunused-function"
#pragma GCC diagnostic ignored "-Wunused-but-set-variable"
#pragma GCC diagnostic ignored "-Wmisleading-indentation"
#pragma GCC diagnostic ignored "-Wsign-compare"
#include "stb/stb_image.h"
#pragma GCC diagnostic pop
#endif
#endif

// Document parsing = the reference's own parser, cgltf.h (scene_loader.cpp:233-235, 338-341: cgltf_parse_file / cgltf_load_buffers, the
// accessor readers and cgltf_node_transform_world), taken the same way: through the include path, nothing copied. Without it (or with
// VHR_GLTF_PARSER=own in the environment) this file's own JSON / .glb / accessor reader does the same job; tests/test_scene_loader_cpu.py
// runs every case through both and compares the results field by field.
#if defined(__has_include)
#if __has_include("cgltf/cgltf.h")
#define VHR_HAVE_CGLTF 1
#define CGLTF_IMPLEMENTATION
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wunused-function"
#pragma GCC diagnostic ignored "-Wsign-compare"
#pragma GCC diagnostic ignored "-Wmissing-field-initializers"
#include "cgltf/cgltf.h"
#pragma GCC diagnostic pop
#endif
#endif

namespace SceneLoader {
bool HasCgltf() {
#ifdef VHR_HAVE_CGLTF
    return true;
#else
    return false;
#endif
}
bool HasStbImage() {
#ifdef VHR_HAVE_STB_IMAGE
    return true;
#else
    return false;
#endif
}
namespace {

[[noreturn]] void bad(const std::string &msg) { throw VhrHostError{VHR_ERR_INVALID, "scene loader: " + msg}; }

// ---------------------------------------------------------------------------------------------------------------------
// JSON (RFC 8259) — a small recursive-descent reader; numbers are kept as double, object keys keep file order
// ---------------------------------------------------------------------------------------------------------------------
struct Json {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    double number = 0.0;
    bool boolean = false;
    std::string string;
    std::vector<Json> array;
    std::vector<std::pair<std::string, Json>> object;

    const Json *find(const char *key) const {
        if (type != Object) return nullptr;
        for (auto &kv : object)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool has(const char *key) const { return find(key) != nullptr; }
    const Json &at(const char *key) const {
        const Json *j = find(key);
        if (!j) bad(std::string("missing key '") + key + "'");
        return *j;
    }
    double num(const char *key, double fallback) const {
        const Json *j = find(key);
        return (j && j->type == Number) ? j->number : fallback;
    }
    int integer(const char *key, int fallback) const { return (int)num(key, fallback); }
    std::string str(const char *key, const std::string &fallback = "") const {
        const Json *j = find(key);
        return (j && j->type == String) ? j->string : fallback;
    }
    size_t size() const { return type == Array ? array.size() : 0; }
};

struct JsonReader {
    const char *p, *end;
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    bool lit(const char *s) {
        size_t n = strlen(s);
        if ((size_t)(end - p) >= n && !memcmp(p, s, n)) { p += n; return true; }
        return false;
    }
    static void utf8(std::string &out, uint32_t c) {
        if (c < 0x80) out += (char)c;
        else if (c < 0x800) { out += (char)(0xC0 | (c >> 6)); out += (char)(0x80 | (c & 0x3F)); }
        else if (c < 0x10000) { out += (char)(0xE0 | (c >> 12)); out += (char)(0x80 | ((c >> 6) & 0x3F)); out += (char)(0x80 | (c & 0x3F)); }
        else { out += (char)(0xF0 | (c >> 18)); out += (char)(0x80 | ((c >> 12) & 0x3F)); out += (char)(0x80 | ((c >> 6) & 0x3F)); out += (char)(0x80 | (c & 0x3F)); }
    }
    uint32_t hex4() {
        if (end - p < 4) bad("JSON: truncated \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; ++i) {
            char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= c - '0';
            else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
            else bad("JSON: bad \\u escape");
        }
        return v;
    }
    std::string parse_string() {
        std::string s;
        ++p;   // opening quote
        while (true) {
            if (p >= end) bad("JSON: unterminated string");
            char c = *p++;
            if (c == '"') break;
            if (c != '\\') { s += c; continue; }
            if (p >= end) bad("JSON: unterminated escape");
            char e = *p++;
            switch (e) {
                case '"': s += '"'; break;
                case '\\': s += '\\'; break;
                case '/': s += '/'; break;
                case 'b': s += '\b'; break;
                case 'f': s += '\f'; break;
                case 'n': s += '\n'; break;
                case 'r': s += '\r'; break;
                case 't': s += '\t'; break;
                case 'u': {
                    uint32_t c1 = hex4();
                    if (c1 >= 0xD800 && c1 < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                        p += 2;
                        uint32_t c2 = hex4();
                        c1 = 0x10000 + ((c1 - 0xD800) << 10) + (c2 - 0xDC00);
                    }
                    utf8(s, c1);
                    break;
                }
                default: bad("JSON: unknown escape");
            }
        }
        return s;
    }
    Json parse_value(int depth) {
        if (depth > 256) bad("JSON: nesting too deep");
        ws();
        if (p >= end) bad("JSON: unexpected end");
        Json j;
        char c = *p;
        if (c == '{') {
            j.type = Json::Object;
            ++p; ws();
            if (p < end && *p == '}') { ++p; return j; }
            while (true) {
                ws();
                if (p >= end || *p != '"') bad("JSON: expected a key");
                std::string key = parse_string();
                ws();
                if (p >= end || *p != ':') bad("JSON: expected ':'");
                ++p;
                j.object.emplace_back(std::move(key), parse_value(depth + 1));
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; break; }
                bad("JSON: expected ',' or '}'");
            }
        } else if (c == '[') {
            j.type = Json::Array;
            ++p; ws();
            if (p < end && *p == ']') { ++p; return j; }
            while (true) {
                j.array.push_back(parse_value(depth + 1));
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; break; }
                bad("JSON: expected ',' or ']'");
            }
        } else if (c == '"') {
            j.type = Json::String;
            j.string = parse_string();
        } else if (lit("true")) { j.type = Json::Bool; j.boolean = true; }
        else if (lit("false")) { j.type = Json::Bool; j.boolean = false; }
        else if (lit("null")) { j.type = Json::Null; }
        else {
            char *stop = nullptr;
            std::string tmp(p, (size_t)std::min<ptrdiff_t>(end - p, 64));
            double v = strtod(tmp.c_str(), &stop);
            if (stop == tmp.c_str()) bad("JSON: unexpected character");
            p += stop - tmp.c_str();
            j.type = Json::Number;
            j.number = v;
        }
        return j;
    }
};

Json parse_json(const char *text, size_t size) {
    JsonReader r{text, text + size};
    Json j = r.parse_value(0);
    r.ws();
    if (r.p != r.end) bad("JSON: trailing characters");
    return j;
}

// ---------------------------------------------------------------------------------------------------------------------
// files, data URIs
// ---------------------------------------------------------------------------------------------------------------------
std::vector<uint8_t> read_file(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) bad("cannot open '" + path + "'");
    std::vector<uint8_t> data;
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + n);
    fclose(f);
    return data;
}
std::string parent_dir(const std::string &path) {
    size_t k = path.find_last_of("/\\");
    return k == std::string::npos ? std::string() : path.substr(0, k + 1);
}
std::string file_name(const std::string &path) {
    size_t k = path.find_last_of("/\\");
    return k == std::string::npos ? path : path.substr(k + 1);
}
std::vector<uint8_t> base64(const char *s, size_t n) {
    std::vector<uint8_t> out;
    uint32_t acc = 0;
    int bits = 0;
    for (size_t i = 0; i < n; ++i) {
        char c = s[i];
        int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A';
        else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
        else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+' || c == '-') v = 62;
        else if (c == '/' || c == '_') v = 63;
        else if (c == '=') break;
        else continue;
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back((uint8_t)(acc >> bits)); }
    }
    return out;
}
std::string percent_decode(const std::string &s) {
    std::string o;
    for (size_t i = 0; i < s.size(); ++i) {
        if (s[i] == '%' && i + 2 < s.size() + 0 && isxdigit((unsigned char)s[i + 1]) && isxdigit((unsigned char)s[i + 2])) {
            o += (char)strtol(s.substr(i + 1, 2).c_str(), nullptr, 16);
            i += 2;
        } else o += s[i];
    }
    return o;
}
// uri of a buffer or an image: "data:...;base64,XXXX" or a path relative to the .gltf
std::vector<uint8_t> load_uri(const std::string &uri, const std::string &dir) {
    if (uri.compare(0, 5, "data:") == 0) {
        size_t k = uri.find(";base64,");
        if (k == std::string::npos) bad("data URI without base64 payload");
        return base64(uri.data() + k + 8, uri.size() - k - 8);
    }
    return read_file(dir + percent_decode(uri));
}

// ---------------------------------------------------------------------------------------------------------------------
// PNG (ISO/IEC 15948): chunks, zlib stream (system zlib), the five scanline filters, every colour type at 8 / 16 bits and
// the sub-byte grey / palette depths; no Adam7 interlace
// ---------------------------------------------------------------------------------------------------------------------
uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

}  // namespace

bool DecodePNG(const uint8_t *data, size_t size, uint32_t &width, uint32_t &height, std::vector<uint8_t> &rgba, std::string &error) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (size < 8 || memcmp(data, sig, 8)) { error = "not a PNG (JPEG and other formats need a decoder this loader does not have)"; return false; }
    size_t pos = 8;
    uint32_t W = 0, H = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    bool end_seen = false;
    while (pos + 12 <= size && !end_seen) {
        uint32_t len = be32(data + pos);
        const uint8_t *tag = data + pos + 4, *body = data + pos + 8;
        if ((size_t)len > size - pos - 12) { error = "PNG: truncated chunk"; return false; }
        if (!memcmp(tag, "IHDR", 4)) {
            if (len < 13) { error = "PNG: short IHDR"; return false; }
            W = be32(body); H = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
        } else if (!memcmp(tag, "PLTE", 4)) palette.assign(body, body + len);
        else if (!memcmp(tag, "tRNS", 4)) trns.assign(body, body + len);
        else if (!memcmp(tag, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(tag, "IEND", 4)) end_seen = true;
        pos += 12 + (size_t)len;
    }
    if (W == 0 || H == 0 || W > 32768 || H > 32768) { error = "PNG: bad extent"; return false; }
    if (interlace) { error = "PNG: Adam7 interlace is not supported"; return false; }
    int channels;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: error = "PNG: bad colour type"; return false;
    }
    if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4))) || (ctype == 3 && depth == 16)) {
        error = "PNG: bad bit depth"; return false;
    }
    const size_t bpp_bits = (size_t)channels * depth, stride = (W * bpp_bits + 7) / 8, bpp = std::max<size_t>(1, bpp_bits / 8);
    std::vector<uint8_t> raw((stride + 1) * H);
    uLongf out_len = (uLongf)raw.size();
    int zr = uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || out_len != raw.size()) { error = "PNG: zlib stream is corrupt"; return false; }
    // unfilter in place
    std::vector<uint8_t> prev(stride, 0);
    for (uint32_t y = 0; y < H; ++y) {
        uint8_t *row = raw.data() + (stride + 1) * y;
        const int ft = row[0];
        uint8_t *cur = row + 1;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int pred;
            switch (ft) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: {
                    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: error = "PNG: bad filter type"; return false;
            }
            cur[i] = (uint8_t)(cur[i] + pred);
        }
        memcpy(prev.data(), cur, stride);
    }
    // expand to RGBA8 (what stbi_load(..., STBI_rgb_alpha) hands the reference: 16-bit samples keep their high byte,
    // grey replicates, missing alpha = 255, tRNS applies)
    width = W; height = H;
    rgba.assign((size_t)W * H * 4, 255);
    auto sample = [&](const uint8_t *cur, size_t idx) -> uint32_t {      // idx-th sample of the row, raw value
        if (depth == 8) return cur[idx];
        if (depth == 16) return ((uint32_t)cur[2 * idx] << 8) | cur[2 * idx + 1];
        const size_t bit = idx * depth;
        return (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
    };
    auto to8 = [&](uint32_t v) -> uint8_t {
        if (depth == 8) return (uint8_t)v;
        if (depth == 16) return (uint8_t)(v >> 8);
        return (uint8_t)(v * 255u / ((1u << depth) - 1u));
    };
    for (uint32_t y = 0; y < H; ++y) {
        const uint8_t *cur = raw.data() + (stride + 1) * y + 1;
        uint8_t *o = rgba.data() + (size_t)y * W * 4;
        for (uint32_t x = 0; x < W; ++x, o += 4) {
            if (ctype == 3) {
                const uint32_t k = sample(cur, x);
                if ((size_t)k * 3 + 2 < palette.size()) { o[0] = palette[k * 3]; o[1] = palette[k * 3 + 1]; o[2] = palette[k * 3 + 2]; }
                else { o[0] = o[1] = o[2] = 0; }
                o[3] = k < trns.size() ? trns[k] : 255;
            } else if (ctype == 0 || ctype == 4) {
                const uint32_t g = sample(cur, (size_t)x * channels);
                o[0] = o[1] = o[2] = to8(g);
                if (ctype == 4) o[3] = to8(sample(cur, (size_t)x * 2 + 1));
                else if (trns.size() >= 2 && g == (((uint32_t)trns[0] << 8) | trns[1])) o[3] = 0;
            } else {
                const uint32_t r = sample(cur, (size_t)x * channels), g = sample(cur, (size_t)x * channels + 1), b = sample(cur, (size_t)x * channels + 2);
                o[0] = to8(r); o[1] = to8(g); o[2] = to8(b);
                if (ctype == 6) o[3] = to8(sample(cur, (size_t)x * 4 + 3));
                else if (trns.size() >= 6 && r == (((uint32_t)trns[0] << 8) | trns[1]) && g == (((uint32_t)trns[2] << 8) | trns[3]) &&
                         b == (((uint32_t)trns[4] << 8) | trns[5])) o[3] = 0;
            }
        }
    }
    return true;
}

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// small column-major matrix helpers (glm conventions: m[c*4+r])
// ---------------------------------------------------------------------------------------------------------------------
struct Mat4 { float m[16]; };
Mat4 identity() { Mat4 r{}; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }
Mat4 mul(const Mat4 &a, const Mat4 &b) {       // a * b
    Mat4 r{};
    for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 4; ++rr) {
            float acc = 0.0f;
            for (int k = 0; k < 4; ++k) acc += a.m[k * 4 + rr] * b.m[c * 4 + k];
            r.m[c * 4 + rr] = acc;
        }
    return r;
}
// inverse of a rigid transform T * R (what the camera set-up produces, scene_loader.cpp:62-65)
Mat4 inverse_rigid(const Mat4 &t) {
    Mat4 r = identity();
    for (int c = 0; c < 3; ++c)
        for (int rr = 0; rr < 3; ++rr) r.m[c * 4 + rr] = t.m[rr * 4 + c];
    for (int rr = 0; rr < 3; ++rr) r.m[12 + rr] = -(r.m[0 + rr] * t.m[12] + r.m[4 + rr] * t.m[13] + r.m[8 + rr] * t.m[14]);
    return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// glTF document
// ---------------------------------------------------------------------------------------------------------------------
struct Accessor {
    const uint8_t *base = nullptr;   // first element
    size_t stride = 0, count = 0;
    int component_type = 0;          // 5120 BYTE .. 5126 FLOAT
    int components = 0;              // 1 (SCALAR) .. 16 (MAT4)
    bool normalized = false;
};
int component_size(int t) {
    switch (t) {
        case 5120: case 5121: return 1;
        case 5122: case 5123: return 2;
        case 5125: case 5126: return 4;
    }
    bad("accessor: unknown componentType");
}
int type_components(const std::string &t) {
    if (t == "SCALAR") return 1;
    if (t == "VEC2") return 2;
    if (t == "VEC3") return 3;
    if (t == "VEC4") return 4;
    if (t == "MAT2") return 4;
    if (t == "MAT3") return 9;
    if (t == "MAT4") return 16;
    bad("accessor: unknown type '" + t + "'");
}

struct Document {
    Json root;
    std::string dir;
    std::vector<std::vector<uint8_t>> buffers;
    std::vector<int> parent;                 // node -> parent node (-1 = root)

    const Json &list(const char *key) const {
        static const Json empty = [] { Json j; j.type = Json::Array; return j; }();
        const Json *j = root.find(key);
        return (j && j->type == Json::Array) ? *j : empty;
    }
    void view_bytes(int view_index, const uint8_t *&ptr, size_t &len, size_t &stride) const {
        const Json &views = list("bufferViews");
        if (view_index < 0 || (size_t)view_index >= views.size()) bad("bufferView index out of range");
        const Json &v = views.array[view_index];
        const int b = v.integer("buffer", -1);
        if (b < 0 || (size_t)b >= buffers.size()) bad("buffer index out of range");
        const size_t off = (size_t)v.num("byteOffset", 0), n = (size_t)v.num("byteLength", 0);
        if (off + n > buffers[b].size()) bad("bufferView exceeds its buffer");
        ptr = buffers[b].data() + off; len = n; stride = (size_t)v.num("byteStride", 0);
    }
    Accessor accessor(int index) const {
        const Json &accs = list("accessors");
        if (index < 0 || (size_t)index >= accs.size()) bad("accessor index out of range");
        const Json &a = accs.array[index];
        if (a.has("sparse")) bad("sparse accessors are not supported");
        Accessor r;
        r.component_type = a.integer("componentType", 0);
        r.components = type_components(a.str("type"));
        r.count = (size_t)a.num("count", 0);
        const Json *nz = a.find("normalized");
        r.normalized = nz && nz->type == Json::Bool && nz->boolean;
        const uint8_t *p; size_t len, stride;
        view_bytes(a.integer("bufferView", -1), p, len, stride);
        const size_t off = (size_t)a.num("byteOffset", 0), elem = (size_t)component_size(r.component_type) * r.components;
        r.stride = stride ? stride : elem;
        if (r.count && off + (r.count - 1) * r.stride + elem > len) bad("accessor exceeds its bufferView");
        r.base = p + off;
        return r;
    }
};
// cgltf_accessor_read_float: component -> float, normalised integers scaled to [0,1] / [-1,1]
void read_float(const Accessor &a, size_t i, float *out, int n) {
    const uint8_t *e = a.base + i * a.stride;
    for (int k = 0; k < n; ++k) {
        if (k >= a.components) { out[k] = 0.0f; continue; }
        switch (a.component_type) {
            case 5126: { float f; memcpy(&f, e + 4 * k, 4); out[k] = f; break; }
            case 5120: { int8_t v = (int8_t)e[k]; out[k] = a.normalized ? std::max((float)v / 127.0f, -1.0f) : (float)v; break; }
            case 5121: { uint8_t v = e[k]; out[k] = a.normalized ? (float)v / 255.0f : (float)v; break; }
            case 5122: { int16_t v; memcpy(&v, e + 2 * k, 2); out[k] = a.normalized ? std::max((float)v / 32767.0f, -1.0f) : (float)v; break; }
            case 5123: { uint16_t v; memcpy(&v, e + 2 * k, 2); out[k] = a.normalized ? (float)v / 65535.0f : (float)v; break; }
            case 5125: { uint32_t v; memcpy(&v, e + 4 * k, 4); out[k] = (float)v; break; }
        }
    }
}
// cgltf_accessor_read_index
uint32_t read_index(const Accessor &a, size_t i) {
    const uint8_t *e = a.base + i * a.stride;
    switch (a.component_type) {
        case 5121: return e[0];
        case 5123: { uint16_t v; memcpy(&v, e, 2); return v; }
        case 5125: { uint32_t v; memcpy(&v, e, 4); return v; }
    }
    bad("index accessor must be UNSIGNED_BYTE / SHORT / INT");
}

void load_document(const char *path, Document &doc) {
    std::vector<uint8_t> file = read_file(path);
    doc.dir = parent_dir(path);
    std::vector<uint8_t> glb_bin;
    bool have_bin = false;
    if (file.size() >= 12 && !memcmp(file.data(), "glTF", 4)) {
        // .glb container: 12-byte header, then chunks (length, type, data); first JSON, optional BIN
        auto le32 = [&](size_t o) { uint32_t v; memcpy(&v, file.data() + o, 4); return v; };
        if (le32(4) != 2) bad(".glb: only container version 2 is supported");
        size_t pos = 12;
        bool have_json = false;
        while (pos + 8 <= file.size()) {
            const uint32_t len = le32(pos), type = le32(pos + 4);
            if ((size_t)len > file.size() - pos - 8) bad(".glb: truncated chunk");
            if (type == 0x4E4F534A && !have_json) { doc.root = parse_json((const char *)file.data() + pos + 8, len); have_json = true; }
            else if (type == 0x004E4942 && !have_bin) { glb_bin.assign(file.data() + pos + 8, file.data() + pos + 8 + len); have_bin = true; }
            pos += 8 + (size_t)len;
        }
        if (!have_json) bad(".glb: no JSON chunk");
    } else {
        doc.root = parse_json((const char *)file.data(), file.size());
    }
    if (doc.root.type != Json::Object) bad("glTF: the document is not a JSON object");
    const Json *asset = doc.root.find("asset");
    if (!asset || asset->str("version").compare(0, 2, "2.") != 0) bad("glTF: asset.version 2.x required");
    // cgltf_load_buffers (scene_loader.cpp:235)
    const Json &bufs = doc.list("buffers");
    for (size_t i = 0; i < bufs.size(); ++i) {
        const Json &b = bufs.array[i];
        if (b.has("uri")) doc.buffers.push_back(load_uri(b.str("uri"), doc.dir));
        else if (i == 0 && have_bin) doc.buffers.push_back(glb_bin);
        else bad("buffer without uri outside a .glb");
        if (doc.buffers.back().size() < (size_t)b.num("byteLength", 0)) bad("buffer is shorter than its byteLength");
    }
    const Json &nodes = doc.list("nodes");
    doc.parent.assign(nodes.size(), -1);
    for (size_t i = 0; i < nodes.size(); ++i) {
        const Json *ch = nodes.array[i].find("children");
        if (!ch) continue;
        for (const Json &c : ch->array) {
            const int k = (int)c.number;
            if (k < 0 || (size_t)k >= nodes.size()) bad("node child index out of range");
            doc.parent[k] = (int)i;
        }
    }
}

Mat4 node_local(const Json &n) {
    Mat4 r = identity();
    if (const Json *m = n.find("matrix")) {
        if (m->size() != 16) bad("node.matrix must have 16 elements");
        for (int i = 0; i < 16; ++i) r.m[i] = (float)m->array[i].number;
        return r;
    }
    float t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
    if (const Json *j = n.find("translation")) for (int i = 0; i < 3 && i < (int)j->size(); ++i) t[i] = (float)j->array[i].number;
    if (const Json *j = n.find("rotation")) for (int i = 0; i < 4 && i < (int)j->size(); ++i) q[i] = (float)j->array[i].number;
    if (const Json *j = n.find("scale")) for (int i = 0; i < 3 && i < (int)j->size(); ++i) s[i] = (float)j->array[i].number;
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    r.m[0] = (1 - 2 * y * y - 2 * z * z) * s[0]; r.m[1] = (2 * x * y + 2 * z * w) * s[0]; r.m[2] = (2 * x * z - 2 * y * w) * s[0];
    r.m[4] = (2 * x * y - 2 * z * w) * s[1]; r.m[5] = (1 - 2 * x * x - 2 * z * z) * s[1]; r.m[6] = (2 * y * z + 2 * x * w) * s[1];
    r.m[8] = (2 * x * z + 2 * y * w) * s[2]; r.m[9] = (2 * y * z - 2 * x * w) * s[2]; r.m[10] = (1 - 2 * x * x - 2 * y * y) * s[2];
    r.m[12] = t[0]; r.m[13] = t[1]; r.m[14] = t[2];
    return r;
}
// cgltf_node_transform_world
Mat4 node_world(const Document &doc, int node) {
    const Json &nodes = doc.list("nodes");
    Mat4 w = node_local(nodes.array[node]);
    int guard = 0;
    for (int p = doc.parent[node]; p >= 0; p = doc.parent[p]) {
        w = mul(node_local(nodes.array[p]), w);
        if (++guard > 4096) bad("node hierarchy has a cycle");
    }
    return w;
}

// scene_loader.cpp:8-38
VkFilter GetVkFilter(int filter) {
    switch (filter) {
        case 0x2600: case 0x2700: case 0x2701: return VK_FILTER_NEAREST;
        case 0x2601: case 0x2702: case 0x2703: return VK_FILTER_LINEAR;
        default: return VK_FILTER_LINEAR;        // reference: assert(false), then LINEAR
    }
}
VkSamplerAddressMode GetVkAddressMode(int mode) {
    switch (mode) {
        case 0x812F: return VK_SAMPLER_ADDRESS_MODE_CLAMP_TO_EDGE;
        case 0x812D: return VK_SAMPLER_ADDRESS_MODE_CLAMP_TO_BORDER;
        case 0x2901: return VK_SAMPLER_ADDRESS_MODE_REPEAT;
        case 0x8370: return VK_SAMPLER_ADDRESS_MODE_MIRRORED_REPEAT;
        default: return VK_SAMPLER_ADDRESS_MODE_CLAMP_TO_EDGE;   // reference: assert(false), then CLAMP_TO_EDGE
    }
}

// scene_loader.cpp:43-72: the camera node
void setup_camera(Scene &scene, const Mat4 &M, float yfov, float aspect, float znear) {
    // VkUtils::InfiniteReverseDepthProjection (vulkan_utils.h:494-503)
    const float scale = 1.0f / tanf(yfov * 0.5f);
    Camera &c = scene.camera;
    memset(c.perspective, 0, sizeof(c.perspective));
    c.perspective[0] = scale / aspect; c.perspective[5] = scale; c.perspective[11] = -1.0f; c.perspective[14] = znear;
    // glm::extractEulerAngleYXZ, then transform = T * glm::yawPitchRoll(yaw, pitch, roll) (drops any scale)
    const float T1 = atan2f(M.m[8], M.m[10]);
    const float C2 = sqrtf(M.m[1] * M.m[1] + M.m[5] * M.m[5]);
    const float T2 = atan2f(-M.m[9], C2);
    const float S1 = sinf(T1), C1 = cosf(T1);
    const float T3 = atan2f(S1 * M.m[6] - C1 * M.m[4], C1 * M.m[0] - S1 * M.m[2]);
    const float ch = cosf(T1), sh = sinf(T1), cp = cosf(T2), sp = sinf(T2), cb = cosf(T3), sb = sinf(T3);
    Mat4 R = identity();
    R.m[0] = ch * cb + sh * sp * sb; R.m[1] = sb * cp; R.m[2] = -sh * cb + ch * sp * sb;
    R.m[4] = -ch * sb + sh * sp * cb; R.m[5] = cb * cp; R.m[6] = sb * sh + ch * sp * cb;
    R.m[8] = sh * cp; R.m[9] = -sp; R.m[10] = ch * cp;
    Mat4 Tm = identity();
    Tm.m[12] = M.m[12]; Tm.m[13] = M.m[13]; Tm.m[14] = M.m[14];
    const Mat4 TR = mul(Tm, R), view = inverse_rigid(TR);
    memcpy(c.transform, TR.m, sizeof(TR.m));
    memcpy(c.view, view.m, sizeof(view.m));
    c.yaw = T1; c.pitch = T2; c.roll = T3;
}

// scene_loader.cpp:74-100: a KHR_lights_punctual directional light
void setup_directional_light(Scene &scene, const Mat4 &M, const float col[3]) {
    // glm::decompose -> rotation; direction = normalize(rot * (0,0,-1)) = minus the normalised third basis vector
    float d[3] = {-M.m[8], -M.m[9], -M.m[10]};
    const float len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (len > 0) { d[0] /= len; d[1] /= len; d[2] /= len; }
    DirectionalLight &dl = scene.directional_light;
    // glm::ortho(-8, 8, -8, 8, 12, 0.1) with GLM_FORCE_DEPTH_ZERO_TO_ONE (right-handed, depth 0..1; pch.h:39)
    Mat4 P = identity();
    const float l = -8, r = 8, b = -8, t = 8, zn = 12.0f, zf = 0.1f;
    P.m[0] = 2 / (r - l); P.m[5] = 2 / (t - b); P.m[10] = -1 / (zf - zn);
    P.m[12] = -(r + l) / (r - l); P.m[13] = -(t + b) / (t - b); P.m[14] = -zn / (zf - zn);
    // glm::lookAt(-dir * 12, 0, (0,1,0)), right-handed
    const float eye[3] = {-d[0] * 12.0f, -d[1] * 12.0f, -d[2] * 12.0f};
    float f[3] = {-eye[0], -eye[1], -eye[2]};
    const float fl = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    for (float &v : f) v /= fl;
    const float up[3] = {0, 1, 0};
    float s[3] = {f[1] * up[2] - f[2] * up[1], f[2] * up[0] - f[0] * up[2], f[0] * up[1] - f[1] * up[0]};
    const float sl = sqrtf(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    for (float &v : s) v /= sl;
    const float u[3] = {s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]};
    Mat4 V = identity();
    V.m[0] = s[0]; V.m[4] = s[1]; V.m[8] = s[2];
    V.m[1] = u[0]; V.m[5] = u[1]; V.m[9] = u[2];
    V.m[2] = -f[0]; V.m[6] = -f[1]; V.m[10] = -f[2];
    V.m[12] = -(s[0] * eye[0] + s[1] * eye[1] + s[2] * eye[2]);
    V.m[13] = -(u[0] * eye[0] + u[1] * eye[1] + u[2] * eye[2]);
    V.m[14] = f[0] * eye[0] + f[1] * eye[1] + f[2] * eye[2];
    const Mat4 PV = mul(P, V);
    memcpy(dl.projview, PV.m, sizeof(PV.m));
    dl.direction[0] = d[0]; dl.direction[1] = d[1]; dl.direction[2] = d[2]; dl.direction[3] = 0.0f;
    dl.color[0] = col[0]; dl.color[1] = col[1]; dl.color[2] = col[2]; dl.color[3] = 1.0f;
    const float intensity = scene.name == "Pica.glb" ? 2.0f : 30.0f;                  // :98 (the glTF intensity is ignored)
    for (float &v : dl.intensity) v = intensity;
}

// scene_loader.cpp:317-329: no directional light in the file
void default_directional_light(Scene &scene) {
    DirectionalLight &dl = scene.directional_light;
    dl = DirectionalLight{};
    dl.direction[0] = 0.0f; dl.direction[1] = -1.0f; dl.direction[2] = 0.01f; dl.direction[3] = 0.0f;
    dl.color[0] = dl.color[1] = dl.color[2] = 1.0f; dl.color[3] = 0.0f;
}

// scene_loader.cpp:284-290: four channels whatever the file holds
void decode_image(const std::vector<uint8_t> &bytes, ParsedTexture &pt, const std::string &what) {
    std::string err;
    bool decoded = false;
#ifdef VHR_HAVE_STB_IMAGE
    {
        int x = 0, y = 0, n = 0;
        if (uint8_t *px = stbi_load_from_memory(bytes.data(), (int)bytes.size(), &x, &y, &n, STBI_rgb_alpha)) {
            pt.width = (uint32_t)x; pt.height = (uint32_t)y;
            pt.rgba.assign(px, px + (size_t)x * y * 4);
            stbi_image_free(px);
            decoded = true;
        } else {
            err = std::string("stb_image: ") + (stbi_failure_reason() ? stbi_failure_reason() : "unknown failure");
        }
    }
#endif
    if (!decoded && !DecodePNG(bytes.data(), bytes.size(), pt.width, pt.height, pt.rgba, err)) bad(what + ": " + err);
}

int texture_of(const Json *info) {      // textureInfo object -> texture index or -1
    if (!info || info->type != Json::Object) return -1;
    return info->integer("index", -1);
}

}  // namespace

static void ParseSceneOwn(const char *path, ParsedScene &out) {
    Document doc;
    load_document(path, doc);
    out = ParsedScene();
    Scene &scene = out.scene;
    scene.name = file_name(path);                         // scene_loader.cpp:338
    const Json &meshes = doc.list("meshes"), &materials = doc.list("materials"), &textures = doc.list("textures");
    const Json &images = doc.list("images"), &samplers = doc.list("samplers"), &nodes = doc.list("nodes");
    const Json &cameras = doc.list("cameras");

    auto material_of = [&](const Json &prim) -> const Json * {
        const int m = prim.integer("material", -1);
        return (m >= 0 && (size_t)m < materials.size()) ? &materials.array[m] : nullptr;
    };
    // ---- textures to upload and their formats (scene_loader.cpp:239-275): base colour sRGB, everything else UNORM; the
    // first use decides. Order = first appearance scanning meshes -> primitives.
    std::vector<std::pair<int, VkFormat>> to_upload;
    auto want = [&](int tex, VkFormat fmt) {
        if (tex < 0) return;
        if ((size_t)tex >= textures.size()) bad("texture index out of range");
        for (auto &t : to_upload)
            if (t.first == tex) return;
        to_upload.emplace_back(tex, fmt);
    };
    for (const Json &mesh : meshes.array)
        for (const Json &prim : mesh.at("primitives").array) {
            const Json *mat = material_of(prim);
            if (!mat) continue;
            if (const Json *pbr = mat->find("pbrMetallicRoughness")) {
                want(texture_of(pbr->find("baseColorTexture")), VK_FORMAT_R8G8B8A8_SRGB);
                want(texture_of(pbr->find("metallicRoughnessTexture")), VK_FORMAT_R8G8B8A8_UNORM);
            }
            want(texture_of(mat->find("normalTexture")), VK_FORMAT_R8G8B8A8_UNORM);
        }
    // ---- decode + describe (scene_loader.cpp:277-309). The reference keys its lookup table by the IMAGE (its name
    // pointer): two textures sharing an image resolve to whichever was uploaded last; kept.
    std::map<int, int> slot_of_image;
    for (auto &tu : to_upload) {
        const Json &tex = textures.array[tu.first];
        const int img = tex.integer("source", -1);
        if (img < 0 || (size_t)img >= images.size()) bad("texture without a valid image source");
        const Json &image = images.array[img];
        std::vector<uint8_t> bytes;
        if (image.has("uri")) bytes = load_uri(image.str("uri"), doc.dir);
        else {
            const uint8_t *p; size_t len, stride;
            doc.view_bytes(image.integer("bufferView", -1), p, len, stride);
            bytes.assign(p, p + len);
        }
        ParsedTexture pt;
        decode_image(bytes, pt, "image " + std::to_string(img));
        pt.format = tu.second;
        pt.name = image.str("name", image.str("uri"));
        const int smp = tex.integer("sampler", -1);
        if (smp >= 0 && (size_t)smp < samplers.size()) {
            const Json &s = samplers.array[smp];
            // cgltf leaves absent filters at 0 (reference: assert, LINEAR) and defaults the wrap modes to REPEAT (10497)
            pt.sampler = SamplerInfo{GetVkFilter(s.integer("magFilter", 0)), GetVkFilter(s.integer("minFilter", 0)),
                                     GetVkAddressMode(s.integer("wrapS", 0x2901)), GetVkAddressMode(s.integer("wrapT", 0x2901))};
        }
        slot_of_image[img] = (int)out.textures.size();
        out.textures.push_back(std::move(pt));
    }
    auto slot_for = [&](int tex) -> int {
        if (tex < 0) return -1;
        const int img = textures.array[tex].integer("source", -1);
        auto it = slot_of_image.find(img);
        return it == slot_of_image.end() ? -1 : it->second;
    };

    // ---- nodes in file order (scene_loader.cpp:311-315 -> ParseNode :40-231) ---------------------------------------------
    bool have_directional_light = false;
    const Json *lights = nullptr;
    if (const Json *ext = doc.root.find("extensions"))
        if (const Json *lp = ext->find("KHR_lights_punctual")) lights = lp->find("lights");
    for (size_t ni = 0; ni < nodes.size(); ++ni) {
        const Json &node = nodes.array[ni];
        if (node.has("camera")) {                                                              // :43-72
            const int ci = node.integer("camera", -1);
            if (ci < 0 || (size_t)ci >= cameras.size()) bad("camera index out of range");
            const Json &cam = cameras.array[ci];
            if (cam.str("type") != "perspective") bad("only perspective cameras are supported (scene_loader.cpp:44)");
            const Json &persp = cam.at("perspective");
            const float yfov = (float)persp.num("yfov", 1.0), aspect = (float)persp.num("aspectRatio", 1.0), znear = (float)persp.num("znear", 0.1);
            setup_camera(scene, node_world(doc, (int)ni), yfov, aspect, znear);
            continue;
        }
        int light_index = -1;
        if (const Json *ext = node.find("extensions"))
            if (const Json *lp = ext->find("KHR_lights_punctual")) light_index = lp->integer("light", -1);
        if (light_index >= 0 && lights && (size_t)light_index < lights->size() && lights->array[light_index].str("type") == "directional") {   // :74-100
            const Json &light = lights->array[light_index];
            float col[3] = {1, 1, 1};
            if (const Json *c = light.find("color")) for (int i = 0; i < 3 && i < (int)c->size(); ++i) col[i] = (float)c->array[i].number;
            setup_directional_light(scene, node_world(doc, (int)ni), col);
            have_directional_light = true;
            continue;
        }
        if (!node.has("mesh")) continue;                                                       // :102-104
        const int mi = node.integer("mesh", -1);
        if (mi < 0 || (size_t)mi >= meshes.size()) bad("mesh index out of range");
        const Mat4 transform = node_world(doc, (int)ni);
        Mesh mesh;
        for (const Json &prim : meshes.array[mi].at("primitives").array) {
            if (prim.integer("mode", 4) != 4) bad("only triangle-list primitives are supported (scene_loader.cpp:112)");
            const uint32_t vertex_offset = (uint32_t)out.vertices.size(), index_offset = (uint32_t)out.indices.size();
            const Json &attrs = prim.at("attributes");
            auto attr = [&](const char *name, int comps, Accessor &acc) -> bool {
                const Json *a = attrs.find(name);
                if (!a) return false;
                acc = doc.accessor((int)a->number);
                if (acc.components != comps) bad(std::string(name) + ": unexpected accessor type");
                return true;
            };
            Accessor pos, nrm, tan, uv0, uv1;
            if (!attr("POSITION", 3, pos)) bad("primitive without POSITION (scene_loader.cpp:149)");
            const bool has_n = attr("NORMAL", 3, nrm), has_t = attr("TANGENT", 4, tan), has_uv0 = attr("TEXCOORD_0", 2, uv0), has_uv1 = attr("TEXCOORD_1", 2, uv1);
            for (size_t j = 0; j < pos.count; ++j) {                                           // :150-173
                Vertex v{};
                read_float(pos, j, v.pos, 3);
                if (has_n && j < nrm.count) read_float(nrm, j, v.normal, 3);
                if (has_t && j < tan.count) read_float(tan, j, v.tangent, 4);
                if (has_uv0 && j < uv0.count) read_float(uv0, j, v.uv0, 2);
                if (has_uv1 && j < uv1.count) read_float(uv1, j, v.uv1, 2);
                out.vertices.push_back(v);
            }
            if (!prim.has("indices")) bad("primitive without indices (scene_loader.cpp:175)");
            const Accessor idx = doc.accessor(prim.integer("indices", -1));
            for (size_t j = 0; j < idx.count; ++j) {
                const uint32_t k = read_index(idx, j);
                if (k >= pos.count) bad("index exceeds the primitive's vertex count");
                out.indices.push_back(k);
            }
            // :182-218
            Material material{{1.0f, 1.0f, 1.0f, 1.0f}, -1, -1, -1, 1.0f, 1.0f, 0, 0.0f};
            if (const Json *mat = material_of(prim)) {
                const Json *pbr = mat->find("pbrMetallicRoughness");
                static const Json none = [] { Json j; j.type = Json::Object; return j; }();
                if (!pbr) pbr = &none;                    // cgltf fills the defaults: factors 1
                const int albedo = texture_of(pbr->find("baseColorTexture"));
                if (albedo >= 0) material.base_color_texture = slot_for(albedo);
                else if (const Json *f = pbr->find("baseColorFactor")) {
                    for (int i = 0; i < 4 && i < (int)f->size(); ++i) material.base_color[i] = (float)f->array[i].number;
                }
                const int mr = texture_of(pbr->find("metallicRoughnessTexture"));
                if (mr >= 0) material.metallic_roughness_texture = slot_for(mr);
                material.metallic_factor = (float)pbr->num("metallicFactor", 1.0);
                material.roughness_factor = (float)pbr->num("roughnessFactor", 1.0);
                const int nt = texture_of(mat->find("normalTexture"));
                if (nt >= 0) {
                    material.normal_map = slot_for(nt);
                    if (!has_t) bad("normal map without vertex tangents (scene_loader.cpp:212-213)");
                }
                if (mat->str("alphaMode", "OPAQUE") == "MASK") {
                    material.alpha_mask = 1;
                    material.alpha_cutoff = (float)mat->num("alphaCutoff", 0.5);
                }
            }
            Primitive p{};
            memcpy(p.transform, transform.m, sizeof(p.transform));
            p.material = material;
            p.vertex_offset = vertex_offset; p.index_offset = index_offset; p.index_count = (uint32_t)idx.count;
            mesh.primitives.push_back(p);
        }
        scene.meshes.push_back(std::move(mesh));
    }
    if (!have_directional_light) default_directional_light(scene);                             // :317-329
}

#ifdef VHR_HAVE_CGLTF
// ---------------------------------------------------------------------------------------------------------------------
// The same flattening on cgltf's document (the reference's ParseglTF / ParseNode, scene_loader.cpp:40-334)
// ---------------------------------------------------------------------------------------------------------------------
static void ParseSceneCgltf(const char *path, ParsedScene &out) {
    cgltf_options options{};
    cgltf_data *data = nullptr;
    const cgltf_result pr = cgltf_parse_file(&options, path, &data);                              // :338-341
    if (pr != cgltf_result_success) bad(pr == cgltf_result_file_not_found ? std::string("cannot open '") + path + "'" : "cgltf_parse_file: error " + std::to_string((int)pr));
    struct Guard { cgltf_data *d; ~Guard() { cgltf_free(d); } } guard{data};
    const cgltf_result lr = cgltf_load_buffers(&options, data, path);                             // :234-235
    if (lr != cgltf_result_success) bad("cgltf_load_buffers: error " + std::to_string((int)lr));
    if (cgltf_validate(data) != cgltf_result_success) bad("cgltf_validate: accessor / buffer view / index out of range");
    out = ParsedScene();
    Scene &scene = out.scene;
    scene.name = file_name(path);                                                                 // :338
    const std::string dir = parent_dir(path);

    // ---- textures to upload and their formats (:239-275): base colour sRGB, everything else UNORM; the first use decides
    std::vector<std::pair<const cgltf_texture *, VkFormat>> to_upload;
    auto want = [&](const cgltf_texture *tex, VkFormat fmt) {
        if (!tex) return;
        for (auto &t : to_upload)
            if (t.first == tex) return;
        to_upload.emplace_back(tex, fmt);
    };
    for (cgltf_size i = 0; i < data->meshes_count; ++i)
        for (cgltf_size j = 0; j < data->meshes[i].primitives_count; ++j) {
            const cgltf_material *mat = data->meshes[i].primitives[j].material;
            if (!mat) continue;
            if (mat->has_pbr_metallic_roughness) {
                want(mat->pbr_metallic_roughness.base_color_texture.texture, VK_FORMAT_R8G8B8A8_SRGB);
                want(mat->pbr_metallic_roughness.metallic_roughness_texture.texture, VK_FORMAT_R8G8B8A8_UNORM);
            }
            want(mat->normal_texture.texture, VK_FORMAT_R8G8B8A8_UNORM);
        }
    // ---- decode + describe (:277-309); the lookup table is keyed by the IMAGE, as in the reference
    std::map<const cgltf_image *, int> slot_of_image;
    for (auto &tu : to_upload) {
        const cgltf_texture *tex = tu.first;
        const cgltf_image *image = tex->image;
        if (!image) bad("texture without a valid image source");
        std::vector<uint8_t> bytes;
        if (image->uri) bytes = load_uri(image->uri, dir);
        else if (image->buffer_view && image->buffer_view->buffer && image->buffer_view->buffer->data) {
            const uint8_t *p = (const uint8_t *)image->buffer_view->buffer->data + image->buffer_view->offset;
            bytes.assign(p, p + image->buffer_view->size);
        } else bad("image without uri or bufferView");
        ParsedTexture pt;
        decode_image(bytes, pt, "image " + std::to_string((int)(image - data->images)));
        pt.format = tu.second;
        pt.name = image->name ? image->name : (image->uri ? image->uri : "");
        if (const cgltf_sampler *smp = tex->sampler)
            pt.sampler = SamplerInfo{GetVkFilter(smp->mag_filter), GetVkFilter(smp->min_filter), GetVkAddressMode(smp->wrap_s), GetVkAddressMode(smp->wrap_t)};
        slot_of_image[image] = (int)out.textures.size();
        out.textures.push_back(std::move(pt));
    }
    auto slot_for = [&](const cgltf_texture *tex) -> int {
        if (!tex) return -1;
        auto it = slot_of_image.find(tex->image);
        return it == slot_of_image.end() ? -1 : it->second;
    };

    // ---- nodes in file order (:311-315 -> ParseNode :40-231)
    bool have_directional_light = false;
    for (cgltf_size ni = 0; ni < data->nodes_count; ++ni) {
        const cgltf_node &node = data->nodes[ni];
        Mat4 world;
        cgltf_node_transform_world(&node, world.m);
        if (node.camera) {                                                                        // :43-72
            if (node.camera->type != cgltf_camera_type_perspective) bad("only perspective cameras are supported (scene_loader.cpp:44)");
            const cgltf_camera_perspective &pc = node.camera->data.perspective;
            setup_camera(scene, world, pc.yfov, pc.aspect_ratio != 0.0f ? pc.aspect_ratio : 1.0f, pc.znear);
            continue;
        }
        if (node.light && node.light->type == cgltf_light_type_directional) {                     // :74-100
            setup_directional_light(scene, world, node.light->color);
            have_directional_light = true;
            continue;
        }
        if (!node.mesh) continue;                                                                 // :102-104
        Mesh mesh;
        for (cgltf_size pi = 0; pi < node.mesh->primitives_count; ++pi) {
            const cgltf_primitive &prim = node.mesh->primitives[pi];
            if (prim.type != cgltf_primitive_type_triangles) bad("only triangle-list primitives are supported (scene_loader.cpp:112)");
            const uint32_t vertex_offset = (uint32_t)out.vertices.size(), index_offset = (uint32_t)out.indices.size();
            const cgltf_accessor *pos = nullptr, *nrm = nullptr, *tan = nullptr, *uv0 = nullptr, *uv1 = nullptr;
            for (cgltf_size a = 0; a < prim.attributes_count; ++a) {                             // :123-147
                const cgltf_attribute &at = prim.attributes[a];
                auto expect = [&](cgltf_type t, const char *name) { if (at.data->type != t) bad(std::string(name) + ": unexpected accessor type"); };
                if (at.type == cgltf_attribute_type_position) { expect(cgltf_type_vec3, "POSITION"); pos = at.data; }
                else if (at.type == cgltf_attribute_type_normal) { expect(cgltf_type_vec3, "NORMAL"); nrm = at.data; }
                else if (at.type == cgltf_attribute_type_tangent) { expect(cgltf_type_vec4, "TANGENT"); tan = at.data; }
                else if (at.type == cgltf_attribute_type_texcoord && at.index == 0) { expect(cgltf_type_vec2, "TEXCOORD_0"); uv0 = at.data; }
                else if (at.type == cgltf_attribute_type_texcoord && at.index == 1) { expect(cgltf_type_vec2, "TEXCOORD_1"); uv1 = at.data; }
            }
            if (!pos) bad("primitive without POSITION (scene_loader.cpp:149)");
            if (pos->is_sparse || (nrm && nrm->is_sparse) || (tan && tan->is_sparse) || (uv0 && uv0->is_sparse) || (uv1 && uv1->is_sparse))
                bad("sparse accessors are not supported");
            for (cgltf_size j = 0; j < pos->count; ++j) {                                         // :150-173
                Vertex v{};
                cgltf_accessor_read_float(pos, j, v.pos, 3);
                if (nrm && j < nrm->count) cgltf_accessor_read_float(nrm, j, v.normal, 3);
                if (tan && j < tan->count) cgltf_accessor_read_float(tan, j, v.tangent, 4);
                if (uv0 && j < uv0->count) cgltf_accessor_read_float(uv0, j, v.uv0, 2);
                if (uv1 && j < uv1->count) cgltf_accessor_read_float(uv1, j, v.uv1, 2);
                out.vertices.push_back(v);
            }
            if (!prim.indices) bad("primitive without indices (scene_loader.cpp:175)");
            if (prim.indices->component_type != cgltf_component_type_r_8u && prim.indices->component_type != cgltf_component_type_r_16u &&
                prim.indices->component_type != cgltf_component_type_r_32u)
                bad("index accessor must be UNSIGNED_BYTE / SHORT / INT");
            for (cgltf_size j = 0; j < prim.indices->count; ++j) {
                const uint32_t k = (uint32_t)cgltf_accessor_read_index(prim.indices, j);
                if (k >= pos->count) bad("index exceeds the primitive's vertex count");
                out.indices.push_back(k);
            }
            Material material{{1.0f, 1.0f, 1.0f, 1.0f}, -1, -1, -1, 1.0f, 1.0f, 0, 0.0f};        // :182-218
            if (const cgltf_material *mat = prim.material) {
                const cgltf_pbr_metallic_roughness &pbr = mat->pbr_metallic_roughness;        // cgltf fills the defaults: factors 1
                if (pbr.base_color_texture.texture) material.base_color_texture = slot_for(pbr.base_color_texture.texture);
                else if (mat->has_pbr_metallic_roughness) memcpy(material.base_color, pbr.base_color_factor, sizeof(material.base_color));
                if (pbr.metallic_roughness_texture.texture) material.metallic_roughness_texture = slot_for(pbr.metallic_roughness_texture.texture);
                if (mat->has_pbr_metallic_roughness) { material.metallic_factor = pbr.metallic_factor; material.roughness_factor = pbr.roughness_factor; }
                if (mat->normal_texture.texture) {
                    material.normal_map = slot_for(mat->normal_texture.texture);
                    if (!tan) bad("normal map without vertex tangents (scene_loader.cpp:212-213)");
                }
                if (mat->alpha_mode == cgltf_alpha_mode_mask) {
                    material.alpha_mask = 1;
                    material.alpha_cutoff = mat->alpha_cutoff;
                }
            }
            Primitive p{};
            memcpy(p.transform, world.m, sizeof(p.transform));
            p.material = material;
            p.vertex_offset = vertex_offset; p.index_offset = index_offset; p.index_count = (uint32_t)prim.indices->count;
            mesh.primitives.push_back(p);
        }
        scene.meshes.push_back(std::move(mesh));
    }
    if (!have_directional_light) default_directional_light(scene);                                // :317-329
}
#endif

void ParseScene(const char *path, ParsedScene &out) {
#ifdef VHR_HAVE_CGLTF
    const char *which = getenv("VHR_GLTF_PARSER");
    if (!which || strcmp(which, "own") != 0) { ParseSceneCgltf(path, out); return; }
#endif
    ParseSceneOwn(path, out);
}

Scene LoadScene(ResourceManager &resource_manager, const char *path) {
    ParsedScene parsed;
    try {
        ParseScene(path, parsed);
    } catch (const VhrHostError &e) {
        // scene_loader.cpp:344-346: print and hand back an empty scene
        printf("Error Parsing glTF 2.0 File (%s)\n", e.message.c_str());
        Scene empty;
        empty.name = file_name(path);
        return empty;
    }
    // textures first, then the geometry (scene_loader.cpp:277-331); slots are whatever UploadTexture hands out
    std::vector<int> slot(parsed.textures.size(), -1);
    for (size_t i = 0; i < parsed.textures.size(); ++i) {
        ParsedTexture &t = parsed.textures[i];
        slot[i] = (int)resource_manager.UploadTextureFromData(t.width, t.height, t.rgba.data(), t.format, &t.sampler);
    }
    auto rebase = [&](int32_t &idx) { if (idx >= 0) idx = slot[(size_t)idx]; };
    for (Mesh &mesh : parsed.scene.meshes)
        for (Primitive &p : mesh.primitives) {
            rebase(p.material.base_color_texture);
            rebase(p.material.metallic_roughness_texture);
            rebase(p.material.normal_map);
        }
    resource_manager.UpdateGeometry(parsed.vertices, parsed.indices, parsed.scene);
    return parsed.scene;
}

}  // namespace SceneLoader
