// OBJ / MTL ingest straight into the arrays b200rt_scene_create takes (include/b200rt.h, "ingest").
//
// Replaces Utils::parse_obj (source/utils.cpp:16-98): rapidobj::ParseFile + rapidobj::Triangulate, then the loop that fills
// ParsedOBJ { triangles, material_indices (+1: slot 0 is the built-in default material), emissive_triangle_indices, materials }.
// Written from the format and from parse_obj's observable behaviour, not from rapidobj's code:
//   * faces keep file order; triangles pass through; a quad is cut along its shorter diagonal (0-2 if |p0 - p2|^2 < |p1 - p3|^2,
//     else 1-3) exactly as rapidobj::Triangulate does, so the triangle list — and with it every primitive index, light pick and
//     material lookup downstream — matches the reference's; polygons with more than 4 vertices are fanned from their first vertex
//     (rapidobj ear-clips them: a documented divergence, none of the bundled scenes has one)
//   * materials: Ke, Kd, Pm, Pr, illum; roughness clamped to >= 1e-2 (:82); illum 0 (no PBR extension) -> roughness 1, metalness 0
//     (:84-92); slot 0 = { emission (1, 0, 1), diffuse 0, metalness 0, roughness 1 } (:75)
//   * a triangle is emissive when its material's Ke has a positive component (:62-69)
// Host code; one pass over the text per file.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "obj_ingest.h"

namespace b200rt {

namespace {

bool read_file(const std::string& path, std::string& out)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const size_t got = n > 0 ? std::fread(&out[0], 1, (size_t)n, f) : 0;
    std::fclose(f);
    out.resize(got);
    return true;
}

inline const char* skip_ws(const char* p, const char* end) { while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) p++; return p; }
inline const char* line_end(const char* p, const char* end) { while (p < end && *p != '\n') p++; return p; }

inline bool parse_float(const char*& p, const char* end, float& v)
{
    p = skip_ws(p, end);
    if (p >= end || *p == '\n') return false;
    char* q = nullptr;
    v = std::strtof(p, &q);
    if (q == p) return false;
    p = q;
    return true;
}

std::string rest_of_line(const char* p, const char* end)
{
    p = skip_ws(p, end);
    const char* e = line_end(p, end);
    while (e > p && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\r')) e--;
    return std::string(p, e);
}

struct Mtl { float ke[3] = { 0, 0, 0 }, kd[3] = { 0, 0, 0 }; float metallic = 0.0f, roughness = 0.0f; int illum = 0; };

bool parse_mtl(const std::string& path, std::vector<Mtl>& mats, std::unordered_map<std::string, int>& index, std::string& err)
{
    std::string text;
    if (!read_file(path, text)) { err = "cannot open material library " + path; return false; }
    const char* p = text.data();
    const char* end = p + text.size();
    Mtl* cur = nullptr;
    while (p < end)
    {
        p = skip_ws(p, end);
        const char* le = line_end(p, end);
        auto key = [&](const char* k) { const size_t n = std::strlen(k); return (size_t)(le - p) > n && std::strncmp(p, k, n) == 0 && (p[n] == ' ' || p[n] == '\t'); };
        if (key("newmtl"))
        {
            const std::string name = rest_of_line(p + 6, end);
            index[name] = (int)mats.size();
            mats.emplace_back();
            cur = &mats.back();
        }
        else if (cur)
        {
            const char* q = p + 2;
            if (key("Ke")) { for (int i = 0; i < 3; i++) parse_float(q, le, cur->ke[i]); }
            else if (key("Kd")) { for (int i = 0; i < 3; i++) parse_float(q, le, cur->kd[i]); }
            else if (key("Pm")) parse_float(q, le, cur->metallic);
            else if (key("Pr")) parse_float(q, le, cur->roughness);
            else if (key("illum")) { q = p + 5; float v = 0; if (parse_float(q, le, v)) cur->illum = (int)v; }
        }
        p = le < end ? le + 1 : end;
    }
    return true;
}

// one "v/vt/vn" reference of a face: only the position index matters here; returns false at the end of the line
inline bool parse_face_vertex(const char*& p, const char* le, long n_positions, long& index_out)
{
    p = skip_ws(p, le);
    if (p >= le) return false;
    char* q = nullptr;
    const long v = std::strtol(p, &q, 10);
    if (q == p) return false;
    p = q;
    while (p < le && *p != ' ' && *p != '\t' && *p != '\r') p++;          // skip /vt/vn
    index_out = v > 0 ? v - 1 : n_positions + v;                          // 1-based, or relative to the vertices read so far
    return true;
}

} // namespace

bool load_obj(const std::string& path, ParsedObjArrays& out, std::string& err)
{
    out = ParsedObjArrays();
    std::string text;
    if (!read_file(path, text)) { err = "cannot open " + path; return false; }
    const std::string dir = path.find_last_of("/\\") == std::string::npos ? std::string(".") : path.substr(0, path.find_last_of("/\\"));

    std::vector<float> pos;
    pos.reserve(text.size() / 24);
    std::vector<Mtl> mats;
    std::unordered_map<std::string, int> mat_index;
    int cur_mat = -1;
    std::vector<long> poly;
    const char* p = text.data();
    const char* end = p + text.size();
    while (p < end)
    {
        p = skip_ws(p, end);
        const char* le = line_end(p, end);
        if (le - p >= 2 && p[0] == 'v' && (p[1] == ' ' || p[1] == '\t'))
        {
            const char* q = p + 1;
            float v[3] = { 0, 0, 0 };
            for (int i = 0; i < 3; i++) if (!parse_float(q, le, v[i])) { err = "bad vertex line"; return false; }
            pos.insert(pos.end(), v, v + 3);
        }
        else if (le - p >= 2 && p[0] == 'f' && (p[1] == ' ' || p[1] == '\t'))
        {
            const char* q = p + 1;
            poly.clear();
            long idx;
            const long n_pos = (long)(pos.size() / 3);
            while (parse_face_vertex(q, le, n_pos, idx))
            {
                if (idx < 0 || idx >= n_pos) { err = "face references vertex " + std::to_string(idx + 1) + " of " + std::to_string(n_pos); return false; }
                poly.push_back(idx);
            }
            if (poly.size() < 3) { err = "face with fewer than 3 vertices"; return false; }
            auto P = [&](long i) { return &pos[3 * (size_t)i]; };
            auto emit = [&](long a, long b, long c) {
                const float* pa = P(a); const float* pb = P(b); const float* pc = P(c);
                out.tri_xyz9.insert(out.tri_xyz9.end(), pa, pa + 3);
                out.tri_xyz9.insert(out.tri_xyz9.end(), pb, pb + 3);
                out.tri_xyz9.insert(out.tri_xyz9.end(), pc, pc + 3);
                out.tri_material.push_back(cur_mat + 1);          // +1: slot 0 is the default material (utils.cpp:57-60)
            };
            if (poly.size() == 3) emit(poly[0], poly[1], poly[2]);
            else if (poly.size() == 4)
            {
                const float *p0 = P(poly[0]), *p1 = P(poly[1]), *p2 = P(poly[2]), *p3 = P(poly[3]);
                const float ax = p0[0] - p2[0], ay = p0[1] - p2[1], az = p0[2] - p2[2];
                const float bx = p1[0] - p3[0], by = p1[1] - p3[1], bz = p1[2] - p3[2];
                const float d02 = ax * ax + ay * ay + az * az, d13 = bx * bx + by * by + bz * bz;
                if (d02 < d13) { emit(poly[0], poly[1], poly[2]); emit(poly[0], poly[2], poly[3]); }
                else { emit(poly[0], poly[1], poly[3]); emit(poly[1], poly[2], poly[3]); }
            }
            else
                for (size_t k = 1; k + 1 < poly.size(); k++) emit(poly[0], poly[k], poly[k + 1]);
        }
        else if (le - p > 7 && std::strncmp(p, "usemtl", 6) == 0 && (p[6] == ' ' || p[6] == '\t'))
        {
            const std::string name = rest_of_line(p + 6, end);
            auto it = mat_index.find(name);
            if (it == mat_index.end()) { err = "material '" + name + "' is not defined by any mtllib"; return false; }
            cur_mat = it->second;
        }
        else if (le - p > 7 && std::strncmp(p, "mtllib", 6) == 0 && (p[6] == ' ' || p[6] == '\t'))
        {
            const std::string name = rest_of_line(p + 6, end);
            if (!parse_mtl(dir + "/" + name, mats, mat_index, err)) return false;      // mandatory, like MaterialLibrary::Default()
        }
        p = le < end ? le + 1 : end;
    }

    // SimpleMaterial records: emission rgba, diffuse rgba, metalness, roughness (simple_material.h:6-13; Color's alpha defaults to 1)
    out.materials10.assign({ 1.0f, 0.0f, 1.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 1.0f });
    for (const Mtl& m : mats)
    {
        float rough = std::max(1.0e-2f, m.roughness), metal = m.metallic;
        if (m.illum == 0) { rough = 1.0f; metal = 0.0f; }
        const float rec[10] = { m.ke[0], m.ke[1], m.ke[2], 1.0f, m.kd[0], m.kd[1], m.kd[2], 1.0f, metal, rough };
        out.materials10.insert(out.materials10.end(), rec, rec + 10);
    }
    for (size_t t = 0; t < out.tri_material.size(); t++)
    {
        const int mi = out.tri_material[t] - 1;
        if (mi >= 0 && (mats[(size_t)mi].ke[0] > 0 || mats[(size_t)mi].ke[1] > 0 || mats[(size_t)mi].ke[2] > 0)) out.emissive_tri.push_back((int)t);
    }
    return true;
}

} // namespace b200rt
