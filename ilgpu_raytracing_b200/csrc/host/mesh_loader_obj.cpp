// mesh_loader_obj.cpp — the asset path of the reference, mirrored for the C++ engine:
//   MeshLoaderOBJ.Load      Engine/MeshLoaderOBJ.cs:67-254   (OBJ faces / usemtl / mtllib, fan triangulation, texture binding)
//   LoadMtl                 Engine/MeshLoaderOBJ.cs:319-441  (newmtl, Kd, map_Kd, map_d, d, Tr, Ni, illum)
//   LoadTgaBGRA             Engine/MeshLoaderOBJ.cs:504-593  (TGA types 2 / 3 / 10, 8 / 24 / 32 bpp, both origins)
//   Scene.LoadObjInstance   Engine/Scene.cs:144-256          (append to the scene lists; every material flattens its own copy of its textures)
//
// What is NOT here: the reference decodes every non-TGA image through System.Drawing.Bitmap (MeshLoaderOBJ.cs:466-482), a
// platform library that is not part of the repository.  The lossless formats are decoded natively - uncompressed 24 / 32-bit BMP
// here, 8-bit PNG of every colour type in png_decode.cpp; anything else (JPEG, GIF, TIFF: lossy or platform-defined decoders)
// raises InvalidDataException instead of guessing.  Console diagnostics are omitted.
#include <cctype>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <unordered_map>

#include "engine.h"

namespace ILGPU_Raytracing {
namespace Engine {

namespace {

inline Float3 F3(float x, float y, float z) { Float3 r; r.X = x; r.Y = y; r.Z = z; return r; }

bool file_exists(const std::string& p) { std::ifstream f(p, std::ios::binary); return f.good(); }

// Path.Combine(baseDir, rel): a rooted second part wins; no separator normalisation (a Windows-style "textures\\a.tga" stays one name)
std::string path_combine(const std::string& a, const std::string& b) {
    if (b.empty()) return a;
    if (b[0] == '/') return b;
    if (a.empty()) return b;
    return (a.back() == '/') ? a + b : a + "/" + b;
}
std::string dir_of(const std::string& path) {
    size_t k = path.find_last_of('/');
    if (k == std::string::npos) return ".";
    return k == 0 ? "/" : path.substr(0, k);
}
std::string trim(const std::string& s) {   // string.Trim(): white space at both ends
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) a++;
    while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}
bool starts_with(const std::string& s, const char* p) { return s.compare(0, strlen(p), p) == 0; }
std::string lower(std::string s) { for (auto& c : s) c = (char)std::tolower((unsigned char)c); return s; }

// float.Parse / int.Parse with CultureInfo.InvariantCulture: the whole token must be a number
float parse_float(const std::string& tok) {
    std::string t = trim(tok);
    if (t.empty()) throw FormatException("empty number");
    char* end = nullptr; errno = 0;
    float v = strtof(t.c_str(), &end);
    if (end != t.c_str() + t.size()) throw FormatException("not a number: '" + t + "'");
    return v;
}
int parse_int(const std::string& tok) {
    std::string t = trim(tok);
    if (t.empty()) throw FormatException("empty integer");
    char* end = nullptr; errno = 0;
    long v = strtol(t.c_str(), &end, 10);
    if (end != t.c_str() + t.size()) throw FormatException("not an integer: '" + t + "'");
    return (int)v;
}
// Parse3 / Parse2 (:275-293): fields separated by ' ' only
void parse_fields(const std::string& s, float* out, int n) {
    size_t i = 0;
    for (int k = 0; k < n; k++) {
        while (i < s.size() && s[i] == ' ') i++;
        size_t j = i; while (j < s.size() && s[j] != ' ') j++;
        out[k] = parse_float(s.substr(i, j - i));
        i = j;
    }
}
int parse_one_index(const std::string& s, int countSoFar) {   // :313-317
    int val = parse_int(s);
    return val > 0 ? (val - 1) : (countSoFar + val);
}
void parse_face_vvt(const std::string& tok, int vCount, int tCount, int* v, int* t) {   // :295-311
    size_t s1 = tok.find('/');
    if (s1 == std::string::npos) { *v = parse_one_index(tok, vCount); *t = 0; return; }
    *v = parse_one_index(tok.substr(0, s1), vCount);
    std::string rest = tok.substr(s1 + 1);
    size_t s2 = rest.find('/');
    if (s2 == std::string::npos) *t = parse_one_index(rest, tCount);
    else { std::string vt = rest.substr(0, s2); *t = !vt.empty() ? parse_one_index(vt, tCount) : 0; }
}
// StreamReader.ReadLine: lines end at \n, \r or \r\n
bool read_line(std::istream& in, std::string& line) {
    line.clear();
    int c = in.get();
    if (c == EOF) return false;
    while (c != EOF) {
        if (c == '\n') break;
        if (c == '\r') { if (in.peek() == '\n') in.get(); break; }
        line.push_back((char)c);
        c = in.get();
    }
    return true;
}

MaterialRecord default_material() {   // :262-273
    MaterialRecord m; m.Kd = F3(0.8f, 0.8f, 0.8f); m.HasDiffuseMap = 0; m.DiffuseTexIndex = -1; m.Shading = RT_SHADING_LAMBERT; m.IOR = 1.0f;
    m.HasAlphaMap = 0; m.AlphaTexIndex = -1; m.TwoSided = 0; m.AlphaCutoff = 0.5f;
    return m;
}

// Dictionary<string, T> as the loader uses it: enumeration in first-insertion order, assignment to an existing key keeps its place
template <typename T> struct OrderedMap {
    std::vector<std::pair<std::string, T>> items;
    std::unordered_map<std::string, size_t> index;
    bool ignoreCase = false;
    std::string key(const std::string& k) const { return ignoreCase ? lower(k) : k; }
    T* find(const std::string& k) { auto it = index.find(key(k)); return it == index.end() ? nullptr : &items[it->second].second; }
    void set(const std::string& k, const T& v) {
        auto it = index.find(key(k));
        if (it == index.end()) { index[key(k)] = items.size(); items.push_back({k, v}); } else items[it->second].second = v;
    }
};

void load_mtl(const std::string& mtlPath, const std::string& baseDir, OrderedMap<MaterialRecord>& dict, OrderedMap<std::string>& diffuse, OrderedMap<std::string>& alpha) {   // :319-441
    std::ifstream in(mtlPath, std::ios::binary);
    if (!in) throw FileNotFoundException("MTL file not found: " + mtlPath);
    bool have = false; std::string cur; MaterialRecord m = default_material();
    std::string line;
    while (read_line(in, line)) {
        if (line.empty() || line[0] == '#') continue;
        if (starts_with(line, "newmtl ")) {
            if (have) dict.set(cur, m);
            cur = trim(line.substr(7)); have = true;
            m = default_material();
        } else if (starts_with(line, "Kd ")) {
            float f[3]; parse_fields(trim(line.substr(3)), f, 3); m.Kd = F3(f[0], f[1], f[2]);
        } else if (starts_with(line, "map_Kd ")) {
            std::string raw = trim(line.substr(7));
            if (have) diffuse.set(cur, path_combine(baseDir, raw));
            m.HasDiffuseMap = 1;
        } else if (starts_with(line, "map_d ")) {
            std::string raw = trim(line.substr(6));
            if (have) alpha.set(cur, path_combine(baseDir, raw));
            m.HasAlphaMap = 1; m.TwoSided = 1;
        } else if (starts_with(line, "d ")) {
            float d = parse_float(line.substr(2));
            if (d < 0.999f) { m.TwoSided = 1; m.AlphaCutoff = 0.5f; }
        } else if (starts_with(line, "Tr ")) {
            float tr = parse_float(line.substr(3));
            float d = 1.0f - tr;
            if (d < 0.999f) { m.TwoSided = 1; m.AlphaCutoff = 0.5f; }
        } else if (starts_with(line, "Ni ")) {
            std::string s = trim(line.substr(3));
            size_t i = 0; while (i < s.size() && s[i] == ' ') i++;
            size_t j = i; while (j < s.size() && s[j] != ' ') j++;
            m.IOR = parse_float(s.substr(i, j - i));
            if (m.IOR <= 0.0f) m.IOR = 1.0f;
        } else if (starts_with(line, "illum ")) {
            int model = parse_int(line.substr(6));
            if (model >= 5) m.Shading = RT_SHADING_GLASS; else if (model >= 3) m.Shading = RT_SHADING_MIRROR; else m.Shading = RT_SHADING_LAMBERT;
        }
    }
    if (have) dict.set(cur, m);
}

struct ByteReader {
    const std::vector<unsigned char>& d; size_t p = 0;
    unsigned char u8() { if (p >= d.size()) throw EndOfStreamException("unexpected end of image file"); return d[p++]; }
    unsigned u16() { unsigned a = u8(), b = u8(); return a | (b << 8); }
    unsigned u32() { unsigned a = u16(), b = u16(); return a | (b << 16); }
};
std::vector<unsigned char> read_all(const std::string& file) {
    std::ifstream f(file, std::ios::binary);
    return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

TextureSrc load_tga_bgra(const std::string& file) {   // :504-593
    std::vector<unsigned char> bytes = read_all(file);
    ByteReader br{bytes};
    unsigned idLength = br.u8(), colorMapType = br.u8(), imageType = br.u8();
    br.u16(); br.u16(); br.u8();   // colour-map spec
    br.u16(); br.u16();            // origin
    int w = (int)br.u16(), h = (int)br.u16();
    unsigned pixelDepth = br.u8(), imageDesc = br.u8();
    for (unsigned i = 0; i < idLength; i++) br.u8();
    if (colorMapType != 0) throw InvalidDataException("TGA colorMapType not supported: " + file);
    bool topOrigin = (imageDesc & 0x20) != 0;
    int bpp = pixelDepth == 32 ? 4 : (pixelDepth == 24 ? 3 : (pixelDepth == 8 ? 1 : 0));
    if (bpp == 0) throw InvalidDataException("TGA pixelDepth not supported: " + file);
    TextureSrc tex; tex.Path = file; tex.Width = w; tex.Height = h; tex.BGRA.assign((size_t)w * h * 4, 0);
    auto writePixel = [&](int i, unsigned char b, unsigned char g, unsigned char r, unsigned char a) {
        int px = i % w, py = i / w;
        int yOut = topOrigin ? py : (h - 1 - py);
        size_t dst = ((size_t)yOut * w + px) * 4;
        tex.BGRA[dst] = b; tex.BGRA[dst + 1] = g; tex.BGRA[dst + 2] = r; tex.BGRA[dst + 3] = a;
    };
    auto readPixel = [&](unsigned char* q) {
        if (bpp == 4) { q[0] = br.u8(); q[1] = br.u8(); q[2] = br.u8(); q[3] = br.u8(); }
        else if (bpp == 3) { q[0] = br.u8(); q[1] = br.u8(); q[2] = br.u8(); q[3] = 255; }
        else { unsigned char y8 = br.u8(); q[0] = q[1] = q[2] = y8; q[3] = 255; }
    };
    int total = w * h;
    if (imageType == 2 || imageType == 3) {
        for (int i = 0; i < total; i++) { unsigned char q[4]; readPixel(q); writePixel(i, q[0], q[1], q[2], q[3]); }
    } else if (imageType == 10) {
        int i = 0;
        while (i < total) {
            unsigned packet = br.u8();
            int count = (int)(packet & 0x7F) + 1;
            if (packet & 0x80) { unsigned char q[4]; readPixel(q); for (int k = 0; k < count && i < total; k++, i++) writePixel(i, q[0], q[1], q[2], q[3]); }
            else for (int k = 0; k < count && i < total; k++, i++) { unsigned char q[4]; readPixel(q); writePixel(i, q[0], q[1], q[2], q[3]); }
        }
    } else throw InvalidDataException("TGA imageType not supported: " + file);
    return tex;
}

// What `new Bitmap(file)` + ExtractBitmapBGRA (:466-502) yields for an uncompressed 24 / 32-bit BMP: rows top-down, B G R A bytes
// (24-bit sources get alpha 255, as the Format32bppArgb conversion gives them).
TextureSrc load_bmp_bgra(const std::string& file) {
    std::vector<unsigned char> bytes = read_all(file);
    ByteReader br{bytes};
    if (br.u8() != 'B' || br.u8() != 'M') throw InvalidDataException("not a BMP file: " + file);
    br.u32(); br.u32();
    unsigned dataOffset = br.u32(), hdrSize = br.u32();
    if (hdrSize < 40) throw InvalidDataException("BMP header not supported: " + file);
    int w = (int)br.u32(); int hRaw = (int)br.u32();
    br.u16(); unsigned bits = br.u16(), compression = br.u32();
    if ((bits != 24 && bits != 32) || (compression != 0 && !(bits == 32 && compression == 3))) throw InvalidDataException("only uncompressed 24 / 32-bit BMP is decoded natively: " + file);
    bool bottomUp = hRaw > 0; int h = bottomUp ? hRaw : -hRaw;
    if (w <= 0 || h <= 0) throw InvalidDataException("bad BMP size: " + file);
    size_t stride = ((size_t)w * (bits / 8) + 3) & ~(size_t)3;
    if ((size_t)dataOffset + stride * h > bytes.size()) throw EndOfStreamException("BMP pixel data truncated: " + file);
    TextureSrc tex; tex.Path = file; tex.Width = w; tex.Height = h; tex.BGRA.assign((size_t)w * h * 4, 0);
    for (int y = 0; y < h; y++) {
        const unsigned char* row = bytes.data() + dataOffset + stride * (size_t)(bottomUp ? (h - 1 - y) : y);
        for (int x = 0; x < w; x++) {
            const unsigned char* p = row + (size_t)x * (bits / 8);
            unsigned char* q = &tex.BGRA[((size_t)y * w + x) * 4];
            q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; q[3] = bits == 32 ? p[3] : 255;
        }
    }
    return tex;
}

bool try_load_texture_bgra(const std::string& file, TextureSrc* tex) {   // :443-482
    if (!file_exists(file)) return false;
    size_t dot = file.find_last_of('.');
    std::string ext = dot == std::string::npos ? "" : lower(file.substr(dot));
    if (ext == ".tga") *tex = load_tga_bgra(file);
    else if (ext == ".bmp") *tex = load_bmp_bgra(file);
    else if (ext == ".png") *tex = load_png_bgra(file, read_all(file));
    else throw InvalidDataException("texture '" + file + "': the reference decodes this format through System.Drawing.Bitmap, which this build does not have; use PNG, TGA or BMP");
    return true;
}

}   // namespace

MeshHost MeshLoaderOBJ::Load(const std::string& path, float scale, bool flipWinding) {   // :67-254
    std::ifstream in(path, std::ios::binary);
    if (!in) throw FileNotFoundException("OBJ file not found: " + path);
    std::string baseDir = dir_of(path);
    MeshHost mesh;
    std::vector<Float3> tempPositions; std::vector<Float2> tempTex; std::vector<int> faceV, faceT;
    std::string mtlLibPath; int currentMtl = -1;
    OrderedMap<int> mtlNameToIndex;
    std::string line;
    while (read_line(in, line)) {
        if (line.empty() || line[0] == '#') continue;
        if (starts_with(line, "v ")) {
            float f[3]; parse_fields(trim(line.substr(2)), f, 3);
            tempPositions.push_back(F3(f[0] * scale, f[1] * scale, f[2] * scale));
        } else if (starts_with(line, "vt ")) {
            float f[2]; parse_fields(trim(line.substr(3)), f, 2);
            Float2 t; t.X = f[0]; t.Y = f[1]; tempTex.push_back(t);
        } else if (starts_with(line, "f ")) {
            faceV.clear(); faceT.clear();
            std::string s = trim(line.substr(2));
            size_t i = 0;
            while (i < s.size()) {
                while (i < s.size() && s[i] == ' ') i++;
                if (i >= s.size()) break;
                size_t j = i; while (j < s.size() && s[j] != ' ') j++;
                std::string tok = s.substr(i, j - i);
                if (!tok.empty()) { int v, t; parse_face_vvt(tok, (int)tempPositions.size(), (int)tempTex.size(), &v, &t); faceV.push_back(v); faceT.push_back(t); }
                i = j + 1;
            }
            if (faceV.size() >= 3)
                for (size_t k = 1; k + 1 < faceV.size(); k++) {
                    MeshTri tri; MeshTriUV tuv;
                    if (!flipWinding) { tri.i0 = faceV[0]; tri.i1 = faceV[k]; tri.i2 = faceV[k + 1]; tuv.t0 = faceT[0]; tuv.t1 = faceT[k]; tuv.t2 = faceT[k + 1]; }
                    else { tri.i0 = faceV[0]; tri.i1 = faceV[k + 1]; tri.i2 = faceV[k]; tuv.t0 = faceT[0]; tuv.t1 = faceT[k + 1]; tuv.t2 = faceT[k]; }
                    mesh.Triangles.push_back(tri); mesh.TriUVs.push_back(tuv);
                    mesh.TriMaterialIndex.push_back(currentMtl < 0 ? 0 : currentMtl);
                }
        } else if (starts_with(line, "mtllib ")) {
            std::string rel = trim(line.substr(7));
            if (!rel.empty()) mtlLibPath = path_combine(baseDir, rel);
        } else if (starts_with(line, "usemtl ")) {
            std::string name = trim(line.substr(7));
            if (!name.empty()) {
                if (int* idx = mtlNameToIndex.find(name)) currentMtl = *idx;
                else { currentMtl = (int)mesh.Materials.size(); mtlNameToIndex.set(name, currentMtl); mesh.Materials.push_back(default_material()); }
            }
        }
    }
    mesh.Positions = tempPositions; mesh.Texcoords = tempTex;

    // load and merge the MTL materials (:176-198)
    std::vector<std::pair<int, std::string>> materialTexPath, alphaTexPath;   // Dictionary<int,string>, insertion order
    auto setIdx = [](std::vector<std::pair<int, std::string>>& d, int k, const std::string& v) { for (auto& e : d) if (e.first == k) { e.second = v; return; } d.push_back({k, v}); };
    if (!trim(mtlLibPath).empty() && file_exists(mtlLibPath)) {
        OrderedMap<MaterialRecord> loaded; OrderedMap<std::string> diffuseMap, alphaMap;
        load_mtl(mtlLibPath, baseDir, loaded, diffuseMap, alphaMap);
        for (auto& kv : loaded.items) {
            if (int* idx = mtlNameToIndex.find(kv.first)) mesh.Materials[(size_t)*idx] = kv.second;
            else { int idx2 = (int)mesh.Materials.size(); mtlNameToIndex.set(kv.first, idx2); mesh.Materials.push_back(kv.second); }
        }
        for (auto& kv : diffuseMap.items) if (int* mi = mtlNameToIndex.find(kv.first)) setIdx(materialTexPath, *mi, kv.second);
        for (auto& kv : alphaMap.items) if (int* mi = mtlNameToIndex.find(kv.first)) setIdx(alphaTexPath, *mi, kv.second);
    }
    // bind textures (:200-251): one load per distinct path (case-insensitive), missing files clear the flags
    OrderedMap<int> texPathToIndex; texPathToIndex.ignoreCase = true;
    for (auto& kv : materialTexPath) {
        int matIndex = kv.first; const std::string& p = kv.second; int texIndex;
        if (int* ti = texPathToIndex.find(p)) texIndex = *ti;
        else {
            TextureSrc tex;
            if (!try_load_texture_bgra(p, &tex)) { mesh.Materials[(size_t)matIndex].HasDiffuseMap = 0; mesh.Materials[(size_t)matIndex].DiffuseTexIndex = -1; continue; }
            texIndex = (int)mesh.Textures.size(); mesh.Textures.push_back(tex); texPathToIndex.set(p, texIndex);
        }
        mesh.Materials[(size_t)matIndex].HasDiffuseMap = 1; mesh.Materials[(size_t)matIndex].DiffuseTexIndex = texIndex;
    }
    for (auto& kv : alphaTexPath) {
        int matIndex = kv.first; const std::string& p = kv.second; int texIndex;
        if (int* ti = texPathToIndex.find(p)) texIndex = *ti;
        else {
            TextureSrc tex;
            if (!try_load_texture_bgra(p, &tex)) { mesh.Materials[(size_t)matIndex].HasAlphaMap = 0; mesh.Materials[(size_t)matIndex].AlphaTexIndex = -1; continue; }
            texIndex = (int)mesh.Textures.size(); mesh.Textures.push_back(tex); texPathToIndex.set(p, texIndex);
        }
        mesh.Materials[(size_t)matIndex].HasAlphaMap = 1; mesh.Materials[(size_t)matIndex].AlphaTexIndex = texIndex; mesh.Materials[(size_t)matIndex].TwoSided = 1;
    }
    return mesh;
}

void Scene::LoadObjInstance(const std::string& objPath, const Affine3x4& objectToWorld, float uniformScale) {   // Scene.cs:144-256
    if (trim(objPath).empty() || !file_exists(objPath)) throw FileNotFoundException("OBJ file not found: " + objPath);
    MeshHost mesh = MeshLoaderOBJ::Load(objPath, uniformScale, /*flipWinding*/ false);
    if (mesh.Triangles.empty() || mesh.Positions.empty()) throw InvalidOperationException("OBJ has no triangles: " + objPath);
    // the reference indexes _hMaterials[baseMat + 0] for faces without usemtl even when the file defines no material at all
    // (an out-of-range read on its device); refuse instead
    if (mesh.Materials.empty()) throw InvalidOperationException("OBJ defines no material (no usemtl / mtllib): " + objPath);
    if (mesh.Texcoords.empty()) throw InvalidOperationException("OBJ has no texture coordinates; faces index vt 0 (MeshLoaderOBJ.cs:297): " + objPath);
    // materials: each one flattens its own copy of its textures into the texel array, diffuse before alpha (:182-228)
    std::vector<MaterialRecord> remapped;
    for (MaterialRecord m : mesh.Materials) {
        auto flatten = [&](const TextureSrc& src) {
            std::vector<RGBA32> px((size_t)src.Width * src.Height);
            for (size_t p = 0; p < px.size(); p++) { px[p].B = src.BGRA[4 * p]; px[p].G = src.BGRA[4 * p + 1]; px[p].R = src.BGRA[4 * p + 2]; px[p].A = src.BGRA[4 * p + 3]; }
            return AddTexture(src.Width, src.Height, px.data());
        };
        if (m.HasDiffuseMap != 0 && m.DiffuseTexIndex >= 0 && m.DiffuseTexIndex < (int)mesh.Textures.size()) { m.DiffuseTexIndex = flatten(mesh.Textures[(size_t)m.DiffuseTexIndex]); m.HasDiffuseMap = 1; }
        else { m.HasDiffuseMap = 0; m.DiffuseTexIndex = -1; }
        if (m.HasAlphaMap != 0 && m.AlphaTexIndex >= 0 && m.AlphaTexIndex < (int)mesh.Textures.size()) { m.AlphaTexIndex = flatten(mesh.Textures[(size_t)m.AlphaTexIndex]); m.HasAlphaMap = 1; }
        else { m.HasAlphaMap = 0; m.AlphaTexIndex = -1; }
        remapped.push_back(m);
    }
    LoadMeshInstance(mesh.Positions.data(), (int)mesh.Positions.size(), mesh.Triangles.data(), (int)mesh.Triangles.size(), mesh.Texcoords.data(), (int)mesh.Texcoords.size(),
                     mesh.TriUVs.data(), mesh.TriMaterialIndex.data(), remapped.data(), (int)remapped.size(), objectToWorld);
}

}   // namespace Engine
}   // namespace ILGPU_Raytracing
