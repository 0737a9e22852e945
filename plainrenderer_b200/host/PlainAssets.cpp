// PlainAssets.cpp - loaders for what the reference's asset pipeline writes, behind include/plain_assets.h:
//   .plain scenes   Plain/src/Common/ModelLoadSaveBinary.cpp:40-231 (layout), Scene.h:6-9 (ObjectBinary),
//                   MeshProcessing.cpp:20-120 (index width, 28-byte vertices with the float position first)
//   R16F 3-D .dds   ImageIO.cpp:342-431 (reader), :433-520 (writer): "DDS " + 124-byte header + 20-byte DX10 header
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include "SdfBakeCommon.h"
#include "plain_assets.h"

namespace {
thread_local std::string g_error;
int failAsset(const std::string& m) { g_error = m; return 1; }

struct Mesh {
    plain_mesh_info info{};
    std::vector<uint32_t> indices;
    std::vector<float> positions;
    std::vector<unsigned char> vertices;  // the 28-byte vertices as stored (MeshProcessing.cpp:52-106)
};
bool readAll(const char* path, std::vector<unsigned char>& out) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) return false;
    const std::streamsize n = f.tellg();
    f.seekg(0);
    out.resize((size_t)n);
    return n == 0 || (bool)f.read((char*)out.data(), n);
}
const uint32_t kVertexBytes = 28;            // VertexInput.h:27-31: position 12, uv 4, normal 4, tangent 4, bitangent 4
const uint32_t kPlainMagic = 0x424d6c50u;    // "PlMB"
const uint32_t kDdsMagic = 0x20534444u;      // "DDS "
const uint32_t kDxgiR16Float = 54;
}  // namespace

struct plain_scene {
    std::vector<plain_scene_object> objects;
    std::vector<Mesh> meshes;
};

extern "C" {

const char* PLAIN_ASSET(last_error)(void) { return g_error.c_str(); }

int PLAIN_ASSET(scene_load)(const char* path, plain_scene** out) {
    std::vector<unsigned char> d;
    if (!readAll(path, d)) return failAsset(std::string("scene_load: cannot read ") + path);
    size_t p = 0;
    auto need = [&](size_t n) { return p + n <= d.size(); };
    auto rd = [&](void* dst, size_t n) { std::memcpy(dst, d.data() + p, n); p += n; };
    // ModelFileHeader {u32 magic; size_t objectCount; size_t meshCount}: 24 bytes with the padding after the magic number
    if (!need(24)) return failAsset("scene_load: truncated header");
    uint32_t magic;
    uint64_t objectCount, meshCount;
    rd(&magic, 4); p += 4; rd(&objectCount, 8); rd(&meshCount, 8);
    if (magic != kPlainMagic) return failAsset("scene_load: not a .plain file (magic number)");
    if (objectCount > d.size() / 72 || meshCount > d.size() / 48) return failAsset("scene_load: implausible object/mesh count");
    plain_scene* s = new plain_scene();
    for (uint64_t i = 0; i < objectCount; i++) {  // ObjectBinary {mat4; size_t meshIndex}: 72 bytes
        if (!need(72)) { delete s; return failAsset("scene_load: truncated object table"); }
        plain_scene_object o;
        rd(o.model_matrix, 64); rd(&o.mesh_index, 8);
        s->objects.push_back(o);
    }
    for (uint64_t m = 0; m < meshCount; m++) {
        Mesh mesh;
        if (!need(32)) { delete s; return failAsset("scene_load: truncated mesh header"); }
        rd(&mesh.info.index_count, 4); rd(&mesh.info.vertex_count, 4);
        rd(mesh.info.bb_min, 12); rd(mesh.info.bb_max, 12);  // AxisAlignedBoundingBox {vec3 min; vec3 max}, AABB.h:4-7
        char* paths[4] = {mesh.info.albedo_path, mesh.info.normal_path, mesh.info.specular_path, mesh.info.sdf_path};
        for (int k = 0; k < 4; k++) {
            uint32_t len;
            if (!need(4)) { delete s; return failAsset("scene_load: truncated path"); }
            rd(&len, 4);
            if (!need(len)) { delete s; return failAsset("scene_load: truncated path"); }
            if (len > 255) { delete s; return failAsset("scene_load: a texture / SDF path is longer than 255 bytes (plain_mesh_info holds 256)"); }
            const size_t n = len;
            std::memcpy(paths[k], d.data() + p, n);
            paths[k][n] = 0;
            p += len;
        }
        if (!need(12)) { delete s; return failAsset("scene_load: truncated mesh"); }
        rd(mesh.info.mean_albedo, 12);
        const bool wide = !(mesh.info.index_count < 65535u);  // MeshProcessing.cpp:28: 16-bit indices when indexCount < uint16 max
        const size_t indexBytes = (size_t)mesh.info.index_count * (wide ? 4 : 2), vertexBytes = (size_t)mesh.info.vertex_count * kVertexBytes;
        if (!need(indexBytes + vertexBytes)) { delete s; return failAsset("scene_load: truncated index/vertex data"); }
        mesh.indices.resize(mesh.info.index_count);
        for (uint32_t i = 0; i < mesh.info.index_count; i++) {
            if (wide) { uint32_t v; std::memcpy(&v, d.data() + p + 4 * (size_t)i, 4); mesh.indices[i] = v; }
            else { uint16_t v; std::memcpy(&v, d.data() + p + 2 * (size_t)i, 2); mesh.indices[i] = v; }
        }
        p += indexBytes;
        mesh.positions.resize((size_t)mesh.info.vertex_count * 3);
        for (uint32_t i = 0; i < mesh.info.vertex_count; i++) std::memcpy(&mesh.positions[3 * (size_t)i], d.data() + p + (size_t)i * kVertexBytes, 12);
        mesh.vertices.assign(d.data() + p, d.data() + p + vertexBytes);
        p += vertexBytes;
        s->meshes.push_back(std::move(mesh));
    }
    if (p != d.size()) { delete s; return failAsset("scene_load: trailing bytes after the last mesh"); }
    *out = s;
    return 0;
}
void PLAIN_ASSET(scene_destroy)(plain_scene* s) { delete s; }
int PLAIN_ASSET(scene_counts)(const plain_scene* s, uint64_t* objects, uint64_t* meshes) {
    *objects = s->objects.size(); *meshes = s->meshes.size();
    return 0;
}
int PLAIN_ASSET(scene_object)(const plain_scene* s, uint64_t i, plain_scene_object* out) {
    if (i >= s->objects.size()) return failAsset("scene_object: index out of range");
    *out = s->objects[i];
    return 0;
}
int PLAIN_ASSET(scene_mesh_info)(const plain_scene* s, uint64_t m, plain_mesh_info* out) {
    if (m >= s->meshes.size()) return failAsset("scene_mesh_info: index out of range");
    *out = s->meshes[m].info;
    return 0;
}
int PLAIN_ASSET(scene_mesh_geometry)(const plain_scene* s, uint64_t m, float* positions, uint32_t* indices) {
    if (m >= s->meshes.size()) return failAsset("scene_mesh_geometry: index out of range");
    const Mesh& mesh = s->meshes[m];
    if (positions) std::memcpy(positions, mesh.positions.data(), mesh.positions.size() * 4);
    if (indices) std::memcpy(indices, mesh.indices.data(), mesh.indices.size() * 4);
    return 0;
}

int PLAIN_ASSET(scene_mesh_vertices)(const plain_scene* s, uint64_t m, void* vertices) {
    if (m >= s->meshes.size()) return failAsset("scene_mesh_vertices: index out of range");
    std::memcpy(vertices, s->meshes[m].vertices.data(), s->meshes[m].vertices.size());
    return 0;
}

// ---- DDS: header offsets of the fields used (standard DDS_HEADER after the 4-byte magic) ----
static int ddsHeader(const std::vector<unsigned char>& d, uint32_t extent[3]) {
    if (d.size() < 148) return failAsset("dds: file shorter than the DDS + DX10 headers");
    uint32_t magic, fourCC, dxgi;
    std::memcpy(&magic, d.data(), 4);
    std::memcpy(&extent[1], d.data() + 12, 4);  // height
    std::memcpy(&extent[0], d.data() + 16, 4);  // width
    std::memcpy(&extent[2], d.data() + 24, 4);  // depth
    std::memcpy(&fourCC, d.data() + 84, 4);
    std::memcpy(&dxgi, d.data() + 128, 4);
    if (magic != kDdsMagic) return failAsset("dds: magic number");
    if (fourCC != 0x30315844u || dxgi != kDxgiR16Float) return failAsset("dds: only DX10 / DXGI_FORMAT_R16_FLOAT bricks are supported");
    if (extent[2] < 1) extent[2] = 1;
    if (d.size() != 148 + (size_t)extent[0] * extent[1] * extent[2] * 2) return failAsset("dds: payload size does not match the extent");
    return 0;
}
int PLAIN_ASSET(dds_r16f_info)(const char* path, uint32_t extent[3]) {
    std::vector<unsigned char> d;
    if (!readAll(path, d)) return failAsset(std::string("dds: cannot read ") + path);
    return ddsHeader(d, extent);
}
int PLAIN_ASSET(dds_r16f_load)(const char* path, uint16_t* out, size_t capacity) {
    std::vector<unsigned char> d;
    uint32_t e[3];
    if (!readAll(path, d)) return failAsset(std::string("dds: cannot read ") + path);
    if (ddsHeader(d, e)) return 1;
    const size_t n = (size_t)e[0] * e[1] * e[2];
    if (capacity < n) return failAsset("dds: output buffer too small");
    std::memcpy(out, d.data() + 148, n * 2);
    return 0;
}
int PLAIN_ASSET(dds_r16f_save)(const char* path, const uint32_t extent[3], const uint16_t* texels) {
    unsigned char h[148];
    std::memset(h, 0, sizeof(h));
    auto put = [&](size_t off, uint32_t v) { std::memcpy(h + off, &v, 4); };
    put(0, kDdsMagic);
    put(4, 124);                                  // dwSize
    put(8, 0x1 | 0x2 | 0x4 | 0x1000 | 0x800000);  // caps | height | width | pixelformat | depth
    put(12, extent[1]); put(16, extent[0]);
    put(20, extent[0] * 2);                       // pitch
    put(24, extent[2]);
    put(28, 1);                                   // mip count
    put(76, 32); put(80, 0x4);                    // pixel format: size, DDPF_FOURCC
    put(84, 0x30315844u);                         // "DX10"
    put(108, 0x1000);                             // DDSCAPS_TEXTURE
    put(112, 0x200000);                           // DDSCAPS2_VOLUME
    put(128, kDxgiR16Float);
    put(132, 4);                                  // D3D10_RESOURCE_DIMENSION_TEXTURE3D
    put(140, 1);                                  // array size
    std::ofstream f(path, std::ios::binary);
    if (!f) return failAsset(std::string("dds: cannot write ") + path);
    f.write((const char*)h, sizeof(h));
    f.write((const char*)texels, (std::streamsize)((size_t)extent[0] * extent[1] * extent[2] * 2));
    return f ? 0 : failAsset("dds: write failed");
}

void PLAIN_ASSET(sdf_resolution)(const float bbMin[3], const float bbMax[3], uint32_t out[3]) { sdfbake::brickResolution(bbMin, bbMax, out); }

}  // extern "C"
