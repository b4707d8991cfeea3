// vtrace_host.hpp — C++ mirror of the reference's Rust host side above the C ABI, for building
// compiled (non-Python) hosts against librender when rustc is not available:
//   Renderer, TextureUploadQueue, GPUInstance   <- src/render.rs
//   SceneGraph + depth-first flattening         <- src/scene.rs:18-181
//   RawDynamicChunk, Color, load_magica_voxel   <- src/voxel/{common,rawchunk,magica_voxel}.rs
//   perspective / look_at / translate / scale   <- glm-rs 0.2.3 call sites (src/render.rs:190,208-214)
// Same names, argument meaning and failure behaviour (where Rust panics, this throws).
// Header-only; link with -lrender (include/vtrace_abi.h).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../include/vtrace_abi.h"

namespace vtrace {

// ---- glm-rs style column-major matrices ---------------------------------------------------
struct Vec3 { float x, y, z; };
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline Vec3 normalize(Vec3 a) { float l = std::sqrt(dot(a, a)); return {a.x / l, a.y / l, a.z / l}; }

struct Mat4 {
    float c[4][4]; // c[col][row]
    static Mat4 identity() { Mat4 m{}; for (int i = 0; i < 4; ++i) m.c[i][i] = 1.0f; return m; }
};
inline Mat4 operator*(const Mat4& a, const Mat4& b) {
    Mat4 r{};
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i)
            r.c[j][i] = a.c[0][i] * b.c[j][0] + a.c[1][i] * b.c[j][1] + a.c[2][i] * b.c[j][2] + a.c[3][i] * b.c[j][3];
    return r;
}
inline Mat4 perspective(float fovy, float aspect, float near_, float far_) { // glm::ext::perspective (RH, -1..1, no Y flip)
    const float ys = 1.0f / std::tan(fovy / 2.0f), xs = ys / aspect;
    Mat4 m{};
    m.c[0][0] = xs; m.c[1][1] = ys;
    m.c[2][2] = (far_ + near_) / (near_ - far_); m.c[2][3] = -1.0f;
    m.c[3][2] = (2.0f * far_ * near_) / (near_ - far_);
    return m;
}
inline Mat4 look_at(Vec3 eye, Vec3 center, Vec3 up) { // glm::ext::look_at (RH)
    const Vec3 f = normalize(center - eye), s = normalize(cross(f, up)), u = cross(s, f);
    Mat4 m = Mat4::identity();
    m.c[0][0] = s.x; m.c[1][0] = s.y; m.c[2][0] = s.z;
    m.c[0][1] = u.x; m.c[1][1] = u.y; m.c[2][1] = u.z;
    m.c[0][2] = -f.x; m.c[1][2] = -f.y; m.c[2][2] = -f.z;
    m.c[3][0] = -dot(s, eye); m.c[3][1] = -dot(u, eye); m.c[3][2] = dot(f, eye);
    return m;
}
inline Mat4 translate(const Mat4& m, Vec3 v) {
    Mat4 r = m;
    for (int i = 0; i < 4; ++i) r.c[3][i] = m.c[0][i] * v.x + m.c[1][i] * v.y + m.c[2][i] * v.z + m.c[3][i];
    return r;
}
inline Mat4 scale(const Mat4& m, Vec3 v) {
    Mat4 r = m;
    for (int i = 0; i < 4; ++i) { r.c[0][i] = m.c[0][i] * v.x; r.c[1][i] = m.c[1][i] * v.y; r.c[2][i] = m.c[2][i] * v.z; }
    return r;
}

// ---- src/voxel ---------------------------------------------------------------------------
struct Color { // src/voxel/common.rs:65-87
    uint8_t r, g, b, a;
    static Color from_uint(uint32_t x) { return {uint8_t(x), uint8_t(x >> 8), uint8_t(x >> 16), uint8_t(x >> 24)}; }
};
static_assert(sizeof(Color) == 4, "Color is 4 bytes");

class RawDynamicChunk { // src/voxel/rawchunk.rs:254-320
public:
    RawDynamicChunk(size_t dx, size_t dy, size_t dz, Color fill) : dim_x_(dx), dim_y_(dy), dim_z_(dz), data_(dx * dy * dz, fill) {}
    Color* at_mut(int x, int y, int z) {
        if (x < 0 || y < 0 || z < 0 || size_t(x) >= dim_x_ || size_t(y) >= dim_y_ || size_t(z) >= dim_z_) return nullptr;
        return &data_[size_t(z) + dim_z_ * (size_t(y) + dim_y_ * size_t(x))]; // :292
    }
    const Color* get_raw() const { return data_.data(); }
    std::pair<int, int> dim_x() const { return {0, int(dim_x_)}; }
    std::pair<int, int> dim_y() const { return {0, int(dim_y_)}; }
    std::pair<int, int> dim_z() const { return {0, int(dim_z_)}; }

private:
    size_t dim_x_, dim_y_, dim_z_;
    std::vector<Color> data_;
};

// src/voxel/magica_voxel.rs:18-44 over dot_vox 4.1.0 semantics: one chunk per model; voxel.i = file
// index - 1; palette[k] = k-th RGBA quad; axes (x, size.y - z - 1, y).  Throws where Rust unwraps.
inline std::vector<RawDynamicChunk> load_magica_voxel(const std::string& filepath) {
    std::ifstream f(filepath, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + filepath);
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    auto u32 = [&](size_t o) { return uint32_t(buf[o]) | uint32_t(buf[o + 1]) << 8 | uint32_t(buf[o + 2]) << 16 | uint32_t(buf[o + 3]) << 24; };
    if (buf.size() < 20 || std::memcmp(buf.data(), "VOX ", 4) != 0 || std::memcmp(buf.data() + 8, "MAIN", 4) != 0)
        throw std::runtime_error("not a MagicaVoxel file: " + filepath);
    struct Model { uint32_t sx, sy, sz; size_t xyzi; };
    std::vector<Model> models;
    size_t rgba = 0, pending_size = 0;
    for (size_t off = 20; off + 12 <= buf.size();) {
        const uint32_t n = u32(off + 4), m = u32(off + 8);
        const size_t body = off + 12;
        if (body + n + m > buf.size()) break;
        if (!std::memcmp(&buf[off], "SIZE", 4)) pending_size = body;
        else if (!std::memcmp(&buf[off], "XYZI", 4) && pending_size) { models.push_back({u32(pending_size), u32(pending_size + 4), u32(pending_size + 8), body}); pending_size = 0; }
        else if (!std::memcmp(&buf[off], "RGBA", 4) && !rgba) rgba = body;
        off = body + n + m;
    }
    // no RGBA chunk: MagicaVoxel's published default palette (index 0 unused, 6x6x6 colour cube without black, ramps of
    // red / green / blue / grey), laid out like an RGBA chunk (entry k = colour of file index k + 1); dot_vox's own copy of
    // the table is not available offline, so this is an unpinned restatement
    std::vector<uint32_t> palette(256, 0);
    if (rgba) {
        for (size_t k = 0; k < 256; ++k) palette[k] = u32(rgba + 4 * k);
    } else {
        static const uint32_t lv[6] = {0xff, 0xcc, 0x99, 0x66, 0x33, 0x00}, ramp[10] = {0xee, 0xdd, 0xbb, 0xaa, 0x88, 0x77, 0x55, 0x44, 0x22, 0x11};
        uint32_t t[257] = {};
        for (uint32_t k = 0; k < 215; ++k) t[k + 1] = 0xff000000u | lv[k % 6] << 16 | lv[(k / 6) % 6] << 8 | lv[k / 36];
        for (uint32_t j = 0; j < 10; ++j) {
            t[216 + j] = 0xff000000u | ramp[j];
            t[226 + j] = 0xff000000u | ramp[j] << 8;
            t[236 + j] = 0xff000000u | ramp[j] << 16;
            t[246 + j] = 0xff000000u | ramp[j] * 0x010101u;
        }
        for (size_t k = 0; k < 256; ++k) palette[k] = t[k + 1];
    }
    std::vector<RawDynamicChunk> chunks;
    for (const Model& md : models) {
        RawDynamicChunk chunk(md.sx, md.sy, md.sz, Color::from_uint(0));
        const uint32_t nv = u32(md.xyzi);
        for (uint32_t v = 0; v < nv; ++v) {
            const uint8_t* q = &buf[md.xyzi + 4 + 4 * size_t(v)];
            Color* c = chunk.at_mut(q[0], int(md.sy) - int(q[2]) - 1, q[1]);
            if (!c) throw std::runtime_error("voxel outside chunk");
            const uint32_t i = q[3] ? q[3] - 1u : 0u;
            *c = Color::from_uint(palette[i]);
        }
        chunks.push_back(std::move(chunk));
    }
    return chunks;
}

// ---- src/scene.rs ------------------------------------------------------------------------
using TextureHandle = uint32_t;

class SceneGraph {
public:
    SceneGraph() : model_(Mat4::identity()), is_child_(false), total_children_(0), texture_handle_(0) {}
    static SceneGraph new_child(const Mat4& model, TextureHandle h) {
        SceneGraph s;
        s.model_ = model; s.is_child_ = true; s.texture_handle_ = h;
        return s;
    }
    Mat4& get_model_mut() { return model_; }
    uint32_t num_total_children() const { return is_child_ ? 1u : total_children_; }
    void add_child(SceneGraph child) {
        if (!is_child_) {
            total_children_ += child.num_total_children();
            children_.push_back(std::move(child));
        } else { // a child becomes a parent holding its former self (src/scene.rs:84-104)
            SceneGraph self_child = new_child(Mat4::identity(), texture_handle_);
            is_child_ = false;
            total_children_ = 1 + child.num_total_children();
            children_.clear();
            children_.push_back(std::move(self_child));
            children_.push_back(std::move(child));
        }
    }
    // depth-first flattening with the model stack multiplied down (src/scene.rs:146-181)
    void flatten(std::vector<std::pair<Mat4, TextureHandle>>& out) const { flatten(Mat4::identity(), true, out); }

private:
    void flatten(const Mat4& parent, bool root, std::vector<std::pair<Mat4, TextureHandle>>& out) const {
        const Mat4 m = root ? model_ : parent * model_;
        if (is_child_) { out.emplace_back(m, texture_handle_); return; }
        for (const SceneGraph& c : children_) c.flatten(m, false, out);
    }
    Mat4 model_;
    bool is_child_;
    uint32_t total_children_;
    TextureHandle texture_handle_;
    std::vector<SceneGraph> children_;
};

// ---- src/render.rs -----------------------------------------------------------------------
struct GPUInstance { // src/render.rs:32-36,74-78
    Mat4 model;
    static GPUInstance from_model(const Mat4& model, uint32_t texture_id) {
        GPUInstance g{model};
        std::memcpy(&g.model.c[3][3], &texture_id, 4);
        return g;
    }
};
static_assert(sizeof(GPUInstance) == 64, "GPUInstance is 64 bytes");

class TextureUploadQueue { // src/render.rs:150-175
public:
    TextureHandle add_texture(std::shared_ptr<RawDynamicChunk> texture) {
        const TextureHandle h = num_textures_added_++;
        queue_.emplace_back(std::move(texture), h);
        return h;
    }
    bool pop(std::pair<std::shared_ptr<RawDynamicChunk>, TextureHandle>& out) {
        if (queue_.empty()) return false;
        out = std::move(queue_.front());
        queue_.pop_front();
        return true;
    }

private:
    std::deque<std::pair<std::shared_ptr<RawDynamicChunk>, TextureHandle>> queue_;
    TextureHandle num_textures_added_ = 0;
};

class Renderer {
public:
    Renderer() { // src/render.rs:183-205
        const uint64_t code = entry();
        if (code != 0) throw std::runtime_error(std::string("ERROR: renderer initialization failed: ") + vt_last_error());
        fov_ = 80.0f / 180.0f * 3.1415926f;
        perspective_ = create_perspective(fov_, 1.0f);
        camera_ = create_camera({0, 0, 0}, {1, 0, 0});
    }
    ~Renderer() { cleanup(); } // Drop, src/render.rs:316-320
    Renderer(const Renderer&) = delete;
    Renderer& operator=(const Renderer&) = delete;

    static Mat4 create_perspective(float fov, float aspect) { return perspective(fov, aspect, 0.01f, 10000.0f); }
    static Mat4 create_camera(Vec3 position, Vec3 direction) { return look_at(position, position + direction, {0.0f, 1.0f, 0.0f}); }
    const user_input* get_input_data_pointer() const { return ::get_input_data_pointer(); }

    void update_instances(const SceneGraph& scene) { // src/render.rs:220-238
        std::vector<std::pair<Mat4, TextureHandle>> flat;
        scene.flatten(flat);
        auto* ptr = reinterpret_cast<GPUInstance*>(start_update_instances(scene.num_total_children()));
        if (!ptr) throw std::runtime_error("ERROR: Updating instances failed");
        uint32_t true_instance_count = 0;
        for (const auto& [model, handle] : flat) {
            auto it = texture_handle_lookup_.find(handle);
            if (it == texture_handle_lookup_.end()) continue; // texture not resident yet
            ptr[true_instance_count++] = GPUInstance::from_model(model, it->second);
        }
        if (end_update_instances(true_instance_count) != 0) throw std::runtime_error("ERROR: Updating instances failed");
    }

    // src/render.rs:240-313: uploads at most ONE queued texture, renders, then rebuilds the matrices
    bool render_tick(Vec3 pos, Vec3 dir, TextureUploadQueue& queue) {
        std::pair<std::shared_ptr<RawDynamicChunk>, TextureHandle> item;
        if (queue.pop(item)) {
            const auto dx = item.first->dim_x(), dy = item.first->dim_y(), dz = item.first->dim_z();
            const int32_t id = add_texture(reinterpret_cast<const uint8_t*>(item.first->get_raw()), uint32_t(dx.second - dx.first),
                                           uint32_t(dy.second - dy.first), uint32_t(dz.second - dz.first));
            if (id < 0) throw std::runtime_error("ERROR: Adding texture failed");
            texture_handle_lookup_[item.second] = uint32_t(id);
        }
        used_perspective_ = perspective_;
        used_camera_ = camera_;
        render_tick_info info{&perspective_, &camera_};
        const bool ok = ::render_tick(&window_width_, &window_height_, &info) == 0;
        if (ok) { rendered_perspective_ = used_perspective_; rendered_camera_ = used_camera_; }
        if (window_width_ != prev_window_width_ || window_height_ != prev_window_height_) {
            perspective_ = create_perspective(fov_, float(window_width_) / float(window_height_));
            prev_window_width_ = window_width_;
            prev_window_height_ = window_height_;
        }
        camera_ = create_camera(pos, dir);
        ++frame_num_;
        return ok;
    }
    int32_t window_width() const { return window_width_; }
    int32_t window_height() const { return window_height_; }
    size_t frame_num() const { return frame_num_; }
    // the matrices the most recent successfully rendered frame used (frame k renders frame k-1's pose)
    const Mat4& rendered_perspective() const { return rendered_perspective_; }
    const Mat4& rendered_camera() const { return rendered_camera_; }

private:
    int32_t window_width_ = 1, window_height_ = 1, prev_window_width_ = 1, prev_window_height_ = 1;
    float fov_ = 0.0f;
    size_t frame_num_ = 0;
    Mat4 perspective_{}, camera_{}, used_perspective_{}, used_camera_{}, rendered_perspective_{}, rendered_camera_{};
    std::unordered_map<TextureHandle, uint32_t> texture_handle_lookup_;
};

} // namespace vtrace
