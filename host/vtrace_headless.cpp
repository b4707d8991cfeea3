// vtrace_headless.cpp — a compiled host over librender, shaped like the reference's src/main.rs +
// src/world.rs: loads AncientTemple.vox and Treasure.vox, queues them for upload (one per tick,
// src/render.rs:246), builds the 11x11 entity grid scene graph every frame (src/world.rs:143-161),
// and drives Renderer::update_instances / Renderer::render_tick until the library stops the loop
// (VT_MAX_FRAMES).  Prints an FNV-1a hash of the final RGBA8 frame and optionally writes a PPM.
//
//   VT_WIDTH=640 VT_HEIGHT=360 VT_MAX_FRAMES=6 ./vtrace_headless <assets dir> [out.ppm]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "vtrace_host.hpp"

using namespace vtrace;

struct WorldState { // src/world.rs:28-76 (camera + entity registry; terrain paging is out of scope)
    Vec3 camera_position{0.0f, -2.0f, 0.0f};
    float camera_theta = 0.588f, camera_phi = 2.2f;
    TextureHandle treasure = 0, temple = 0;

    WorldState(TextureUploadQueue& q, const std::string& assets) {
        temple = q.add_texture(std::make_shared<RawDynamicChunk>(load_magica_voxel(assets + "/AncientTemple.vox").at(0)));
        treasure = q.add_texture(std::make_shared<RawDynamicChunk>(load_magica_voxel(assets + "/Treasure.vox").at(0)));
    }
    Vec3 get_camera_direction() const { // src/world.rs:75-81
        return {std::cos(camera_theta) * std::sin(camera_phi), std::cos(camera_phi), std::sin(camera_theta) * std::sin(camera_phi)};
    }
    SceneGraph update(float dt, const user_input& in) { // src/world.rs:88-161
        camera_theta += 0.002f * float(in.mouse_x - in.last_mouse_x); // SENSITIVITY * mouse delta (zero when headless)
        camera_phi -= 0.002f * float(in.mouse_y - in.last_mouse_y);
        camera_theta += 0.01f * dt;                                   // scripted pan so consecutive frames differ
        SceneGraph scene, scene_entities;
        for (int x = -5; x <= 5; ++x)
            for (int z = -5; z <= 5; ++z) {
                const Mat4 model = translate(Mat4::identity(), {float(x) * 1.5f, -5.0f, float(z) * 1.5f});
                scene_entities.add_child(SceneGraph::new_child(model, (uint32_t(x + z + 10) % 2 == 0) ? treasure : temple));
            }
        scene.add_child(std::move(scene_entities));
        return scene;
    }
};

int main(int argc, char** argv) {
    const std::string assets = argc > 1 ? argv[1] : "tests/golden/assets";
    try {
        Renderer renderer;
        TextureUploadQueue queue;
        WorldState world(queue, assets);
        const user_input* input_ptr = renderer.get_input_data_pointer(); // src/main.rs:31
        SceneGraph scene = world.update(0.0f, *input_ptr);
        bool code = true;
        while (code) { // src/main.rs:36-55 (the render thread is joined every frame, so this is the same order)
            const Vec3 pos = world.camera_position, dir = world.get_camera_direction();
            renderer.update_instances(scene);
            code = renderer.render_tick(pos, dir, queue);
            scene = world.update(1.0f, *input_ptr); // src/main.rs:52
        }
        const int w = renderer.window_width(), h = renderer.window_height();
        std::vector<uint8_t> rgba(size_t(w) * h * 4);
        if (vt_read_color(rgba.data(), rgba.size()) != int64_t(rgba.size())) { std::fprintf(stderr, "read-back failed: %s\n", vt_last_error()); return 2; }
        uint64_t hash = 1469598103934665603ull;
        for (uint8_t b : rgba) { hash ^= b; hash *= 1099511628211ull; }
        vt_stats st{};
        vt_get_stats(&st);
        std::printf("frames=%zu size=%dx%d iterations=%llu fnv1a=%016llx\n", renderer.frame_num(), w, h,
                    (unsigned long long)st.iterations, (unsigned long long)hash);
        auto dump = [](const char* name, const Mat4& m) {
            std::printf("%s=", name);
            for (int c = 0; c < 4; ++c)
                for (int r = 0; r < 4; ++r) { uint32_t u; std::memcpy(&u, &m.c[c][r], 4); std::printf("%08x%s", u, (c == 3 && r == 3) ? "\n" : ","); }
        };
        dump("P", renderer.rendered_perspective());
        dump("V", renderer.rendered_camera());
        if (argc > 2) {
            FILE* f = std::fopen(argv[2], "wb");
            if (f) {
                std::fprintf(f, "P6\n%d %d\n255\n", w, h);
                for (size_t p = 0; p < size_t(w) * h; ++p) std::fwrite(&rgba[4 * p], 1, 3, f);
                std::fclose(f);
            }
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
