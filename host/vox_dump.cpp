// vox_dump.cpp — prints, for every model of a .vox file, its dimensions and the FNV-1a 64 hash of the RGBA bytes the
// engine would hand to add_texture (the C++ mirror of src/voxel/magica_voxel.rs in vtrace_host.hpp).  CPU only:
//   g++ -std=c++17 -O2 -I../include -o vox_dump vox_dump.cpp && ./vox_dump file.vox
#include <cstdio>

#include "vtrace_host.hpp"

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: vox_dump file.vox\n"); return 2; }
    try {
        const auto chunks = vtrace::load_magica_voxel(argv[1]);
        for (size_t m = 0; m < chunks.size(); ++m) {
            const auto& c = chunks[m];
            const size_t n = size_t(c.dim_x().second) * size_t(c.dim_y().second) * size_t(c.dim_z().second) * 4;
            const uint8_t* p = reinterpret_cast<const uint8_t*>(c.get_raw());
            uint64_t h = 1469598103934665603ull;
            for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
            std::printf("model %zu dims %d %d %d fnv1a %016llx\n", m, c.dim_x().second, c.dim_y().second, c.dim_z().second, (unsigned long long)h);
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
