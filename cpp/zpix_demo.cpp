// zpix_demo.cpp -- the reference's canonical caller pattern (example/convert.zig:59-64:
// load -> rgbaPixels -> free) through the C++ host layer.  Prints width, height and an FNV-1a hash
// of the RGBA bytes per file; exit code 0 when every file decoded.
#include <cstdio>

#include "zpix.hpp"

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s file.jpg [...]\n", argv[0]);
        return 64;
    }
    try {
        zpix::jpeg::Context ctx;
        std::vector<std::string> paths(argv + 1, argv + argc);
        auto res = zpix::jpeg::loadBatch(ctx, paths);
        int bad = 0;
        for (size_t i = 0; i < res.size(); i++) {
            if (res[i].status) {
                std::printf("%s error.%s\n", paths[i].c_str(), zpx_error_name(res[i].status));
                bad++;
                continue;
            }
            const auto& px = res[i].img.rgbaPixels();
            uint64_t h = 1469598103934665603ull;
            for (uint8_t b : px) h = (h ^ b) * 1099511628211ull;
            auto r = res[i].img.bounds();
            std::printf("%s %dx%d fnv1a=%016llx\n", paths[i].c_str(), r.dX(), r.dY(), (unsigned long long)h);
        }
        return bad ? 1 : 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 2;
    }
}
