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
        if (paths[0] == "--native") {
            // the batch dispatcher (zpix.fromBuffers): native variants, hash of the .pixels buffer
            paths.erase(paths.begin());
            std::vector<std::vector<uint8_t>> files;
            std::vector<std::pair<const uint8_t*, size_t>> bufs;
            for (auto& p : paths) files.push_back(zpix::jpeg::readFile(p));
            for (auto& f : files) bufs.push_back({f.data(), f.size()});
            bufs.push_back({reinterpret_cast<const uint8_t*>("\x89PNG\r\n\x1a\n"), 8});
            auto res = zpix::fromBuffers(ctx, bufs);
            if (res.back().status != zpix::kUnknownImageFormat) return 3;
            for (size_t i = 0; i + 1 < res.size(); i++) {
                if (res[i].status) return 1;
                const std::vector<uint8_t>& px = std::visit([](auto const& m) -> const std::vector<uint8_t>& { return m.pixels; }, res[i].img.v);
                uint64_t h = 1469598103934665603ull;
                for (uint8_t b : px) h = (h ^ b) * 1099511628211ull;
                std::printf("%s variant=%d len=%zu fnv1a=%016llx\n", paths[i].c_str(), (int)res[i].img.v.index(), px.size(), (unsigned long long)h);
            }
            return 0;
        }
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
