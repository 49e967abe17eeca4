// Host-only check of ofps_b200::LumaFileSource (include/ofps_b200.hpp): raw "WxH@FPS:path" and .y4m inputs.
#include <cstdio>
#include <vector>

#include "ofps_b200.hpp"

using namespace ofps_b200;

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    try {
        LumaFileSource s(argv[1]);
        std::vector<uint8_t> y((size_t)s.width() * s.height());
        unsigned long long sum = 0;
        int n = 0;
        while (s.next(y.data())) {
            for (uint8_t v : y) sum = sum * 31 + v;
            n++;
        }
        std::printf("%d %d %.3f %d %llu\n", s.width(), s.height(), s.framerate(), n, sum);
        return 0;
    } catch (const Error& e) {
        std::printf("ERROR %d %s\n", e.code, e.what());
        return 1;
    }
}
