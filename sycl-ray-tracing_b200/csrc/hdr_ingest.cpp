// Radiance RGBE (.hdr) reader for environment maps: replaces the stbi_loadf call of Utils::read_image_float
// (source/utils.cpp:100-124) for the one format main.cpp feeds it (main.cpp:35). Written from the format description:
// "#?RADIANCE" / "#?RGBE" header, FORMAT=32-bit_rle_rgbe, a blank line, "-Y h +X w", then scanlines either flat (4 bytes per
// pixel) or new-style run-length encoded (2 2 hi lo, then the four channels one after the other as runs / literals).
// Decoding to float follows what the reference's loader produces: channel = mantissa * 2^(e - 136); e == 0 -> 0.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "obj_ingest.h"

namespace b200rt {

bool load_hdr(const std::string& path, bool flip_y, std::vector<float>& rgb, int& width, int& height, std::string& err)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    std::vector<unsigned char> data;
    {
        std::fseek(f, 0, SEEK_END);
        const long n = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        data.resize(n > 0 ? (size_t)n : 0);
        if (n > 0 && std::fread(data.data(), 1, (size_t)n, f) != (size_t)n) { std::fclose(f); err = "short read"; return false; }
        std::fclose(f);
    }
    size_t p = 0;
    auto next_line = [&](std::string& line) {
        line.clear();
        while (p < data.size() && data[p] != '\n') line.push_back((char)data[p++]);
        if (p < data.size()) p++;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        return p <= data.size();
    };
    std::string line;
    next_line(line);
    if (line != "#?RADIANCE" && line != "#?RGBE") { err = "not a Radiance HDR file"; return false; }
    bool format_ok = false;
    for (;;)
    {
        if (p >= data.size()) { err = "truncated header"; return false; }
        next_line(line);
        if (line.empty()) break;
        if (line == "FORMAT=32-bit_rle_rgbe") format_ok = true;
    }
    if (!format_ok) { err = "unsupported HDR format (need 32-bit_rle_rgbe)"; return false; }
    next_line(line);
    int h = 0, w = 0;
    if (std::sscanf(line.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) { err = "unsupported HDR orientation: " + line; return false; }
    width = w; height = h;
    std::vector<unsigned char> rgbe((size_t)w * h * 4);
    std::vector<unsigned char> scan((size_t)w * 4);
    for (int y = 0; y < h; y++)
    {
        unsigned char* row = &rgbe[(size_t)y * w * 4];
        const bool rle = w >= 8 && w < 32768 && p + 4 <= data.size() && data[p] == 2 && data[p + 1] == 2 && !(data[p + 2] & 0x80) &&
                         ((data[p + 2] << 8) | data[p + 3]) == w;
        if (!rle)
        {
            if (p + (size_t)w * 4 > data.size()) { err = "truncated pixel data"; return false; }
            std::memcpy(row, &data[p], (size_t)w * 4);
            p += (size_t)w * 4;
            continue;
        }
        p += 4;
        for (int c = 0; c < 4; c++)
        {
            int x = 0;
            while (x < w)
            {
                if (p >= data.size()) { err = "truncated run-length data"; return false; }
                int count = data[p++];
                if (count > 128)
                {
                    count -= 128;
                    if (x + count > w || p >= data.size()) { err = "corrupt run"; return false; }
                    const unsigned char v = data[p++];
                    for (int i = 0; i < count; i++) scan[(size_t)(x++) * 4 + c] = v;
                }
                else
                {
                    if (count == 0 || x + count > w || p + (size_t)count > data.size()) { err = "corrupt literal run"; return false; }
                    for (int i = 0; i < count; i++) scan[(size_t)(x++) * 4 + c] = data[p++];
                }
            }
        }
        std::memcpy(row, scan.data(), (size_t)w * 4);
    }
    rgb.resize((size_t)w * h * 3);
    for (int y = 0; y < h; y++)
    {
        const int src_y = flip_y ? h - 1 - y : y;          // file row 0 is the top; the reference loads with the vertical flip on
        for (int x = 0; x < w; x++)
        {
            const unsigned char* q = &rgbe[((size_t)src_y * w + x) * 4];
            float* o = &rgb[((size_t)y * w + x) * 3];
            if (q[3] != 0)
            {
                const float f1 = std::ldexp(1.0f, (int)q[3] - (128 + 8));
                o[0] = q[0] * f1; o[1] = q[1] * f1; o[2] = q[2] * f1;
            }
            else o[0] = o[1] = o[2] = 0.0f;
        }
    }
    return true;
}

} // namespace b200rt
