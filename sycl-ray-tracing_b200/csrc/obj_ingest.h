// Host-side asset ingest (obj_ingest.cpp, hdr_ingest.cpp): files -> the arrays b200rt_scene_create[_multi] takes.
#pragma once
#include <string>
#include <vector>

namespace b200rt {

// ParsedOBJ (include/parsed_obj.h) as flat arrays: 9 floats per triangle, one material index per triangle (slot 0 = the default
// material), 10 floats per material (== SimpleMaterial), original indices of the emissive triangles
struct ParsedObjArrays
{
    std::vector<float> tri_xyz9;
    std::vector<int> tri_material;
    std::vector<float> materials10;
    std::vector<int> emissive_tri;
};

bool load_obj(const std::string& path, ParsedObjArrays& out, std::string& err);

// Radiance .hdr (RGBE, flat or new-style RLE scanlines) -> RGB float triplets, row 0 = bottom row when flip_y (what
// Utils::read_image_float hands stbi_loadf, utils.cpp:100-124): mantissa * 2^(exponent - 136), 0 when the exponent byte is 0
bool load_hdr(const std::string& path, bool flip_y, std::vector<float>& rgb, int& width, int& height, std::string& err);

} // namespace b200rt
