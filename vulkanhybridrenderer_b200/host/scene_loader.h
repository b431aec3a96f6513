// scene_loader.h — glTF 2.0 scene path of the host, B200 build.
//
// Mirrors src/scene/scene_loader.{h,cpp} of the reference (SceneLoader::LoadScene :336-349, ParseglTF :233-334,
// ParseNode :40-231): same flattening of nodes -> meshes -> primitives into the global Vertex[] / uint32 indices[] /
// Primitive[] arrays ResourceManager::UpdateGeometry consumes, same material mapping, same texture formats (base colour
// sRGB, the rest UNORM) and glTF-sampler -> SamplerInfo mapping, same camera and directional-light set-up.
// The reference parses with cgltf and decodes images with stb_image (third-party, vendored there, not copied here): both are used
// through the include path when the library is built next to the reference tree (HasCgltf / HasStbImage); the loader also has its
// own small JSON / .glb / accessor reader and a PNG decoder on zlib, which are the fall-back and the cross-check of the tests.
//
// Parsing is split from uploading so the loader can be exercised without a GPU:
//   ParseScene(path, out)          file -> ParsedScene (pure host work)
//   LoadScene(resource_manager, path)   ParseScene + UploadTextureFromData per texture + UpdateGeometry (reference order)
#pragma once
#include <string>
#include <vector>

#include "render_graph.h"

namespace SceneLoader {
// true when the library was built against the reference's vendored cgltf.h: ParseScene then parses with it (VHR_GLTF_PARSER=own in the
// environment selects this file's own reader instead; without cgltf.h that reader is the only one)
bool HasCgltf();
// true when the library was built against the reference's vendored stb_image.h (JPEG / PNG / ... textures); false: own PNG decoder only
bool HasStbImage();

struct ParsedTexture {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgba;          // tightly packed R8G8B8A8 (stbi_load(..., STBI_rgb_alpha))
    VkFormat format = VK_FORMAT_R8G8B8A8_UNORM;
    SamplerInfo sampler{VK_FILTER_LINEAR, VK_FILTER_LINEAR, VK_SAMPLER_ADDRESS_MODE_REPEAT, VK_SAMPLER_ADDRESS_MODE_REPEAT};
    std::string name;                   // image name or uri (ResourceManager::TagImage)
};

struct ParsedScene {
    Scene scene;                        // meshes with texture indices RELATIVE to `textures` (0..n-1); LoadScene rebases them
    std::vector<Vertex> vertices;
    std::vector<uint32_t> indices;
    std::vector<ParsedTexture> textures;   // upload order = first use scanning meshes -> primitives (scene_loader.cpp:242-275)
};

// Throws VhrHostError on malformed input (the reference prints "Error Parsing glTF 2.0 File" and returns an empty scene,
// scene_loader.cpp:344-346; LoadScene below keeps that behaviour).
void ParseScene(const char *path, ParsedScene &out);

Scene LoadScene(ResourceManager &resource_manager, const char *path);

// PNG (8/16-bit, every colour type, non-interlaced) -> RGBA8. Exposed for tests.
bool DecodePNG(const uint8_t *data, size_t size, uint32_t &width, uint32_t &height, std::vector<uint8_t> &rgba, std::string &error);

}  // namespace SceneLoader
