// host_types.h — host-side vocabulary of the render-graph boundary, B200 build.
//
// Same names and meaning as the reference's declarations so that HybridRenderPath::RegisterPath reads the same:
//   TransientResource / TransientImage / TransientImageType   <- src/rendering_backend/vulkan_common.h:233-268
//   *PipelineDescription, ComputeKernel, PushConstantDescription, callbacks  <- vulkan_common.h:206-336
//   VkUtils::CreateTransient*Image helpers                     <- src/rendering_backend/vulkan_utils.h:347-453
//   push-constant PODs, PerFrameData, Vertex, Primitive        <- src/rendering_backend/glsl_common.h:31-99
// There is no Vulkan here: VkFormat is reduced to the four formats the hot path uses (values = the Vulkan enum's),
// pipelines are names looked up in the CUDA library, images are linear device buffers owned by libvhr_b200.so.
#pragma once
#include <stdint.h>

#include <array>
#include <functional>
#include <string>
#include <variant>
#include <vector>

#include "vhr_b200.h"

enum VkFormat : int {
    VK_FORMAT_UNDEFINED = 0,
    VK_FORMAT_R8G8B8A8_UNORM = VHR_FORMAT_R8G8B8A8_UNORM,
    VK_FORMAT_R8G8B8A8_SRGB = VHR_FORMAT_R8G8B8A8_SRGB,
    VK_FORMAT_B8G8R8A8_UNORM = VHR_FORMAT_B8G8R8A8_UNORM,
    VK_FORMAT_B8G8R8A8_SRGB = VHR_FORMAT_B8G8R8A8_SRGB,
    VK_FORMAT_R16G16_SFLOAT = VHR_FORMAT_R16G16_SFLOAT,
    VK_FORMAT_R16G16B16A16_SFLOAT = VHR_FORMAT_R16G16B16A16_SFLOAT,
    VK_FORMAT_D32_SFLOAT = VHR_FORMAT_D32_SFLOAT,
};

// vulkan_common.h:21-26; the enumerators carry the Vulkan values
enum VkFilter : int { VK_FILTER_NEAREST = VHR_FILTER_NEAREST, VK_FILTER_LINEAR = VHR_FILTER_LINEAR };
enum VkSamplerAddressMode : int {
    VK_SAMPLER_ADDRESS_MODE_REPEAT = VHR_ADDRESS_MODE_REPEAT,
    VK_SAMPLER_ADDRESS_MODE_MIRRORED_REPEAT = VHR_ADDRESS_MODE_MIRRORED_REPEAT,
    VK_SAMPLER_ADDRESS_MODE_CLAMP_TO_EDGE = VHR_ADDRESS_MODE_CLAMP_TO_EDGE,
    VK_SAMPLER_ADDRESS_MODE_CLAMP_TO_BORDER = VHR_ADDRESS_MODE_CLAMP_TO_BORDER,
};
struct SamplerInfo {
    VkFilter mag_filter;
    VkFilter min_filter;
    VkSamplerAddressMode address_mode_u;
    VkSamplerAddressMode address_mode_v;
};

namespace glmlite {
struct uvec2 { uint32_t x, y; };
struct ivec2 { int32_t x, y; };
}  // namespace glmlite

// ---- glsl_common.h PODs (layouts verified against the reference header: SURVEY Appendix C) -----------------------
struct DirectionalLight { float projview[16]; float direction[4]; float color[4]; float intensity[4]; };
struct PerFrameData {
    float camera_view[16], camera_proj[16], camera_view_inverse[16], camera_proj_inverse[16], camera_viewproj_inverse[16];
    float camera_view_prev_frame[16], camera_proj_prev_frame[16];
    DirectionalLight directional_light;
    float display_size[2], display_size_inverse[2];
    uint32_t frame_index;
    int32_t blue_noise_texture_index;
};
static_assert(sizeof(PerFrameData) == 584, "PerFrameData must match glsl_common.h:59-72");
struct Vertex { float pos[3]; float normal[3]; float tangent[4]; float uv0[2]; float uv1[2]; };
static_assert(sizeof(Vertex) == 56, "Vertex must match glsl_common.h:74-80");
struct Material {
    float base_color[4];
    int32_t base_color_texture, metallic_roughness_texture, normal_map;
    float metallic_factor, roughness_factor;
    int32_t alpha_mask;
    float alpha_cutoff;
};
struct Primitive { float transform[16]; Material material; uint32_t vertex_offset, index_offset, index_count; };
static_assert(sizeof(Primitive) == 120, "Primitive must match glsl_common.h:82-99");
struct Mesh { std::vector<Primitive> primitives; };
// vulkan_common.h:33-41,68-73; matrices are glm column-major (m[c*4+r])
struct Camera {
    float perspective[16];
    float transform[16];
    float view[16];
    float yaw, pitch, roll;
};
struct Scene {
    std::string name;
    Camera camera{};
    DirectionalLight directional_light{};
    std::vector<Mesh> meshes;
};

struct SVGFPushConstants {
    glmlite::ivec2 integrated_shadow_and_ao;
    int prev_frame_normals_and_object_ids;
    int shadow_and_ao_history;
    int shadow_and_ao_moments_history;
    int atrous_step;
};
static_assert(sizeof(SVGFPushConstants) == 24, "SVGFPushConstants must match glsl_common.h:31-39");
struct SSAOPushConstants { float radius; };
struct SSRPushConstants { float ray_distance, step_size, thickness; int bsearch_steps; };

// ---- render-graph declarations -----------------------------------------------------------------------------------
struct PushConstantDescription { uint32_t size; uint32_t shader_stage; };
inline constexpr PushConstantDescription PUSHCONSTANTS_NONE{0, 0};

enum class TransientResourceType { Image, Buffer };
enum class TransientImageType { AttachmentImage, SampledImage, StorageImage };
struct ClearValue { float color[4]; };
struct TransientImage {
    TransientImageType type;
    uint32_t width, height;      // 0,0 = swapchain-sized (render_graph.cpp:962-966)
    VkFormat format;
    uint32_t binding;
    ClearValue clear_value;
    bool multisampled;
};
struct TransientResource {
    TransientResourceType type;
    const char *name;
    TransientImage image;
};

struct GraphicsPipelineDescription {
    const char *name;
    const char *vertex_shader;
    const char *fragment_shader;
    PushConstantDescription push_constants = PUSHCONSTANTS_NONE;
    std::vector<int> specialization_constants;
};
struct HitShader { const char *closest_hit = nullptr; const char *any_hit = nullptr; };
struct RaytracingPipelineDescription {
    const char *name;
    const char *raygen_shader;
    std::vector<const char *> miss_shaders;
    std::vector<HitShader> hit_shaders;
};
struct ComputeKernel { const char *shader; };
struct ComputePipelineDescription {
    std::vector<ComputeKernel> kernels;
    PushConstantDescription push_constant_description = PUSHCONSTANTS_NONE;
};

class GraphicsExecutionContext;
using GraphicsExecutionCallback = std::function<void(GraphicsExecutionContext &)>;
using ExecuteGraphicsCallback = std::function<void(std::string, GraphicsExecutionCallback)>;
using GraphicsPassCallback = std::function<void(ExecuteGraphicsCallback)>;
class RaytracingExecutionContext;
using RaytracingExecutionCallback = std::function<void(RaytracingExecutionContext &)>;
using ExecuteRaytracingCallback = std::function<void(std::string, RaytracingExecutionCallback)>;
using RaytracingPassCallback = std::function<void(ExecuteRaytracingCallback)>;
class ComputeExecutionContext;
using ComputePassCallback = std::function<void(ComputeExecutionContext &)>;

struct GraphicsPassDescription { std::vector<GraphicsPipelineDescription> pipeline_descriptions; GraphicsPassCallback callback; };
struct RaytracingPassDescription { RaytracingPipelineDescription pipeline_description; RaytracingPassCallback callback; };
struct ComputePassDescription { ComputePipelineDescription pipeline_description; ComputePassCallback callback; };
struct RenderPassDescription {
    const char *name;
    std::vector<TransientResource> dependencies;
    std::vector<TransientResource> outputs;
    std::variant<GraphicsPassDescription, RaytracingPassDescription, ComputePassDescription> description;
};

namespace VkUtils {
inline ClearValue ClearColor(float r, float g, float b, float a) { return ClearValue{{r, g, b, a}}; }
inline ClearValue ClearDepth(float d) { return ClearValue{{d, 0.0f, 0.0f, 0.0f}}; }
inline TransientResource MakeImage(TransientImageType type, const char *name, uint32_t w, uint32_t h, VkFormat format, uint32_t binding,
                                   ClearValue clear = ClearValue{{0, 0, 0, 0}}) {
    return TransientResource{TransientResourceType::Image, name, TransientImage{type, w, h, format, binding, clear, false}};
}
inline TransientResource CreateTransientAttachmentImage(const char *name, VkFormat format, uint32_t binding, ClearValue clear) {
    return MakeImage(TransientImageType::AttachmentImage, name, 0, 0, format, binding, clear);
}
inline TransientResource CreateTransientAttachmentImage(const char *name, uint32_t w, uint32_t h, VkFormat format, uint32_t binding, ClearValue clear) {
    return MakeImage(TransientImageType::AttachmentImage, name, w, h, format, binding, clear);
}
inline TransientResource CreateTransientSampledImage(const char *name, VkFormat format, uint32_t binding) {
    return MakeImage(TransientImageType::SampledImage, name, 0, 0, format, binding);
}
inline TransientResource CreateTransientSampledImage(const char *name, uint32_t w, uint32_t h, VkFormat format, uint32_t binding) {
    return MakeImage(TransientImageType::SampledImage, name, w, h, format, binding);
}
inline TransientResource CreateTransientStorageImage(const char *name, VkFormat format, uint32_t binding) {
    return MakeImage(TransientImageType::StorageImage, name, 0, 0, format, binding);
}
// vulkan_utils.h:444-453: the swapchain image; in the headless build an image named RENDER_OUTPUT in the swapchain's
// format, B8G8R8A8_SRGB (vulkan_context.cpp:331)
inline TransientResource CreateTransientRenderOutput(uint32_t binding) {
    return MakeImage(TransientImageType::AttachmentImage, "RENDER_OUTPUT", 0, 0, VK_FORMAT_B8G8R8A8_SRGB, binding);
}
}  // namespace VkUtils

// VK_CHECK analogue (vulkan_common.h:4-7): print and abort the operation; the C entry points turn it into a status.
struct VhrHostError { int status; std::string message; };
#define VHR_CHECK(expr)                                                                                          \
    do {                                                                                                         \
        int rc__ = (expr);                                                                                       \
        if (rc__ < 0) throw VhrHostError{rc__, std::string(#expr) + ": " + vhr_last_error()};                    \
    } while (0)
#define VHR_ASSERT(cond, msg)                                                                                    \
    do {                                                                                                         \
        if (!(cond)) throw VhrHostError{VHR_ERR_INVALID, std::string("assert(" #cond "): ") + (msg)};            \
    } while (0)
