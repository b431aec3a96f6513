// hybrid_render_path.cpp — see hybrid_render_path.h. Node table of HybridRenderPath::RegisterPath
// (reference: src/render_paths/hybrid_render_path.cpp:12-381), written table-first: the image declarations are data,
// the pass bodies are the per-frame call sequences of the reference lambdas.
#include "hybrid_render_path.h"

#include <utility>

void RenderPath::Build() {
    VHR_CHECK(vhr_context_synchronize(resource_manager.ctx));      // vkDeviceWaitIdle
    render_graph.DestroyResources();
    RegisterPath(render_graph, resource_manager);
    render_graph.Build();
}
void RenderPath::Rebuild() {
    VHR_CHECK(vhr_context_synchronize(resource_manager.ctx));
    DeregisterPath(render_graph, resource_manager);
    Build();
}

namespace {
// image names of the hybrid path (Appendix B of SURVEY.md)
constexpr const char *kAlbedo = "Albedo";
constexpr const char *kNormals = "World Space Normals and Object IDs";
constexpr const char *kMotion = "Motion Vectors and Metallic Roughness";
constexpr const char *kDepth = "Depth";
constexpr const char *kShadowMap = "Shadow Map";
constexpr const char *kRt = "Raytraced Shadows and Ambient Occlusion";
constexpr const char *kRefl = "Raytraced Reflections";
constexpr const char *kDenoised = "Denoised Raytraced Shadows and Ambient Occlusion";
constexpr const char *kSsaoRaw = "Screen Space Ambient Occlusion Raw";
constexpr const char *kSsao = "Screen Space Ambient Occlusion";
constexpr const char *kSsr = "Screen Space Reflections";
constexpr VkFormat F4 = VK_FORMAT_R16G16B16A16_SFLOAT, F2 = VK_FORMAT_R16G16_SFLOAT, D32 = VK_FORMAT_D32_SFLOAT, BGRA8 = VK_FORMAT_B8G8R8A8_UNORM;

using VkUtils::CreateTransientAttachmentImage;
using VkUtils::CreateTransientSampledImage;
using VkUtils::CreateTransientStorageImage;

inline uint32_t groups(uint32_t n) { return n / 8 + (n % 8 != 0); }

// Draw loop shared by the two rasterised geometry passes (hybrid_render_path.cpp:36-54, 78-96): one indexed draw per
// primitive, object id = flat primitive index. Without a rasteriser the calls only count.
void draw_scene(GraphicsExecutionContext &ec) {
    ec.BindGlobalVertexAndIndexBuffers();
    int object_id = 0;
    for (Mesh &mesh : ec.resource_manager.scene.meshes)
        for (Primitive &primitive : mesh.primitives) {
            struct { float normal_matrix[12]; int object_id; } pc{};   // HybridPushConstants stand-in
            pc.object_id = object_id++;
            ec.PushConstants(pc);
            ec.DrawIndexed(primitive.index_count, 1, primitive.index_offset, primitive.vertex_offset, 0);
        }
}
}  // namespace

void HybridRenderPath::RegisterPath(RenderGraph &rg, ResourceManager &rm) {
    const bool any_rt = shadow_mode == SHADOW_MODE_RAYTRACED || ambient_occlusion_mode == AMBIENT_OCCLUSION_MODE_RAYTRACED ||
                        reflection_mode == REFLECTION_MODE_RAYTRACED;

    // ---- G-Buffer Pass (:13-56) — rasterised in the reference; an external producer here -------------------------
    rg.AddGraphicsPass("G-Buffer Pass", {},
                       {CreateTransientAttachmentImage(kAlbedo, BGRA8, 0, VkUtils::ClearColor(0, 0, 0, 0)),
                        CreateTransientAttachmentImage(kNormals, F4, 1, VkUtils::ClearColor(0, 0, 0, 0)),
                        CreateTransientAttachmentImage(kMotion, F4, 2, VkUtils::ClearColor(0, 0, -1, -1)),
                        CreateTransientAttachmentImage(kDepth, D32, 3, VkUtils::ClearDepth(0))},
                       {GraphicsPipelineDescription{"G-Buffer Pipeline", "hybrid_render_path/gbuf.vert", "hybrid_render_path/gbuf.frag",
                                                    PushConstantDescription{68, 0}}},
                       [](ExecuteGraphicsCallback execute_pipeline) { execute_pipeline("G-Buffer Pipeline", draw_scene); });

    if (shadow_mode == SHADOW_MODE_RASTERIZED) {
        // ---- Shadow Map Pass (:58-99). Note the reference's `else if`: with rasterised shadows the Raytrace Pass is
        // not registered even when AO / reflections are ray traced. Kept. -------------------------------
        rg.AddGraphicsPass("Shadow Map Pass", {}, {CreateTransientAttachmentImage(kShadowMap, 4096, 4096, D32, 0, VkUtils::ClearDepth(0))},
                           {GraphicsPipelineDescription{"Shadow Map Pass Pipeline", "hybrid_render_path/depth_prepass.vert",
                                                        "hybrid_render_path/depth_prepass.frag", PushConstantDescription{68, 0}}},
                           [](ExecuteGraphicsCallback execute_pipeline) { execute_pipeline("Shadow Map Pass Pipeline", draw_scene); });
    } else if (any_rt) {
        // ---- Raytrace Pass (:101-136) -------------------------------------------------------------------------------
        RaytracingPipelineDescription pipe{"Raytrace Pipeline", "hybrid_render_path/raygen.rgen",
                                           {"hybrid_render_path/miss.rmiss", "hybrid_render_path/reflection_miss.rmiss"},
                                           {HitShader{"hybrid_render_path/reflection_hit.rchit", nullptr}}};
        rg.AddRaytracingPass("Raytrace Pass", {CreateTransientSampledImage(kNormals, F4, 0), CreateTransientSampledImage(kDepth, D32, 1)},
                             {CreateTransientStorageImage(kRt, F2, 2), CreateTransientStorageImage(kRefl, F4, 3)}, pipe,
                             [&rm](ExecuteRaytracingCallback execute_pipeline) {
                                 execute_pipeline("Raytrace Pipeline", [&rm](RaytracingExecutionContext &ec) { ec.TraceRays(rm.width, rm.height); });
                             });
    }

    if (ambient_occlusion_mode == AMBIENT_OCCLUSION_MODE_SSAO) {
        // ---- SSAO Pass + SSAO Blur Pass (:138-200). The reference attaches the push constants to the blur pass, whose
        // shader has no push-constant block, and dispatches ssao.comp without any: reproduced. -------------
        ssao_push_constants = SSAOPushConstants{0.75f};
        rg.AddComputePass("SSAO Pass", {CreateTransientSampledImage(kNormals, F4, 0), CreateTransientSampledImage(kDepth, D32, 1)},
                          {CreateTransientStorageImage(kSsaoRaw, F4, 2)}, ComputePipelineDescription{{ComputeKernel{"hybrid_render_path/ssao.comp"}}},
                          [](ComputeExecutionContext &ec) {
                              glmlite::uvec2 s = ec.GetDisplaySize();
                              ec.Dispatch("hybrid_render_path/ssao.comp", groups(s.x), groups(s.y), 1);
                          });
        rg.AddComputePass("SSAO Blur Pass", {CreateTransientStorageImage(kSsaoRaw, F4, 0)}, {CreateTransientStorageImage(kSsao, F4, 1)},
                          ComputePipelineDescription{{ComputeKernel{"hybrid_render_path/ssao_blur.comp"}}, PushConstantDescription{sizeof(SSAOPushConstants), 0}},
                          [this](ComputeExecutionContext &ec) {
                              glmlite::uvec2 s = ec.GetDisplaySize();
                              ec.Dispatch("hybrid_render_path/ssao_blur.comp", groups(s.x), groups(s.y), 1, ssao_push_constants);
                          });
    }

    if (reflection_mode == REFLECTION_MODE_SSR) {
        // ---- SSR Pass (:202-243) ----------------------------------------------------------------------------------------
        ssr_push_constants = SSRPushConstants{25.0f, 0.1f, 0.5f, 10};
        rg.AddComputePass("SSR Pass",
                          {CreateTransientSampledImage(kAlbedo, BGRA8, 0), CreateTransientSampledImage(kNormals, F4, 1),
                           CreateTransientSampledImage(kMotion, F4, 2), CreateTransientSampledImage(kDepth, D32, 3)},
                          {CreateTransientStorageImage(kSsr, F4, 4)},
                          ComputePipelineDescription{{ComputeKernel{"hybrid_render_path/ssr.comp"}}, PushConstantDescription{sizeof(SSRPushConstants), 0}},
                          [this](ComputeExecutionContext &ec) {
                              glmlite::uvec2 s = ec.GetDisplaySize();
                              ec.Dispatch("hybrid_render_path/ssr.comp", groups(s.x), groups(s.y), 1, ssr_push_constants);
                          });
    }

    if (denoise_shadow_and_ao && any_rt) {
        // ---- SVGF Denoise Pass (:245-331) -----------------------------------------------------------------------------
        auto &pc = svgf_push_constants;
        pc.integrated_shadow_and_ao.x = (int)rm.UploadNewStorageImage(rm.width, rm.height, F4);
        pc.integrated_shadow_and_ao.y = (int)rm.UploadNewStorageImage(rm.width, rm.height, F4);
        pc.prev_frame_normals_and_object_ids = (int)rm.UploadNewStorageImage(rm.width, rm.height, F4);
        pc.shadow_and_ao_history = (int)rm.UploadNewStorageImage(rm.width, rm.height, F4);
        pc.shadow_and_ao_moments_history = (int)rm.UploadNewStorageImage(rm.width, rm.height, F2);
        svgf_textures_created = true;
        // B200 mode of the node (not in the reference): the pass body below stays the reference's call sequence; the library answers
        // svgf.comp + the step-1 a-trous dispatch with one fused kernel and the three blits by aliasing buffers copy-on-write
        VHR_CHECK(vhr_set_option(rm.ctx, VHR_OPT_SVGF_FUSED, svgf_fused ? 1 : 0));
        VHR_CHECK(vhr_set_option(rm.ctx, VHR_OPT_BLIT_ALIAS, svgf_fused ? 1 : 0));

        rg.AddComputePass(
            "SVGF Denoise Pass",
            {CreateTransientStorageImage(kNormals, F4, 0), CreateTransientStorageImage(kMotion, F4, 1), CreateTransientSampledImage(kDepth, D32, 2),
             CreateTransientStorageImage(kRt, F2, 3)},
            {CreateTransientStorageImage(kDenoised, F4, 4)},
            ComputePipelineDescription{{ComputeKernel{"hybrid_render_path/svgf.comp"}, ComputeKernel{"hybrid_render_path/svgf_atrous_filter.comp"}},
                                       PushConstantDescription{sizeof(SVGFPushConstants), 0}},
            [this](ComputeExecutionContext &ec) {
                auto &pc = svgf_push_constants;
                const glmlite::uvec2 s = ec.GetDisplaySize();
                const uint32_t gx = groups(s.x), gy = groups(s.y);
                ec.Dispatch("hybrid_render_path/svgf.comp", gx, gy, 1, pc);               // temporal accumulation + variance
                for (int i = 0; i < 5; ++i) {                                             // five a-trous iterations, step 2^i
                    pc.atrous_step = 1 << i;
                    ec.Dispatch("hybrid_render_path/svgf_atrous_filter.comp", gx, gy, 1, pc);
                    if (i == 0) ec.BlitImageStorageToStorage(pc.integrated_shadow_and_ao.y, pc.shadow_and_ao_history);   // history = 1st iteration
                    std::swap(pc.integrated_shadow_and_ao.x, pc.integrated_shadow_and_ao.y);
                }
                ec.BlitImageTransientToStorage(kNormals, pc.prev_frame_normals_and_object_ids);
                ec.BlitImageStorageToTransient(pc.integrated_shadow_and_ao.y, kDenoised);   // = iteration 3's output
                std::swap(pc.integrated_shadow_and_ao.x, pc.integrated_shadow_and_ao.y);
            });
    }

    // ---- Composition Pass (:333-379): consumer of the hot path; writes RENDER_OUTPUT, which anchors the execution order.
    rg.AddGraphicsPass(
        "Composition Pass",
        {CreateTransientSampledImage(kAlbedo, BGRA8, 0), CreateTransientSampledImage(kNormals, F4, 1), CreateTransientSampledImage(kMotion, F4, 2),
         CreateTransientSampledImage(kDepth, D32, 3), CreateTransientSampledImage(kShadowMap, 4096, 4096, D32, 4), CreateTransientSampledImage(kSsao, F4, 5),
         CreateTransientSampledImage(kSsr, F4, 6),
         denoise_shadow_and_ao ? CreateTransientSampledImage(kDenoised, F4, 7) : CreateTransientSampledImage(kRt, F2, 7),
         CreateTransientSampledImage(kRefl, F4, 8)},
        {VkUtils::CreateTransientRenderOutput(0)},
        {GraphicsPipelineDescription{"Composition Pipeline", "hybrid_render_path/composition.vert", "hybrid_render_path/composition.frag", PUSHCONSTANTS_NONE,
                                     {shadow_mode, ambient_occlusion_mode, reflection_mode}}},
        [](ExecuteGraphicsCallback execute_pipeline) {
            execute_pipeline("Composition Pipeline", [](GraphicsExecutionContext &ec) { ec.Draw(3, 1, 0, 0); });
        });
}

void HybridRenderPath::DeregisterPath(RenderGraph &, ResourceManager &rm) {
    if (!svgf_textures_created) return;     // hybrid_render_path.cpp:383-392
    rm.DestroyStorageImage(svgf_push_constants.integrated_shadow_and_ao.x);
    rm.DestroyStorageImage(svgf_push_constants.integrated_shadow_and_ao.y);
    rm.DestroyStorageImage(svgf_push_constants.prev_frame_normals_and_object_ids);
    rm.DestroyStorageImage(svgf_push_constants.shadow_and_ao_history);
    rm.DestroyStorageImage(svgf_push_constants.shadow_and_ao_moments_history);
    svgf_textures_created = false;
}
