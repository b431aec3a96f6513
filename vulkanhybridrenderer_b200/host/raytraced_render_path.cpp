// raytraced_render_path.cpp — see raytraced_render_path.h (reference: src/render_paths/raytraced_render_path.cpp:11-78).
#include "raytraced_render_path.h"

void RaytracedRenderPath::RegisterPath(RenderGraph &rg, ResourceManager &rm) {
    // the pipeline's shader set is chosen when the path is (re)built; the CUDA library keys it by this option
    VHR_CHECK(vhr_set_option(rm.ctx, VHR_OPT_RAYTRACED_ALPHA_TEST, use_anyhit_shader ? 1 : 0));
    RaytracingPipelineDescription pipe{
        "Raytracing Pipeline",
        use_anyhit_shader ? "raytraced_render_path/raygen_test_alpha.rgen" : "raytraced_render_path/raygen.rgen",
        {"raytraced_render_path/miss.rmiss", "raytraced_render_path/shadow_miss.rmiss"},
        {use_anyhit_shader ? HitShader{"raytraced_render_path/closesthit_test_alpha.rchit", "raytraced_render_path/shadow_anyhit.rahit"}
                           : HitShader{"raytraced_render_path/closesthit.rchit", nullptr}}};
    rg.AddRaytracingPass("Raytracing Pass", {}, {VkUtils::CreateTransientStorageImage("RaytracedOutput", VK_FORMAT_B8G8R8A8_UNORM, 0)}, pipe,
                         [&rm](ExecuteRaytracingCallback execute_pipeline) {
                             execute_pipeline("Raytracing Pipeline", [&rm](RaytracingExecutionContext &ec) { ec.TraceRays(rm.width, rm.height); });
                         });
    rg.AddGraphicsPass("Composition Pass", {VkUtils::CreateTransientSampledImage("RaytracedOutput", VK_FORMAT_B8G8R8A8_UNORM, 0)},
                       {VkUtils::CreateTransientRenderOutput(0)},
                       {GraphicsPipelineDescription{"Composition Pipeline", "raytraced_render_path/composition.vert", "raytraced_render_path/composition.frag"}},
                       [](ExecuteGraphicsCallback execute_pipeline) {
                           execute_pipeline("Composition Pipeline", [](GraphicsExecutionContext &ec) { ec.Draw(3, 1, 0, 0); });
                       });
}

void RaytracedRenderPath::DeregisterPath(RenderGraph &, ResourceManager &) {}
