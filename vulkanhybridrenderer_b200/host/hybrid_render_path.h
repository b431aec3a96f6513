// hybrid_render_path.h — the hybrid path's node declarations on the B200 render graph.
// Mirrors src/render_paths/render_path.{h,cpp} and src/render_paths/hybrid_render_path.{h,cpp}: same pass names, image
// names, formats, bindings, kernel keys, push-constant sizes, mode switches and per-frame call sequence, so a pass here
// is a drop-in for the corresponding reference node. ImGuiDrawSettings has no counterpart (UI is out of scope); the
// public mode members below stand in for its radio buttons — change them and call Rebuild(), like the reference does.
#pragma once
#include "render_graph.h"

enum ShadowMode { SHADOW_MODE_RAYTRACED = 0, SHADOW_MODE_RASTERIZED = 1, SHADOW_MODE_OFF = 2 };
enum AmbientOcclusionMode { AMBIENT_OCCLUSION_MODE_RAYTRACED = 0, AMBIENT_OCCLUSION_MODE_SSAO = 1, AMBIENT_OCCLUSION_MODE_OFF = 2 };
enum ReflectionMode { REFLECTION_MODE_RAYTRACED = 0, REFLECTION_MODE_SSR = 1, REFLECTION_MODE_OFF = 2 };

class RenderPath {
public:
    RenderPath(RenderGraph &render_graph, ResourceManager &resource_manager) : render_graph(render_graph), resource_manager(resource_manager) {}
    virtual ~RenderPath() = default;
    void Build();      // render_path.cpp:14-20
    void Rebuild();    // render_path.cpp:22-27
    virtual void RegisterPath(RenderGraph &render_graph, ResourceManager &resource_manager) = 0;
    virtual void DeregisterPath(RenderGraph &render_graph, ResourceManager &resource_manager) = 0;

protected:
    RenderGraph &render_graph;
    ResourceManager &resource_manager;
};

class HybridRenderPath : public RenderPath {
public:
    using RenderPath::RenderPath;
    void RegisterPath(RenderGraph &render_graph, ResourceManager &resource_manager) override;
    void DeregisterPath(RenderGraph &render_graph, ResourceManager &resource_manager) override;

    // defaults of hybrid_render_path.h:32-35
    int shadow_mode = SHADOW_MODE_RAYTRACED;
    int ambient_occlusion_mode = AMBIENT_OCCLUSION_MODE_OFF;
    int reflection_mode = REFLECTION_MODE_OFF;
    bool denoise_shadow_and_ao = false;
    // B200 addition (not in the reference): VHR_OPT_SVGF_FUSED + VHR_OPT_BLIT_ALIAS for the SVGF node — same call sequence, one fused
    // temporal + a-trous-0 kernel and copy-free blits inside the library.
    bool svgf_fused = false;

    SVGFPushConstants svgf_push_constants{};
    bool svgf_textures_created = false;
    SSRPushConstants ssr_push_constants{};
    SSAOPushConstants ssao_push_constants{};
};
