// raytraced_render_path.h — the fully ray-traced render path's node declarations on the B200 render graph.
// Mirrors src/render_paths/raytraced_render_path.{h,cpp}: "Raytracing Pass" (pipeline "Raytracing Pipeline": raygen + two miss
// shaders + one hit group, or the *_test_alpha shaders with the any-hit alpha test) writing "RaytracedOutput", and the
// "Composition Pass" that copies it to RENDER_OUTPUT.
#pragma once
#include "hybrid_render_path.h"

class RaytracedRenderPath : public RenderPath {
public:
    using RenderPath::RenderPath;
    void RegisterPath(RenderGraph &render_graph, ResourceManager &resource_manager) override;
    void DeregisterPath(RenderGraph &render_graph, ResourceManager &resource_manager) override;

    int use_anyhit_shader = 0;      // raytraced_render_path.h:14 (the "Alpha test for shadows" radio buttons, :80-91)
};
