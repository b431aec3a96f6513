"""The checker the GPU parity tests compare the CUDA kernels with.

For every shader pass it covers this is **oracle/_ref** — the reference's own GLSL files compiled for the CPU (oracle/make_ref.py,
oracle/ref_shim.h): svgf.comp, svgf_atrous_filter.comp, ssao.comp, ssao_blur.comp, ssr.comp, composition.frag and the "Raytrace
Pipeline" (raygen.rgen + miss shaders + reflection_hit.rchit, with traceRayEXT bound to the oracle's double-precision BVH).
The library is built in the development container (where /root/reference exists) and travels to the GPU box with the snapshot.

The hand-written oracle (oracle/*.cpp) remains the checker for what the reference shaders cannot express: the parametrised ray
pass (ao_spp != 2, ray kinds switched off — raygen.rgen is hard-wired to 2 samples and all three kinds), the G-buffer producer,
hit distances, ray counts, the fully ray-traced path. tests/test_ref_pinning_cpu.py holds it bit-for-bit equal to oracle/_ref
wherever both apply, so a comparison with either is a comparison with the reference text.

Same names as oracle_lib: `import checker as O`.
"""
import numpy as np

import oracle_lib as _O
import ref_lib as _R
from oracle_lib import *  # noqa: F401,F403  (lib, set_num_threads, helpers used by a few tests)
from oracle_lib import _h, _p  # noqa: F401

USING_REF = _R.available()
NAME = "oracle/_ref (reference GLSL compiled for the CPU)" if USING_REF else "oracle/ (hand-written restatement; oracle/_ref library absent)"
print(f"[checker] shader passes are compared with {NAME}")

_impl = _R if USING_REF else _O
svgf_temporal = _impl.svgf_temporal
svgf_atrous = _impl.svgf_atrous
SvgfState = _impl.SvgfState
ssao = _impl.ssao
ssao_blur = _impl.ssao_blur
ssr = _impl.ssr
composition = _impl.composition


class OracleScene(_O.OracleScene):
    """oracle_lib.OracleScene whose raygen() runs the compiled reference pipeline whenever the call IS the reference's configuration
    (2 AO samples, all three ray kinds); hit distances and ray counts still come from the oracle's own loop."""

    def __init__(self, scene):
        super().__init__(scene)
        self._scene = scene

    def raygen(self, pfd, depth, normals, ao_spp=2, flags=7, rows=None, want_t=False):
        if not (USING_REF and ao_spp == 2 and flags == 7):
            return super().raygen(pfd, depth, normals, ao_spp=ao_spp, flags=flags, rows=rows, want_t=want_t)
        out = _R.raygen(self._scene, self, pfd, depth, normals, rows=rows)
        if want_t:
            extra = super().raygen(pfd, depth, normals, ao_spp=ao_spp, flags=flags, rows=rows, want_t=True)
            out["refl_t"] = extra["refl_t"]
            out["rays"] = extra["rays"]
        else:
            H, W = depth.shape[:2]
            y0, y1 = (0, H) if rows is None else rows
            out["rays"] = int(np.count_nonzero(np.asarray(depth)[y0:y1] != 0.0)) * 4     # shadow + 2 AO + reflection per lit pixel (Q3)
        return out
