"""CPU tests of the oracle (test infrastructure) — run with -m "not gpu".

The reference ships no tests or golden vectors for this path (parity unpinned). What pins the C oracle
here is (1) the known-answer vectors of SURVEY Appendix A.3, derived by hand from common.glsl:47-68; (2) a second,
independent restatement of the GLSL written in numpy in THIS file, directly from the shader text
(svgf_atrous_filter.comp:17-103, svgf.comp:16-145, ssao_blur.comp:11-26, common.glsl:29-93), evaluated on small
images; (3) the committed fixtures under tests/golden/ (tools/make_golden.py) which freeze today's oracle outputs so a
later edit of the oracle cannot drift silently.
"""
import numpy as np
import pytest

import helpers as Hh
import oracle_lib as O
from vulkanhybridrenderer_b200 import camera, scenes
from vulkanhybridrenderer_b200 import types as T

f32 = np.float32


# ---------------------------------------------------------------------------------------------------------------
# common.glsl
# ---------------------------------------------------------------------------------------------------------------
def py_seed_thread(seed):          # common.glsl:47-56
    seed &= 0xFFFFFFFF
    seed = (seed ^ 61) ^ (seed >> 16)
    seed = (seed * 9) & 0xFFFFFFFF
    seed = seed ^ (seed >> 4)
    seed = (seed * 0x27d4eb2d) & 0xFFFFFFFF
    seed = seed ^ (seed >> 15)
    return seed


def py_random(state):              # common.glsl:58-64
    state ^= (state << 13) & 0xFFFFFFFF
    state ^= state >> 17
    state ^= (state << 5) & 0xFFFFFFFF
    return state


def py_random01(state):            # common.glsl:66-68
    state = py_random(state)
    return state, np.array(0x3f800000 | (state >> 9), np.uint32).view(np.float32) - f32(1.0)


KATS = {
    0: (0xc0a9496a, [0.847836018, 0.639855146, 0.355129719, 0.101476431]),
    1: (0x27922c9d, [0.133637309, 0.616365790, 0.992120028, 0.394519806]),
    16221: (0x6fff9630, [0.226605177, 0.090955853, 0.430844665, 0.511248469]),
}


@pytest.mark.parametrize("seed", sorted(KATS))
def test_rng_known_answers(seed):
    import ctypes as C
    want_state, want = KATS[seed]
    assert py_seed_thread(seed) == want_state
    assert O.lib().vo_seed_thread(seed) == want_state
    st_py = want_state
    st_c = C.c_uint32(want_state)
    for w in want:
        st_py, r = py_random01(st_py)
        rc = O.lib().vo_random01(C.byref(st_c))
        assert abs(float(r) - w) < 5e-9 and float(r) == rc
        assert st_c.value == st_py


def test_rng_pixel_seed_rule():
    # raygen.rgen:17 (Q4): seed = (y * LaunchSize.y + x) * frame_index — x=7, y=5, H=1080, frame 3 => 16221
    assert (5 * 1080 + 7) * 3 == 16221


def test_half_conversion_matches_numpy_rte():
    allh = np.arange(65536, dtype=np.uint16)
    vals = allh.view(np.float16).astype(np.float32)
    for h in allh[::97]:
        got = O.lib().vo_h2f(int(h))
        want = float(vals[h])
        assert (np.isnan(got) and np.isnan(want)) or got == want
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.standard_normal(4000).astype(f32) * f32(10), rng.uniform(-7e4, 7e4, 500).astype(f32),
                         np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e-8, 6e-8, 2.98e-8, 6.1e-5, np.inf, -np.inf], f32),
                         (rng.uniform(1, 2, 500).astype(f32))])
    for x in xs:
        got = O.lib().vo_f2h(float(x))
        want = int(np.array(x, f32).astype(np.float16).view(np.uint16))
        assert got == want, (x, hex(got), hex(want))


def test_sampling_helpers():
    rng = np.random.default_rng(1)
    out3 = np.zeros(3, f32)
    out9 = np.zeros(9, f32)
    for _ in range(200):
        u0, u1 = f32(rng.uniform()), f32(rng.uniform())
        # common.glsl:29-34
        O.lib().vo_uniform_sample_cone(float(u0), float(u1), 0.999995, O._p(out3))
        ct = (f32(1) - u0) + u0 * f32(0.999995)
        st = np.sqrt(f32(1) - ct * ct)
        phi = u1 * f32(2 * np.pi)
        np.testing.assert_allclose(out3, [np.cos(phi) * st, np.sin(phi) * st, ct], rtol=0, atol=2e-7)
        # common.glsl:37-42
        O.lib().vo_cosine_hemisphere(float(u0), float(u1), O._p(out3))
        r = np.sqrt(u0)
        np.testing.assert_allclose(out3, [r * np.cos(f32(2 * np.pi) * u1), r * np.sin(f32(2 * np.pi) * u1), np.sqrt(f32(1) - u0)], atol=2e-7)
        assert abs(np.linalg.norm(out3) - 1) < 1e-6
        # common.glsl:80-93: columns (b1, b2, n) orthonormal
        n = rng.standard_normal(3); n = (n / np.linalg.norm(n)).astype(f32)
        O.lib().vo_onb(O._p(n), O._p(out9))
        M = out9.reshape(3, 3)          # column-major: M[c] = column c
        np.testing.assert_allclose(M[2], n, atol=0)
        np.testing.assert_allclose(M @ M.T, np.eye(3), atol=5e-5)   # Frisvad amplifies rounding as n.z -> -1
    n = np.array([0, 0, -1], f32)       # the z < -0.9999999 branch
    O.lib().vo_onb(O._p(n), O._p(out9))
    np.testing.assert_array_equal(out9.reshape(3, 3), [[0, -1, 0], [-1, 0, 0], [0, 0, -1]])


def test_struct_sizes():
    # glsl_common.h sizes measured with g++ on the reference header
    want = {0: 584, 1: 56, 2: 44, 3: 120, 4: 112}   # PerFrameData, Vertex, Material, Primitive, DirectionalLight
    for k, v in want.items():
        assert O.lib().vo_sizeof(k) == v


# ---------------------------------------------------------------------------------------------------------------
# Independent numpy restatement of svgf_atrous_filter.comp (vectorised over the image, fp32)
# ---------------------------------------------------------------------------------------------------------------
def np_atrous(normals_h, integ_h, step):
    H, W = normals_h.shape[:2]
    n = normals_h.astype(f32)
    I = integ_h.astype(f32)
    ids = np.trunc(n[..., 3]).astype(np.int32)
    yy, xx = np.mgrid[0:H, 0:W]

    def shifted(a, dx, dy, fill=0):
        out = np.full_like(a, fill)
        ys, xs = yy + dy, xx + dx
        ok = (ys >= 0) & (ys < H) & (xs >= 0) & (xs < W)
        out[ok] = a[ys[ok], xs[ok]]
        return out, ok

    gauss = np.array([[1 / 16, 1 / 8, 1 / 16], [1 / 8, 1 / 4, 1 / 8], [1 / 16, 1 / 8, 1 / 16]], f32)
    var = np.zeros((H, W, 2), f32)
    for y in (-1, 0, 1):
        for x in (-1, 0, 1):
            q, ok = shifted(I[..., 2:4], x, y)
            var = var + np.where(ok[..., None], gauss[y + 1, x + 1] * q, f32(0)).astype(f32)
    aw = np.array([1 / 16, 1 / 4, 3 / 8, 1 / 4, 1 / 16], f32)
    sum_w = np.ones((H, W, 2), f32)
    acc = I.copy()
    den = f32(4.0) * np.sqrt(var) + f32(1e-6)
    for y in range(-2, 3):
        for x in range(-2, 3):
            if x == 0 and y == 0:
                continue
            q, ok = shifted(I, x * step, y * step)
            nq, _ = shifted(n, x * step, y * step)
            idq, _ = shifted(ids, x * step, y * step, fill=-12345)
            kern = f32(aw[y + 2] * aw[x + 2])
            d = (n[..., 0] * nq[..., 0] + n[..., 1] * nq[..., 1]) + n[..., 2] * nq[..., 2]
            with np.errstate(all="ignore"):
                wn = np.where(d > 0, np.power(np.maximum(d, f32(0)), f32(128.0)), f32(0)).astype(f32)   # Q10
            wid = (ids == idq).astype(f32)
            w0 = (kern * wn * wid).astype(f32)
            e = np.abs(I[..., 0:2] - q[..., 0:2]) / den
            w = (w0[..., None] * np.exp(-e)).astype(f32)
            w = np.where(ok[..., None], w, f32(0))
            sum_w = sum_w + w
            acc = acc + np.concatenate([w, w * w], -1) * q
    out = acc / np.concatenate([sum_w, sum_w * sum_w], -1)
    return out.astype(np.float16)


@pytest.mark.parametrize("step", [1, 2, 4, 16])
def test_atrous_oracle_vs_numpy_restatement(step):
    W, H = 96, 56
    _, _, frames = Hh.scene_and_gbuffer(W, H, tris=6000)
    pfd, g = frames[0]
    integ = Hh.noise_integrated(H, W, seed=7)
    ref = O.svgf_atrous(pfd, g["normals"], integ, step)
    mine = np_atrous(g["normals"], integ, step)
    s = Hh.compare(mine, ref)
    # two fp32 evaluations with different summation order / libm: agreement to ~1 half ulp
    assert s["nan_mismatch"] == 0 and s["max_abs"] <= 1e-3 and s["exact"] > 0.98, s


def test_atrous_properties():
    W, H = 64, 40
    _, _, frames = Hh.scene_and_gbuffer(W, H, tris=6000)
    pfd, g = frames[0]
    # constant image with zero variance stays constant wherever a normal exists; sky pixels (n = 0) pass through (Q10)
    integ = np.zeros((H, W, 4), np.float16)
    integ[..., 0] = 0.5; integ[..., 1] = 0.25
    out = O.svgf_atrous(pfd, g["normals"], integ, 2)
    np.testing.assert_array_equal(out[..., :2], integ[..., :2])
    # a pixel whose object id differs from every neighbour is returned unchanged (centre weight 1, others 0)
    normals = g["normals"].copy()
    normals[..., 3] = np.arange(W * H, dtype=np.float32).reshape(H, W) % 2048   # unique-ish ids within any 5x5 window
    noisy = Hh.noise_integrated(H, W, seed=9)
    out = O.svgf_atrous(pfd, normals, noisy, 1)
    np.testing.assert_array_equal(out, noisy)


# ---------------------------------------------------------------------------------------------------------------
# Independent restatement of svgf.comp for the zero-motion case + structural properties of the pass
# ---------------------------------------------------------------------------------------------------------------
def test_temporal_zero_history_passthrough_and_q2():
    W, H = 48, 32
    _, _, frames = Hh.scene_and_gbuffer(W, H, tris=6000)
    pfd, g = frames[0]
    rng = np.random.default_rng(3)
    rt = np.stack([rng.integers(0, 2, (H, W)), rng.integers(0, 3, (H, W)) * 0.5], -1).astype(np.float16)
    zero4 = np.zeros((H, W, 4), np.float16)
    zero2 = np.zeros((H, W, 2), np.float16)
    # frame 0: zero prev-normals reject every reprojection (Q14) => pass-through, variance 0, moments (l, l^2)
    integ, mom = O.svgf_temporal(pfd, g["normals"], g["motion"], rt, zero4, zero4, zero2)
    np.testing.assert_array_equal(integ[..., :2], rt)
    assert not integ[..., 2:].any()
    np.testing.assert_array_equal(mom[..., 0], rt[..., 0])
    np.testing.assert_array_equal(mom[..., 1], (rt[..., 0].astype(f32) ** 2).astype(np.float16))
    # frame 1, zero motion, history = constant c, prev normals = current: interior pixels of large flat regions get
    # mix(c, cur, 0.2); AO variance follows Q2: moments (0.2a, 0.8 + 0.2 a^2)
    motion = np.zeros((H, W, 4), np.float16)
    hist = np.zeros((H, W, 4), np.float16); hist[..., 0] = 0.5; hist[..., 1] = 0.25
    integ, mom = O.svgf_temporal(pfd, g["normals"], motion, rt, g["normals"], hist, mom)
    n = g["normals"].astype(f32)
    ids = n[..., 3]
    same = np.ones((H, W), bool)
    for dy in (0, 1):
        for dx in (0, 1):
            sh = np.roll(np.roll(ids, -dy, 0), -dx, 1)
            nn = np.roll(np.roll(n[..., :3], -dy, 0), -dx, 1)
            same &= (sh == ids) & (np.sum(nn * n[..., :3], -1) >= np.cos(np.pi / 4))
    same[-1, :] = False; same[:, -1] = False
    same &= np.linalg.norm(n[..., :3], axis=-1) > 0.5
    assert same.sum() > 100
    cur = rt.astype(f32)
    want_s = (f32(0.5) * f32(0.8) + cur[..., 0] * f32(0.2)).astype(np.float16)
    want_a = (f32(0.25) * f32(0.8) + cur[..., 1] * f32(0.2)).astype(np.float16)
    assert np.max(np.abs(integ[..., 0][same].astype(f32) - want_s[same].astype(f32))) <= 5e-4
    assert np.max(np.abs(integ[..., 1][same].astype(f32) - want_a[same].astype(f32))) <= 5e-4
    a = cur[..., 1]
    am0, am1 = f32(0.2) * a, f32(0.8) + f32(0.2) * a * a
    want_var_a = np.maximum(0, am1 - am0 * am0)
    assert np.max(np.abs(integ[..., 3][same].astype(f32) - want_var_a[same])) <= 1e-3


def test_svgf_pass_bookkeeping_q1():
    """Denoised == output of iteration index 3; history == iteration 0 output; prev normals == normals (hybrid_render_path.cpp:299-328)."""
    W, H = 64, 40
    _, _, frames = Hh.scene_and_gbuffer(W, H, tris=6000)
    st = O.SvgfState(W, H)
    rng = np.random.default_rng(4)
    for f in range(2):
        pfd, g = frames[f]
        rt = np.stack([rng.integers(0, 2, (H, W)), rng.integers(0, 3, (H, W)) * 0.5], -1).astype(np.float16)
        den, iters, temporal = st.run(pfd, g["normals"], g["motion"], rt)
        np.testing.assert_array_equal(den, iters[3])
        # the pass is the composition of the two kernels
        it = temporal
        for i in range(5):
            it = O.svgf_atrous(pfd, g["normals"], it, 1 << i)
            np.testing.assert_array_equal(it, iters[i])


# ---------------------------------------------------------------------------------------------------------------
# SSAO blur restatement (ssao_blur.comp:11-26) and ray oracle vs brute force
# ---------------------------------------------------------------------------------------------------------------
def test_ssao_blur_vs_numpy():
    W, H = 40, 30
    rng = np.random.default_rng(5)
    raw = np.repeat(rng.uniform(0, 1, (H, W, 1)).astype(np.float16), 4, -1)
    pfd = np.zeros((), T.PerFrameData)
    pfd["display_size"] = (W, H)
    got = O.ssao_blur(pfd, raw).astype(f32)
    x = raw[..., 0].astype(np.float64)
    pad = np.zeros((H + 12, W + 12)); pad[6:-6, 6:-6] = x
    want = sum(pad[6 + dy:6 + dy + H, 6 + dx:6 + dx + W] for dy in range(-6, 7) for dx in range(-6, 7)) / 169.0
    assert np.max(np.abs(got[..., 0] - want)) <= 1e-3
    for c in range(1, 4):
        np.testing.assert_array_equal(got[..., c], got[..., 0])


def test_ray_oracle_vs_brute_force():
    sc = scenes.tiny_scene()
    osc = O.OracleScene(sc)
    tris = Hh.world_triangles(sc)
    assert osc.num_triangles == len(tris) == sc.num_triangles
    rng = np.random.default_rng(6)
    lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
    n_hit = 0
    for _ in range(300):
        o = rng.uniform(lo - 0.5, hi + 0.5).astype(f32)
        d = rng.standard_normal(3); d = (d / np.linalg.norm(d)).astype(f32)
        tmax = float(rng.choice([5.0, 1e4]))
        bf_any, bf_t, margin = Hh.brute_force_hits(tris, o, d, 0.01, tmax)
        if margin < 1e-6:
            continue    # grazing: either answer is legitimate
        assert osc.trace_any(o, d, 0.01, tmax) == bf_any
        cl = osc.trace_closest(o, d, 0.01, tmax)
        assert (cl is not None) == bf_any
        if cl is not None:
            n_hit += 1
            assert abs(cl[0][0] - bf_t) <= 1e-5 * max(1.0, bf_t)
    assert n_hit > 30


def test_raygen_sky_and_band_consistency():
    W, H = 64, 48
    sc = scenes.tiny_scene(width=W, height=H)
    osc = O.OracleScene(sc)
    seq = camera.FrameSequencer(W, H, sc.light)
    pfd = seq.next(sc.camera)
    g = osc.gbuffer(pfd, W, H)
    full = osc.raygen(pfd, g["depth"], g["normals"])
    sky = g["depth"] == 0
    assert sky.any() and (~sky).any()
    assert np.all(full["shadow_ao"][sky].astype(f32) == 1.0)          # raygen.rgen:20-24
    assert not full["reflections"][sky].any()
    assert full["rays"] == int((~sky).sum()) * 4                        # 1 shadow + 2 AO + 1 reflection (Q3)
    # row bands stitch to the full frame (the multi-GPU split relies on this)
    top = osc.raygen(pfd, g["depth"], g["normals"], rows=(0, 20))
    bot = osc.raygen(pfd, g["depth"], g["normals"], rows=(20, H))
    np.testing.assert_array_equal(np.concatenate([top["shadow_ao"][:20], bot["shadow_ao"][20:]]), full["shadow_ao"])
    np.testing.assert_array_equal(np.concatenate([top["reflections"][:20], bot["reflections"][20:]]), full["reflections"])


def test_oracle_thread_count_can_be_set_explicitly():
    """bench.py's CPU arm sets the OpenMP thread count itself (torchrun exports OMP_NUM_THREADS=1 to every rank) and reports what
    the runtime then uses."""
    before = O.set_num_threads(0)
    assert before >= 1
    assert O.set_num_threads(2) == 2
    assert O.set_num_threads(0) == 2
    O.set_num_threads(before)


# ---------------------------------------------------------------------------------------------------------------
# Independent scalar transcription of the WHOLE of svgf.comp (motion, 2x2 taps, 3x3 retry), written from the shader text
# ---------------------------------------------------------------------------------------------------------------
def _svgf_comp_scalar(pfd, normals, motion, rt, prev_normals, history, moments):
    """svgf.comp:16-145 one pixel at a time in float32, separate multiplies and adds, GLSL conversions (int() and ivec2() truncate,
    fract = x - floor(x), mix(a, b, t) = a (1 - t) + b t). The moments image is RG16F: a load returns (x, y, 0, 1), a store keeps .xy."""
    H, W = rt.shape[:2]
    dsx, dsy = f32(pfd["display_size"][0]), f32(pfd["display_size"][1])
    n32, m32, r32 = normals.astype(f32), motion.astype(f32), rt.astype(f32)
    pn32, h32, mo32 = prev_normals.astype(f32), history.astype(f32), moments.astype(f32)
    integ = np.zeros((H, W, 4), np.float16)
    mom_out = np.zeros((H, W, 2), np.float16)
    one, cos_pi_4 = f32(1), f32(0.70710678118654752440084)

    def valid(px, py, cur_id, cur_n):
        if px < 0 or py < 0 or f32(px) >= dsx or f32(py) >= dsy:
            return False
        p = pn32[py, px]
        if cur_id != int(p[3]):
            return False
        d = f32(f32(f32(cur_n[0] * p[0]) + f32(cur_n[1] * p[1])) + f32(cur_n[2] * p[2]))
        return not (d < cos_pi_4)

    for cy in range(H):
        for cx in range(W):
            cur_n, cur_id = n32[cy, cx, :3], int(n32[cy, cx, 3])
            mvx, mvy = m32[cy, cx, 0], m32[cy, cx, 1]
            cs, ca = r32[cy, cx, 0], r32[cy, cx, 1]
            pcx = f32(f32(f32(cx) - f32(mvx * dsx)) + f32(0.5))
            pcy = f32(f32(f32(cy) - f32(mvy * dsy)) + f32(0.5))
            x, y = f32(pcx - np.floor(pcx)), f32(pcy - np.floor(pcy))
            ax, ay = int(pcx), int(pcy)                                     # ivec2(): towards zero
            w4 = [f32(f32(one - x) * f32(one - y)), f32(x * f32(one - y)), f32(f32(one - x) * y), f32(x * y)]
            ps = pa = f32(0); psm = [f32(0), f32(0)]; pam = [f32(0), f32(0)]; s = f32(0)
            for i, (ox, oy) in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
                sx, sy = ax + ox, ay + oy
                if valid(sx, sy, cur_id, cur_n):
                    hv, mv = h32[sy, sx], (mo32[sy, sx, 0], mo32[sy, sx, 1], f32(0), f32(1))
                    ps = f32(ps + f32(w4[i] * hv[0])); pa = f32(pa + f32(w4[i] * hv[1]))
                    psm = [f32(psm[0] + f32(w4[i] * mv[0])), f32(psm[1] + f32(w4[i] * mv[1]))]
                    pam = [f32(pam[0] + f32(w4[i] * mv[2])), f32(pam[1] + f32(w4[i] * mv[3]))]
                    s = f32(s + w4[i])
            ok = s > f32(1e-6)
            if not ok:                                                      # accumulators are NOT reset (svgf.comp:81-97)
                for oy in (-1, 0, 1):
                    for ox in (-1, 0, 1):
                        sx, sy = ax + ox, ay + oy
                        if valid(sx, sy, cur_id, cur_n):
                            hv, mv = h32[sy, sx], (mo32[sy, sx, 0], mo32[sy, sx, 1], f32(0), f32(1))
                            ps = f32(ps + hv[0]); pa = f32(pa + hv[1])
                            psm = [f32(psm[0] + mv[0]), f32(psm[1] + mv[1])]
                            pam = [f32(pam[0] + mv[2]), f32(pam[1] + mv[3])]
                            s = f32(s + one)
                ok = s > f32(1e-6)
            sm = [cs, f32(cs * cs)]; am = [ca, f32(ca * ca)]
            mix = lambda a, b, t: f32(f32(a * f32(one - t)) + f32(b * t))
            if ok:
                al = f32(0.2)
                ps = f32(ps / s); pa = f32(pa / s)
                psm = [f32(psm[0] / s), f32(psm[1] / s)]; pam = [f32(pam[0] / s), f32(pam[1] / s)]
                sm = [mix(psm[0], sm[0], al), mix(psm[1], sm[1], al)]
                am = [mix(pam[0], am[0], al), mix(pam[1], am[1], al)]
                vs = max(f32(0), f32(sm[1] - f32(sm[0] * sm[0]))); va = max(f32(0), f32(am[1] - f32(am[0] * am[0])))
                integ[cy, cx] = (mix(ps, cs, al), mix(pa, ca, al), vs, va)
            else:
                vs = max(f32(0), f32(sm[1] - f32(sm[0] * sm[0]))); va = max(f32(0), f32(am[1] - f32(am[0] * am[0])))
                integ[cy, cx] = (cs, ca, vs, va)
            mom_out[cy, cx] = (sm[0], sm[1])
    return integ, mom_out


def test_temporal_oracle_vs_scalar_transcription_with_motion():
    W, H = 40, 28
    _, _, frames = Hh.scene_and_gbuffer(W, H, tris=6000, moving=True)
    (_, g0), (pfd, g1) = frames[0], frames[1]
    rng = np.random.default_rng(9)
    rt = np.stack([rng.integers(0, 2, (H, W)), rng.integers(0, 3, (H, W)) * 0.5], -1).astype(np.float16)
    hist = np.zeros((H, W, 4), np.float16)
    hist[..., :2] = rng.uniform(0, 1, (H, W, 2))
    mom = rng.uniform(0, 1, (H, W, 2)).astype(np.float16)
    # larger motion than the 0.05-unit camera step gives, so taps leave the image and the 3x3 retry is exercised
    motion = g1["motion"].copy()
    motion[..., 0] = (motion[..., 0].astype(f32) * f32(6) + f32(0.02)).astype(np.float16)
    motion[..., 1] = (motion[..., 1].astype(f32) * f32(6) - f32(0.03)).astype(np.float16)
    got_i, got_m = O.svgf_temporal(pfd, g1["normals"], motion, rt, g0["normals"], hist, mom)
    want_i, want_m = _svgf_comp_scalar(pfd, g1["normals"], motion, rt, g0["normals"], hist, mom)
    valid_px = (got_i[..., 0] != rt[..., 0]) | (got_i[..., 2] != 0)
    assert 0.2 < valid_px.mean() < 0.98                  # both outcomes of the reprojection occur
    np.testing.assert_array_equal(got_i.view(np.uint16), want_i.view(np.uint16))
    np.testing.assert_array_equal(got_m.view(np.uint16), want_m.view(np.uint16))


# ---------------------------------------------------------------------------------------------------------------
# Independent scalar transcription of ssao.comp (texel-corner uv, LINEAR / REPEAT taps, 16 disk samples)
# ---------------------------------------------------------------------------------------------------------------
def _texture_linear_repeat(img, u, v):
    """texture() through the default sampler (LINEAR, REPEAT, one mip): Vulkan spec 16.6-16.8 in float32."""
    H, W = img.shape[:2]
    # filter coordinate in fixed point with 8 fractional bits, rounded to nearest (subTexelPrecisionBits = 8)
    x, y = f32(f32(u * f32(W)) - f32(0.5)), f32(f32(v * f32(H)) - f32(0.5))
    x, y = f32(np.floor(f32(f32(x * f32(256.0)) + f32(0.5))) * f32(0.00390625)), f32(np.floor(f32(f32(y * f32(256.0)) + f32(0.5))) * f32(0.00390625))
    fx, fy = np.floor(x), np.floor(y)
    a, b = f32(x - fx), f32(y - fy)
    x0, y0 = int(fx) % W, int(fy) % H
    x1, y1 = (int(fx) + 1) % W, (int(fy) + 1) % H
    t00, t10, t01, t11 = (img[y0, x0].astype(f32), img[y0, x1].astype(f32), img[y1, x0].astype(f32), img[y1, x1].astype(f32))
    oma, omb = f32(f32(1) - a), f32(f32(1) - b)
    return ((oma * omb).astype(f32) * t00 + (a * omb).astype(f32) * t10 + (oma * b).astype(f32) * t01 + (a * b).astype(f32) * t11).astype(f32)


def _ssao_comp_scalar(pfd, depth, normals, radius):
    H, W = depth.shape
    inv = np.asarray(pfd["camera_proj_inverse"], f32).reshape(4, 4).T          # column-major storage -> math layout
    view3 = np.asarray(pfd["camera_view"], f32).reshape(4, 4).T[:3, :3]
    dsi = np.asarray(pfd["display_size_inverse"], f32)
    ds_y, frame = np.uint32(int(pfd["display_size"][1])), np.uint32(int(pfd["frame_index"]))
    out = np.zeros((H, W), f32)

    def gl_max(a, b):          # max() with a NaN operand returns the other one on NVIDIA hardware (a sample that lands on the sky:
        return b if a != a else max(a, b)      # depth 0 -> w = 0 -> inf / NaN position), the behaviour the oracle restates

    def view_pos(d, u, v):
        p = inv @ np.array([f32(u * f32(2) - f32(1)), f32(v * f32(2) - f32(1)), d, f32(1)], f32)
        return (p[:3] / p[3]).astype(f32)

    with np.errstate(all="ignore"):
        for gy in range(H):
            for gx in range(W):
                u, v = f32(f32(gx) * dsi[0]), f32(f32(gy) * dsi[1])
                d = _texture_linear_repeat(depth[..., None], u, v)[0]
                if d == 0:
                    continue
                P = view_pos(d, u, v)
                N = (view3 @ _texture_linear_repeat(normals, u, v)[:3]).astype(f32)
                pr = f32(f32(radius) / P[2])
                s = np.uint32((np.uint32(gy) * ds_y + np.uint32(gx)) * frame)
                s = np.uint32((s ^ np.uint32(61)) ^ (s >> np.uint32(16))); s = np.uint32(s * np.uint32(9))
                s = np.uint32(s ^ (s >> np.uint32(4))); s = np.uint32(s * np.uint32(0x27d4eb2d)); s = np.uint32(s ^ (s >> np.uint32(15)))

                def rnd01():
                    nonlocal s
                    s = np.uint32(s ^ np.uint32(s << np.uint32(13))); s = np.uint32(s ^ (s >> np.uint32(17))); s = np.uint32(s ^ np.uint32(s << np.uint32(5)))
                    return f32(np.array([np.uint32(0x3f800000) | (s >> np.uint32(9))], np.uint32).view(f32)[0] - f32(1))
                acc = f32(0)
                for _ in range(16):
                    ang = f32(f32(rnd01() * f32(2)) * f32(np.pi))
                    dist = f32(rnd01() * pr)
                    su, sv = f32(u + f32(np.cos(ang) * dist)), f32(v + f32(np.sin(ang) * dist))
                    V = (view_pos(_texture_linear_repeat(depth[..., None], su, sv)[0], su, sv) - P).astype(f32)
                    acc = f32(acc + f32(gl_max(f32(f32(V @ N) - f32(1e-4)), f32(0)) / f32(f32(V @ V) + f32(1e-4))))
                out[gy, gx] = gl_max(f32(f32(1) - f32(f32(0.125) * acc)), f32(0))
    return out


def test_ssao_oracle_vs_scalar_transcription():
    W, H = 36, 24
    _, _, frames = Hh.scene_and_gbuffer(W, H, tris=6000)
    pfd, g = frames[1]                                   # frame_index > 0: per-pixel seeds differ (Q5)
    got = O.ssao(pfd, g["depth"], g["normals"], 0.75).astype(f32)[..., 0]
    want = _ssao_comp_scalar(pfd, g["depth"], g["normals"], 0.75)
    diff = np.abs(got - want)
    # same algorithm, different libm and summation order: a sample whose bilinear tap straddles a depth edge amplifies an ulp of
    # its position through znear / depth, so a few pixels may move; everything else agrees to fp16 rounding
    assert (diff <= 1e-3).mean() >= 0.99, (diff <= 1e-3).mean()
    assert (got == want.astype(np.float16).astype(f32)).mean() >= 0.97      # measured: every pixel equal after the fp16 store
    assert (got > 0).mean() > 0.3 and (got < 1).mean() > 0.3     # a real AO image, not a constant


# ---------------------------------------------------------------------------------------------------------------
# Independent transcription of raygen.rgen's ray generation (seed rule, cone / hemisphere samples, Frisvad basis, origin offset)
# with brute-force float64 visibility: pins which rays the oracle traces, not only how it intersects them
# ---------------------------------------------------------------------------------------------------------------
def _onb(n):                                           # common.glsl:80-93, columns M[0], M[1], M[2]
    if n[2] < f32(-0.9999999):
        return np.array([0, -1, 0], f32), np.array([-1, 0, 0], f32), n
    a = f32(f32(1) / f32(f32(1) + n[2]))
    b = f32(f32(-n[0] * n[1]) * a)
    return (np.array([f32(f32(1) - f32(f32(n[0] * n[0]) * a)), b, -n[0]], f32),
            np.array([b, f32(f32(1) - f32(f32(n[1] * n[1]) * a)), -n[1]], f32), n)


def _raygen_rays_scalar(pfd, depth, normals, x, y):
    """raygen.rgen:14-55 for one pixel -> [(origin, direction, tmax)] for the shadow ray and the two AO rays, or None for sky."""
    H, W = depth.shape
    d = depth[y, x]                                                          # LINEAR sampler at the texel centre = the texel
    if d == 0:
        return None
    u, v = f32(f32(x + f32(0.5)) / f32(W)), f32(f32(y + f32(0.5)) / f32(H))
    inv = np.asarray(pfd["camera_viewproj_inverse"], f32).reshape(4, 4).T
    p = inv @ np.array([f32(u * f32(2) - f32(1)), f32(v * f32(2) - f32(1)), d, f32(1)], f32)
    P = (p[:3] / p[3]).astype(f32)
    L = (-np.asarray(pfd["directional_light"]["direction"], f32)[:3]).astype(f32)
    N = normals[y, x, :3].astype(f32)
    origin = (P + N * f32(0.1)).astype(f32)
    s = np.uint32((np.uint32(y) * np.uint32(H) + np.uint32(x)) * np.uint32(int(pfd["frame_index"])))     # LaunchSize.y, not .x
    s = np.uint32((s ^ np.uint32(61)) ^ (s >> np.uint32(16))); s = np.uint32(s * np.uint32(9))
    s = np.uint32(s ^ (s >> np.uint32(4))); s = np.uint32(s * np.uint32(0x27d4eb2d)); s = np.uint32(s ^ (s >> np.uint32(15)))
    state = [s]

    def rnd01():
        t = state[0]
        t = np.uint32(t ^ np.uint32(t << np.uint32(13))); t = np.uint32(t ^ (t >> np.uint32(17))); t = np.uint32(t ^ np.uint32(t << np.uint32(5)))
        state[0] = t
        return f32(np.array([np.uint32(0x3f800000) | (t >> np.uint32(9))], np.uint32).view(f32)[0] - f32(1))
    two_pi = f32(6.28318530717958647692528)
    r1, r2 = rnd01(), rnd01()
    ct = f32(f32(f32(1) - r1) + f32(r1 * f32(0.999995)))
    st = f32(np.sqrt(f32(f32(1) - f32(ct * ct))))
    phi = f32(r2 * two_pi)
    cone = np.array([f32(np.cos(phi) * st), f32(np.sin(phi) * st), ct], f32)
    cone = (cone / f32(np.sqrt(f32(cone @ cone)))).astype(f32)
    m0, m1, m2 = _onb(L)
    rays = [(origin, (m0 * cone[0] + m1 * cone[1] + m2 * cone[2]).astype(f32), 10000.0)]
    for _ in range(2):
        r1, r2 = rnd01(), rnd01()
        hx = f32(f32(np.sqrt(r1)) * f32(np.cos(f32(two_pi * r2))))
        hy = f32(f32(np.sqrt(r1)) * f32(np.sin(f32(two_pi * r2))))
        hz = f32(np.sqrt(f32(f32(1) - r1)))
        m0, m1, m2 = _onb(N)
        rays.append((origin, (m0 * hx + m1 * hy + m2 * hz).astype(f32), 5.0))
    return rays


def test_raygen_oracle_vs_scalar_ray_generation_and_brute_force():
    W, H = 40, 30
    sc = scenes.tiny_scene(width=W, height=H)
    osc = O.OracleScene(sc)
    tris = Hh.world_triangles(sc)
    seq = camera.FrameSequencer(W, H, sc.light)
    seq.next(sc.camera)
    pfd = seq.next(sc.camera)                              # frame_index 1: per-pixel seeds (frame 0 seeds every pixel alike, Q5)
    assert int(pfd["frame_index"]) > 0
    g = osc.gbuffer(pfd, W, H)
    got = osc.raygen(pfd, g["depth"], g["normals"], ao_spp=2, flags=3)["shadow_ao"].astype(f32)
    checked = agree_s = agree_a = 0
    for y in range(0, H, 2):
        for x in range(0, W, 2):
            with np.errstate(over="ignore"):             # uint32 wrap-around is the point of the hash
                rays = _raygen_rays_scalar(pfd, g["depth"], g["normals"], x, y)
            if rays is None:
                assert got[y, x, 0] == 1.0 and got[y, x, 1] == 1.0
                continue
            res, grazing = [], False
            for o, d, tmax in rays:
                hit, _, margin = Hh.brute_force_hits(tris, o, d, 0.01, tmax)
                grazing |= margin < 1e-5
                res.append(0.0 if hit else 1.0)            # miss.rmiss writes 1 (visible), a hit leaves the payload at 0
            if grazing:
                continue
            checked += 1
            agree_s += got[y, x, 0] == res[0]
            agree_a += got[y, x, 1] == (res[1] + res[2]) / 2
    assert checked > 100
    assert agree_s == checked and agree_a == checked, (checked, agree_s, agree_a)


# ---------------------------------------------------------------------------------------------------------------
# Independent transcription of the reflection ray + reflection_hit.rchit (float64, constant materials)
# ---------------------------------------------------------------------------------------------------------------
def _closest_hit_f64(tris64, o, d, tmin, tmax):
    e1 = tris64[:, 1] - tris64[:, 0]; e2 = tris64[:, 2] - tris64[:, 0]
    pv = np.cross(d[None, :], e2)
    det = np.einsum("ij,ij->i", e1, pv)
    ok = np.abs(det) > 0
    inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
    tv = o[None, :] - tris64[:, 0]
    u = np.einsum("ij,ij->i", tv, pv) * inv
    qv = np.cross(tv, e1)
    v = np.einsum("j,ij->i", d, qv) * inv
    t = np.einsum("ij,ij->i", e2, qv) * inv
    inside = ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > tmin) & (t < tmax)
    if not inside.any():
        return None
    tt = np.where(inside, t, np.inf)
    k = int(np.argmin(tt))
    second = np.partition(tt, 1)[1] if len(tt) > 1 else np.inf
    edge = min(u[k], v[k], 1 - u[k] - v[k])
    return k, float(u[k]), float(v[k]), float(tt[k]), float(second - tt[k]), float(edge)


def _reflection_hit_f64(pfd, sc, g, prim_id, b1, b2, cam):
    """reflection_hit.rchit:10-72 in float64 for a primitive with constant material (no textures)."""
    p = sc.primitives[g]
    base = int(p["index_offset"]) + 3 * prim_id
    vi = [int(p["vertex_offset"]) + int(sc.indices[base + k]) for k in range(3)]
    bary = np.array([1.0 - b1 - b2, b1, b2])
    normal = sum(sc.vertices["normal"][vi[k]].astype(np.float64) * bary[k] for k in range(3))      # not transformed, not normalised
    lp = sum(sc.vertices["pos"][vi[k]].astype(np.float64) * bary[k] for k in range(3))
    M = np.asarray(p["transform"], np.float64).reshape(4, 4).T
    position = (M @ np.append(lp, 1.0))[:3]
    m = p["material"]
    assert int(m["base_color_texture"]) == -1 and int(m["metallic_roughness_texture"]) == -1
    albedo = np.asarray(m["base_color"], np.float64)[:3]
    metallic, roughness = float(m["metallic_factor"]), float(m["roughness_factor"])
    V = cam - position; V /= np.linalg.norm(V)
    L = -np.asarray(pfd["directional_light"]["direction"], np.float64)[:3]
    N = normal
    Hv = L + V; Hv /= np.linalg.norm(Hv)
    roughness = min(max(roughness, 0.04), 1.0); metallic = min(max(metallic, 0.0), 1.0)
    f0 = 0.04 * (1 - metallic) + albedo * metallic
    hv = max(Hv @ V, 0.0)
    F = f0 + (1 - f0) * (1 - hv) ** 5
    a2 = roughness * roughness
    ndh = max(N @ Hv, 0.0)
    f = ndh * ndh * (a2 - 1) + 1
    D = a2 / (np.pi * f * f)
    k = (roughness + 1) ** 2 * 0.125
    ndv, ndl = max(N @ V, 0.0), max(N @ L, 0.0)
    G = ndv / (ndv * (1 - k) + k) * (ndl / (ndl * (1 - k) + k))
    spec = D * G * F / max(4.0 * ndv * ndl, 1e-6)
    diff = (1 - F) * (1 - metallic) * albedo / np.pi
    li = np.asarray(pfd["directional_light"]["intensity"], np.float64)[:3]
    lc = np.asarray(pfd["directional_light"]["color"], np.float64)[:3]
    return albedo * (0.2 / np.pi) + (diff + spec) * ndl * li * lc


def test_reflection_oracle_vs_scalar_transcription():
    W, H = 40, 30
    sc = scenes.tiny_scene(width=W, height=H)
    osc = O.OracleScene(sc)
    tris64 = Hh.world_triangles(sc).astype(np.float64)
    counts = np.array([int(p["index_count"]) // 3 for p in sc.primitives])
    first = np.concatenate([[0], np.cumsum(counts)])
    seq = camera.FrameSequencer(W, H, sc.light)
    pfd = seq.next(sc.camera)
    g = osc.gbuffer(pfd, W, H)
    got = osc.raygen(pfd, g["depth"], g["normals"], ao_spp=2, flags=4)["reflections"].astype(np.float64)
    inv = np.asarray(pfd["camera_viewproj_inverse"], np.float64).reshape(4, 4).T
    cam = np.asarray(pfd["camera_view_inverse"], np.float64).reshape(4, 4).T[:3, 3]
    n_hit = n_miss = 0
    for y in range(H):
        for x in range(W):
            d = float(g["depth"][y, x])
            if d == 0:
                assert not got[y, x].any()
                continue
            u, v = (x + 0.5) / W, (y + 0.5) / H
            p = inv @ np.array([u * 2 - 1, v * 2 - 1, d, 1.0])
            P = p[:3] / p[3]
            N = g["normals"][y, x, :3].astype(np.float64)
            I = P - cam; I /= np.linalg.norm(I)
            R = I - 2.0 * (N @ I) * N
            hit = _closest_hit_f64(tris64, P + 0.1 * N, R, 0.01, 10000.0)
            if hit is None:
                if not got[y, x].any():
                    n_miss += 1
                continue
            k, b1, b2, t, gap, edge = hit
            if gap < 1e-4 * max(1.0, t) or edge < 1e-4:
                continue                                    # two surfaces at the same distance / an edge: either answer is legitimate
            gi = int(np.searchsorted(first, k, side="right") - 1)
            want = _reflection_hit_f64(pfd, sc, gi, k - int(first[gi]), b1, b2, cam)
            assert got[y, x, 3] == 1.0, (x, y)
            np.testing.assert_allclose(got[y, x, :3], want, rtol=3e-3, atol=2e-3, err_msg=f"pixel {x},{y}")
            n_hit += 1
    assert n_hit > 60 and n_miss > 60, (n_hit, n_miss)


# ---------------------------------------------------------------------------------------------------------------
# Independent transcription of the G-buffer encodings (gbuf.vert:19-28, gbuf.frag:17-59, clear values hybrid_render_path.cpp:16-19)
# ---------------------------------------------------------------------------------------------------------------
def test_gbuffer_oracle_vs_scalar_transcription():
    W, H = 40, 30
    sc = scenes.tiny_scene(width=W, height=H)
    osc = O.OracleScene(sc)
    tris64 = Hh.world_triangles(sc).astype(np.float64)
    counts = np.array([int(p["index_count"]) // 3 for p in sc.primitives])
    first = np.concatenate([[0], np.cumsum(counts)])
    seq = camera.FrameSequencer(W, H, sc.light)
    cam = sc.camera
    seq.next(cam)
    cam.set_pose(cam.position + np.array([0.05, 0.0, 0.01]), cam.yaw + 0.002, cam.pitch)      # previous != current camera
    pfd = seq.next(cam)
    g = osc.gbuffer(pfd, W, H)
    col = lambda name: np.asarray(pfd[name], np.float64).reshape(4, 4).T
    vp, vp_prev, vp_inv = col("camera_proj") @ col("camera_view"), col("camera_proj_prev_frame") @ col("camera_view_prev_frame"), col("camera_viewproj_inverse")
    eye = col("camera_view_inverse")[:3, 3]
    dsi = np.asarray(pfd["display_size_inverse"], np.float64)
    n_hit = n_sky = 0
    for y in range(H):
        for x in range(W):
            u, v = (x + 0.5) / W, (y + 0.5) / H
            far = vp_inv @ np.array([u * 2 - 1, v * 2 - 1, 0.5, 1.0])                           # any depth on the pixel's line of sight
            d = far[:3] / far[3] - eye
            d /= np.linalg.norm(d)
            hit = _closest_hit_f64(tris64, eye, d, 0.0, 1e9)
            if hit is None:
                # clear values: albedo 0, normals / id 0, motion (0, 0, -1, -1), depth 0
                if g["depth"][y, x] == 0:
                    assert not g["albedo"][y, x].any() and not g["normals"][y, x].any()
                    assert tuple(g["motion"][y, x].astype(f32)) == (0.0, 0.0, -1.0, -1.0)
                    n_sky += 1
                continue
            k, b1, b2, t, gap, edge = hit
            if gap < 1e-4 * max(1.0, t) or edge < 1e-3:
                continue                                                                      # silhouettes: coverage may differ
            gi = int(np.searchsorted(first, k, side="right") - 1)
            p = sc.primitives[gi]
            base = int(p["index_offset"]) + 3 * (k - int(first[gi]))
            vi = [int(p["vertex_offset"]) + int(sc.indices[base + j]) for j in range(3)]
            bary = np.array([1.0 - b1 - b2, b1, b2])
            lp = sum(sc.vertices["pos"][vi[j]].astype(np.float64) * bary[j] for j in range(3))
            ln = sum(sc.vertices["normal"][vi[j]].astype(np.float64) * bary[j] for j in range(3))
            M = np.asarray(p["transform"], np.float64).reshape(4, 4).T
            world = M @ np.append(lp, 1.0)
            clip, clip_prev = vp @ world, vp_prev @ world
            nrm = np.linalg.inv(M[:3, :3]).T @ ln                                              # glm::inverseTranspose(mat3(transform))
            nrm /= np.linalg.norm(nrm)
            m = p["material"]
            assert g["depth"][y, x] > 0, (x, y)
            np.testing.assert_allclose(g["depth"][y, x], clip[2] / clip[3], rtol=2e-4, err_msg=f"depth {x},{y}")
            np.testing.assert_allclose(g["normals"][y, x, :3].astype(np.float64), nrm, atol=2e-3, err_msg=f"normal {x},{y}")
            assert float(g["normals"][y, x, 3]) == float(gi)
            want_mv = np.array([x + 0.5, y + 0.5]) * dsi - ((clip_prev[:2] / clip_prev[3]) * 0.5 + 0.5)
            np.testing.assert_allclose(g["motion"][y, x, :2].astype(np.float64), want_mv, atol=1e-3, err_msg=f"motion {x},{y}")
            np.testing.assert_allclose(g["motion"][y, x, 2:].astype(np.float64), [float(m["metallic_factor"]), float(m["roughness_factor"])], atol=1e-3)
            want_bgra = np.round(np.asarray(m["base_color"], np.float64)[[2, 1, 0, 3]] * 255.0)
            assert np.all(np.abs(g["albedo"][y, x].astype(np.float64) - want_bgra) <= 1), (x, y)
            n_hit += 1
    assert n_hit > 300 and n_sky > 20, (n_hit, n_sky)


# ---------------------------------------------------------------------------------------------------------------
# Independent transcription of the fully ray-traced path (raytraced_render_path/raygen.rgen, closesthit.rchit, miss shaders)
# ---------------------------------------------------------------------------------------------------------------
def test_raytraced_path_oracle_vs_scalar_transcription():
    W, H = 40, 30
    sc = scenes.tiny_scene(width=W, height=H)
    osc = O.OracleScene(sc)
    tris64 = Hh.world_triangles(sc).astype(np.float64)
    counts = np.array([int(p["index_count"]) // 3 for p in sc.primitives])
    first = np.concatenate([[0], np.cumsum(counts)])
    seq = camera.FrameSequencer(W, H, sc.light)
    pfd = seq.next(sc.camera)
    got = osc.raytraced(pfd, W, H)                                   # B8G8R8A8_UNORM
    col = lambda name: np.asarray(pfd[name], np.float64).reshape(4, 4).T
    view_inv, proj_inv = col("camera_view_inverse"), col("camera_proj_inverse")
    light_dir = -np.asarray(pfd["directional_light"]["direction"], np.float64)[:3]
    li = np.asarray(pfd["directional_light"]["intensity"], np.float64)[:3]
    lc = np.asarray(pfd["directional_light"]["color"], np.float64)[:3]
    to_bgra8 = lambda rgba: np.round(np.clip(np.asarray(rgba, np.float64), 0, 1)[[2, 1, 0, 3]] * 255.0)
    n_hit = n_miss = n_lit = n_shadowed = 0
    for y in range(H):
        for x in range(W):
            ndc = np.array([(x + 0.5) / W, (y + 0.5) / H]) * 2.0 - 1.0
            origin = (view_inv @ np.array([0, 0, 0, 1.0]))[:3]
            target = proj_inv @ np.array([ndc[0], ndc[1], 1.0, 1.0])
            tdir = target[:3] / np.linalg.norm(target[:3])
            direction = (view_inv @ np.append(tdir, 0.0))[:3]
            hit = _closest_hit_f64(tris64, origin, direction, 0.1, 10000.0)
            if hit is None:
                if np.all(np.abs(got[y, x].astype(np.float64) - to_bgra8([0.3, 0.8, 0.2, 1.0])) <= 1):
                    n_miss += 1
                continue
            k, b1, b2, t, gap, edge = hit
            if gap < 1e-4 * max(1.0, t) or edge < 1e-3:
                continue
            gi = int(np.searchsorted(first, k, side="right") - 1)
            p = sc.primitives[gi]
            base = int(p["index_offset"]) + 3 * (k - int(first[gi]))
            vi = [int(p["vertex_offset"]) + int(sc.indices[base + j]) for j in range(3)]
            bary = np.array([1.0 - b1 - b2, b1, b2])
            normal = sum(sc.vertices["normal"][vi[j]].astype(np.float64) * bary[j] for j in range(3))
            lp = sum(sc.vertices["pos"][vi[j]].astype(np.float64) * bary[j] for j in range(3))
            position = (np.asarray(p["transform"], np.float64).reshape(4, 4).T @ np.append(lp, 1.0))[:3]
            albedo = np.asarray(p["material"]["base_color"], np.float64)[:3]
            shadow = _closest_hit_f64(tris64, position, light_dir, 0.1, 10000.0)
            if shadow is not None and shadow[5] < 1e-3:
                continue                                             # the shadow ray grazes an edge
            ambient = albedo / np.pi
            if shadow is None:
                rgb = ambient + max(normal @ light_dir, 0.0) * albedo * li * lc
                n_lit += 1
            else:
                rgb = ambient
                n_shadowed += 1
            want = to_bgra8(np.append(rgb, 1.0))
            assert np.all(np.abs(got[y, x].astype(np.float64) - want) <= 1), (x, y, got[y, x], want)
            n_hit += 1
    assert n_hit > 300 and n_miss > 20 and n_lit > 20 and n_shadowed > 20, (n_hit, n_miss, n_lit, n_shadowed)


def test_pixel_ray_regenerator_reproduces_the_masks_and_brute_force_agrees_with_the_bvh():
    """vo_raygen_pixel_rays (used to classify GPU / checker mask mismatches) generates exactly vo_raygen's rays: tracing them one by one gives
    vo_raygen's image; vo_brute_force (every triangle, double precision) agrees with the BVH query on every ray that does not graze."""
    W, H = 64, 40
    sc, osc, frames = Hh.scene_and_gbuffer(W, H, tris=4000)
    pfd, g = frames[1]
    for spp in (2, 4):
        ref = osc.raygen(pfd, g["depth"], g["normals"], ao_spp=spp, flags=3)["shadow_ao"].astype(np.float32)
        got = np.ones((H, W, 2), np.float32)
        checked = 0
        for y in range(H):
            for x in range(W):
                rays = O.raygen_pixel_rays(pfd, g["depth"], g["normals"], x, y, spp)
                if len(rays) == 0:
                    continue
                vis = [0.0 if osc.trace_any(r[:3], r[4:7], float(r[3]), float(r[7])) else 1.0 for r in rays]
                got[y, x] = (vis[0], np.float32(sum(vis[1:])) / np.float32(spp))
                if (x + 7 * y) % 29 == 0:
                    for r, v in zip(rays, vis):
                        hit, _, margin = osc.brute_force(r[:3], r[4:7], r[3], r[7])
                        assert margin < 1e-6 or hit == (v == 0.0)
                        checked += 1
        assert np.array_equal(got.astype(np.float16), ref.astype(np.float16)), f"regenerated rays give another image (ao_spp {spp})"
        assert checked > 100


def test_ssao_kernel_shortcuts_restated_in_numpy():
    """The two exact shortcuts of ssao_kernels.cu's sample loop, restated in numpy against the oracle's formulas (ssao.comp:36-44).
    sample_depth_fixed: texel and weight from the integer floor(fma(t, 256, 0.5)) are bit-identical to bilinear_setup's float formula
    for |t| < 2^14. sincos_turn: within one ulp of the correctly rounded sine / cosine on the angles random01 * 2 * PI can take."""
    f32 = np.float32
    rng = np.random.default_rng(1)

    def oracle_setup(u, n):
        uu = ((u * f32(n)).astype(f32) - f32(0.5)).astype(f32)
        s = (np.floor(((uu * f32(256)).astype(f32) + f32(0.5)).astype(f32)) * f32(0.00390625)).astype(f32)
        fl = np.floor(s)
        return fl.astype(np.int64), (s - fl).astype(f32), uu

    def kernel_setup(t):
        k = np.floor((t.astype(np.float64) * 256 + 0.5).astype(f32)).astype(np.int64)         # F2I.FLOOR(fma(t, 256, 0.5))
        magic = (np.uint32(0x4B000000) | (k & 255).astype(np.uint32)).view(f32)
        return k >> 8, (magic.astype(np.float64) * 0.00390625 - 32768.0).astype(f32)           # fma(magic, 1/256, -2^15)

    for n in (1920, 1080, 101, 59, 3840):
        u = np.concatenate([rng.uniform(-2, 3, 400_000).astype(f32), (np.arange(0, n * 4) / f32(n * 2)).astype(f32),
                            rng.uniform(-1e-3, 1e-3, 50_000).astype(f32), rng.uniform(-8, 8, 50_000).astype(f32)])
        i0, a, t = oracle_setup(u, n)
        j0, b = kernel_setup(t)
        ok = np.abs(t) < 16384
        assert ok.sum() > 0.9 * ok.size and (i0[ok] == j0[ok]).all() and (a[ok] == b[ok]).all(), n

    def fma(a, b, c):
        return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)

    r = (rng.integers(0, 1 << 23, 1_000_000).astype(np.uint32) | np.uint32(0x3F800000)).view(f32) - f32(1)
    ang = ((r * f32(2)).astype(f32) * f32(np.pi)).astype(f32)
    assert (ang == (r * f32(2.0 * float(f32(np.pi)))).astype(f32)).all()                          # fl(fl(2 r) PI) = fl(r (2 PI))
    m = fma(ang, f32(0.636619772), f32(12582912.0))
    q = m.view(np.int32) & 3
    qf = (m - f32(12582912.0)).astype(f32)
    f = fma(qf, f32(-1.57079637), ang)
    f = fma(qf, f32(4.37113883e-8), f)
    f2 = (f * f).astype(f32)
    sp = fma(fma(fma(f2, f32(-1.9515295891e-4), f32(8.3321608736e-3)), f2, f32(-1.6666654611e-1)), (f * f2).astype(f32), f)
    cp = fma(fma(fma(fma(f2, f32(2.443315711809948e-5), f32(-1.388731625493765e-3)), f2, f32(4.166664568298827e-2)), f2, f32(-0.5)), f2, f32(1.0))
    s = np.where(q & 1, cp, sp)
    c = np.where(q & 1, sp, cp)
    s = np.where(q & 2, -s, s)
    c = np.where((q + 1) & 2, -c, c)
    for got, want in ((s, np.sin(ang.astype(np.float64))), (c, np.cos(ang.astype(np.float64)))):
        assert np.abs(got - want).max() < 1.2e-7 * 0.75
        ulps = np.abs(got.view(np.int32).astype(np.int64) - want.astype(f32).view(np.int32))
        big = np.abs(want) > 1e-3                                                                  # next to a zero one ulp is tiny
        assert ulps[big].max() <= 1


def test_shared_reciprocal_division_restated_in_numpy():
    """div_exact of vhr_common.cuh (the screen-space kernels' a / w without __fdiv_rn's slow path), restated in numpy: one Newton step on
    a reciprocal that is up to one ulp off, the quotient, one residual correction — against numpy's correctly rounded float32 division
    on 2e7 operand pairs with exponents in [-30, 30]; and the zero / infinite / NaN operand rule a * rcp(w)."""
    f32 = np.float32
    rng = np.random.default_rng(5)
    N = 20_000_000

    def fma(a, b, c):
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)

    a = (rng.standard_normal(N) * np.exp2(rng.integers(-30, 30, N))).astype(f32)
    w = (rng.standard_normal(N) * np.exp2(rng.integers(-30, 30, N))).astype(f32)
    w[w == 0] = 1
    r0 = (f32(1) / w).astype(f32)
    pert = rng.integers(-1, 2, N)
    r0 = np.where(pert == 0, r0, np.nextafter(r0, np.where(pert > 0, f32(np.inf), f32(-np.inf)).astype(f32))).astype(f32)
    r = fma(r0, fma(-w, r0, np.ones(N, f32)), r0)
    q0 = (a * r).astype(f32)
    q = fma(r, fma(-w, q0, a), q0)
    assert (q == (a / w).astype(f32)).all()
    with np.errstate(all="ignore"):
        sa = np.array([0.0, 1.5, -2.0, np.inf, -np.inf, np.nan, 0.0, -0.0, 3.0, np.inf], f32)
        sw = np.array([0.0, 0.0, 0.0, 0.0, np.inf, 1.0, np.inf, np.inf, -np.inf, np.inf], f32)
        rcp = (f32(1) / sw).astype(f32)
        got, want = (sa * rcp).astype(f32), (sa / sw).astype(f32)
    assert (np.isnan(got) == np.isnan(want)).all() and (got[~np.isnan(want)] == want[~np.isnan(want)]).all()
    assert (np.signbit(got[~np.isnan(want)]) == np.signbit(want[~np.isnan(want)])).all()
