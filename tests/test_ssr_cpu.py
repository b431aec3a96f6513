"""Pins the SSR oracle (oracle/oracle_ssr.cpp) against an independent scalar transcription of ssr.comp written from the
shader text (float32 numpy scalars, pure-Python loops: small frames only), plus properties the shader implies. CPU only."""
import numpy as np
import pytest

import helpers as Hh
import oracle_lib as O

f32 = np.float32
PI_INVERSE = f32(0.31830988618379067153776)
PI = f32(3.14159265358979323846264)


def _wrap(i, n):
    return i % n      # python modulo is already non-negative: REPEAT addressing


def _taps(u, v, W, H):
    uu = f32(f32(u * f32(W)) - f32(0.5)); vv = f32(f32(v * f32(H)) - f32(0.5))
    uu, vv = f32(np.floor(f32(f32(uu * f32(256.0)) + f32(0.5))) * f32(0.00390625)), f32(np.floor(f32(f32(vv * f32(256.0)) + f32(0.5))) * f32(0.00390625))      # 8 fractional bits of sub-texel precision
    fx, fy = np.floor(uu), np.floor(vv)
    a, b = f32(uu - fx), f32(vv - fy)
    ix = int(fx) if np.isfinite(fx) and abs(fx) < 1e9 else 0
    iy = int(fy) if np.isfinite(fy) and abs(fy) < 1e9 else 0
    return _wrap(ix, W), _wrap(ix + 1, W), _wrap(iy, H), _wrap(iy + 1, H), a, b


def _lerp(a, b, t00, t10, t01, t11):
    one = f32(1)
    return f32(f32(f32(f32(f32(one - a) * f32(one - b)) * t00) + f32(f32(a * f32(one - b)) * t10)) + f32(f32(f32(one - a) * b) * t01)) + f32(f32(a * b) * t11)


def _tex(img, u, v):
    """texture() of an [H, W, C] float32 image through the default sampler (LINEAR, REPEAT), Vulkan float weights."""
    H, W = img.shape[:2]
    x0, x1, y0, y1, a, b = _taps(u, v, W, H)
    with np.errstate(all="ignore"):
        return np.array([f32(_lerp(a, b, img[y0, x0, c], img[y0, x1, c], img[y1, x0, c], img[y1, x1, c])) for c in range(img.shape[2])], f32)


def _mul44(m, x):       # m[c][r], glm column-major; left-to-right sums of rounded products
    with np.errstate(all="ignore"):
        return np.array([f32(f32(f32(f32(m[0, r] * x[0]) + f32(m[1, r] * x[1])) + f32(m[2, r] * x[2])) + f32(m[3, r] * x[3])) for r in range(4)], f32)


def _dot(a, b):
    with np.errstate(all="ignore"):
        return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))


def _normalize(a):
    with np.errstate(all="ignore"):
        return (a / np.sqrt(_dot(a, a))).astype(f32)


def _world_pos(inv, depth, u, v):        # glsl_common.h:117-122
    p = _mul44(inv, np.array([f32(u * f32(2)) - f32(1), f32(v * f32(2)) - f32(1), depth, f32(1)], f32))
    with np.errstate(all="ignore"):
        return (p[:3] / p[3]).astype(f32)


def scalar_ssr(pfd, g, ray_distance, step_size, thickness, bsearch_steps, pixels):
    """ssr.comp:61-137 for the listed (x, y) pixels."""
    H, W = g["depth"].shape
    a8 = g["albedo"].astype(f32)
    albedo_img = (np.stack([a8[..., 2], a8[..., 1], a8[..., 0]], -1) / f32(255)).astype(f32)
    depth_img = g["depth"].astype(f32)[..., None]
    normals_img = g["normals"].astype(f32)
    motion_img = g["motion"].astype(f32)
    proj = np.asarray(pfd["camera_proj"], f32).reshape(4, 4)
    view = np.asarray(pfd["camera_view"], f32).reshape(4, 4)
    inv = np.asarray(pfd["camera_viewproj_inverse"], f32).reshape(4, 4)
    # proj * view: column c of the product = proj * (column c of view)
    pv = np.stack([_mul44(proj, view[c]) for c in range(4)])
    cam = np.asarray(pfd["camera_view_inverse"], f32).reshape(4, 4)[3, :3]
    dsi = np.asarray(pfd["display_size_inverse"], f32)
    L = (-np.asarray(pfd["directional_light"]["direction"], f32)[:3]).astype(f32)
    li = np.asarray(pfd["directional_light"]["intensity"], f32)[:3]
    lc = np.asarray(pfd["directional_light"]["color"], f32)[:3]
    step_size, thickness = f32(step_size), f32(thickness)
    n_steps = int(f32(ray_distance) / step_size)

    def to_uv(p):
        c = _mul44(pv, np.array([p[0], p[1], p[2], f32(1)], f32))
        with np.errstate(all="ignore"):
            return f32(f32(f32(c[0] / c[3]) * f32(0.5)) + f32(0.5)), f32(f32(f32(c[1] / c[3]) * f32(0.5)) + f32(0.5))

    def dist(a, b):
        d = (a - b).astype(f32)
        return np.sqrt(_dot(d, d))

    def probe(P, rd, offset):
        with np.errstate(all="ignore"):
            rp = (P + (rd * offset).astype(f32)).astype(f32)
            u, v = to_uv(rp)
            sp = _world_pos(inv, _tex(depth_img, u, v)[0], u, v)
            return f32(dist(cam, rp) - dist(cam, sp)), u, v

    def lighting(u, v):
        with np.errstate(all="ignore"):
            albedo = _tex(albedo_img, u, v)
            position = _world_pos(inv, _tex(depth_img, u, v)[0], u, v)
            mr = _tex(motion_img, u, v)[2:]
            V = _normalize((cam - position).astype(f32))
            N = _tex(normals_img, u, v)[:3]
            Hh_ = _normalize((L + V).astype(f32))
            metallic = f32(min(max(mr[0], f32(0)), f32(1)))
            rough = f32(min(max(mr[1], f32(0.04)), f32(1)))
            f0 = (f32(0.04) * f32(f32(1) - metallic) + albedo * metallic).astype(f32)
            o = f32(f32(1) - max(_dot(Hh_, V), f32(0)))
            F = (f0 + (f32(1) - f0) * o * o * o * o * o).astype(f32)
            diffuse = ((f32(1) - F) * f32(f32(1) - metallic) * albedo / PI).astype(f32)
            a2 = f32(rough * rough)
            nh = max(_dot(N, Hh_), f32(0))
            ff = f32(f32(f32(nh * nh) * f32(a2 - f32(1))) + f32(1))
            D = f32(a2 / f32(f32(PI * ff) * ff))
            k = f32(f32(f32(rough + f32(1)) * f32(rough + f32(1))) * f32(0.125))
            nv, nl = max(_dot(N, V), f32(0)), max(_dot(N, L), f32(0))
            G = f32(f32(nv / f32(f32(nv * f32(f32(1) - k)) + k)) * f32(nl / f32(f32(nl * f32(f32(1) - k)) + k)))
            spec = ((f32(D * G) * F) / max(f32(f32(f32(4) * nv) * nl), f32(1e-6))).astype(f32)
            amb = (albedo * f32(PI_INVERSE * f32(0.2))).astype(f32)
            return (amb + (((diffuse + spec).astype(f32) * nl).astype(f32) * li).astype(f32) * lc).astype(f32)

    out = {}
    for (x, y) in pixels:
        u0, v0 = f32(f32(x) * dsi[0]), f32(f32(y) * dsi[1])
        with np.errstate(all="ignore"):
            P = _world_pos(inv, _tex(depth_img, u0, v0)[0], u0, v0)
            N = _tex(normals_img, u0, v0)[:3]
            I = _normalize((P - cam).astype(f32))
            rd = _normalize((I - (N * f32(f32(2) * _dot(N, I))).astype(f32)).astype(f32))
        found, prev_step, final_step = False, f32(0), f32(0)
        for i in range(n_steps):
            offset = f32(step_size * f32(i))
            delta, _, _ = probe(P, rd, offset)
            if delta > f32(0.3) and delta < thickness:
                final_step, found = offset, True
                break
            prev_step = offset
        if not found:
            out[(x, y)] = np.zeros(4, f32)
            continue
        mid = f32(f32(prev_step + final_step) * f32(0.5))
        fu, fv = f32(0), f32(0)
        for i in range(bsearch_steps):
            delta, fu, fv = probe(P, rd, mid)
            if delta > f32(0.3) and delta < thickness:
                mid = f32(f32(prev_step + mid) * f32(0.5))
            else:
                mid, prev_step = f32(mid + f32(mid - prev_step)), mid
        out[(x, y)] = np.append(lighting(fu, fv), f32(1))
    return out


@pytest.fixture(scope="module")
def frame():
    W, H = 96, 64
    sc, osc, frames = Hh.scene_and_gbuffer(W, H, tris=6000)
    return W, H, frames[1]


def test_ssr_oracle_vs_scalar_transcription(frame):
    W, H, (pfd, g) = frame
    params = dict(ray_distance=8.0, step_size=0.1, thickness=0.5, bsearch_steps=6)
    ref = O.ssr(pfd, g["albedo"], g["normals"], g["motion"], g["depth"], **params).astype(np.float32)
    hits = np.argwhere(ref[..., 3] > 0)
    assert len(hits) > 50, "degenerate SSR test frame: no reflections found"
    rng = np.random.default_rng(0)
    pick = [tuple(int(c) for c in hits[i][::-1]) for i in rng.choice(len(hits), 40, replace=False)]
    misses = np.argwhere(ref[..., 3] == 0)
    pick += [tuple(int(c) for c in misses[i][::-1]) for i in rng.choice(len(misses), 12, replace=False)]
    got = scalar_ssr(pfd, g, pixels=pick, **params)
    exact = 0
    for (x, y), val in got.items():
        want = ref[y, x]
        half = val.astype(np.float16).astype(np.float32)
        assert np.allclose(half, want, rtol=2e-3, atol=1e-3), ((x, y), half, want)
        exact += int(np.array_equal(half, want))
    assert exact >= len(pick) - 2, f"only {exact}/{len(pick)} pixels bit-identical to the scalar transcription"


def test_ssr_properties(frame):
    W, H, (pfd, g) = frame
    args = (pfd, g["albedo"], g["normals"], g["motion"], g["depth"])
    base = O.ssr(*args, ray_distance=8.0, step_size=0.1, thickness=0.5, bsearch_steps=6)
    # thickness <= 0.3: the acceptance window (0.3, thickness) is empty -> nothing is ever found
    assert not O.ssr(*args, ray_distance=8.0, step_size=0.1, thickness=0.3, bsearch_steps=6).any()
    # zero march steps -> nothing found
    assert not O.ssr(*args, ray_distance=0.05, step_size=0.1, thickness=0.5, bsearch_steps=6).any()
    # found pixels carry alpha 1, the others stay (0,0,0,0)
    a = base[..., 3].astype(np.float32)
    assert set(np.unique(a)) <= {0.0, 1.0}
    assert not base[a == 0].any()
    # the march decides "found"; the binary search only moves the shading point
    other = O.ssr(*args, ray_distance=8.0, step_size=0.1, thickness=0.5, bsearch_steps=0)
    assert np.array_equal(other[..., 3], base[..., 3])
    # bsearch_steps = 0 shades final_uv = (0,0): one colour for every found pixel
    found = other[other[..., 3] > 0]
    assert len(np.unique(found.view(np.uint16).reshape(-1, 4), axis=0)) == 1
    # row bands stitch to the full frame
    top = O.ssr(*args, ray_distance=8.0, step_size=0.1, thickness=0.5, bsearch_steps=6, rows=(0, 24))
    bot = O.ssr(*args, ray_distance=8.0, step_size=0.1, thickness=0.5, bsearch_steps=6, rows=(24, H))
    assert np.array_equal(np.concatenate([top[:24], bot[24:]]).view(np.uint16), base.view(np.uint16))


def test_window_estimate_stays_inside_its_error_bound():
    """ssr_kernels.cu decides the two comparisons of a march step from an ESTIMATE of delta (MUFU reciprocal, fused dot products, rsqrt
    instead of three IEEE quotients and two correctly rounded square roots) whenever the estimate is further than 32 * 2^-24 * (d1 + d2 +
    |sp|_1) from both thresholds. Restated in numpy with worst-case MUFU errors (reciprocal +-1 ulp, rsqrt +-2 ulp) on random probes at
    four scene scales: the estimate never leaves 8 * 2^-24 * (...) of the oracle-order value — a quarter of the band the kernel uses."""
    f32 = np.float32
    rng = np.random.default_rng(11)
    N = 1_000_000

    def r32(x):
        return x.astype(f32)

    def fma(a, b, c):
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)

    def pert(x, ulps):
        return (x.view(np.int32) + rng.integers(-ulps, ulps + 1, x.shape).astype(np.int32)).view(f32)

    def dist_rn(a, b):
        d = r32(a - b)
        s = r32(r32(r32(d[:, 0] * d[:, 0]) + r32(d[:, 1] * d[:, 1])) + r32(d[:, 2] * d[:, 2]))
        return r32(np.sqrt(s.astype(np.float64)))

    for scale in (1e-2, 1.0, 30.0, 1000.0):
        cam = r32(rng.standard_normal((N, 3)) * scale)
        rp = r32(cam + rng.standard_normal((N, 3)) * scale * rng.uniform(0.001, 3, (N, 1)))
        w = r32(rng.uniform(1e-4, 10, N) * rng.choice([-1, 1], N))
        q = r32(r32(cam + rng.standard_normal((N, 3)) * scale * rng.uniform(0.001, 3, (N, 1))) * w[:, None])
        exact = r32(dist_rn(cam, rp) - dist_rn(cam, r32(q / w[:, None])))
        s = r32(q * pert(r32(1.0 / w.astype(np.float64)), 1)[:, None])
        e, f = r32(cam - s), r32(cam - rp)
        d2q = fma(e[:, 0], e[:, 0], fma(e[:, 1], e[:, 1], r32(e[:, 2] * e[:, 2])))
        d1q = fma(f[:, 0], f[:, 0], fma(f[:, 1], f[:, 1], r32(f[:, 2] * f[:, 2])))
        d2 = r32(d2q * pert(r32(1 / np.sqrt(d2q.astype(np.float64))), 2))
        d1 = r32(d1q * pert(r32(1 / np.sqrt(d1q.astype(np.float64))), 2))
        est = r32(d1 - d2)
        unit = (d1 + d2 + np.abs(s).sum(1)).astype(np.float64) * 2.0 ** -24
        assert (np.abs(est.astype(np.float64) - exact.astype(np.float64)) <= 8 * unit).all(), scale
