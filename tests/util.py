"""Shared comparison helpers for the parity tests."""
import numpy as np


def rel_rmse(img, ref):
    den = float(np.sqrt((ref[..., :3].astype(np.float64) ** 2).mean()))
    num = float(np.sqrt(((img[..., :3].astype(np.float64) - ref[..., :3].astype(np.float64)) ** 2).mean()))
    return num / max(den, 1e-30)


def render_pair(pkg, scene, test_api, ref_api, w, h, depth, flags=0, spp=1, frame=0, crop=None):
    """Uploads `scene` to both contexts, renders the same uniform on both, returns a dict of comparisons."""
    scene.upload(test_api)
    scene.upload(ref_api)
    u = scene.uniform(test_api, w, h, frame, depth)
    return compare_frames(pkg, test_api, ref_api, u, w, h, flags, spp, crop)


def compare_frames(pkg, test_api, ref_api, u, w, h, flags=0, spp=1, crop=None):
    img = test_api.render_frame(u, test_api.opts(w, h, spp, flags, crop))
    ref = ref_api.render_frame(u, ref_api.opts(w, h, spp, flags, crop))
    a = {k: test_api.get_aov(k, w, h) for k in (pkg.AOV_PRIM_ID, pkg.AOV_INST_ID, pkg.AOV_HIT_T)}
    b = {k: ref_api.get_aov(k, w, h) for k in (pkg.AOV_PRIM_ID, pkg.AOV_INST_ID, pkg.AOV_HIT_T)}
    if crop:
        x0, y0, cw, ch = crop
        sl = (slice(y0, y0 + ch), slice(x0, x0 + cw))
    else:
        sl = (slice(None), slice(None))
    same_id = (a[pkg.AOV_PRIM_ID][sl] == b[pkg.AOV_PRIM_ID][sl]) & (a[pkg.AOV_INST_ID][sl] == b[pkg.AOV_INST_ID][sl])
    same_t = a[pkg.AOV_HIT_T][sl].view(np.uint32) == b[pkg.AOV_HIT_T][sl].view(np.uint32)
    return dict(img=img, ref=ref, id_agreement=float(same_id.mean()), t_agreement=float(same_t.mean()),
                rmse=rel_rmse(img[sl], ref[sl]), bit_exact=bool(np.array_equal(img.view(np.uint32), ref.view(np.uint32))),
                stats=test_api.get_stats(), ref_stats=ref_api.get_stats())


def random_rays(n, seed, lo, hi, tmax=1e32):
    """n rays: origins uniform in the box [lo, hi] (inflated), directions uniform on the sphere."""
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    c, e = (lo + hi) / 2, (hi - lo) / 2
    o = c + (rng.random((n, 3), dtype=np.float32) * 2 - 1) * e * 1.3
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = o, 0.001, d, tmax
    return rays
