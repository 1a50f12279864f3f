"""
Diffuse interreflection (rt.set_uint("path_seg_range", 2, 4), moon_renderer.py:583 - paths of up to four
segments) on the CUDA path against the float64 oracle, which follows the same paths (same random dimensions, same
cosine-distributed directions) with its own exhaustive tracer.
"""
import numpy as np
import pytest

from helpers import image_metrics, make_gpu, make_oracle, sun_at_phase

pytestmark = pytest.mark.gpu


def _scene(W=720, H=360, seed=3):
    from moonrtx_b200.synth import synth_ldem, synth_color
    from moonrtx_b200.data_loader import downscale_elevation, color_texture
    elev, _ = downscale_elevation(synth_ldem(W, H, seed=seed, craters=60), 1)
    # exaggerated relief (x 12) so that crater walls face each other steeply and the second-order light is not tiny
    elev = (1.0 + (elev - elev.mean()) * 12.0).astype(np.float32)
    elev /= elev.max()
    tex = color_texture(synth_color(256, 128), 2.2, 1)
    tex[..., :3] = np.maximum(tex[..., :3], 150)              # a bright surface: interreflection that shows
    return elev, tex


@pytest.mark.parametrize("spp,seg,phase", [(1, (2, 3), 70.0), (4, (2, 4), 95.0)])
def test_interreflection_matches_oracle(spp, seg, phase):
    elev, tex = _scene()
    kw = dict(light_pos=sun_at_phase(phase), fov=1.6, path_seg_range=seg)
    W, H = 160, 120
    rt = make_gpu(elev, W, H, texture=tex, debug_hits=False, **kw)
    orc = make_oracle(elev, W, H, texture=tex, jitter=spp > 1, **kw)
    if spp > 1:
        rt.set_param(max_accumulation_frames=spp, min_accumulation_step=spp)
    img = rt.render_cycle().copy()
    acc = rt.get_accum_buffer().copy()
    o = orc.render(nsamples=spp)
    ref = orc.tonemap(o["accum"])
    mae, psnr = image_metrics(img, ref)
    assert mae <= 0.25 and psnr >= 45.0, (mae, psnr)
    # the radiance itself, pixel by pixel: paths agree except where float32 shading sends a bounce ray to the other side of
    # a silhouette - a handful of pixels
    oa = o["accum"]
    rel = np.abs(acc[..., :3] - oa[..., :3]).max(axis=2) / (np.abs(oa[..., :3]).max(axis=2) + 1e-3 * oa[..., :3].max())
    assert np.mean(rel > 0.02) < 0.01, float(np.mean(rel > 0.02))
    # ... and there IS second-order light: the same frame with direct light only is darker, most of all in the shadows
    rt.set_uint("path_seg_range", 2, 2)
    rt.render_cycle()
    direct = rt.get_accum_buffer().copy()
    gain = acc[..., :3].sum(axis=2) - direct[..., :3].sum(axis=2)
    assert gain.min() >= -1e-4 * acc[..., :3].max() and gain.sum() > 0.005 * direct[..., :3].sum()
    shadowed = (direct[..., :3].sum(axis=2) == 0) & (direct[..., 3] > 0) & (rt.get_hit_buffer()[..., 3] > 0)
    assert int(shadowed.sum()) > 50 and int((gain[shadowed] > 0).sum()) >= 3       # (most unlit pixels are the night side; one path in two or three goes on)
    od = make_oracle(elev, W, H, texture=tex, jitter=spp > 1, **{**kw, "path_seg_range": (2, 2)}).render(nsamples=spp)["accum"]
    assert abs(gain.sum() / (oa[..., :3].sum() - od[..., :3].sum()) - 1.0) < 0.02        # the added light as a whole, within 2 %
    rt.close()


def test_bounces_leave_the_direct_path_alone_and_count_their_rays():
    elev, tex = _scene()
    kw = dict(light_pos=sun_at_phase(80.0), fov=1.6)
    rt = make_gpu(elev, 160, 120, texture=tex, **kw)
    rt.counters(reset=True)
    rt.render_cycle()
    c0 = rt.counters()
    hits0 = rt.get_hit_records_f64().copy()
    rt.set_uint("path_seg_range", 2, 4)
    rt.counters(reset=True)
    rt.render_cycle()
    c2 = rt.counters()
    assert np.array_equal(rt.get_hit_records_f64(), hits0)                   # camera hits are what they were
    assert c2["primary_rays"] == c0["primary_rays"] and c2["primary_hits"] == c0["primary_hits"]
    assert c2["shadow_rays"] > c0["shadow_rays"] and c2["node_visits"] > 1.15 * c0["node_visits"]
    rt.close()
