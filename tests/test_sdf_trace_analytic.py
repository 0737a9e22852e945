"""The SDF sphere tracer (SDF.inc:101-184, shared by sdfDiffuseTrace.comp and sdfDebugVisualisation.comp) of the oracle against analytic
geometry: primary rays through the brick of a box (an exact box distance field sampled into an R16F brick) must hit where a ray / box
intersection says, with the hit face's normal - independent of the C++ restatement. Run through the frontend's SDF debug frame
(RenderFrontend.cpp:321-340), modes 3 (normals) and 4 (march count)."""
import numpy as np

from conftest import CAMERA, decode_r11g11b10


def box_brick(half, res=32):
    """exact signed distance of the box [-half, half] sampled at the texel centres of a brick spanning the padded box (padSDFBoundingBox)"""
    half = np.asarray(half, np.float64)
    pad = np.maximum(2 * half * 0.075, 0.5)
    ext = half + pad
    axes = [((np.arange(res) + 0.5) / res * 2 - 1) * ext[k] for k in range(3)]
    z, y, x = np.meshgrid(axes[2], axes[1], axes[0], indexing="ij")
    q = np.stack([np.abs(x) - half[0], np.abs(y) - half[1], np.abs(z) - half[2]], -1)
    d = np.linalg.norm(np.maximum(q, 0), axis=-1) + np.minimum(q.max(-1), 0)
    return d.astype(np.float16).view(np.uint16)


def test_primary_rays_hit_the_box_where_the_analytic_intersection_is(ffi, oracle):
    w, h = 160, 90
    half = np.array([1.0, 0.8, 1.3])
    centre = np.array([-7.0, -1.5, 0.9])
    out = {}
    for mode in (3, 4):
        s = ffi.default_settings(oracle, w, h, sdf_debug_mode=mode, sun_direction_deg=(40.0, 35.0))
        fe = ffi.Frontend(oracle, s)
        mesh = fe.register_sdf_mesh(box_brick(half), -half, half, (0.6, 0.6, 0.6))
        M = np.eye(4, dtype=np.float32)
        M[:3, 3] = centre
        fe.set_scene([(mesh, M.T.ravel(), centre - half, centre + half)])
        fe.set_exposure(2e-5)
        zeros4, zeros16 = np.zeros(w * h * 4, np.uint8), np.zeros(w * h * 16, np.uint8)
        for f in range(2):  # the pass list is recorded before the frame's camera is set (main.cpp:79-90): the culling frustum is last frame's
            fe.render_frame(ffi.camera(*CAMERA), (f + 1) / 60.0, 1 / 60.0, zeros4, zeros4, zeros4, zeros16, None)
        out[mode] = decode_r11g11b10(fe.backend.read_image(fe.image("post0"), 0, np.uint32).reshape(h, w))
        g = fe.global_shader_info()
        fe.close()
    # the rays of sdfDebugVisualisation.comp:77-85
    pos, fwd, right, up = (np.array(v, np.float64) for v in CAMERA)
    ys, xs = np.mgrid[0:h, 0:w]
    pc = (np.stack([xs / w, ys / h], -1) - 0.5) * 2
    Vd = -fwd + g.cameraTanFovHalf * pc[..., 1:2] * up - g.cameraTanFovHalf * g.cameraAspectRatio * pc[..., 0:1] * right
    d = -Vd / np.linalg.norm(Vd, axis=-1, keepdims=True)
    o = pos + g.nearPlane * d - centre
    with np.errstate(divide="ignore", invalid="ignore"):
        t1, t2 = (-half - o) / d, (half - o) / d
    tn, tf = np.minimum(t1, t2), np.maximum(t1, t2)
    t_enter, t_exit = tn.max(-1), tf.min(-1)
    hit = (t_exit > np.maximum(t_enter, 0))
    axis = tn.argmax(-1)
    hp = o + d * t_enter[..., None]
    normal = np.zeros((h, w, 3))
    np.put_along_axis(normal, axis[..., None], -np.sign(np.take_along_axis(d, axis[..., None], -1)), -1)
    # pixels well inside a face: the hit point at least 0.35 m from the face's borders (the trilinear brick rounds the edges)
    other = np.ones((h, w, 3), bool)
    np.put_along_axis(other, axis[..., None], False, -1)
    margin = np.where(other, half - np.abs(hp), np.inf).min(-1)
    interior = hit & (margin > 0.35)
    assert interior.sum() > 300
    got_n = out[3][interior] * 2 - 1                       # mode 3: N * 0.5 + 0.5 (6 / 5 mantissa bits)
    assert np.abs(got_n - normal[interior]).max() < 0.06
    assert (out[4][interior][:, 0] > 0).all()              # mode 4: march count / 128, at least one step to reach the surface
    # silhouette: no hit two pixels outside the analytic box, a hit two pixels inside
    from scipy.ndimage import binary_dilation, binary_erosion
    far_outside, well_inside = ~binary_dilation(hit, iterations=2), binary_erosion(hit, iterations=2)
    is_box = (out[3] != out[4]).any(-1)                    # a miss shows the same sky-view LUT sample in both modes, a hit the normal / the count
    assert is_box[well_inside].all() and not is_box[far_outside].any()
