"""The oracle's restatement of the reference's first-party geometry against the reference's OWN
sources -- src/utils/image_geom_util.cpp, include/utils/image_geom_util.h, include/backend/equirectangular_camera.h --
compiled with container-only OpenCV stubs into oracle/_ref/libref_geom.so (golden copy: tests/golden/geom_ref.npz).
Same compiler, same flags, same expressions: bit-exact."""
import numpy as np
import pytest


def test_pinhole_projection_bit_exact_against_reference_source(oracle, golden):
    g = golden("geom_ref.npz")
    K4 = g["K4"]
    for i, p in enumerate(g["P"]):
        uv, px, Jp = oracle.geom_pinhole(p, K4)
        assert np.array_equal(uv, g["uv"][i]) and np.array_equal(px, g["px"][i]) and np.array_equal(Jp, g["Jproj"][i])
    assert np.array_equal(g["Jintr"][0], [[K4[0], 0.0], [0.0, K4[1]]])      # applyIntrinsics' Jacobian = diag(fx, fy)


def test_cross2matrix_bit_exact(oracle, golden):
    g = golden("geom_ref.npz")
    for v, M in zip(g["V"], g["cross"]):
        assert np.array_equal(oracle.geom_cross2matrix(v), M)


def test_equirectangular_projection_bit_exact_against_reference_header(oracle, golden):
    g = golden("geom_ref.npz")
    W = g["W"]
    n = len(W)
    for i in range(len(g["eq_px"])):
        w, (PW, PH) = (W[i], (1280, 720)) if i < n else (W[i - n], (4096, 2048))
        px, J = oracle.geom_equirect(w, PW, PH)
        assert np.array_equal(px, g["eq_px"][i]), i
        assert np.array_equal(J, g["eq_J"][i]), i


def test_live_against_reference_sources(oracle):
    if not oracle.have_ref_geom():
        pytest.skip("oracle/_ref/libref_geom.so not built (reference tree absent)")
    rng = np.random.default_rng(5)
    K4 = np.array([200.0, 201.0, 120.0, 90.0])
    for _ in range(2000):
        p = np.array([rng.normal(0, 0.5), rng.normal(0, 0.5), rng.uniform(0.2, 3.0)])
        a, b = oracle.geom_pinhole(p, K4), oracle.ref_geom_pinhole(p, K4)
        assert all(np.array_equal(x, y) for x, y in zip(a, b[:3]))
        w = rng.normal(0, 1, 3)
        pa, Ja = oracle.geom_equirect(w, 2048, 1024)
        pb, Jb = oracle.ref_geom_equirect(w, 2048, 1024)
        assert np.array_equal(pa, pb) and np.array_equal(Ja, Jb)
