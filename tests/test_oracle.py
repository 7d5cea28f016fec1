"""Pin the CPU oracle (oracle/esr_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import os

import numpy as np
import pytest

from oracle import esr_oracle as O

ARCHS = [(-1, "imdn"), (0, "rfdn"), (4, "rlfn"), (18, "bsrn")]
GOLDEN = ARCHS + [(22, "rfdn40"), (40, "rfdn_pruned"), (26, "imdn_nb7"), (3, "fmen")]   # (model id, golden file tag): RFDN at nf = 40, pruned RFDN, IMDN nb = 7, FMEN


def _weights(golden_dir, mid):
    return O.load_weights(os.path.join(golden_dir, "weights", O.MODELS[mid]["weights"] + ".npz"))


@pytest.mark.parametrize("mid,arch", GOLDEN)
def test_oracle_matches_reference_small_inputs(golden_dir, mid, arch):
    z = np.load(os.path.join(golden_dir, f"ref_{arch}_small.npz"))
    dr = float(z["data_range"])
    assert dr == O.MODELS[mid]["data_range"]
    assert str(z["name"]) == O.MODELS[mid]["name"]
    w = _weights(golden_dir, mid)
    for i in range(4):
        x, y = z[f"x{i}"], z[f"y{i}"]
        yo = O.forward(O.MODELS[mid]["arch"], w, x, dtype=np.float64)
        assert yo.shape == y.shape == (x.shape[0], 3, 4 * x.shape[2], 4 * x.shape[3])
        # the golden is the reference's fp32 output; its own fp32-vs-fp64 noise is <= 1.3e-6 of range
        # on test.bmp and up to 5.7e-6 on uniform-noise inputs (BSRN, 64x64)
        assert np.abs(yo - y).max() / dr < 1e-5, (arch, i)


@pytest.mark.parametrize("mid,arch", [(0, "rfdn"), (18, "bsrn")])
def test_oracle_matches_reference_test_bmp_256(golden_dir, mid, arch):
    z = np.load(os.path.join(golden_dir, f"ref_{arch}_256.npz"))
    img = np.load(os.path.join(golden_dir, "test_bmp.npz"))["img"]
    dr = float(z["data_range"])
    x = O.uint2tensor4(img, dr)
    y = O.forward(arch, _weights(golden_dir, mid), x, dtype=np.float32)
    assert y.shape == (1, 3, 1024, 1024)
    for (a, b), crop in zip(z["crops_yx"], z["crops"]):
        assert np.abs(y[0, :, a:a + 32, b:b + 32] - crop).max() / dr < 1e-5
    assert np.abs(y[0, :, ::16, ::16] - z["sub16"]).max() / dr < 1e-5
    np.testing.assert_allclose(y.astype(np.float64).sum(axis=(0, 2, 3)), z["sum_c"], rtol=1e-6)
    u8 = O.tensor2uint(y, dr)[::8, ::8]
    assert (u8 != z["uint8_sub"]).mean() < 1e-3


def test_oracle_tiled_forward_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "ref_rfdn_tiled.npz"))
    y = O.forward_tiled("rfdn", _weights(golden_dir, 0), z["x"], tile=32, tile_overlap=8, dtype=np.float64)
    assert np.abs(y - z["y"]).max() / 255.0 < 5e-6


def test_esa_minimum_extent_raises(golden_dir):
    # H=14 -> conv2 gives 6 < 7 pool window: PyTorch raises; so does the oracle
    w = _weights(golden_dir, 0)
    with pytest.raises(ValueError):
        O.forward("rfdn", w, np.zeros((1, 3, 14, 20), np.float32))


def test_primitives_against_torch():
    torch = pytest.importorskip("torch")
    F = torch.nn.functional
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 6, 17, 13))
    w = rng.standard_normal((5, 6, 3, 3))
    b = rng.standard_normal(5)
    for s, p in [(1, 1), (2, 0), (1, 0)]:
        ref = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), s, p).numpy()
        np.testing.assert_allclose(O.conv2d(x, w, b, s, p), ref, atol=1e-12)
    wd = rng.standard_normal((6, 1, 3, 3))
    ref = F.conv2d(torch.from_numpy(x), torch.from_numpy(wd), None, 1, 1, groups=6).numpy()
    np.testing.assert_allclose(O.conv2d(x, wd, None, 1, 1, groups=6), ref, atol=1e-12)
    xp = rng.standard_normal((1, 3, 31, 25))
    np.testing.assert_array_equal(O.max_pool2d(xp, 7, 3), F.max_pool2d(torch.from_numpy(xp), 7, 3).numpy())
    xs = rng.standard_normal((1, 3, 9, 5))
    ref = F.interpolate(torch.from_numpy(xs), (64, 47), mode="bilinear", align_corners=False).numpy()
    np.testing.assert_allclose(O.interpolate_bilinear(xs, (64, 47)), ref, atol=1e-13)
    xps = rng.standard_normal((2, 48, 4, 5))
    np.testing.assert_array_equal(O.pixel_shuffle(xps, 4), F.pixel_shuffle(torch.from_numpy(xps), 4).numpy())
    np.testing.assert_allclose(O.gelu(xs), F.gelu(torch.from_numpy(xs)).numpy(), atol=1e-14)
    np.testing.assert_allclose(O.leaky_relu(xs, 0.05), F.leaky_relu(torch.from_numpy(xs), 0.05).numpy(), atol=0)


@pytest.mark.parametrize("mid,arch", ARCHS)
def test_torch_port_matches_reference_small_inputs(golden_dir, mid, arch):
    """oracle/esr_oracle_torch.py (the CPU-baseline port timed by bench.py) against the same goldens."""
    torch = pytest.importorskip("torch")
    from oracle import esr_oracle_torch as OT

    z = np.load(os.path.join(golden_dir, f"ref_{arch}_small.npz"))
    dr = float(z["data_range"])
    w = OT.prepare(_weights(golden_dir, mid))
    for i in range(4):
        y = OT.forward(arch, w, z[f"x{i}"]).numpy()
        # same ATen kernels as the reference: agreement is at fp32 round-off of the thread partition
        assert np.abs(y - z[f"y{i}"]).max() / dr < 1e-5, (arch, i)
    yd = OT.forward(arch, OT.prepare(_weights(golden_dir, mid), torch.float64), z["x3"], torch.float64).numpy()
    yn = O.forward(arch, _weights(golden_dir, mid), z["x3"], dtype=np.float64)
    assert np.abs(yd - yn).max() / dr < 1e-11     # the two restatements agree in fp64


def test_imresize_np_matches_reference(golden_dir):
    """MATLAB-bicubic resize of the DIV2K LR protocol (utils/utils_image.py:704-774) on test.bmp: x1/4 of the whole
    image, of an odd-sized crop, and a x2 enlargement (no antialiasing branch)."""
    img = np.load(os.path.join(golden_dir, "test_bmp.npz"))["img"].astype(np.float32) / 255.
    z = np.load(os.path.join(golden_dir, "test_bmp_lr.npz"))
    for got, want in [(O.imresize_np(img, 1 / 4), z["lr64"]), (O.imresize_np(img[:130, :77], 1 / 4), z["lr_odd"]),
                      (O.imresize_np(img[:40, :36], 2), z["up2"])]:
        assert got.shape == want.shape and got.dtype == np.float32
        assert np.abs(got - want).max() < 2e-6


@pytest.mark.parametrize("mid,arch,shape", [(0, "rfdn", (339, 510)), (18, "bsrn", (270, 480))])
def test_torch_port_matches_reference_at_baseline_shapes(golden_dir, mid, arch, shape):
    """BASELINE.json configs[2] / configs[4] LR shapes: the port bench.py times (and the GPU tests use as the
    full-image reference at these sizes) against crops / a strided subsample / channel sums of the reference's output."""
    pytest.importorskip("torch")
    from oracle import esr_oracle_torch as OT

    z = np.load(os.path.join(golden_dir, f"ref_{arch}_{shape[0]}x{shape[1]}.npz"))
    dr = float(z["data_range"])
    x = O.shaped_input(int(z["seed"]), shape[0], shape[1], dr)
    y = OT.forward(arch, OT.prepare(_weights(golden_dir, mid)), x).numpy()
    assert y.shape == (1, 3, 4 * shape[0], 4 * shape[1])
    for (a, b), crop in zip(z["crops_yx"], z["crops"]):
        assert np.abs(y[0, :, a:a + 32, b:b + 32] - crop).max() / dr < 1e-5
    assert np.abs(y[0, :, ::16, ::16] - z["sub16"]).max() / dr < 1e-5
    np.testing.assert_allclose(y.astype(np.float64).sum(axis=(0, 2, 3)), z["sum_c"], rtol=2e-6)
