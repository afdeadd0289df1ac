"""Pins the oracle's restatement of cv::GaussianBlur (the only hot-path arithmetic that is not in
/root/reference, helpers.cpp:287,294) against the real OpenCV in this image (cv2 4.13)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
cv2.setNumThreads(1)

# incremental sigmas of the default pyramid (S=3), the first-level blur, S=10 steps, per-patch blurs
SIGMAS = [1.5198685, 1.226274, 1.545008, 1.946588, 2.452547, 0.61725, 0.8, 0.62195, 2.6707, 7.9]


def ksize(sigma):
    n = int(2.0 * 3.0 * float(np.float32(sigma)) + 1.0)
    return n + (n % 2 == 0)


@pytest.mark.parametrize("kind", ["port", "ref"])
@pytest.mark.parametrize("shape", [(96, 160), (135, 240), (64, 72)])
def test_blur_bit_exact_vs_opencv_on_vector_columns(kind, shape, port_oracle, ref_oracle):
    orc = port_oracle if kind == "port" else ref_oracle
    rng = np.random.default_rng(3)
    img = (rng.random(shape) * 255).astype(np.float32)
    for sigma in SIGMAS:
        s = float(np.float32(sigma))
        n = ksize(s)
        want = cv2.GaussianBlur(img, (n, n), s, sigmaY=s, borderType=cv2.BORDER_REPLICATE)
        got = orc.gaussian_blur(img, s)
        assert shape[1] % 4 == 0
        assert np.array_equal(got, want), (kind, sigma, n, np.abs(got - want).max())


@pytest.mark.parametrize("shape", [(67, 33), (19, 19), (43, 43), (25, 57)])
def test_blur_odd_widths_differ_only_in_opencv_scalar_tail(shape, port_oracle):
    """OpenCV finishes the last (width mod 4) columns with scalar code that rounds differently; the
    oracle uses the vector-loop formula everywhere. Bound the difference and localise it."""
    rng = np.random.default_rng(4)
    img = (rng.random(shape) * 255).astype(np.float32)
    for sigma in SIGMAS:
        s = float(np.float32(sigma))
        n = ksize(s)
        want = cv2.GaussianBlur(img, (n, n), s, sigmaY=s, borderType=cv2.BORDER_REPLICATE)
        got = port_oracle.gaussian_blur(img, s)
        assert np.abs(got - want).max() <= 1e-4
        tail = shape[1] % 4
        body = shape[1] - tail
        assert np.array_equal(got[:, :body], want[:, :body]), (sigma, n)


def test_blur_port_equals_ref(port_oracle, ref_oracle):
    rng = np.random.default_rng(5)
    for shape in [(67, 33), (19, 19), (480, 640), (13, 211)]:
        img = (rng.random(shape) * 255).astype(np.float32)
        for sigma in SIGMAS + [0.45, 12.3]:
            a = port_oracle.gaussian_blur(img, sigma)
            b = ref_oracle.gaussian_blur(img, sigma)
            assert np.array_equal(a, b), (shape, sigma)
