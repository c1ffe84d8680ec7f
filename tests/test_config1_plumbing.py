"""BASELINE.json configs[0]: "Thordec single 64x64 I-frame bitstream on CPU -> bit-exact YUV (plumbing, no GPU)".  The reference's own CPU-runnable case: the compiled
reference encoder writes a one-I-frame 64x64 stream, the compiled reference decoder decodes it, and the decoded YUV — read back through thor_b200/yuvio.py —
equals the encoder's reconstruction sample for sample.  No GPU, no libthor_b200: this pins the container / file plumbing the GPU tests build on
(tests/test_dropin_link.py decodes the same kind of stream with Thordec_b200)."""
import importlib.util
import os
import subprocess

import numpy as np
import pytest

from test_dropin_link import HDB, REF, synth_yuv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("yuvio", os.path.join(ROOT, "thor_b200", "yuvio.py"))
Y = importlib.util.module_from_spec(spec)
spec.loader.exec_module(Y)
needs = pytest.mark.skipif(not (os.path.exists(os.path.join(REF, "Thorenc")) and os.path.exists(os.path.join(REF, "Thordec"))), reason="oracle/_ref not built")


@needs
@pytest.mark.parametrize("bitdepth", [8, 10])
def test_single_64x64_intra_frame_roundtrip(tmp_path, bitdepth):
    tmp = str(tmp_path)
    w = h = 64
    synth_yuv(os.path.join(tmp, "in.yuv"), w, h, 1, bitdepth)
    extra = ["-bitdepth", str(bitdepth), "-input_bitdepth", str(bitdepth)] if bitdepth != 8 else []
    # tiny frames + CDEF corrupt the reference's frame header: CDEF off
    r = subprocess.run([os.path.join(REF, "Thorenc")] + HDB + ["-cdef", "0", "-if", os.path.join(tmp, "in.yuv"), "-of", os.path.join(tmp, "a.bit"), "-rf", os.path.join(tmp, "rec.yuv"),
                                                              "-width", str(w), "-height", str(h), "-n", "1", "-qp", "32", "-f", "30"] + extra, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    r = subprocess.run([os.path.join(REF, "Thordec"), os.path.join(tmp, "a.bit"), os.path.join(tmp, "dec.yuv")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    assert os.path.getsize(os.path.join(tmp, "dec.yuv")) == Y.frame_bytes(w, h, bitdepth)
    with open(os.path.join(tmp, "dec.yuv"), "rb") as fd, open(os.path.join(tmp, "rec.yuv"), "rb") as fr, open(os.path.join(tmp, "in.yuv"), "rb") as fi:
        dec = Y.read_yuv_frame(fd, w, h, bitdepth, bitdepth)
        rec = Y.read_yuv_frame(fr, w, h, bitdepth, bitdepth)
        src = Y.read_yuv_frame(fi, w, h, bitdepth, bitdepth)
    for p in range(3):
        assert np.array_equal(dec[p], rec[p]), "decoder output differs from the encoder's reconstruction (plane %d)" % p
    mse = float(np.mean((dec[0].astype(np.float64) - src[0].astype(np.float64)) ** 2)) / (1 << (2 * (bitdepth - 8)))
    assert mse < 40.0, "the decoded frame is not a coded version of the source (luma MSE %.1f at 8-bit scale)" % mse
