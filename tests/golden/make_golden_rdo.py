#!/usr/bin/env python
"""Generates tests/golden/rdo_jobs_hdb_128x136.xz: the RD-loop jobs of a 9-frame 128x136 HDB encode (I, P, seven hierarchical B frames with interpolated
reference and bipred; CDEF off: tiny frames + CDEF corrupt the reference's header) WITH the decisions of the reference's own process_block(), written by the
reference encoder + observing shim (oracle/_ref/Thorenc_capture, TB_RDO_DUMP).  The frame_NNN.job files are concatenated (8-byte length prefix each) and xz-compressed.
Run in the development container (needs /root/reference compiled into oracle/_ref):  python tests/golden/make_golden_rdo.py"""
import lzma
import os
import struct
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from test_dropin_link import HDB, REF, synth_yuv  # noqa: E402

W, H, N = 128, 136, 9
OUT = os.path.join(HERE, "rdo_jobs_hdb_128x136.xz")


def capture(tmp):
    synth_yuv(os.path.join(tmp, "in.yuv"), W, H, N)
    r = subprocess.run([os.path.join(REF, "Thorenc_capture")] + HDB + ["-cdef", "0", "-if", os.path.join(tmp, "in.yuv"), "-of", os.path.join(tmp, "o.bit"), "-width", str(W),
                                                                       "-height", str(H), "-n", str(N), "-qp", "32", "-f", "30"], capture_output=True, text=True,
                       env=dict(os.environ, TB_RDO_DUMP=tmp))
    assert r.returncode == 0, r.stderr[-800:]
    return [open(os.path.join(tmp, f), "rb").read() for f in sorted(os.listdir(tmp)) if f.endswith(".job")]


def pack(jobs):
    return lzma.compress(b"".join(struct.pack("<Q", len(j)) + j for j in jobs), preset=9 | lzma.PRESET_EXTREME)


def unpack(blob, directory):
    raw, o, k = lzma.decompress(blob), 0, 0
    while o < len(raw):
        n = struct.unpack_from("<Q", raw, o)[0]
        open(os.path.join(directory, "frame_%03d.job" % k), "wb").write(raw[o + 8:o + 8 + n])
        o += 8 + n
        k += 1
    return k


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as tmp:
        jobs = capture(tmp)
    blob = pack(jobs)
    open(OUT, "wb").write(blob)
    print("%s: %d jobs, %d bytes raw, %d bytes compressed" % (OUT, len(jobs), sum(len(j) for j in jobs), len(blob)))
