"""the wire formats either side of the hot path — planar YUV / y4m frames in, reconstructed frames out — for hosts that are not the reference's C
(bench.py, thor_b200/segments.py, test harnesses).  4:2:0 only (every BASELINE.json configuration).  No library is loaded here.

  read_yuv_frame / write_yuv_frame   common/common_frame.c:478-654 (read_yuv_frame, write_yuv_frame): plane order Y, U, V; one or two bytes per sample (little endian)
                                     by the FILE's bit depth; conversion between the file's bit depth (input_bitdepth) and the codec's (bitdepth) exactly as the
                                     reference does it: up-shift on read, rounded down-shift with saturation on write
  parse_y4m_header                   enc/strings.c:376-449: geometry, frame rate, chroma format / bit depth from a YUV4MPEG2 header; every frame is preceded by "FRAME\\n"

Planes may be read straight into caller-owned arrays (`out=`: e.g. views of pinned staging buffers that tb_rdo_batch_upload() / tb_frame_upload() copy from).
"""
import numpy as np


class Y4MHeader:
    def __init__(self, width, height, fps_num, fps_den, subsample, input_bitdepth, file_headerlen, frame_headerlen=6):
        self.width, self.height, self.fps_num, self.fps_den = width, height, fps_num, fps_den
        self.subsample, self.input_bitdepth, self.file_headerlen, self.frame_headerlen = subsample, input_bitdepth, file_headerlen, frame_headerlen

    @property
    def frame_rate(self):
        return self.fps_num / self.fps_den


def parse_y4m_header(buf):
    """enc/strings.c:376-449.  buf: the first bytes of the file (>= the header line).  Returns None if it is not a y4m file; raises ValueError for what the
    reference rejects (interlaced input, corrupt header)."""
    if not buf.startswith(b"YUV4MPEG2 "):
        return None
    end = buf.find(b"\n")
    if end < 0 or buf[end:end + 7] != b"\nFRAME\n":
        raise ValueError("Corrupt Y4M file")
    w = h = 0
    num, den, sub, bd = 30, 1, 420, 8
    for tok in buf[10:end].split(b" "):
        if not tok:
            continue
        k, v = tok[:1], tok[1:].decode("ascii", "replace")
        if k == b"W":
            w = int(v)
        elif k == b"H":
            h = int(v)
        elif k == b"F":
            a, b = v.split(":")
            num, den = int(a), int(b)
        elif k == b"I":
            if not v.startswith("p"):
                raise ValueError("Only progressive input supported")
        elif k == b"C":
            if v.startswith("mono"):
                sub = 400
            else:
                digits = ""
                while v and v[0].isdigit():
                    digits, v = digits + v[0], v[1:]
                sub = int(digits)
                if v.startswith("p"):  # e.g. C420p10
                    d2 = ""
                    v = v[1:]
                    while v and v[0].isdigit():
                        d2, v = d2 + v[0], v[1:]
                    if d2:
                        bd = int(d2)
        # 'A' (aspect) and 'X' (extensions) carry nothing the hot path needs
    return Y4MHeader(w, h, num, den, sub, bd, end + 1)


def frame_bytes(width, height, input_bitdepth):
    return (width * height * 3 // 2) * (2 if input_bitdepth > 8 else 1)


def read_yuv_frame(f, width, height, input_bitdepth=8, bitdepth=8, sample_bytes=None, out=None, frame_headerlen=0):
    """One 4:2:0 frame from the binary file object `f` at its current position -> [Y, U, V] arrays of uint8 (sample_bytes 1) or uint16 (2) at the CODEC's bit depth.
    sample_bytes defaults to the reference's rule (frame_bitdepth = 16 when either bit depth exceeds 8).  `out`: three writable arrays of the plane shapes to fill."""
    if sample_bytes is None:
        sample_bytes = 2 if max(input_bitdepth, bitdepth) > 8 else 1
    if frame_headerlen:
        f.read(frame_headerlen)
    fdt = np.dtype("<u2") if input_bitdepth > 8 else np.dtype("u1")
    odt = np.uint16 if sample_bytes == 2 else np.uint8
    planes = []
    for p, (ph, pw) in enumerate(((height, width), (height >> 1, width >> 1), (height >> 1, width >> 1))):
        raw = np.frombuffer(f.read(ph * pw * fdt.itemsize), dtype=fdt)
        if raw.size != ph * pw:
            raise EOFError("Error reading %s from file" % "YUV"[p])
        v = raw.reshape(ph, pw).astype(np.int32)
        if bitdepth > input_bitdepth:
            v = v << (bitdepth - input_bitdepth)
        elif bitdepth < input_bitdepth:  # (y + round) >> shift, round = 1 << (shift - 1) as common_frame.c:485 computes it for this direction: 0
            v = v >> (input_bitdepth - bitdepth)
        dst = out[p] if out is not None else np.empty((ph, pw), odt)
        dst[...] = v.astype(odt)
        planes.append(dst)
    return planes


def write_yuv_frame(f, planes, input_bitdepth=8, bitdepth=8):
    """[Y, U, V] at the codec's bit depth -> the file's bit depth (common/common_frame.c:552-654): rounded down-shift with saturation, or up-shift."""
    fdt = np.dtype("<u2") if input_bitdepth > 8 else np.dtype("u1")
    maxv = (1 << input_bitdepth) - 1
    for pl in planes:
        v = np.asarray(pl).astype(np.int32)
        if bitdepth > input_bitdepth:
            v = np.minimum((v + (1 << (bitdepth - input_bitdepth - 1))) >> (bitdepth - input_bitdepth), maxv)
        elif bitdepth < input_bitdepth:
            v = v << (input_bitdepth - bitdepth)
        f.write(np.ascontiguousarray(v.astype(fdt)).tobytes())


class YuvReader:
    """Sequential frames of a .yuv / .y4m file (enc/mainenc.c:330-348: file_headerlen + n * (frame_headerlen + frame size))."""

    def __init__(self, path, width=0, height=0, input_bitdepth=8, bitdepth=None, skip=0):
        self.f = open(path, "rb")
        hdr = parse_y4m_header(self.f.read(256))
        self.file_headerlen = self.frame_headerlen = 0
        if hdr is not None:
            if hdr.subsample != 420:
                raise ValueError("only 4:2:0 input is supported here (got %d)" % hdr.subsample)
            width, height, input_bitdepth = hdr.width, hdr.height, hdr.input_bitdepth
            self.file_headerlen, self.frame_headerlen = hdr.file_headerlen, hdr.frame_headerlen
        self.width, self.height, self.input_bitdepth = width, height, input_bitdepth
        self.bitdepth = bitdepth if bitdepth is not None else input_bitdepth
        self.header = hdr
        self.f.seek(self.file_headerlen + skip * (self.frame_headerlen + frame_bytes(width, height, input_bitdepth)))

    def read(self, out=None):
        return read_yuv_frame(self.f, self.width, self.height, self.input_bitdepth, self.bitdepth, out=out, frame_headerlen=self.frame_headerlen)

    def close(self):
        self.f.close()
