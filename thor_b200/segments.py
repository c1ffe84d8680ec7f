"""Multi-GPU sharding of an encode by intra-period segments — host-side plumbing only.

The Thor bitstream is a sequence of self-delimiting frame chunks (4-byte big-endian length + payload,
enc/putbits.c:45-83).  With an intra period P, everything coded after I-frame k*P depends only on frames >= k*P
(enc/mainenc.c:504-518), so rank g can encode input frames [g*S, g*S + S] (S = a multiple of P; one extra frame so
the trailing B frames see the next I frame) with the UNMODIFIED host encoder started at `-skip g*S`.  Its chunks are
byte-identical to the monolithic run's chunks for the same coding positions except for two things, both fixed here at
gather time:
  * the first chunk of every segment but the first is the warm-up I frame (it also carries the sequence header,
    enc/mainenc.c:219-226) -> dropped;
  * `-skip` renumbers frames from 0, and the frame header codes frame_num in 16 bits (enc/write_bits.c:115)
    -> patched by + g*S;
  * the monolithic encoder's sliding reference window still holds the previous segment's trailing B frames, so a
    reference index that points at the segment's own I frame is larger by num_reorder_pics there -> patched (fixed-width
    6-bit fields, enc/write_bits.c:110-113).  Everything after the frame header is bit-identical (verified on the
    reference: tests/test_segments.py).
The only communication is scatter of raw YUV and gather of chunk bytes (torch.distributed: NCCL on GPUs, gloo in the
CPU tests); there is no data-path collective.
"""
import os
import struct
import subprocess
import tempfile

import numpy as np


def plan_segments(n_frames, intra_period, world):
    """-> per rank list of (skip, n) ranges, each a whole number of intra periods (+1 frame), in rank order.
    n_frames counts display frames 0..n_frames-1 with frame 0 the first I frame."""
    if intra_period <= 0:
        return [[(0, n_frames)]] + [[] for _ in range(world - 1)]  # no intra period: the path does not shard (replicas only)
    nseg = max(1, (n_frames - 1 + intra_period - 1) // intra_period)
    segs = [(k * intra_period, min(intra_period + 1, n_frames - k * intra_period)) for k in range(nseg)]
    per = (nseg + world - 1) // world
    return [segs[r * per:(r + 1) * per] for r in range(world)]


def split_chunks(stream):
    out, i = [], 0
    while i < len(stream):
        (ln,) = struct.unpack(">I", stream[i:i + 4])
        out.append(stream[i:i + 4 + ln])
        i += 4 + ln
    assert i == len(stream), "truncated bitstream"
    return out


def parse_frame_header(payload):
    """fixed-length fields of write_frame_header (enc/write_bits.c:98-121): 1 bit frame type, 8 qp, 4 num_intra_modes,
    [2 bits num_ref-1 and 6 bits (ref_array[r]+1) per reference if not an I frame], 16 bits frame_num.
    -> (is_inter, [bit offset of each 6-bit reference field], bit offset of frame_num)"""
    inter = payload[0] >> 7
    if not inter:
        return 0, [], 13
    num_ref = ((payload[1] & 0x06) >> 1) + 1  # bits 13..14
    return 1, [15 + 6 * r for r in range(num_ref)], 15 + 6 * num_ref


def frame_num_bit_offset(payload):
    return parse_frame_header(payload)[2]


def _get_bits(word, nbits_total, off, n):
    return (word >> (nbits_total - off - n)) & ((1 << n) - 1)


def _set_bits(word, nbits_total, off, n, val):
    sh = nbits_total - off - n
    return (word & ~(((1 << n) - 1) << sh)) | ((val & ((1 << n) - 1)) << sh)


def patch_chunk(chunk, frame_delta, position=None, tail=0):
    """Rewrite the header of one frame chunk coded by a segment worker so that it equals the monolithic encoder's:
    frame_num += frame_delta; and, for the chunk at coding position `position` (1 = first frame after the segment's
    warm-up I frame), a reference index equal to position-1 points at that I frame, which in the monolithic run sits
    `tail` slots further back in the sliding reference window (the previous segment's trailing B frames are coded between
    the I frame and this frame, enc/encode_frame.c:823-835) -> index += tail."""
    if frame_delta == 0 and not tail:
        return chunk
    payload = bytearray(chunk[4:])
    inter, ref_offs, fn_off = parse_frame_header(payload)
    nbytes = (fn_off + 16 + 7) // 8
    nb = nbytes * 8
    word = int.from_bytes(payload[:nbytes], "big")
    word = _set_bits(word, nb, fn_off, 16, _get_bits(word, nb, fn_off, 16) + frame_delta)
    if tail and position is not None:
        for off in ref_offs:
            v = _get_bits(word, nb, off, 6)  # ref_array[r] + 1; 0 = interpolated frame
            if v - 1 == position - 1:
                word = _set_bits(word, nb, off, 6, v + tail)
    payload[:nbytes] = word.to_bytes(nbytes, "big")
    return chunk[:4] + bytes(payload)


def patch_frame_num(chunk, delta):
    return patch_chunk(chunk, delta)


def merge_segments(segments, num_reorder_pics=0):
    """segments: list of (skip, bitstream bytes) -> the monolithic bitstream.  num_reorder_pics = the encoder's
    -num_reorder_pics (B frames per sub-GOP; 0 for low-delay configs)."""
    out = []
    for k, (skip, stream) in enumerate(sorted(segments, key=lambda s: s[0])):
        chunks = split_chunks(stream)
        if k == 0:
            out.extend(chunks)
        else:
            out.extend(patch_chunk(c, skip, pos, num_reorder_pics) for pos, c in enumerate(chunks[1:], start=1))
    return b"".join(out)


def frame_bytes(width, height, sample_bytes=1):
    return (width * height * 3 // 2) * sample_bytes


def encode_sharded(encoder, flags, yuv_path, width, height, n_frames, intra_period, out_path=None, sample_bytes=1, workdir=None, num_reorder_pics=None):
    """Collective call (every rank of the default torch.distributed group): rank 0 reads the raw YUV, scatters each rank's
    frame range, every rank runs `encoder` (e.g. oracle/_ref/Thorenc_b200 = reference host + libthor_b200.so) on its
    segments, rank 0 gathers the chunks and returns/writes the merged stream."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    plan = plan_segments(n_frames, intra_period, world)
    if num_reorder_pics is None:  # take it from the flag list (default 0, enc/strings.c)
        fl = list(flags)
        num_reorder_pics = int(fl[fl.index("-num_reorder_pics") + 1]) if "-num_reorder_pics" in fl else 0
    fb = frame_bytes(width, height, sample_bytes)
    # ---- scatter raw YUV: each rank receives the bytes of the frames its segments cover
    spans = [(segs[0][0], segs[-1][0] + segs[-1][1]) if segs else (0, 0) for segs in plan]
    if rank == 0:
        raw = np.fromfile(yuv_path, dtype=np.uint8, count=n_frames * fb)
        mine = raw[spans[0][0] * fb:spans[0][1] * fb]
        for r in range(1, world):
            a, b = spans[r]
            if b > a:
                dist.send(torch.from_numpy(raw[a * fb:b * fb].copy()).to(dev), dst=r)
    else:
        a, b = spans[rank]
        buf = torch.empty((b - a) * fb, dtype=torch.uint8, device=dev)
        if b > a:
            dist.recv(buf, src=0)
        mine = buf.cpu().numpy()
    # ---- encode my segments with the unmodified host encoder
    results = []
    with tempfile.TemporaryDirectory(dir=workdir) as tmp:
        base = spans[rank][0]
        for (skip, n) in plan[rank]:
            inp, bit = os.path.join(tmp, "seg.yuv"), os.path.join(tmp, "seg.bit")
            mine[(skip - base) * fb:(skip - base + n) * fb].tofile(inp)
            cmd = [encoder] + list(flags) + ["-intra_period", str(intra_period), "-if", inp, "-of", bit, "-width", str(width), "-height", str(height), "-n", str(n)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("encoder failed on rank %d: %s" % (rank, r.stderr[-1000:]))
            results.append((skip, open(bit, "rb").read()))
    # ---- gather (skip, length, bytes) per segment on rank 0
    gathered = [None] * world
    dist.all_gather_object(gathered, [(s, len(b)) for s, b in results])
    merged = None
    if rank == 0:
        segs = list(results)
        for r in range(1, world):
            for (skip, ln) in gathered[r]:
                t = torch.empty(ln, dtype=torch.uint8, device=dev)
                dist.recv(t, src=r)
                segs.append((skip, t.cpu().numpy().tobytes()))
        merged = merge_segments(segs, num_reorder_pics)
        if out_path:
            with open(out_path, "wb") as f:
                f.write(merged)
    else:
        for (skip, b) in results:
            dist.send(torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev), dst=0)
    dist.barrier()
    return merged
