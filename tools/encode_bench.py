#!/usr/bin/env python
"""Real encodes, timed: the reference encoder (oracle/_ref/Thorenc, CPU, one thread) beside the same host objects linked against
libthor_b200.so with the RD loop on the GPU (oracle/_ref/Thorenc_b200_rdo), same input, same flags; the two .bit files and reconstructions
must be identical.  Prints one JSON line per run.  Needs a GPU for the second encoder.

  python tools/encode_bench.py --config hdb --width 1920 --height 1080 --frames 9 [--filters off]
--filters off adds -deblocking 0 -clpf 0 -cdef 0 to BOTH encoders: the stream then isolates the RD loop (the in-loop filters of the
linked encoder still go through the per-call drop-in symbols, which are slow at large sizes)."""
import argparse
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_dropin_link import HDB, LDB, REF, synth_yuv  # noqa: E402


def run(exe, flags, args, tag, tmp, env=None):
    bit, rec = os.path.join(tmp, tag + ".bit"), os.path.join(tmp, tag + "_rec.yuv")
    cmd = [os.path.join(REF, exe)] + flags + ["-if", os.path.join(tmp, "in.yuv"), "-of", bit, "-rf", rec, "-width", str(args.width), "-height", str(args.height),
                                              "-n", str(args.frames), "-qp", "32", "-f", "30"]
    e = dict(os.environ)
    e.update(env or {})
    t = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True, env=e, timeout=args.timeout)
    dt = time.time() - t
    if r.returncode != 0:
        raise SystemExit("%s failed: %s" % (exe, r.stderr[-1500:]))
    return open(bit, "rb").read(), open(rec, "rb").read(), dt, r.stderr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="hdb", choices=["hdb", "ldb"])
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=360)
    ap.add_argument("--frames", type=int, default=9)
    ap.add_argument("--bitdepth", type=int, default=8)
    ap.add_argument("--filters", default="on", choices=["on", "off"])
    ap.add_argument("--timeout", type=int, default=3000)
    ap.add_argument("--tmp", default="/tmp/encode_bench")
    ap.add_argument("--skip-ref", action="store_true")
    args = ap.parse_args()
    os.makedirs(args.tmp, exist_ok=True)
    flags = list(HDB if args.config == "hdb" else LDB)
    if args.filters == "off":
        flags += ["-deblocking", "0", "-clpf", "0", "-cdef", "0"]
    if args.bitdepth != 8:
        flags += ["-bitdepth", str(args.bitdepth), "-input_bitdepth", str(args.bitdepth)]
    synth_yuv(os.path.join(args.tmp, "in.yuv"), args.width, args.height, args.frames, args.bitdepth)
    px = args.width * args.height * args.frames
    out = {"config": args.config, "size": "%dx%d" % (args.width, args.height), "frames": args.frames, "bitdepth": args.bitdepth, "filters": args.filters}
    if not args.skip_ref:
        bit_ref, rec_ref, t_ref, _ = run("Thorenc", flags, args, "ref", args.tmp)
        out.update({"cpu_seconds": round(t_ref, 2), "cpu_mpixel_s": round(px / t_ref / 1e6, 4), "cpu": "reference Thorenc, 1 thread, SIMD"})
    bit, rec, t_gpu, err = run("Thorenc_b200_rdo", flags, args, "gpu", args.tmp, {"TB_RDO_STATS": "1", "TB_RDO_TRACE": "1"})
    m = re.search(r"frames decided by tb_rdo_encode_frame: (\d+), by the reference's host loop: (\d+); seconds in tb_rdo_encode_frame ([\d.]+), in serialisation ([\d.]+)", err)
    per_frame = [float(x) for x in re.findall(r"tb_rdo_encode_frame ([\d.]+) ms", err)]
    for line in err.splitlines():
        if "[tb_rdo prof]" in line:
            print(line, file=sys.stderr)
    out.update({"gpu_seconds_whole_process": round(t_gpu, 2), "gpu_mpixel_s_whole_process": round(px / t_gpu / 1e6, 4),
                "frames_on_device": int(m.group(1)) if m else None, "frames_on_host_loop": int(m.group(2)) if m else None,
                "rd_loop_seconds": float(m.group(3)) if m else None, "serialise_seconds": float(m.group(4)) if m else None,
                "rd_loop_mpixel_s": round(px / float(m.group(3)) / 1e6, 4) if m and float(m.group(3)) > 0 else None,
                "rd_loop_ms_per_frame": per_frame})
    if not args.skip_ref:
        out["bit_exact"] = bool(bit == bit_ref and rec == rec_ref)
    print(json.dumps(out))
    if not args.skip_ref and not out["bit_exact"]:
        raise SystemExit("MISMATCH: the GPU encode differs from the reference")


if __name__ == "__main__":
    main()
