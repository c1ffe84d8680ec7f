#!/usr/bin/env python
"""One launch of rdo_batch_kernel over the inter frames of a captured clip, for ncu / compute-sanitizer:
  ncu --set full --import-source on --clock-control none -k regex:rdo_batch -c 1 -o gpurun_out/rdo python tools/rdo_profile.py --size 640x384
Prints the kernel's own counters (cycles per primitive class, work executed)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="hdb")
    ap.add_argument("--size", default="640x384")
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--copies", type=int, default=1)
    ap.add_argument("--launches", type=int, default=1)
    ap.add_argument("--cache", default="/tmp/thor_b200_bench")
    args = ap.parse_args()
    cfg = bench.Cfg(args.config, tuple(int(v) for v in args.size.split("x")), args.frames or None)
    os.makedirs(args.cache, exist_ok=True)
    jobdir, meta = bench.capture_jobs(cfg, args.cache)
    import thor_b200 as tb
    from thor_b200 import rdo_jobs as RJ
    L = tb.lib
    tb.init(0)
    L.tb_rdo_batch_create.restype = C.c_void_p
    L.tb_rdo_batch_create.argtypes = [C.c_int, C.c_int]
    L.tb_rdo_batch_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.tb_rdo_batch_run.argtypes = [C.c_void_p, C.c_int]
    L.tb_rdo_batch_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.tb_rdo_batch_sync.argtypes = [C.c_void_p]
    L.tb_rdo_batch_grid.argtypes = [C.c_void_p]
    L.tb_rdo_batch_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.tb_rdo_last_error.restype = C.c_char_p
    jobs = [j for j in RJ.load_jobs(jobdir) if j.hdr.frame_type != 0]
    keep = []

    def alloc(n):
        a = np.zeros(max(n, 16), np.uint8); keep.append(a)
        return a.ctypes.data
    hosts = [RJ.HostFrame(j, alloc) for _ in range(args.copies) for j in jobs]
    b = L.tb_rdo_batch_create(len(hosts), cfg.ESZ)
    for s, hf in enumerate(hosts):
        assert L.tb_rdo_batch_upload(b, s, C.byref(hf.desc)) == 0, L.tb_rdo_last_error()
    for _ in range(args.launches):
        t = time.time()
        assert L.tb_rdo_batch_run(b, len(hosts)) == 0, L.tb_rdo_last_error()
        for s, hf in enumerate(hosts):
            assert L.tb_rdo_batch_download(b, s, C.byref(hf.desc)) == 0
        assert L.tb_rdo_batch_sync(b) == 0, L.tb_rdo_last_error()
        dt = time.time() - t
    st = (C.c_uint64 * len(bench.STAT_NAMES))()
    L.tb_rdo_batch_stats(b, st, len(st))
    d = dict(zip(bench.STAT_NAMES, [int(v) for v in st]))
    px = sum(hf.job.pixels for hf in hosts)
    ok = all(all(hf.check().values()) for hf in hosts)
    out = {"size": args.size, "frames": len(hosts), "ctas": int(L.tb_rdo_batch_grid(b)), "seconds": round(dt, 3), "mpixel_s": round(px / dt / 1e6, 4), "parity_ok": ok,
           "cycles_per_search": round(d["cyc_me"] / max(1, d["searches"])), "cycles_per_txfm_chain": round(d["cyc_tx_chain"] / max(1, d["txfm_chains"])),
           "cycles_per_prediction": round(d["cyc_interp"] / max(1, d["predictions"])), "cycles_per_intra": round(d["cyc_intra"] / max(1, d["intra_predictions"])),
           "warp_cycles_per_super_block": round(d["cyc_total"] * (1 - d["cyc_idle"] / max(1, d["cyc_total"])) / max(1, d["super_blocks"])),
           "share": {k[4:]: round(d[k] / max(1, d["cyc_total"]), 4) for k in bench.STAT_NAMES[:10] + bench.STAT_NAMES[23:40]}, "phase": {k[3:]: round(d[k] / max(1, sum(d[n] for n in bench.STAT_NAMES[40:49])), 4) for k in bench.STAT_NAMES[40:49]}, "me_stage": {k[9:]: round(d[k] / max(1, sum(d[n] for n in bench.STAT_NAMES[49:])), 4) for k in bench.STAT_NAMES[49:]}, "work": {k: d[k] for k in bench.STAT_NAMES[11:23]}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
