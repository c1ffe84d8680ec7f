#!/usr/bin/env python
"""Can several encoder PROCESSES (oracle/_ref/Thorenc_b200_rdo: the reference's host objects + the device RD loop) share one GPU concurrently?
Times 1, 4, 8 concurrent encodes of the same clip without and with the CUDA MPS daemon; prints one JSON line."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_dropin_link import HDB, REF, synth_yuv  # noqa: E402

W, H, N = 640, 360, 5
T = "/tmp/mps_probe"
os.makedirs(T, exist_ok=True)
synth_yuv(T + "/in.yuv", W, H, N)
FLAGS = list(HDB) + ["-deblocking", "0", "-clpf", "0", "-cdef", "0"]


def run(k, env):
    t = time.time()
    ps = [subprocess.Popen([os.path.join(REF, "Thorenc_b200_rdo")] + FLAGS + ["-if", T + "/in.yuv", "-of", "%s/o%d.bit" % (T, i), "-rf", "%s/r%d.yuv" % (T, i), "-width", str(W),
                            "-height", str(H), "-n", str(N), "-qp", "32", "-f", "30"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env) for i in range(k)]
    rc = [p.wait() for p in ps]
    dt = time.time() - t
    same = len({open("%s/o%d.bit" % (T, i), "rb").read() for i in range(k)}) == 1
    return {"procs": k, "seconds": round(dt, 2), "mpixel_s": round(k * W * H * N / dt / 1e6, 4), "rc_ok": all(r == 0 for r in rc), "identical": same}


out = {"clip": "%dx%dx%d HDB, filters off" % (W, H, N), "no_mps": [run(k, dict(os.environ)) for k in (1, 4)]}
env = dict(os.environ, CUDA_MPS_PIPE_DIRECTORY="/tmp/mps", CUDA_MPS_LOG_DIRECTORY="/tmp/mps_log")
os.makedirs("/tmp/mps", exist_ok=True)
os.makedirs("/tmp/mps_log", exist_ok=True)
r = subprocess.run(["nvidia-cuda-mps-control", "-d"], env=env, capture_output=True, text=True)
out["mps_daemon_rc"] = r.returncode
if r.returncode == 0:
    time.sleep(2)
    out["mps"] = [run(k, env) for k in (1, 4, 8, 16)]
    subprocess.run(["nvidia-cuda-mps-control"], input="quit\n", env=env, capture_output=True, text=True)
print(json.dumps(out))
