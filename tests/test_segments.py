"""Multi-GPU sharding logic on CPU: 2 ranks over gloo, each encoding its intra-period segments with
the reference encoder as the per-rank worker; the gathered, frame_num-patched stream must equal the monolithic stream."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thor_b200 import segments as sg  # noqa: E402
from test_dropin_link import synth_yuv, HDB  # noqa: E402

THORENC = os.path.join(ROOT, "oracle", "_ref", "Thorenc")
needs_ref = pytest.mark.skipif(not os.path.exists(THORENC), reason="oracle/_ref not built")


def test_plan_and_patch():
    assert sg.plan_segments(33, 16, 2) == [[(0, 17)], [(16, 17)]]
    assert sg.plan_segments(65, 16, 2) == [[(0, 17), (16, 17)], [(32, 17), (48, 17)]]
    assert sg.plan_segments(40, 16, 4) == [[(0, 17)], [(16, 17)], [(32, 8)], []]
    assert sg.plan_segments(17, 0, 2) == [[(0, 17)], []]
    # P-frame header: type=1, qp=0x23, modes=10, num_ref-1=1 (2 refs), refs, then frame_num=5
    bits = "1" + format(0x23, "08b") + format(10, "04b") + "01" + format(3, "06b") + format(1, "06b") + format(5, "016b") + "1011"
    bits += "0" * (-len(bits) % 8)
    payload = int(bits, 2).to_bytes(len(bits) // 8, "big")
    chunk = len(payload).to_bytes(4, "big") + payload
    assert sg.frame_num_bit_offset(payload) == 27
    out = sg.patch_frame_num(chunk, 64)
    got = format(int.from_bytes(out[4:], "big"), "0%db" % len(bits))
    assert int(got[27:43], 2) == 69 and got[:27] == bits[:27] and got[43:] == bits[43:]


def _worker(rank, world, port, tmp, w, h, n, period):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    merged = sg.encode_sharded(THORENC, HDB + ["-qp", "32", "-f", "30"], os.path.join(tmp, "in.yuv"), w, h, n, period,
                               out_path=os.path.join(tmp, "sharded.bit") if rank == 0 else None, workdir=tmp)
    assert (merged is not None) == (rank == 0)
    dist.destroy_process_group()


@needs_ref
def test_two_rank_sharded_encode_equals_monolithic(tmp_path):
    import torch.multiprocessing as mp
    tmp = str(tmp_path)
    w, h, n, period = 64, 64, 33, 16
    synth_yuv(os.path.join(tmp, "in.yuv"), w, h, n, 8, seed=9)
    mono = os.path.join(tmp, "mono.bit")
    r = subprocess.run([THORENC] + HDB + ["-qp", "32", "-f", "30", "-intra_period", str(period), "-if", os.path.join(tmp, "in.yuv"), "-of", mono, "-width", str(w),
                        "-height", str(h), "-n", str(n)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, tmp, w, h, n, period), nprocs=2, join=True)
    a, b = open(mono, "rb").read(), open(os.path.join(tmp, "sharded.bit"), "rb").read()
    assert len(sg.split_chunks(a)) == n
    assert a == b


@needs_ref
def test_three_rank_six_segments(tmp_path):
    """several segments per rank, three ranks: the merge must stitch segments within a rank and across ranks"""
    import torch.multiprocessing as mp
    tmp = str(tmp_path)
    w, h, n, period = 64, 64, 49, 8
    synth_yuv(os.path.join(tmp, "in.yuv"), w, h, n, 8, seed=10)
    mono = os.path.join(tmp, "mono.bit")
    r = subprocess.run([THORENC] + HDB + ["-qp", "32", "-f", "30", "-intra_period", str(period), "-if", os.path.join(tmp, "in.yuv"), "-of", mono, "-width", str(w),
                        "-height", str(h), "-n", str(n)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert [len(p) for p in sg.plan_segments(n, period, 3)] == [2, 2, 2]
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(3, port, tmp, w, h, n, period), nprocs=3, join=True)
    a, b = open(mono, "rb").read(), open(os.path.join(tmp, "sharded.bit"), "rb").read()
    assert len(sg.split_chunks(a)) == n
    assert a == b
