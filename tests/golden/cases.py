"""Golden-vector cases for the Thor hot path and the three back ends that can run them (test infrastructure only):

  RefBackend    the compiled, unmodified reference (oracle/_ref, built from /root/reference) -> used ONLY by
                make_golden.py in the build container to generate tests/golden/*.npz
  OracleBackend oracle/libthor_oracle.so (our plain-C restatement)                      -> CPU test
  GpuBackend    libthor_b200.so through its C ABI                                       -> -m gpu test

Each case is (kind, params dict of ints/floats, inputs dict of numpy arrays) -> outputs dict of numpy arrays.
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from _libs import aligned, P, sdt, sfx  # noqa: E402
from _refstructs import Frame as HFrame, random_blkinfo, CdefStrengths, DeblockData, BLKINFO, Mv  # noqa: E402


def al(a):
    b = aligned(a.shape, a.dtype)
    b[...] = a
    return b


def make_cases(hbd, bd, seed=2026):
    rng = np.random.default_rng(seed + hbd)
    dt = sdt(hbd)
    maxv = (1 << bd) - 1
    cases = []
    yy, xx = np.mgrid[0:96, 0:128]
    plane = np.clip((np.sin(xx / 7.0) + np.cos(yy / 5.0)) * (maxv / 8.0) + maxv / 2.0 + rng.integers(-9, 10, (96, 128)), 0, maxv).astype(dt)
    noise = rng.integers(0, maxv + 1, (96, 128)).astype(dt)
    for i, (w, h) in enumerate([(8, 8), (16, 16), (16, 8), (8, 16), (32, 32), (64, 64), (32, 16), (4, 4), (4, 8)]):
        size = max(w, h)
        src = plane if i % 2 == 0 else noise
        oy, ox = int(rng.integers(4, 96 - size - 6)), int(rng.integers(4, 128 - size - 6))
        cases.append(("sad", dict(w=w, h=h, y=oy + int(rng.integers(-2, 3)), x=ox + int(rng.integers(-2, 3)), fx=int(rng.choice([0, 2, -2])), fy=int(rng.choice([0, 2, -2]))),
                      dict(org=src[oy:oy + size, ox:ox + size].copy(), plane=src)))
    for (w, h) in [(4, 4), (8, 8), (16, 16), (8, 4), (32, 32), (64, 64)]:
        for chroma in (0, 1):
            for _ in range(3):
                frac = 8 if chroma else 4
                xo, yo = int(rng.integers(0, frac)), int(rng.integers(0, frac))
                if xo == 0 and yo == 0:
                    xo = 1
                cases.append(("interp", dict(w=w, h=h, xo=xo, yo=yo, bip=int(rng.integers(0, 3)), chroma=chroma, y=int(rng.integers(4, 96 - h - 6)), x=int(rng.integers(4, 128 - w - 6))),
                              dict(plane=noise)))
    for size in (4, 8, 16, 32, 64, 128):
        for amp in (maxv, 12):
            blk = rng.integers(-amp, amp + 1, (size, size)).astype(np.int16)
            cases.append(("txfm", dict(size=size, fast=int(rng.integers(0, 2)), qp=int(rng.choice([18, 27, 32, 36, 45])), typ=int(rng.integers(0, 4))), dict(block=blk)))
    for size in (4, 8, 16, 32):
        patch = rng.integers(0, maxv + 1, (2 * size + 8, 2 * size + 8)).astype(dt)
        for (xpos, ypos) in [(size, size), (0, size), (size, 0)]:
            cases.append(("intra", dict(size=size, xpos=xpos, ypos=ypos, ur=int(xpos > 0), dl=int(ypos > 0 and xpos > 0)), dict(patch=patch)))
    for n in (8, 16, 32):
        y = plane[:n, :n].copy()
        ry = np.clip(y.astype(int) + rng.integers(-40, 41, (n, n)), 0, maxv).astype(dt)
        ys = y.astype(int).reshape(n // 2, 2, n // 2, 2).sum(axis=(1, 3)) // 4
        u = np.clip(ys * 0.8 + maxv * 0.1 + rng.integers(-3, 4, ys.shape), 0, maxv).astype(dt)
        v = np.clip(maxv - ys * 0.6 + rng.integers(-3, 4, ys.shape), 0, maxv).astype(dt)
        cases.append(("cfl", dict(n=n), dict(y=y, u=u, v=v, ry=ry)))
    # motion search on a 128x96 frame pair
    fw, fh = 128, 96
    refy = plane[:fh, :fw].copy()
    cury = np.clip(np.roll(refy.astype(int), (1, -2), axis=(0, 1)) + rng.integers(-3, 4, (fh, fw)), 0, maxv).astype(dt)
    for speed, bip in ((0, 1), (1, 0), (2, 0)):
        items = []
        for _ in range(12):
            size = int(rng.choice([8, 16, 32]))
            part = int(rng.integers(0, 4))
            bw, bh, ox, oy = size, size, 0, 0
            if part == 1: bh = size // 2; oy = int(rng.integers(0, 2)) * bh
            if part == 2: bw = size // 2; ox = int(rng.integers(0, 2)) * bw
            if part == 3: bw = bh = size // 2; ox = int(rng.integers(0, 2)) * bw; oy = int(rng.integers(0, 2)) * bh
            items.append([size, bw, bh, ox, oy, int(rng.integers(0, fw // size)) * size, int(rng.integers(0, fh // size)) * size, int(rng.integers(0, 2)),
                          int(rng.integers(-20, 20)), int(rng.integers(-20, 20)), int(rng.integers(-20, 20)), int(rng.integers(-20, 20)), int(rng.integers(1, 5))])
        cands = rng.integers(-8, 8, (12, 4, 2)).astype(np.int16)
        lam = rng.uniform(3.0, 40.0, 12)
        cases.append(("me", dict(speed=speed, bip=bip, fw=fw, fh=fh), dict(ref=refy, cur=cury, items=np.array(items, np.int32), cands=cands, lam=lam)))
    # in-loop filter chain on a 128x72 frame
    w, h = 128, 72
    f = HFrame(w, h, bd, hbd, 32, 32)
    f.randomize(rng)
    org = [np.clip(f.plane(p).astype(int) + rng.integers(-5, 6, f.plane(p).shape), 0, maxv).astype(dt) for p in range(3)]
    bi, _ = random_blkinfo(rng, w, h)
    nfb = ((w + 63) // 64) * ((h + 63) // 64)
    cases.append(("filters", dict(w=w, h=h, qp=34), dict(y=f.y.copy(), u=f.u.copy(), v=f.v.copy(), oy=org[0], ou=org[1], ov=org[2], bi=bi.view(np.uint8).reshape(-1).copy(),
                                                          pri=rng.integers(0, 16, (2, nfb)).astype(np.int8), sec=rng.integers(0, 4, (2, nfb)).astype(np.int8))))
    return cases


def _bi_to_dd(bi):
    flat = bi.reshape(-1)
    dd = (DeblockData * len(flat))()
    for i, r in enumerate(flat):
        d = dd[i]
        d.mode = int(r["mode"]); d.cbp.y = int(r["cbp_y"]); d.size = int(r["size"]); d.tb_split = int(r["tb_split"]); d.pb_part = int(r["pb_part"])
        d.inter_pred.mv0.x, d.inter_pred.mv0.y, d.inter_pred.mv1.x, d.inter_pred.mv1.y = int(r["mv0x"]), int(r["mv0y"]), int(r["mv1x"]), int(r["mv1y"])
    return dd


class _Base:
    def __init__(self, hbd, bd):
        self.hbd, self.bd, self.s, self.dt = hbd, bd, sfx(hbd), sdt(hbd)

    def run(self, kind, p, x):
        return getattr(self, "run_" + kind)(p, x)

    def host_frame(self, y, u=None, v=None, pad=160):
        h, w = y.shape
        f = HFrame(w, h, self.bd, self.hbd, pad, pad)
        f.y[...] = y
        if u is not None:
            f.u[...] = u; f.v[...] = v
        return f


class OracleBackend(_Base):
    def __init__(self, hbd, bd):
        super().__init__(hbd, bd)
        from _libs import oracle
        self.O = oracle()

    def fn(self, name):
        return getattr(self.O, "orc_" + name + "_" + self.s)

    def run_sad(self, p, x):
        org, plane = al(x["org"]), al(x["plane"])
        size, st = org.shape[0], plane.shape[1]
        b = P(plane, p["y"] * st + p["x"])
        out = [self.fn("sad")(P(org), b, size, st, p["w"], p["h"])]
        xo = C.c_int(0)
        out += [self.fn("widesad")(P(org), b, size, st, p["w"], p["h"], C.byref(xo)), xo.value]
        out += [self.fn("ssd")(P(org), P(plane, p["y"] * st + (p["x"] & ~31)), size, st, p["w"], p["h"])]
        a, c = C.c_int(0), C.c_int(0)
        out += [self.fn("sad_fasthalf")(P(org), b, size, st, p["w"], p["h"], C.byref(a), C.byref(c)), a.value, c.value]
        a, c = C.c_int(p["fx"]), C.c_int(p["fy"])
        out += [self.fn("sad_fastquarter")(P(org), b, size, st, p["w"], p["h"], C.byref(a), C.byref(c)), a.value, c.value]
        return dict(vals=np.array(out, np.int64))

    def run_interp(self, p, x):
        plane = al(x["plane"])
        st = plane.shape[1]
        o = aligned((p["h"], p["w"]), self.dt)
        if p["chroma"]:
            self.fn("interp_chroma")(p["w"], p["h"], p["xo"], p["yo"], P(o), p["w"], P(plane, p["y"] * st + p["x"]), st, self.bd)
        else:
            self.fn("interp_luma")(p["w"], p["h"], p["xo"], p["yo"], P(o), p["w"], P(plane, p["y"] * st + p["x"]), st, p["bip"], self.bd)
        return dict(out=o.copy())

    def run_txfm(self, p, x):
        O, size, q = self.O, p["size"], min(p["size"], 16)
        blk = al(x["block"])
        cf = aligned((size, size), np.int16, fill=0); cq = aligned((q * q,), np.int16, fill=0)
        rc = aligned((size, size), np.int16, fill=0); rb = aligned((size, size), np.int16, fill=0)
        O.orc_transform(P(blk), P(cf), size, p["fast"], self.bd)
        cbp = O.orc_quantize(P(cf), P(cq), p["qp"], size, p["typ"], None)
        O.orc_dequantize(P(cq), P(rc), p["qp"], size, None)
        O.orc_inverse_transform(P(rc), P(rb), size, self.bd)
        return dict(coeff=cf[:q, :q].copy(), coeffq=cq.copy(), cbp=np.array([cbp]), rcoeff=rc[:q, :q].copy(), rblock=rb.copy())

    def run_intra(self, p, x):
        patch = al(x["patch"])
        st, size = patch.shape[1], p["size"]
        base = (4 + p["ypos"]) * st + 4 + p["xpos"]
        left = aligned((264,), self.dt, fill=0); top = aligned((264,), self.dt, fill=0); tl = aligned((1,), self.dt, fill=0)
        self.fn("make_top_and_left")(P(left), P(top), P(tl), P(patch, base), st, None, 0, 0, 0, p["ypos"], p["xpos"], size, p["ur"], p["dl"], 0, self.bd)
        preds = np.zeros((10, size, size), self.dt)
        for m in range(10):
            o = aligned((size, size), self.dt)
            self.fn("intra_pred")(P(left), P(top), int(tl[0]), p["ypos"], p["xpos"], size, P(o), size, m, self.bd)
            preds[m] = o
        return dict(preds=preds)

    def run_cfl(self, p, x):
        y, u, v, ry = al(x["y"]), al(x["u"]), al(x["v"]), al(x["ry"])
        self.fn("cfl")(P(y), P(u), P(v), P(ry), p["n"], p["n"], p["n"], 1, self.bd)
        return dict(u=u.copy(), v=v.copy())

    def run_me(self, p, x):
        f = self.host_frame(x["ref"])
        self.fn("pad_plane")(P(f.Y, f.origin(0)), f.sy, p["fw"], p["fh"], 160, 160)
        out = []
        for it, cc, lam in zip(x["items"], x["cands"], x["lam"]):
            size, bw, bh, ox, oy, xpos, ypos, sign, mcx, mcy, mpx, mpy, nc = [int(v) for v in it]
            org = aligned((size, size), self.dt)
            org[...] = x["cur"][ypos:ypos + size, xpos:xpos + size]
            m = (C.c_int16 * 2)(0, 0)
            cands = (C.c_int16 * 8)(*[int(v) for v in cc.reshape(-1)])
            cost = self.fn("motion_estimate")(P(org, oy * size + ox), P(f.Y, f.origin(0) + (ypos + oy) * f.sy + xpos + ox), size, f.sy, bw, bh, m,
                                              (C.c_int16 * 2)(mcx, mcy), (C.c_int16 * 2)(mpx, mpy), C.c_double(float(lam)), p["speed"], self.bd, sign, p["fw"], p["fh"],
                                              xpos, ypos, cands, nc, p["bip"])
            out.append([cost, m[0], m[1]])
        return dict(res=np.array(out, np.int64))

    def run_filters(self, p, x):
        O, w, h, qp, bd = self.O, p["w"], p["h"], p["qp"], self.bd
        f = self.host_frame(x["y"], x["u"], x["v"], pad=32)
        bi = al(x["bi"].view(BLKINFO).reshape(h // 4, w // 4))
        out = {}
        self.fn("deblock_y")(P(f.Y, f.origin(0)), f.sy, P(bi), w, h, qp, bd)
        self.fn("deblock_uv")(P(f.U, f.origin(1)), P(f.V, f.origin(1)), f.sc, P(bi), w, h, 1, O.orc_chroma_qp(qp), bd)
        for k, pl in enumerate("yuv"):
            out["db_" + pl] = f.plane(k).copy()
        nfb = x["pri"].shape[1]
        dirs = np.zeros((nfb, 64), np.int32); vars_ = np.zeros((nfb, 64), np.int32)
        pri, sec = al(x["pri"]), al(x["sec"])
        for plane in range(3):
            src = f.full(plane); dst = src.copy()
            self.fn("cdef_plane")(P(src, f.origin(plane)), P(dst, f.origin(plane)), f.stride(plane), w, h, P(bi), w // 4, 1, plane, P(pri[int(plane > 0)]),
                                  P(sec[int(plane > 0)]), 5, 5, P(dirs), P(vars_), bd)
            src[...] = dst
            out["cdef_" + "yuv"[plane]] = f.plane(plane).copy()
        for plane, (fbl, strength) in enumerate(((6, 2), (4, 1), (4, 4))):
            src = f.full(plane); dst = src.copy()
            pw, ph = (w, h) if plane == 0 else (w // 2, h // 2)
            self.fn("clpf_plane")(P(src, f.origin(plane)), P(dst, f.origin(plane)), f.stride(plane), pw, ph, P(bi), w // 4, int(plane != 0), None, fbl, strength, bd, plane, qp)
            src[...] = dst
            out["clpf_" + "yuv"[plane]] = f.plane(plane).copy()
        return out


class RefBackend(_Base):
    """The unmodified reference (SIMD path, use_simd = 1)."""

    def __init__(self, hbd, bd):
        super().__init__(hbd, bd)
        from _libs import ref, ref_enc
        self.R, self.E = ref(), ref_enc(hbd)
        assert self.R is not None and self.E is not None, "oracle/_ref not built"

    def r(self, name):
        return getattr(self.R, name + "_" + self.s)

    def run_sad(self, p, x):
        org, plane = al(x["org"]), al(x["plane"])
        size, st, w, h = org.shape[0], plane.shape[1], p["w"], p["h"]
        b = P(plane, p["y"] * st + p["x"])
        e = lambda n: getattr(self.E, "ref_" + n + "_" + self.s)
        out = [e("sad_calc")(P(org), b, size, st, w, h)]
        xo = C.c_int(0)
        out += [e("widesad_calc")(P(org), b, size, st, w, h, C.byref(xo)), xo.value]
        out += [e("ssd_calc")(P(org), P(plane, p["y"] * st + (p["x"] & ~31)), size, st, w, h)]  # second operand must be aligned for the SIMD path
        a, c = C.c_int(0), C.c_int(0)
        fh = (lambda *q: self.r("sad_calc_fasthalf_simd")(*q, None)) if w > 4 else e("sad_calc_fasthalf")
        out += [fh(P(org), b, size, st, w, h, C.byref(a), C.byref(c)), a.value, c.value]
        a, c = C.c_int(p["fx"]), C.c_int(p["fy"])
        fq = self.r("sad_calc_fastquarter_simd") if w > 4 else e("sad_calc_fastquarter")
        out += [fq(P(org), b, size, st, w, h, C.byref(a), C.byref(c)), a.value, c.value]
        return dict(vals=np.array(out, np.int64))

    def run_interp(self, p, x):
        plane = al(x["plane"])
        st = plane.shape[1]
        o = aligned((p["h"], p["w"]), self.dt)
        if p["chroma"]:
            self.r("get_inter_prediction_chroma_simd")(p["w"], p["h"], p["xo"], p["yo"], P(o), p["w"], P(plane, p["y"] * st + p["x"]), st, self.bd)
        else:
            self.r("get_inter_prediction_luma_simd")(p["w"], p["h"], p["xo"], p["yo"], P(o), p["w"], P(plane, p["y"] * st + p["x"]), st, p["bip"], self.bd)
        return dict(out=o.copy())

    def run_txfm(self, p, x):
        R, size, q = self.R, p["size"], min(p["size"], 16)
        blk = al(x["block"])
        cf = aligned((size, size), np.int16, fill=0); cq = aligned((q * q,), np.int16, fill=0)
        rc = aligned((size, size), np.int16, fill=0); rb = aligned((size, size), np.int16, fill=0)
        R.transform(P(blk), P(cf), size, p["fast"], self.bd)
        cbp = getattr(self.E, "ref_quantize_" + self.s)(P(cf), P(cq), p["qp"], size, p["typ"], None)
        self.r("dequantize")(P(cq), P(rc), p["qp"], size, None)
        R.inverse_transform(P(rc), P(rb), size, self.bd)
        return dict(coeff=cf[:q, :q].copy(), coeffq=cq.copy(), cbp=np.array([cbp]), rcoeff=rc[:q, :q].copy(), rblock=rb.copy())

    def run_intra(self, p, x):
        patch = al(x["patch"])
        st, size = patch.shape[1], p["size"]
        base = (4 + p["ypos"]) * st + 4 + p["xpos"]
        left = aligned((264,), self.dt, fill=0); top = aligned((264,), self.dt, fill=0); tl = aligned((1,), self.dt, fill=0)
        self.r("make_top_and_left")(P(left), P(top), P(tl), P(patch, base), st, None, 0, 0, 0, p["ypos"], p["xpos"], size, p["ur"], p["dl"], 0, self.bd)
        preds = np.zeros((10, size, size), self.dt)
        for m in range(10):
            o = aligned((size, size), self.dt)
            self.r("get_intra_prediction")(P(left), P(top), int(tl[0]), p["ypos"], p["xpos"], size, P(o), size, m, self.bd)
            preds[m] = o
        return dict(preds=preds)

    def run_cfl(self, p, x):
        y, u, v, ry = al(x["y"]), al(x["u"]), al(x["v"]), al(x["ry"])
        self.r("improve_uv_prediction")(P(y), P(u), P(v), P(ry), p["n"], p["n"], p["n"], 1, self.bd)
        return dict(u=u.copy(), v=v.copy())

    def run_me(self, p, x):
        f = self.host_frame(x["ref"])
        self.r("pad_yuv_frame")(C.byref(f.s))
        out = []
        for it, cc, lam in zip(x["items"], x["cands"], x["lam"]):
            size, bw, bh, ox, oy, xpos, ypos, sign, mcx, mcy, mpx, mpy, nc = [int(v) for v in it]
            org = aligned((size, size), self.dt)
            org[...] = x["cur"][ypos:ypos + size, xpos:xpos + size]
            m = (C.c_int16 * 2)(0, 0)
            cands = (C.c_int16 * 8)(*[int(v) for v in cc.reshape(-1)])
            cost = getattr(self.E, "ref_motion_estimate_" + self.s)(P(org, oy * size + ox), P(f.Y, f.origin(0) + (ypos + oy) * f.sy + xpos + ox), size, f.sy, bw, bh, m,
                                                                    (C.c_int16 * 2)(mcx, mcy), (C.c_int16 * 2)(mpx, mpy), C.c_double(float(lam)), p["speed"], self.bd, sign,
                                                                    p["fw"], p["fh"], xpos, ypos, cands, nc, p["bip"])
            out.append([cost, m[0], m[1]])
        return dict(res=np.array(out, np.int64))

    def run_filters(self, p, x):
        from _libs import oracle
        w, h, qp, bd = p["w"], p["h"], p["qp"], self.bd
        f = self.host_frame(x["y"], x["u"], x["v"], pad=32)
        bi = x["bi"].view(BLKINFO).reshape(h // 4, w // 4)
        dd = _bi_to_dd(bi)
        out = {}
        self.r("deblock_frame_y")(C.byref(f.s), dd, w, h, qp, bd)
        self.r("deblock_frame_uv")(C.byref(f.s), dd, w, h, oracle().orc_chroma_qp(qp), bd)
        for k, pl in enumerate("yuv"):
            out["db_" + pl] = f.plane(k).copy()
        nfb = x["pri"].shape[1]
        cst = (CdefStrengths * nfb)()
        for i in range(nfb):
            for pl in range(2):
                cst[i].plane[pl].level = int(x["pri"][pl, i]); cst[i].plane[pl].sec_strength = int(x["sec"][pl, i])
                cst[i].plane[pl].pri_damping = cst[i].plane[pl].sec_damping = 5
        for plane in range(3):
            self.r("cdef_frame")(cst, C.byref(f.s), None, dd, None, 0, bd, plane)
            out["cdef_" + "yuv"[plane]] = f.plane(plane).copy()
        for plane, (fbl, strength) in enumerate(((6, 2), (4, 1), (4, 4))):
            self.r("clpf_frame")(C.byref(f.s), None, dd, None, 0, strength, fbl, bd, plane, qp, None)
            out["clpf_" + "yuv"[plane]] = f.plane(plane).copy()
        return out


class GpuBackend(_Base):
    """libthor_b200.so through its C ABI (drop-in symbols for block cases, batched/frame API for the rest)."""

    def __init__(self, hbd, bd):
        super().__init__(hbd, bd)
        import thor_b200 as tb
        tb.init(0)
        self.tb, self.L, self.esz = tb, tb.lib, 2 if hbd else 1

    def l(self, name):
        return getattr(self.L, name + "_" + self.s)

    def run_sad(self, p, x):
        org, plane = al(x["org"]), al(x["plane"])
        size, st, w, h = org.shape[0], plane.shape[1], p["w"], p["h"]
        b = P(plane, p["y"] * st + p["x"])
        out = [self.l("sad_calc_simd")(P(org), b, size, st, w, h)]
        xo = C.c_int(0)
        out += [self.l("widesad_calc_simd")(P(org), b, size, st, w, h, C.byref(xo)), xo.value]
        bs = P(plane, p["y"] * st + (p["x"] & ~31))
        out += [self.l("ssd_calc_simd")(P(org), bs, size, st, w) if w == h else OracleBackend(self.hbd, self.bd).run_sad(p, x)["vals"][3]]
        a, c = C.c_int(0), C.c_int(0)
        out += [self.l("sad_calc_fasthalf_simd")(P(org), b, size, st, w, h, C.byref(a), C.byref(c)), a.value, c.value]
        a, c = C.c_int(p["fx"]), C.c_int(p["fy"])
        out += [self.l("sad_calc_fastquarter_simd")(P(org), b, size, st, w, h, C.byref(a), C.byref(c)), a.value, c.value]
        return dict(vals=np.array(out, np.int64))

    def run_interp(self, p, x):
        plane = al(x["plane"])
        st = plane.shape[1]
        o = aligned((p["h"], p["w"]), self.dt)
        if p["chroma"]:
            self.l("get_inter_prediction_chroma_simd")(p["w"], p["h"], p["xo"], p["yo"], P(o), p["w"], P(plane, p["y"] * st + p["x"]), st, self.bd)
        else:
            self.l("get_inter_prediction_luma_simd")(p["w"], p["h"], p["xo"], p["yo"], P(o), p["w"], P(plane, p["y"] * st + p["x"]), st, p["bip"], self.bd)
        return dict(out=o.copy())

    def run_txfm(self, p, x):
        L, size, q = self.L, p["size"], min(p["size"], 16)
        blk = al(x["block"])
        cf = aligned((size, size), np.int16, fill=0); cq = aligned((q * q,), np.int16, fill=0)
        rc = aligned((size, size), np.int16, fill=0); rb = aligned((size, size), np.int16, fill=0)
        L.transform_simd(P(blk), P(cf), size, p["fast"], self.bd)
        cbp = L.tb_quantize(P(cf), P(cq), p["qp"], size, p["typ"])
        L.tb_dequantize(P(cq), P(rc), p["qp"], size)
        if size < 64:
            L.inverse_transform_simd(P(rc), P(rb), size, self.bd)
        else:  # host-side 64/128 wrapper of common/transform.c:471-494: 32x32 kernel + replication
            c2 = aligned((32, 32), np.int16, fill=0); b2 = aligned((32, 32), np.int16, fill=0)
            c2[...] = rc[:32, :32]
            L.inverse_transform_simd(P(c2), P(b2), 32, self.bd)
            rb[...] = np.repeat(np.repeat(b2, size // 32, axis=0), size // 32, axis=1)
        return dict(coeff=cf[:q, :q].copy(), coeffq=cq.copy(), cbp=np.array([cbp]), rcoeff=rc[:q, :q].copy(), rblock=rb.copy())

    def run_intra(self, p, x):
        tb, size = self.tb, p["size"]
        patch = np.ascontiguousarray(x["patch"])
        st = patch.shape[1]
        dpatch = tb.DevBuf.from_array(patch)
        items = np.zeros(10, tb.INTRA_ITEM)
        out = tb.DevBuf(10 * size * size * self.esz)
        for m in range(10):
            items[m] = (dpatch.ptr + ((4 + p["ypos"]) * st + 4 + p["xpos"]) * self.esz, out.ptr + m * size * size * self.esz, st, p["xpos"], p["ypos"], size, m, p["ur"], p["dl"])
        d_items = tb.DevBuf.from_array(items)
        tb.check(self.L.tb_intra_batch(d_items.ptr, 10, self.esz, self.bd))
        return dict(preds=out.download(self.dt, (10, size, size)))

    def run_cfl(self, p, x):
        y, u, v, ry = al(x["y"]), al(x["u"]), al(x["v"]), al(x["ry"])
        self.L.tb_improve_uv_prediction(self.esz, P(y), P(u), P(v), P(ry), p["n"], p["n"], p["n"], 1, self.bd)
        return dict(u=u.copy(), v=v.copy())

    def run_me(self, p, x):
        tb, fw, fh = self.tb, p["fw"], p["fh"]
        zero = np.zeros((fh // 2, fw // 2), self.dt)
        drec = tb.Frame(fw, fh, self.esz); dref = tb.Frame(fw, fh, self.esz); dcur = tb.Frame(fw, fh, self.esz)
        drec.upload(x["ref"], zero, zero); dcur.upload(x["cur"], zero, zero)
        tb.check(self.L.tb_create_reference_frame(dref.h, drec.h))
        rptr, rst = dref.plane(0); optr, ost = dcur.plane(0)
        n = len(x["items"])
        items = np.zeros(n, tb.ME_ITEM)
        for i, (it, lam) in enumerate(zip(x["items"], x["lam"])):
            size, bw, bh, ox, oy, xpos, ypos, sign, mcx, mcy, mpx, mpy, nc = [int(v) for v in it]
            items[i] = (optr + ((ypos + oy) * ost + xpos + ox) * self.esz, rptr + ((ypos + oy) * rst + xpos + ox) * self.esz, ost, rst, xpos, ypos, size, bw, bh, sign,
                        mcx, mcy, mpx, mpy, 4 * i, nc, float(lam))
        d_items = tb.DevBuf.from_array(items); d_c = tb.DevBuf.from_array(np.ascontiguousarray(x["cands"], np.int16)); d_out = tb.DevBuf(8 * n)
        tb.check(self.L.tb_motion_estimate_batch(d_items.ptr, n, d_c.ptr, self.esz, self.bd, p["speed"], p["bip"], fw, fh, d_out.ptr))
        r = d_out.download(tb.ME_RESULT, n)
        return dict(res=np.stack([r["cost"].astype(np.int64), r["mvx"].astype(np.int64), r["mvy"].astype(np.int64)], axis=1))

    def run_filters(self, p, x):
        tb, L, w, h, qp, bd = self.tb, self.L, p["w"], p["h"], p["qp"], self.bd
        drec = tb.Frame(w, h, self.esz, 32); dscr = tb.Frame(w, h, self.esz, 32)
        drec.upload(x["y"], x["u"], x["v"])
        dbi = tb.DevBuf.from_array(x["bi"])
        out = {}
        tb.check(L.tb_deblock_frame(drec.h, dbi.ptr, qp, bd))
        for k, a in zip("yuv", drec.download()):
            out["db_" + k] = a
        nfb = x["pri"].shape[1]
        dpri = [tb.DevBuf.from_array(x["pri"][k]) for k in range(2)]; dsec = [tb.DevBuf.from_array(x["sec"][k]) for k in range(2)]
        ddv = tb.DevBuf(nfb * 2 * 64 * 4)
        for plane in range(3):
            tb.check(L.tb_cdef_frame(drec.h, dscr.h, dbi.ptr, dpri[int(plane > 0)].ptr, dsec[int(plane > 0)].ptr, 5, 5, ddv.ptr, bd, plane))
            out["cdef_" + "yuv"[plane]] = drec.download()[plane]
        for plane, (fbl, strength) in enumerate(((6, 2), (4, 1), (4, 4))):
            tb.check(L.tb_clpf_frame(drec.h, dscr.h, dbi.ptr, None, fbl, strength, bd, plane, qp))
            out["clpf_" + "yuv"[plane]] = drec.download()[plane]
        return out


def golden_path(hbd):
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_%s.npz" % sfx(hbd))


def load_golden(hbd):
    z = np.load(golden_path(hbd))
    n = int(z["n"])
    return [{k[len("c%d_" % i):]: z[k] for k in z.files if k.startswith("c%d_" % i)} for i in range(n)]
