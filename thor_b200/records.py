"""numpy record dtypes of the work-item structs in include/thor_b200.h.

Kept free of any library load so that host-only tools (bench.py's CPU arm, list builders) can use the layouts
without dlopen()ing libthor_b200.so:  importlib.util.spec_from_file_location("records", ".../thor_b200/records.py").
"""
import numpy as np

# ---- work-item record layouts (must match include/thor_b200.h) -------------------------------------------------
SAD_ITEM = np.dtype([("a", "u8"), ("b", "u8"), ("astride", "i4"), ("bstride", "i4"), ("width", "u2"), ("height", "u2"), ("pad", "u4")], align=True)
ME_ITEM = np.dtype([("orig", "u8"), ("ref", "u8"), ("ostride", "i4"), ("rstride", "i4"), ("xpos", "i2"), ("ypos", "i2"), ("size", "u1"),
                    ("width", "u1"), ("height", "u1"), ("sign", "u1"), ("mvc_x", "i2"), ("mvc_y", "i2"), ("mvp_x", "i2"), ("mvp_y", "i2"),
                    ("cand_ofs", "i4"), ("ncand", "i4"), ("lambda", "f8")], align=True)
ME_BI_ITEM = np.dtype([("orig", "u8"), ("ref0", "u8"), ("ref1", "u8"), ("ostride", "i4"), ("rstride", "i4"), ("xpos", "i2"), ("ypos", "i2"), ("size", "u1"),
                       ("sign", "u1"), ("pad0", "u1"), ("pad1", "u1"), ("mvc_x", "i2"), ("mvc_y", "i2"), ("mvp_x", "i2"), ("mvp_y", "i2"), ("cand_ofs", "i4"),
                       ("ncand", "i4"), ("lambda", "f8")], align=True)
COMBINE_ITEM = np.dtype([("a", "u8"), ("b", "u8"), ("dst", "u8"), ("astride", "i4"), ("bstride", "i4"), ("dstride", "i4"), ("width", "u2"), ("height", "u2")],
                        align=True)
ME_RESULT = np.dtype([("mvx", "i2"), ("mvy", "i2"), ("cost", "u4")], align=True)
INTERP_ITEM = np.dtype([("ref", "u8"), ("dst", "u8"), ("rstride", "i4"), ("dstride", "i4"), ("xpos", "i2"), ("ypos", "i2"), ("mvx", "i2"),
                        ("mvy", "i2"), ("width", "u1"), ("height", "u1"), ("sign", "u1"), ("chroma", "u1"), ("pic_w", "i2"), ("pic_h", "i2"),
                        ("pad", "u4")], align=True)
TXFM_ITEM = np.dtype([("orig", "u8"), ("pred", "u8"), ("rec", "u8"), ("coeffq", "u8"), ("ostride", "i4"), ("pstride", "i4"), ("rstride", "i4"),
                      ("size", "u1"), ("qp", "u1"), ("coeff_type", "u1"), ("fast", "u1")], align=True)
TXFM_RESULT = np.dtype([("ssd", "u8"), ("cbp", "i4"), ("bits", "i4")], align=True)
TXFM_FAST, TXFM_BITS = 1, 2
INTRA_ITEM = np.dtype([("rec", "u8"), ("dst", "u8"), ("rstride", "i4"), ("xpos", "i2"), ("ypos", "i2"), ("size", "u1"), ("mode", "u1"),
                       ("upright", "u1"), ("downleft", "u1")], align=True)
BLKINFO = np.dtype([("mode", "u1"), ("cbp_y", "u1"), ("size", "u1"), ("tb_split", "u1"), ("pb_part", "u1"), ("pad", "u1", 3), ("mv0x", "i2"),
                    ("mv0y", "i2"), ("mv1x", "i2"), ("mv1y", "i2")])
assert SAD_ITEM.itemsize == 32 and ME_ITEM.itemsize == 56 and ME_RESULT.itemsize == 8 and INTERP_ITEM.itemsize == 48
assert TXFM_ITEM.itemsize == 48 and TXFM_RESULT.itemsize == 16 and INTRA_ITEM.itemsize == 32 and BLKINFO.itemsize == 16

