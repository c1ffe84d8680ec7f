/* ref_exports_frame.c — TEST INFRASTRUCTURE ONLY.  #includes the reference's enc/encode_frame.c in place (from /root/reference, via
 * the include path; nothing is copied) to reach cdef_search and its file-static helper dist_8x8 through trampolines that build
 * the minimal encoder_info_t the function reads (enc/encode_frame.c:228-470). */
#ifdef ORC_HBD
#define SAMPLE uint16_t
#define TEMPLATE(name) name ## _hbd
#define HBD
#define X(name) ref_##name##_hbd
#else
#define X(name) ref_##name##_lbd
#endif
#include <stdint.h>
#include "encode_frame.c"

uint64_t X(dist_8x8)(SAMPLE *dst, int dstride, SAMPLE *src, int sstride, int coeff_shift) { return dist_8x8(dst, dstride, src, sstride, coeff_shift); }

/* returns nb_strength_bits; level/sec: per filter block (raster) the chosen luma and chroma cdef_strength.level / .sec_strength;
 * dirs/vars: 64 ints per filter block; selbits: the put_flc'ed per-block selections as one bit string length */
int X(cdef_search)(yuv_frame_t *rec, yuv_frame_t *org, deblock_data_t *dd, int width, int height, int bitdepth, int qp, double lambda, int cdef_bits, int speed,
                   int *strengths, int *uv_strengths, int *level, int *sec, int *dirs, int *vars, int *stream_bits) {
  static encoder_info_t ei;
  static enc_params par;
  static stream_t st;
  static uint8_t buf[1 << 16];
  memset(&ei, 0, sizeof(ei));
  memset(&par, 0, sizeof(par));
  memset(&st, 0, sizeof(st));
  par.bitdepth = bitdepth;
  par.subsample = 420;
  par.cdef = speed + 1;
  ei.params = &par;
  ei.width = width;
  ei.height = height;
  ei.frame_info.qp = (uint8_t)qp;
  ei.frame_info.lambda = lambda;
  ei.cdef_damping = 5;
  ei.cdef_bits = cdef_bits;
  const int nfb = ((width + 63) >> 6) * ((height + 63) >> 6);
  ei.cdef = (cdef_strengths *)calloc((size_t)nfb, sizeof(cdef_strengths));
  st.bitstream = buf; st.bytesize = sizeof(buf); st.bitrest = 32;
  ei.stream = &st;
  int bits = TEMPLATE(cdef_search)(rec, org, dd, &ei.frame_info, &ei, strengths, uv_strengths, speed);
  for (int i = 0; i < nfb; i++) {
    level[2 * i] = ei.cdef[i].plane[0].level; level[2 * i + 1] = ei.cdef[i].plane[1].level;
    sec[2 * i] = ei.cdef[i].plane[0].sec_strength; sec[2 * i + 1] = ei.cdef[i].plane[1].sec_strength;
    for (int k = 0; k < 64; k++) { dirs[i * 64 + k] = ei.cdef[i].dir[k]; vars[i * 64 + k] = ei.cdef[i].var[k]; }
  }
  *stream_bits = get_bit_pos(&st);
  free(ei.cdef);
  return bits;
}
