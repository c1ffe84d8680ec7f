/* ref_exports.c — TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile (target `ref`) into
 * oracle/_ref/libthorref_enc_{lbd,hbd}.so.  It #includes the reference's enc/encode_block.c FROM
 * /root/reference (include path given on the command line; nothing is copied into this repo) so that the
 * file-static kernels of the reference (quantize, motion_estimate, sad_calc_fasthalf, calc_cbp, ...) can be
 * called from the parity tests through the thin trampolines below. */
#ifdef ORC_HBD
#define SAMPLE uint16_t
#define TEMPLATE(name) name ## _hbd
#define HBD
#define X(name) ref_##name##_hbd
#else
#define X(name) ref_##name##_lbd
#endif
#include <stdint.h>
#include "encode_block.c"

int X(quantize)(int16_t *coeff, int16_t *coeffq, int qp, int size, int type, qmtx_t *wm) { return quantize(coeff, coeffq, qp, size, type, wm); }
void X(get_residual)(int16_t *block, SAMPLE *pblock, SAMPLE *orig, int size, int ps, int os) { get_residual(block, pblock, orig, size, ps, os); }
unsigned X(sad_calc_fasthalf)(const SAMPLE *a, const SAMPLE *b, int as, int bs, int w, int h, int *x, int *y) { return sad_calc_fasthalf(a, b, as, bs, w, h, x, y); }
unsigned X(sad_calc_fastquarter)(const SAMPLE *o, const SAMPLE *r, int os, int rs, int w, int h, int *x, int *y) { return sad_calc_fastquarter(o, r, os, rs, w, h, x, y); }
unsigned X(sad_calc)(SAMPLE *a, SAMPLE *b, int as, int bs, int w, int h) { return sad_calc(a, b, as, bs, w, h); }
unsigned X(widesad_calc)(SAMPLE *a, SAMPLE *b, int as, int bs, int w, int h, int *x) { return widesad_calc(a, b, as, bs, w, h, x); }
uint64_t X(ssd_calc)(SAMPLE *a, SAMPLE *b, int as, int bs, int w, int h) { return ssd_calc(a, b, as, bs, w, h); }
int X(quote_mv_bits)(int dy, int dx) { return quote_mv_bits(dy, dx); }
int X(calc_cbp)(int16_t *block, int size, int thr) { return calc_cbp(block, size, thr); }
int X(motion_estimate)(SAMPLE *orig, SAMPLE *ref, int size, int stride_r, int width, int height, mv_t *mv, mv_t *mvc, mv_t *mvp, double lambda,
                       int encoder_speed, int bitdepth, int sign, int fwidth, int fheight, int xpos, int ypos, mv_t *mvcand, int mvcand_num, int enable_bipred) {
  enc_params p;
  memset(&p, 0, sizeof(p));
  p.encoder_speed = encoder_speed;
  p.bitdepth = bitdepth;
  int n = mvcand_num;
  return motion_estimate(orig, ref, size, stride_r, width, height, mv, mvc, mvp, lambda, &p, sign, fwidth, fheight, xpos, ypos, mvcand, &n, enable_bipred);
}
int X(motion_estimate_bi)(SAMPLE *orig, SAMPLE *ref0, SAMPLE *ref1, int size, int stride_r, int width, int height, mv_t *mv, mv_t *mvc, mv_t *mvp, double lambda,
                          int bitdepth, int sign, int fwidth, int fheight, int xpos, int ypos, mv_t *mvcand, int mvcand_num, int enable_bipred) {
  enc_params p;
  memset(&p, 0, sizeof(p));
  p.bitdepth = bitdepth;
  mv_t list[8];  /* the reference scribbles on entries 0..5 of the list */
  memset(list, 0, sizeof(list));
  for (int i = 0; i < mvcand_num && i < 8; i++) list[i] = mvcand[i];
  int n = mvcand_num;
  return motion_estimate_bi(orig, ref0, ref1, size, stride_r, width, height, mv, mvc, mvp, lambda, &p, sign, fwidth, fheight, xpos, ypos, list, &n, enable_bipred);
}
int X(motion_estimate_sync)(SAMPLE *orig, SAMPLE *ref, int size, int stride_r, int width, int height, mv_t *mv, mv_t *mvc, mv_t *mvp, double lambda, int bitdepth, int sign,
                            int fwidth, int fheight, int xpos, int ypos, mv_t *mvcand, int enable_bipred) {
  enc_params p;
  memset(&p, 0, sizeof(p));
  p.bitdepth = bitdepth;
  int n = 6;
  return motion_estimate_sync(orig, ref, size, stride_r, width, height, mv, mvc, mvp, lambda, &p, sign, fwidth, fheight, xpos, ypos, mvcand, &n, enable_bipred);
}
void X(set_use_simd)(int v) { use_simd = v; }
