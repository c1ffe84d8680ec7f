/* thor_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the cisco/thor per-block hot path, written from the
 * reference's canonical C (file:line cited at every function in thor_oracle_tmpl.h).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm may load this; the product
 * (libthor_b200.so) never links or calls it.
 *
 * PINNING: the reference ships no golden vectors, so this oracle is pinned against the
 * compiled reference itself (oracle/_ref/libthorref.so + libthorref_enc.so, built by oracle/Makefile from
 * /root/reference in place): tests/test_oracle_vs_ref.py compares every function below with the reference
 * symbol on seeded inputs, and tests/golden/ (npz files) holds reference-generated vectors that travel to the GPU box.
 *
 * Every function exists twice: suffix _lbd (SAMPLE = uint8_t) and _hbd (SAMPLE = uint16_t), like the
 * reference's TEMPLATE() scheme (common/types.h:35-38).
 */
#ifndef THOR_ORACLE_H
#define THOR_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* compact per-4x4 block record the frame filters need (subset of deblock_data_t, common/types.h:178-187) */
typedef struct {
  uint8_t mode;      /* block_mode_t: 0 SKIP, 1 INTRA, 2 INTER, 3 BIPRED, 4 MERGE */
  uint8_t cbp_y;     /* cbp.y != 0 */
  uint8_t size;      /* coding block size */
  uint8_t tb_split;
  uint8_t pb_part;   /* part_t: 0 NONE, 1 HOR, 2 VER, 3 QUAD */
  uint8_t pad[3];
  int16_t mv0x, mv0y, mv1x, mv1y;
} orc_blkinfo_t;

typedef struct { int16_t x, y; } orc_mv_t;

/* tables (generated, checked against the reference's exported data in tests) */
const int16_t *orc_dct_matrix(int log2size);           /* common/transform.c:37-241 */
const int     *orc_zigzag(int qsize);                  /* common/common_tables.c:29-62 */
int            orc_chroma_qp(int qp);                  /* common/common_tables.c:65-70 */

/* bit-depth independent */
void orc_transform(const int16_t *block, int16_t *coeff, int size, int fast, int bitdepth);
void orc_inverse_transform(const int16_t *coeff, int16_t *block, int size, int bitdepth);
int  orc_vlc_len(int n, unsigned cn);
int  orc_coeff_bits(const int16_t *coeff, int size, int type);
int  orc_quantize(const int16_t *coeff, int16_t *coeffq, int qp, int size, int coeff_block_type, const uint16_t *wmatrix);
void orc_dequantize(const int16_t *coeff, int16_t *rcoeff, int qp, int size, const uint16_t *wmatrix);
int  orc_calc_cbp(const int16_t *block, int size, int threshold);       /* SIMD semantics, enc/enc_kernels.c:828 */
int  orc_calc_cbp_c(const int16_t *block, int size, int threshold);     /* C semantics, enc/encode_block.c:2182 */
int  orc_check_nz_area(const int16_t *coeff, int size);
int  orc_quote_mv_bits(int mv_diff_y, int mv_diff_x);
void orc_clip_mv(orc_mv_t *mv, int ypos, int xpos, int fwidth, int fheight, int bwidth, int bheight, int sign);
int  orc_clpf_sample(int X, int A, int B, int C, int D, int E, int F, int G, int H, int s, unsigned dmp);
void orc_cdef_filter_block(uint8_t *dst8, uint16_t *dst16, int dstride, const uint16_t *in, int sstride,
                           int pri_strength, int sec_strength, int dir, int pri_damping, int sec_damping,
                           int bsize, int coeff_shift);
int  orc_adjust_strength(int strength, int32_t var);
int  orc_cdef_select(const uint64_t *mse0, const uint64_t *mse1, int sb_count, int speed, int cdef_bits, double lambda, int *strengths, int *uv_strengths, int *selected);

#define ORC_DECL(S, SFX) \
unsigned orc_sad_##SFX(const S *a, const S *b, int astride, int bstride, int width, int height); \
unsigned orc_widesad_##SFX(const S *a, const S *b, int astride, int bstride, int width, int height, int *x); \
uint64_t orc_ssd_##SFX(const S *a, const S *b, int astride, int bstride, int width, int height); \
unsigned orc_sad_fasthalf_##SFX(const S *a, const S *b, int astride, int bstride, int width, int height, int *x, int *y); \
unsigned orc_sad_fastquarter_##SFX(const S *o, const S *r, int os, int rs, int width, int height, int *x, int *y); \
void orc_block_avg_##SFX(S *p, const S *r0, const S *r1, int sp, int s0, int s1, int width, int height); \
void orc_interp_luma_##SFX(int width, int height, int xoff, int yoff, S *qp, int qstride, const S *ip, int istride, int bipred, int bitdepth); \
void orc_interp_chroma_##SFX(int width, int height, int xoff, int yoff, S *qp, int qstride, const S *ip, int istride, int bitdepth); \
void orc_get_inter_prediction_luma_##SFX(S *pblock, const S *ref, int width, int height, int stride, int pstride, const orc_mv_t *mv, int sign, int bipred, int pic_width, int pic_height, int xpos, int ypos, int bitdepth); \
void orc_get_inter_prediction_chroma_##SFX(S *pblock, const S *ref, int width, int height, int stride, int pstride, const orc_mv_t *mv, int sign, int pic_width2, int pic_height2, int xpos, int ypos, int bitdepth); \
void orc_residual_##SFX(int16_t *block, const S *pblock, const S *orig, int size, int pred_stride, int orig_stride); \
void orc_reconstruct_##SFX(const int16_t *block, const S *pblock, S *rec, int size, int pstride, int stride, int bitdepth); \
void orc_make_top_and_left_##SFX(S *left, S *top, S *top_left, const S *rec_frame, int fstride, const S *rblock, int rbstride, int i, int j, int ypos, int xpos, int size, int upright_available, int downleft_available, int tb_split, int bitdepth); \
void orc_intra_pred_##SFX(const S *left, const S *top, S top_left, int ypos, int xpos, int size, S *pblock, int pstride, int intra_mode, int bitdepth); \
void orc_cfl_##SFX(const S *y, S *u, S *v, const S *ry, int n, int cstride, int stride, int sub, int bitdepth); \
void orc_deblock_y_##SFX(S *rec, int stride, const orc_blkinfo_t *bi, int width, int height, int qp, int bitdepth); \
void orc_deblock_uv_##SFX(S *recU, S *recV, int stride, const orc_blkinfo_t *bi, int width, int height, int sub, int qp, int bitdepth); \
void orc_clpf_block_##SFX(const S *src, S *dst, int sstride, int dstride, int x0, int y0, int sizex, int sizey, int bt, unsigned strength, unsigned damping); \
void orc_clpf_plane_##SFX(const S *src, S *dst, int stride, int width, int height, const orc_blkinfo_t *bi, int bi_stride, int sub, const uint8_t *fb_on, int fb_size_log2, unsigned strength, int bitdepth, int plane, int qp); \
void orc_detect_clpf_##SFX(const S *rec, const S *org, int x0, int y0, int width, int height, int ostride, int rstride, int *sum0, int *sum1, unsigned strength, unsigned shift, unsigned size, unsigned dmp); \
void orc_detect_multi_clpf_##SFX(const S *rec, const S *org, int x0, int y0, int width, int height, int ostride, int rstride, int *sum, unsigned shift, unsigned size, unsigned dmp); \
int  orc_cdef_find_dir_##SFX(const S *img, int stride, int32_t *var, int coeff_shift); \
void orc_cdef_prepare_input_##SFX(int sizex, int sizey, int xpos, int ypos, int bt, int padding, uint16_t *src16, int stride16, const S *src, int sstride); \
void orc_cdef_plane_##SFX(const S *src, S *dst, int stride, int width, int height, const orc_blkinfo_t *bi, int bi_stride, int sub, int plane, const int8_t *fb_pri, const int8_t *fb_sec, int pri_damping, int sec_damping, int *dirs, int *vars, int bitdepth); \
void orc_pad_plane_##SFX(S *p, int stride, int w, int h, int pad_hor, int pad_ver); \
void orc_scale_down2x2_##SFX(const S *in, int si, S *out, int so, int wo, int ho); \
void orc_interpolate_frames_##SFX(S *outY, S *outU, S *outV, int so_y, int so_c, const S *r0Y, const S *r0U, const S *r0V, const S *r1Y, const S *r1U, const S *r1V, int sy, int sc, int width, int height, int pad, int ratio, int pos, int max_levels); \
uint64_t orc_dist_8x8_##SFX(const S *dst, int dstride, const S *src, int sstride, int coeff_shift); \
void orc_cdef_search_mse_##SFX(const S *recY, const S *recU, const S *recV, const S *orgY, const S *orgU, const S *orgV, int sy, int sc, int width, int height, const orc_blkinfo_t *bi, int speed, int pri_damping, int bitdepth, uint64_t *mse, int *dirs, int *vars, uint8_t *allskip_out); \
int  orc_motion_estimate_bi_##SFX(const S *orig, const S *ref0, const S *ref1, int size, int stride_r, int width, int height, orc_mv_t *mv, const orc_mv_t *mvc, const orc_mv_t *mvp, double lambda, int bitdepth, int sign, int fwidth, int fheight, int xpos, int ypos, const orc_mv_t *mvcand, int mvcand_num, int enable_bipred); \
int  orc_motion_estimate_sync_##SFX(const S *orig, const S *ref, int size, int stride_r, int width, int height, orc_mv_t *mv, const orc_mv_t *mvc, const orc_mv_t *mvp, double lambda, int bitdepth, int sign, int fwidth, int fheight, int xpos, int ypos, orc_mv_t *mvcand, int enable_bipred); \
void orc_block_combine_##SFX(S *dst, int ds, const S *a, int as, const S *b, int bs, int w, int h, int op, int bitdepth); \
int  orc_motion_estimate_##SFX(const S *orig, const S *ref, int size, int stride_r, int width, int height, orc_mv_t *mv, const orc_mv_t *mvc, const orc_mv_t *mvp, double lambda, int encoder_speed, int bitdepth, int sign, int fwidth, int fheight, int xpos, int ypos, const orc_mv_t *mvcand, int mvcand_num, int enable_bipred);

ORC_DECL(uint8_t, lbd)
ORC_DECL(uint16_t, hbd)

#ifdef __cplusplus
}
#endif
#endif
