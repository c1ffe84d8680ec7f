/* examples/rdo_batch_server.c — INTEGRATION.md §2.0 as real C: a host that has SEVERAL frames ready at once (the B frames of one hierarchy level of one
 * encoder, or the current frames of several encoder instances / intra-period segments) decides them in ONE launch of the device RD loop.
 *
 * Each encoder instance is the reference's own encoder_info_t (enc/mainenc.h:167-213); fill_job() describes its current frame exactly as
 * thor_b200/csrc/tb_rdo_shim.c does for a single frame; the results come back per instance and are serialised by the reference's own write_block()
 * (see the shim's emit_node()).  Compile-checked against the reference's headers: tests/test_abi.py::test_binding_example_compiles_against_reference_headers. */
#include <stdlib.h>
#include <string.h>
#include "global.h"
#include "mainenc.h"
#include "thor_b200.h"

extern const double squared_lambda_QP[52]; /* enc/encode_tables.c */

typedef struct { /* what one instance needs to receive its decisions */
  tb_rdo_blk_t *blk;
  tb_rdo_leaf_t *leaves;
  int32_t *leaf_count;
  int16_t *coeffs;
} rdo_result_t;

static int result_alloc(rdo_result_t *r, int width, int height, int log2_sb_size) {
  const int sb = 1 << log2_sb_size, nsb = ((width + sb - 1) / sb) * ((height + sb - 1) / sb);
  r->blk = calloc((size_t)(height / 4) * (width / 4), sizeof *r->blk);
  r->leaves = calloc((size_t)nsb * TB_RDO_MAX_LEAVES, sizeof *r->leaves);
  r->leaf_count = calloc((size_t)nsb, sizeof *r->leaf_count);
  r->coeffs = calloc((size_t)nsb * TB_RDO_SB_COEFFS, sizeof *r->coeffs);
  return r->blk && r->leaves && r->leaf_count && r->coeffs;
}

/* enc_params / frame_info_t / yuv_frame_t -> tb_rdo_frame_t (host pointers; the library uploads) */
static void fill_job(tb_rdo_frame_t *f, encoder_info_t *e, const rdo_result_t *r, int sample_bytes) {
  const enc_params *p = e->params;
  const frame_info_t *fi = &e->frame_info;
  memset(f, 0, sizeof *f);
  f->width = e->width; f->height = e->height; f->log2_sb_size = p->log2_sb_size; f->bitdepth = p->bitdepth; f->sample_bytes = sample_bytes;
  f->frame_type = fi->frame_type; f->qp = fi->qp; f->num_ref = fi->num_ref; f->interp_ref = fi->interp_ref; f->num_intra_modes = fi->num_intra_modes;
  f->lambda = fi->lambda_coeff * squared_lambda_QP[fi->qp]; /* enc/encode_frame.c:670 with max_delta_qp == 0 */
  f->enable_bipred = p->enable_bipred; f->enable_tb_split = p->enable_tb_split; f->enable_pb_split = p->enable_pb_split; f->encoder_speed = p->encoder_speed;
  f->intra_rdo = p->intra_rdo; f->use_block_contexts = p->use_block_contexts; f->cfl_intra = p->cfl_intra; f->cfl_inter = p->cfl_inter;
  f->early_skip_thr = p->early_skip_thr;
  f->ref_stride[0] = e->ref[0]->stride_y; f->ref_stride[1] = e->ref[0]->stride_c; f->ref_pad = e->ref[0]->pad_hor_y;
  for (int k = 0; k < (int)fi->num_ref; k++) {
    const int ra = fi->ref_array[k];
    yuv_frame_t *ref = ra >= 0 ? e->ref[ra] : e->interp_frames[0];
    f->ref_sign[k] = ref->frame_num > e->rec->frame_num;     /* enc/encode_block.c:1975 */
    f->ref_sign_ge[k] = ref->frame_num >= fi->frame_num;     /* :2282 */
    f->ref[k][0] = ref->y; f->ref[k][1] = ref->u; f->ref[k][2] = ref->v;
  }
  f->orig[0] = e->orig->y; f->orig[1] = e->orig->u; f->orig[2] = e->orig->v; f->orig_stride[0] = e->orig->stride_y; f->orig_stride[1] = e->orig->stride_c;
  f->rec[0] = e->rec->y; f->rec[1] = e->rec->u; f->rec[2] = e->rec->v; f->rec_stride[0] = e->rec->stride_y; f->rec_stride[1] = e->rec->stride_c;
  f->blk = r->blk; f->leaves = r->leaves; f->leaf_count = r->leaf_count; f->coeffs = r->coeffs;
}

/* n encoder instances whose current frames do not depend on each other: one launch decides them all.  Returns TB_OK, or the library's error
 * (TB_ERR_ARG: a configuration the device loop does not cover -> let that instance run the reference's process_block loop; TB_ERR_CUDA: no device) */
int decide_frames_together(encoder_info_t **enc, rdo_result_t *res, int n, int sample_bytes) {
  tb_rdo_frame_t *jobs = malloc((size_t)n * sizeof *jobs);
  if (!jobs) return TB_ERR_ARG;
  for (int i = 0; i < n; i++) {
    if (!res[i].blk && !result_alloc(&res[i], enc[i]->width, enc[i]->height, enc[i]->params->log2_sb_size)) { free(jobs); return TB_ERR_ARG; }
    fill_job(&jobs[i], enc[i], &res[i], sample_bytes);
  }
  const int rc = tb_rdo_encode_frames(jobs, n); /* upload, ONE persistent launch over every super block of every frame, download, synchronise */
  free(jobs);
  return rc;
}

/* the same, keeping the frames resident and the copies asynchronous (pinned host buffers): what bench.py times */
int decide_frames_resident(tb_rdo_batch_t *batch, const tb_rdo_frame_t *jobs, int n) {
  int rc = TB_OK;
  for (int i = 0; i < n && rc == TB_OK; i++) rc = tb_rdo_batch_upload(batch, i, &jobs[i]);
  if (rc == TB_OK) rc = tb_rdo_batch_run(batch, n);
  for (int i = 0; i < n && rc == TB_OK; i++) rc = tb_rdo_batch_download(batch, i, &jobs[i]);
  return rc == TB_OK ? tb_rdo_batch_sync(batch) : rc;
}
