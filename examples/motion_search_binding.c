/* examples/motion_search_binding.c — the reference-side binding of INTEGRATION.md §2.2 as compilable C.
 *
 * search_inter_prediction_params() (enc/encode_block.c:1968-1990) calls motion_estimate() once per (reference, prediction block); a
 * batching host queues those calls and runs them in one launch.  Each queued call carries exactly the arguments of
 * motion_estimate(orig, ref, size, stride_r, width, height, mv, mvc, mvp, lambda, enc_params, sign, fwidth, fheight, xpos, ypos,
 * mvcand, mvcand_num, enable_bipred) (:517); the current and reference frames are the HBM-resident tb_frame_t the filters left there.
 * Results come back in queue order, bit-identical to the per-call path (same visiting order and tie rules).
 *
 * Syntax-checked by tests/test_abi.py against the reference's headers; not part of the library.
 */
#include <string.h>

#include "global.h"
#include "types.h"
#include "thor_b200.h"

typedef struct {
  tb_me_item_t *items_host, *items_dev;   /* pinned / device, capacity `cap` */
  int16_t *cand_host, *cand_dev;          /* integer-pel (x, y) pairs, capacity `cand_cap` pairs */
  tb_me_result_t *res_host, *res_dev;
  int n, cap, ncand, cand_cap;
} tb_me_queue;

int tb_me_queue_create(tb_me_queue *q, int cap, int cand_cap) {
  memset(q, 0, sizeof *q);
  q->cap = cap; q->cand_cap = cand_cap;
  q->items_host = (tb_me_item_t *)tb_malloc_host((size_t)cap * sizeof(tb_me_item_t));
  q->items_dev = (tb_me_item_t *)tb_malloc((size_t)cap * sizeof(tb_me_item_t));
  q->cand_host = (int16_t *)tb_malloc_host((size_t)cand_cap * 2 * sizeof(int16_t));
  q->cand_dev = (int16_t *)tb_malloc((size_t)cand_cap * 2 * sizeof(int16_t));
  q->res_host = (tb_me_result_t *)tb_malloc_host((size_t)cap * sizeof(tb_me_result_t));
  q->res_dev = (tb_me_result_t *)tb_malloc((size_t)cap * sizeof(tb_me_result_t));
  return q->items_host && q->items_dev && q->cand_host && q->cand_dev && q->res_host && q->res_dev ? TB_OK : TB_ERR_CUDA;
}

/* register the candidate list of one (block, reference) — frame_info->mvcand[ref_idx] (:564) — and return its offset */
int tb_me_queue_candidates(tb_me_queue *q, const mv_t *mvcand, int num) {
  if (q->ncand + num > q->cand_cap) return -1;
  const int ofs = q->ncand;
  for (int i = 0; i < num; i++) { q->cand_host[2 * (ofs + i)] = mvcand[i].x; q->cand_host[2 * (ofs + i) + 1] = mvcand[i].y; }
  q->ncand += num;
  return ofs;
}

/* one motion_estimate() call: prediction block (width x height) at offset (ox, oy) inside the size x size coding block at (xpos, ypos) */
int tb_me_queue_add(tb_me_queue *q, const tb_frame_t *cur, const tb_frame_t *ref, int xpos, int ypos, int size, int ox, int oy, int width, int height, mv_t mvc, mv_t mvp,
                    double lambda, int sign, int cand_ofs, int cand_num) {
  if (q->n >= q->cap) return -1;
  int cs, rs;
  const SAMPLE *cy = (const SAMPLE *)tb_frame_plane(cur, 0, &cs), *ry = (const SAMPLE *)tb_frame_plane(ref, 0, &rs);
  tb_me_item_t *it = &q->items_host[q->n];
  it->orig = cy + (size_t)(ypos + oy) * cs + xpos + ox;
  it->ref = ry + (size_t)(ypos + oy) * rs + xpos + ox;
  it->ostride = cs; it->rstride = rs;
  it->xpos = (int16_t)xpos; it->ypos = (int16_t)ypos; it->size = (uint8_t)size; it->width = (uint8_t)width; it->height = (uint8_t)height; it->sign = (uint8_t)sign;
  it->mvc_x = mvc.x; it->mvc_y = mvc.y; it->mvp_x = mvp.x; it->mvp_y = mvp.y;
  it->cand_ofs = cand_ofs; it->ncand = cand_num; it->lambda = lambda;
  return q->n++;
}

/* run everything queued: results[i] = (mv, cost) of call i; encoder_speed / enable_bipred as in enc_params */
int tb_me_queue_run(tb_me_queue *q, int bitdepth, int encoder_speed, int enable_bipred, int fwidth, int fheight) {
  int rc = tb_memcpy_h2d(q->items_dev, q->items_host, (size_t)q->n * sizeof(tb_me_item_t));
  if (rc == TB_OK) rc = tb_memcpy_h2d(q->cand_dev, q->cand_host, (size_t)q->ncand * 2 * sizeof(int16_t));
  if (rc == TB_OK) rc = tb_motion_estimate_batch(q->items_dev, q->n, q->cand_dev, (int)sizeof(SAMPLE), bitdepth, encoder_speed, enable_bipred, fwidth, fheight, q->res_dev);
  if (rc == TB_OK) rc = tb_memcpy_d2h(q->res_host, q->res_dev, (size_t)q->n * sizeof(tb_me_result_t)); /* synchronises */
  q->n = 0; q->ncand = 0;
  return rc;
}
