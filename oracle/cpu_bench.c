/* cpu_bench.c — TEST/BENCH INFRASTRUCTURE ONLY (never linked into libthor_b200.so).
 *
 * Replays the SAME batched work-item lists that bench.py gives the CUDA library on the host's CPU cores, so that
 * the GPU number has the reference's CPU path timed beside it on the same box:
 *   -DCPU_BENCH_REF  -> oracle/_ref/libcpubench_ref.so : every item goes through the UNMODIFIED reference
 *                       (SIMD path, use_simd = 1) linked from oracle/_ref/libthorref*.so          ("kind": "reference")
 *   default          -> oracle/libcpubench_port.so     : the plain-C oracle restatement             ("kind": "port")
 * Work items use the structs of include/thor_b200.h with HOST pointers.  The reference is single-threaded; here
 * independent items are spread over `nthreads` pthreads by a DYNAMIC queue: an atomic cursor hands out small chunks,
 * walking the list from its END (bench.py sorts the lists by block size ascending, so the expensive items go first:
 * longest-processing-time order).  Round 1 used a static contiguous partition, which gave the last threads all the
 * 128x128 items and inflated the GPU/CPU ratio 7-10x.  cpu_bench_imbalance() returns
 * max(thread busy time)/mean(thread busy time) of the last run; bench.py asserts it is < 1.3.  Returns wall-clock seconds.
 */
#define _GNU_SOURCE
#include <malloc.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../include/thor_b200.h"
#include "thor_oracle.h"

#ifdef CPU_BENCH_REF
typedef struct { int16_t x, y; } mv_t;
/* reference symbols (oracle/_ref/libthorref.so and the ref_exports trampolines) */
#define DECL_REF(S, SFX)                                                                                                                          \
  int ref_motion_estimate_##SFX(S *orig, S *ref, int size, int stride_r, int width, int height, mv_t *mv, mv_t *mvc, mv_t *mvp, double lambda,     \
                                int encoder_speed, int bitdepth, int sign, int fwidth, int fheight, int xpos, int ypos, mv_t *mvcand, int mvcand_num, \
                                int enable_bipred);                                                                                               \
  void ref_get_residual_##SFX(int16_t *block, S *pblock, S *orig, int size, int ps, int os);                                                       \
  int ref_quantize_##SFX(int16_t *coeff, int16_t *coeffq, int qp, int size, int type, uint16_t *wm);                                               \
  uint64_t ref_ssd_calc_##SFX(S *a, S *b, int as, int bs, int w, int h);                                                                           \
  void dequantize_##SFX(int16_t *coeff, int16_t *rcoeff, int qp, int size, uint16_t *wm);                                                          \
  void reconstruct_block_##SFX(int16_t *block, S *pblock, S *rec, int size, int pstride, int stride, int bitdepth);                                \
  void get_inter_prediction_luma_##SFX(S *pblock, S *ref, int width, int height, int stride, int pstride, mv_t *mv, int sign, int bipred,          \
                                       int pic_width, int pic_height, int xpos, int ypos, int bitdepth);                                           \
  void get_inter_prediction_chroma_simd_##SFX(int width, int height, int xoff, int yoff, S *qp, int qstride, const S *ip, int istride, int bd);    \
  void make_top_and_left_##SFX(S *left, S *top, S *top_left, S *rec_frame, int fstride, S *rblock, int rbstride, int i, int j, int ypos, int xpos, \
                               int size, int upright, int downleft, int tb_split, int bitdepth);                                                   \
  void get_intra_prediction_##SFX(S *left, S *top, S top_left, int ypos, int xpos, int size, S *pblock, int pstride, int mode, int bitdepth);
DECL_REF(uint8_t, lbd)
DECL_REF(uint16_t, hbd)
void transform(const int16_t *block, int16_t *coeff, int size, int fast, int bitdepth);
void inverse_transform(const int16_t *coeff, int16_t *block, int size, int bitdepth);
extern int use_simd;
#endif

static double now(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

typedef struct {
  int kind, tid, nthreads, n, hbd, bitdepth, speed, bip, fw, fh, chunk;
  long *cursor; /* shared: number of items already handed out (from the end of the list) */
  double busy;  /* this thread's wall time inside the run_* loops */
  double cpu;   /* this thread's CPU time (CLOCK_THREAD_CPUTIME_ID): busy >> cpu means the box did not grant the thread a core */
  const void *items;
  const int16_t *cands;
  void *out;
} job_t;

#define BUF_ALIGN __attribute__((aligned(64)))

static void run_me(const job_t *j, int lo, int hi) {
  const tb_me_item_t *it = (const tb_me_item_t *)j->items;
  tb_me_result_t *out = (tb_me_result_t *)j->out;
  static __thread uint16_t blk[128 * 128] BUF_ALIGN;
  for (int i = lo; i < hi; i++) {
    const tb_me_item_t *q = &it[i];
    const int esz = j->hbd ? 2 : 1, size = q->size;
    /* the reference searches a compact copy of the original block (stride = size, encode_block.c:2457) */
    for (int r = 0; r < q->height; r++) memcpy((char *)blk + (size_t)r * size * esz, (const char *)q->orig + (size_t)r * q->ostride * esz, (size_t)q->width * esz);
    int16_t mv[2], mvc[2] = {q->mvc_x, q->mvc_y}, mvp[2] = {q->mvp_x, q->mvp_y};
    int cost;
#ifdef CPU_BENCH_REF
    if (j->hbd)
      cost = ref_motion_estimate_hbd(blk, (uint16_t *)q->ref, size, q->rstride, q->width, q->height, (mv_t *)mv, (mv_t *)mvc, (mv_t *)mvp, q->lambda, j->speed,
                                     j->bitdepth, q->sign, j->fw, j->fh, q->xpos, q->ypos, (mv_t *)(j->cands + 2 * (size_t)q->cand_ofs), q->ncand, j->bip);
    else
      cost = ref_motion_estimate_lbd((uint8_t *)blk, (uint8_t *)q->ref, size, q->rstride, q->width, q->height, (mv_t *)mv, (mv_t *)mvc, (mv_t *)mvp, q->lambda,
                                     j->speed, j->bitdepth, q->sign, j->fw, j->fh, q->xpos, q->ypos, (mv_t *)(j->cands + 2 * (size_t)q->cand_ofs), q->ncand, j->bip);
#else
    if (j->hbd)
      cost = orc_motion_estimate_hbd(blk, (const uint16_t *)q->ref, size, q->rstride, q->width, q->height, (orc_mv_t *)mv, (orc_mv_t *)mvc, (orc_mv_t *)mvp, q->lambda,
                                     j->speed, j->bitdepth, q->sign, j->fw, j->fh, q->xpos, q->ypos, (const orc_mv_t *)(j->cands + 2 * (size_t)q->cand_ofs), q->ncand,
                                     j->bip);
    else
      cost = orc_motion_estimate_lbd((uint8_t *)blk, (const uint8_t *)q->ref, size, q->rstride, q->width, q->height, (orc_mv_t *)mv, (orc_mv_t *)mvc, (orc_mv_t *)mvp,
                                     q->lambda, j->speed, j->bitdepth, q->sign, j->fw, j->fh, q->xpos, q->ypos, (const orc_mv_t *)(j->cands + 2 * (size_t)q->cand_ofs),
                                     q->ncand, j->bip);
#endif
    out[i].mvx = mv[0];
    out[i].mvy = mv[1];
    out[i].cost = (uint32_t)cost;
  }
}

static void run_txfm(const job_t *j, int lo, int hi) {
  const tb_txfm_item_t *it = (const tb_txfm_item_t *)j->items;
  tb_txfm_result_t *out = (tb_txfm_result_t *)j->out;
  static __thread int16_t block[128 * 128] BUF_ALIGN, coeff[128 * 128] BUF_ALIGN, rcoeff[128 * 128] BUF_ALIGN, rblock[128 * 128] BUF_ALIGN, cq[256] BUF_ALIGN;
  static __thread uint16_t recb[128 * 128] BUF_ALIGN;
  for (int i = lo; i < hi; i++) {
    const tb_txfm_item_t *q = &it[i];
    const int size = q->size, qs = size < 16 ? size : 16, esz = j->hbd ? 2 : 1;
    int cbp;
    uint64_t ssd;
    void *rec = q->rec ? q->rec : (void *)recb;
    const int rstride = q->rec ? q->rstride : size;
#ifdef CPU_BENCH_REF
    if (j->hbd) ref_get_residual_hbd(block, (uint16_t *)q->pred, (uint16_t *)q->orig, size, q->pstride, q->ostride);
    else ref_get_residual_lbd(block, (uint8_t *)q->pred, (uint8_t *)q->orig, size, q->pstride, q->ostride);
    transform(block, coeff, size, q->fast, j->bitdepth);
    cbp = j->hbd ? ref_quantize_hbd(coeff, cq, q->qp, size, q->coeff_type, NULL) : ref_quantize_lbd(coeff, cq, q->qp, size, q->coeff_type, NULL);
    if (cbp) {
      if (j->hbd) dequantize_hbd(cq, rcoeff, q->qp, size, NULL); else dequantize_lbd(cq, rcoeff, q->qp, size, NULL);
      inverse_transform(rcoeff, rblock, size, j->bitdepth);
      if (j->hbd) reconstruct_block_hbd(rblock, (uint16_t *)q->pred, (uint16_t *)rec, size, q->pstride, rstride, j->bitdepth);
      else reconstruct_block_lbd(rblock, (uint8_t *)q->pred, (uint8_t *)rec, size, q->pstride, rstride, j->bitdepth);
    } else
      for (int r = 0; r < size; r++) memcpy((char *)rec + (size_t)r * rstride * esz, (const char *)q->pred + (size_t)r * q->pstride * esz, (size_t)size * esz);
    ssd = j->hbd ? ref_ssd_calc_hbd((uint16_t *)q->orig, (uint16_t *)rec, q->ostride, rstride, size, size)
                 : ref_ssd_calc_lbd((uint8_t *)q->orig, (uint8_t *)rec, q->ostride, rstride, size, size);
#else
    if (j->hbd) orc_residual_hbd(block, (const uint16_t *)q->pred, (const uint16_t *)q->orig, size, q->pstride, q->ostride);
    else orc_residual_lbd(block, (const uint8_t *)q->pred, (const uint8_t *)q->orig, size, q->pstride, q->ostride);
    orc_transform(block, coeff, size, q->fast, j->bitdepth);
    cbp = orc_quantize(coeff, cq, q->qp, size, q->coeff_type, NULL);
    if (cbp) {
      orc_dequantize(cq, rcoeff, q->qp, size, NULL);
      orc_inverse_transform(rcoeff, rblock, size, j->bitdepth);
      if (j->hbd) orc_reconstruct_hbd(rblock, (const uint16_t *)q->pred, (uint16_t *)rec, size, q->pstride, rstride, j->bitdepth);
      else orc_reconstruct_lbd(rblock, (const uint8_t *)q->pred, (uint8_t *)rec, size, q->pstride, rstride, j->bitdepth);
    } else
      for (int r = 0; r < size; r++) memcpy((char *)rec + (size_t)r * rstride * esz, (const char *)q->pred + (size_t)r * q->pstride * esz, (size_t)size * esz);
    ssd = j->hbd ? orc_ssd_hbd((const uint16_t *)q->orig, (const uint16_t *)rec, q->ostride, rstride, size, size)
                 : orc_ssd_lbd((const uint8_t *)q->orig, (const uint8_t *)rec, q->ostride, rstride, size, size);
#endif
    if (q->coeffq) memcpy(q->coeffq, cq, (size_t)qs * qs * 2);
    out[i].ssd = ssd;
    out[i].cbp = cbp;
    out[i].bits = 0;
  }
}

static void run_intra(const job_t *j, int lo, int hi) {
  const tb_intra_item_t *it = (const tb_intra_item_t *)j->items;
  static __thread uint16_t left[2 * 128 + 16] BUF_ALIGN, top[2 * 128 + 16] BUF_ALIGN;
  for (int i = lo; i < hi; i++) {
    const tb_intra_item_t *q = &it[i];
#ifdef CPU_BENCH_REF
    if (j->hbd) {
      uint16_t tl;
      make_top_and_left_hbd(left + 1, top + 1, &tl, (uint16_t *)q->rec, q->rstride, NULL, 0, 0, 0, q->ypos, q->xpos, q->size, q->upright, q->downleft, 0, j->bitdepth);
      get_intra_prediction_hbd(left + 1, top + 1, tl, q->ypos, q->xpos, q->size, (uint16_t *)q->dst, q->size, q->mode, j->bitdepth);
    } else {
      uint8_t tl;
      make_top_and_left_lbd((uint8_t *)left + 1, (uint8_t *)top + 1, &tl, (uint8_t *)q->rec, q->rstride, NULL, 0, 0, 0, q->ypos, q->xpos, q->size, q->upright, q->downleft, 0,
                            j->bitdepth);
      get_intra_prediction_lbd((uint8_t *)left + 1, (uint8_t *)top + 1, tl, q->ypos, q->xpos, q->size, (uint8_t *)q->dst, q->size, q->mode, j->bitdepth);
    }
#else
    if (j->hbd) {
      uint16_t tl;
      orc_make_top_and_left_hbd(left, top, &tl, (const uint16_t *)q->rec, q->rstride, NULL, 0, 0, 0, q->ypos, q->xpos, q->size, q->upright, q->downleft, 0, j->bitdepth);
      orc_intra_pred_hbd(left, top, tl, q->ypos, q->xpos, q->size, (uint16_t *)q->dst, q->size, q->mode, j->bitdepth);
    } else {
      uint8_t tl;
      orc_make_top_and_left_lbd((uint8_t *)left, (uint8_t *)top, &tl, (const uint8_t *)q->rec, q->rstride, NULL, 0, 0, 0, q->ypos, q->xpos, q->size, q->upright, q->downleft,
                                0, j->bitdepth);
      orc_intra_pred_lbd((uint8_t *)left, (uint8_t *)top, tl, q->ypos, q->xpos, q->size, (uint8_t *)q->dst, q->size, q->mode, j->bitdepth);
    }
#endif
  }
}

static void run_interp(const job_t *j, int lo, int hi) {
  const tb_interp_item_t *it = (const tb_interp_item_t *)j->items;
  for (int i = lo; i < hi; i++) {
    const tb_interp_item_t *q = &it[i];
    int16_t mv[2] = {q->mvx, q->mvy};
#ifdef CPU_BENCH_REF
    if (!q->chroma) {
      if (j->hbd) get_inter_prediction_luma_hbd((uint16_t *)q->dst, (uint16_t *)q->ref, q->width, q->height, q->rstride, q->dstride, (mv_t *)mv, q->sign, j->bip, q->pic_w,
                                                q->pic_h, q->xpos, q->ypos, j->bitdepth);
      else get_inter_prediction_luma_lbd((uint8_t *)q->dst, (uint8_t *)q->ref, q->width, q->height, q->rstride, q->dstride, (mv_t *)mv, q->sign, j->bip, q->pic_w, q->pic_h,
                                         q->xpos, q->ypos, j->bitdepth);
    } else {
      /* get_inter_prediction_chroma is file-static in the reference; its body = clamp + integer copy or the SIMD kernel
         (common/inter_prediction.c:65-93) */
      int x = q->sign ? -q->mvx : q->mvx, y = q->sign ? -q->mvy : q->mvy, vf = y & 7, hf = x & 7, vi = y >> 3, hi2 = x >> 3;
      if (vi > q->pic_h - q->ypos) vi = q->pic_h - q->ypos;
      if (vi < -q->xpos - q->height) vi = -q->xpos - q->height;
      if (hi2 > q->pic_w - q->xpos) hi2 = q->pic_w - q->xpos;
      if (hi2 < -q->xpos - q->width) hi2 = -q->xpos - q->width;
      const int esz = j->hbd ? 2 : 1;
      const char *ip = (const char *)q->ref + ((ptrdiff_t)vi * q->rstride + hi2) * esz;
      if (!vf && !hf)
        for (int r = 0; r < q->height; r++) memcpy((char *)q->dst + (size_t)r * q->dstride * esz, ip + (size_t)r * q->rstride * esz, (size_t)q->width * esz);
      else if (q->width > 2) {
        if (j->hbd) get_inter_prediction_chroma_simd_hbd(q->width, q->height, hf, vf, (uint16_t *)q->dst, q->dstride, (const uint16_t *)ip, q->rstride, j->bitdepth);
        else get_inter_prediction_chroma_simd_lbd(q->width, q->height, hf, vf, (uint8_t *)q->dst, q->dstride, (const uint8_t *)ip, q->rstride, j->bitdepth);
      } else {
        if (j->hbd) orc_interp_chroma_hbd(q->width, q->height, hf, vf, (uint16_t *)q->dst, q->dstride, (const uint16_t *)ip, q->rstride, j->bitdepth);
        else orc_interp_chroma_lbd(q->width, q->height, hf, vf, (uint8_t *)q->dst, q->dstride, (const uint8_t *)ip, q->rstride, j->bitdepth);
      }
    }
#else
    if (!q->chroma) {
      if (j->hbd) orc_get_inter_prediction_luma_hbd((uint16_t *)q->dst, (const uint16_t *)q->ref, q->width, q->height, q->rstride, q->dstride, (orc_mv_t *)mv, q->sign, j->bip,
                                                    q->pic_w, q->pic_h, q->xpos, q->ypos, j->bitdepth);
      else orc_get_inter_prediction_luma_lbd((uint8_t *)q->dst, (const uint8_t *)q->ref, q->width, q->height, q->rstride, q->dstride, (orc_mv_t *)mv, q->sign, j->bip, q->pic_w,
                                             q->pic_h, q->xpos, q->ypos, j->bitdepth);
    } else {
      if (j->hbd) orc_get_inter_prediction_chroma_hbd((uint16_t *)q->dst, (const uint16_t *)q->ref, q->width, q->height, q->rstride, q->dstride, (orc_mv_t *)mv, q->sign,
                                                      q->pic_w, q->pic_h, q->xpos, q->ypos, j->bitdepth);
      else orc_get_inter_prediction_chroma_lbd((uint8_t *)q->dst, (const uint8_t *)q->ref, q->width, q->height, q->rstride, q->dstride, (orc_mv_t *)mv, q->sign, q->pic_w,
                                               q->pic_h, q->xpos, q->ypos, j->bitdepth);
    }
#endif
  }
}

static double thread_cpu(void) {
  struct timespec t;
  clock_gettime(CLOCK_THREAD_CPUTIME_ID, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

static void *worker(void *arg) {
  job_t *j = (job_t *)arg;
  const double c0 = thread_cpu();
  for (;;) {
    long taken = __atomic_fetch_add(j->cursor, (long)j->chunk, __ATOMIC_RELAXED);
    if (taken >= j->n) break;
    int hi = j->n - (int)taken, lo = hi - j->chunk < 0 ? 0 : hi - j->chunk;
    double t0 = now();
    switch (j->kind) {
      case 0: run_me(j, lo, hi); break;
      case 1: run_txfm(j, lo, hi); break;
      case 2: run_intra(j, lo, hi); break;
      case 3: run_interp(j, lo, hi); break;
    }
    j->busy += now() - t0;
  }
  j->cpu = thread_cpu() - c0;
  return NULL;
}

static double last_imbalance = 1.0, last_busy_sum = 0.0, last_cpu_sum = 0.0;

/* kind: 0 motion search (tb_me_item_t), 1 transform chain (tb_txfm_item_t), 2 intra (tb_intra_item_t), 3 interpolation
 * (tb_interp_item_t).  Returns elapsed wall-clock seconds. */
double cpu_bench_run(int kind, const void *items, int n, const int16_t *cands, void *out, int hbd, int bitdepth, int speed, int bip, int fw, int fh, int nthreads) {
#ifdef CPU_BENCH_REF
  use_simd = 1;
#endif
  /* The reference allocates and frees its 32 KB scratch blocks inside every kernel call (thor_alloc).  With glibc's default
     trim threshold every free of a thread arena's top chunk returns pages to the kernel and the next call faults them back in;
     across 128 threads that serialises on the process's mmap lock (measured: 16x slower transform chains on the 128-thread
     GPU host than on 8 threads).  That is an artefact of threading a single-threaded program, not of the reference's kernels:
     keep freed memory in the arenas. */
  mallopt(M_TRIM_THRESHOLD, 1 << 30);
  mallopt(M_MMAP_THRESHOLD, 1 << 30);
  mallopt(M_TOP_PAD, 64 << 20);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  static job_t jobs[256];
  long cursor = 0;
  /* chunks small enough that the last one cannot unbalance the run (>= 64 chunks per thread), large enough to keep
     neighbouring items (which share samples) on one core */
  int chunk = n / (nthreads * 64);
  if (chunk < 1) chunk = 1;
  if (chunk > 256) chunk = 256;
  double t0 = now();
  for (int t = 0; t < nthreads; t++) {
    job_t j = {kind, t, nthreads, n, hbd, bitdepth, speed, bip, fw, fh, chunk, &cursor, 0.0, 0.0, items, cands, out};
    jobs[t] = j;
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  double el = now() - t0, mx = 0, sum = 0;
  last_cpu_sum = 0;
  for (int t = 0; t < nthreads; t++) {
    last_cpu_sum += jobs[t].cpu;
    sum += jobs[t].busy;
    if (jobs[t].busy > mx) mx = jobs[t].busy;
  }
  last_imbalance = sum > 0 ? mx / (sum / nthreads) : 1.0;
  last_busy_sum = sum;
  return el;
}

/* max/mean of the per-thread busy times of the last cpu_bench_run (1.0 = perfectly balanced) */
double cpu_bench_imbalance(void) { return last_imbalance; }
/* CPU seconds the threads of the last cpu_bench_run actually consumed (CLOCK_THREAD_CPUTIME_ID) */
double cpu_bench_cpu_seconds(void) { return last_cpu_sum; }
/* core-seconds of the last cpu_bench_run (sum of the per-thread busy times) */
double cpu_bench_core_seconds(void) { return last_busy_sum; }

const char *cpu_bench_kind(void) {
#ifdef CPU_BENCH_REF
  return "reference";
#else
  return "port";
#endif
}
