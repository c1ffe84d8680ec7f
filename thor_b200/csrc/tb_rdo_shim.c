/* tb_rdo_shim.c — the reference-side binding for the device-resident RD loop.
 *
 * This file is INTEGRATION GLUE that lives on the reference's side of the C ABI: it is compiled against the reference's own
 * headers (enc/mainenc.h, enc/write_bits.h, ...; -I /root/reference/{enc,common}) and linked into Thorenc with
 *     -Wl,--wrap=process_block_lbd -Wl,--wrap=process_block_hbd
 * so that enc/encode_frame.c:700-741 (the per-super-block loop, unmodified) reaches __wrap_process_block_*() below.  Nothing of
 * the reference is copied: the bit writer (write_super_mode / write_block, enc/write_bits.c:255-600), find_block_contexts and the
 * stream are the reference's own objects.
 *
 * Flow per frame: at the first super block the whole frame goes to tb_rdo_encode_frame() (include/thor_b200.h: one CUDA launch,
 * super blocks in a wavefront); every process_block() call then only serialises its super block's decisions: reconstruction and
 * deblock_data are copied into the encoder's structures and the quad-tree is written with the reference's bit writer.
 * Configurations the device loop does not cover (see thor_b200.h) fall through to __real_process_block_*(), i.e. the reference's
 * host loop over the per-call drop-in kernels of libthor_b200.so.
 *
 * Linked WITHOUT libthor_b200.so (oracle/_ref/Thorenc_capture: the all-reference CPU encoder + this file) tb_rdo_encode_frame is an unresolved
 * weak symbol and every super block is decided by the reference's own process_block(); the shim then only observes: it times the RD loop
 * (TB_RDO_STATS=1) and, with TB_RDO_DUMP=<dir>, writes every frame's RD-loop JOB (the tb_rdo_frame_t description, source planes, padded
 * reference planes) together with the reference's RESULT (reconstruction before the in-loop filters, per-4x4 block state, process_block's
 * return value per super block) to <dir>/frame_NNN.job — the workload and the expected answers of bench.py and tests/test_gpu_rdo_batch.py.
 *
 * TB_RDO_PROGRESS=<file> (capture link): the file is mapped and updated after every super block with { uint64 pixels decided, double seconds
 * inside process_block, uint64 frames finished } — bench.py's CPU arm reads it to measure the throughput of long-running reference encoders in
 * fixed wall-clock slices without waiting for them to finish.
 *
 * TB_RDO_VERIFY=1 (needs the per-super-block entry of oracle/librdo_hostcheck.so, TEST INFRASTRUCTURE): every super block is decided by
 * tb_rdo_encode_sb() AND by the reference's own process_block() on the same state; bits, reconstruction and deblock_data are compared
 * and the reference's result is kept.  That is how the control flow of tb_rdo.h is pinned against the reference in this container.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include "global.h"
#include "mainenc.h"
#include "encode_block.h"
#include "write_bits.h"
#include "putbits.h"
#include "common_block.h"
#include "thor_b200.h"

extern const double squared_lambda_QP[52];
int __real_process_block_lbd(encoder_info_t *encoder_info, int size, int ypos, int xpos, int qp, int sub);
int __real_process_block_hbd(encoder_info_t *encoder_info, int size, int ypos, int xpos, int qp, int sub);
void find_block_contexts_lbd(int ypos, int xpos, int height, int width, int size, deblock_data_t *deblock_data, block_context_t *block_context, int enable);
void find_block_contexts_hbd(int ypos, int xpos, int height, int width, int size, deblock_data_t *deblock_data, block_context_t *block_context, int enable);
/* optional (weak): only oracle/librdo_hostcheck.so has it */
int tb_rdo_encode_sb(const tb_rdo_frame_t *f, int sbx, int sby) __attribute__((weak));
int tb_rdo_encode_frame(const tb_rdo_frame_t *f) __attribute__((weak)); /* absent in the capture link (no libthor_b200.so) */
const char *tb_rdo_last_error(void) __attribute__((weak));

static struct {
  int w, h, nsb, valid, frame_ok, verify, inited;
  tb_rdo_blk_t *blk;
  tb_rdo_leaf_t *leaves;
  int32_t *leaf_count;
  int16_t *coeffs;
  long sb_total, sb_bad, frames_dev, frames_host;
  double t_rdo, t_emit, t_ref;
  int32_t *sb_cost; /* capture: process_block's return value per super block */
  FILE *dump;
  volatile struct { uint64_t pixels; double t_rd; uint64_t frames; } *progress;
} G;

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void shim_report(void) {
  if (G.verify)
    fprintf(stderr, "[tb_rdo_shim] verify: %ld super blocks compared with the reference's process_block, %ld differ\n", G.sb_total, G.sb_bad);
  if (getenv("TB_RDO_STATS"))
    fprintf(stderr, "[tb_rdo_shim] frames decided by tb_rdo_encode_frame: %ld, by the reference's host loop: %ld; seconds in tb_rdo_encode_frame %.3f, in serialisation %.3f, "
            "in the reference's process_block %.3f\n", G.frames_dev, G.frames_host, G.t_rdo, G.t_emit, G.t_ref);
}

static int supported(const encoder_info_t *e, int qp, int sub, int hbd) {
  const enc_params *p = e->params;
  /* 16-bit samples: the reference's SIMD chroma interpolation treats rectangular frame-edge blocks whose chroma width is 4 mod 8 differently
     from its C code (seen with luma widths 200 and 216); such widths stay with the reference's own loop */
  if (hbd && (e->width & 15)) return 0;
  return !p->sync && !p->qmtx && p->interp_ref != 2 && !p->max_delta_qp && !p->bitrate && p->subsample == 420 && sub == 1 && p->log2_sb_size <= 7 &&
         e->frame_info.num_ref <= TB_RDO_MAX_REF && qp == (int)e->frame_info.qp && !getenv("TB_RDO_OFF");
}

static void ensure_buffers(int w, int h, int log2sb) {
  const int sb = 1 << log2sb, nsb = ((w + sb - 1) / sb) * ((h + sb - 1) / sb);
  if (G.w == w && G.h == h && G.nsb == nsb) return;
  free(G.blk); free(G.leaves); free(G.leaf_count); free(G.coeffs); free(G.sb_cost);
  G.sb_cost = calloc((size_t)nsb, sizeof(int32_t));
  G.w = w; G.h = h; G.nsb = nsb;
  G.blk = calloc((size_t)(h / 4) * (w / 4), sizeof(tb_rdo_blk_t));
  G.leaves = calloc((size_t)nsb * TB_RDO_MAX_LEAVES, sizeof(tb_rdo_leaf_t));
  G.leaf_count = calloc((size_t)nsb, sizeof(int32_t));
  G.coeffs = calloc((size_t)nsb * TB_RDO_SB_COEFFS, sizeof(int16_t));
  if (!G.inited) {
    G.inited = 1;
    G.verify = getenv("TB_RDO_VERIFY") != NULL;
    atexit(shim_report);
    if (getenv("TB_RDO_PROGRESS")) {
      const int fd = open(getenv("TB_RDO_PROGRESS"), O_RDWR | O_CREAT, 0644);
      if (fd < 0 || ftruncate(fd, 64) != 0) { fprintf(stderr, "[tb_rdo_shim] cannot open TB_RDO_PROGRESS file\n"); exit(2); }
      G.progress = mmap(NULL, 64, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      if (G.progress == MAP_FAILED) { fprintf(stderr, "[tb_rdo_shim] cannot map TB_RDO_PROGRESS file\n"); exit(2); }
      close(fd);
    }
  }
}

static void fill_desc(tb_rdo_frame_t *f, encoder_info_t *e, int esz) {
  const enc_params *p = e->params;
  const frame_info_t *fi = &e->frame_info;
  memset(f, 0, sizeof(*f));
  f->width = e->width; f->height = e->height; f->log2_sb_size = p->log2_sb_size; f->bitdepth = p->bitdepth; f->sample_bytes = esz;
  f->frame_type = fi->frame_type; f->qp = fi->qp; f->num_ref = fi->num_ref; f->interp_ref = fi->interp_ref; f->num_intra_modes = fi->num_intra_modes;
  /* block_info->lambda (enc/encode_block.c:2454) = frame_info.lambda when max_delta_qp == 0 (enc/encode_frame.c:670) */
  f->lambda = fi->lambda_coeff * squared_lambda_QP[fi->qp];
  f->enable_bipred = p->enable_bipred; f->enable_tb_split = p->enable_tb_split; f->enable_pb_split = p->enable_pb_split; f->encoder_speed = p->encoder_speed;
  f->intra_rdo = p->intra_rdo; f->use_block_contexts = p->use_block_contexts; f->cfl_intra = p->cfl_intra; f->cfl_inter = p->cfl_inter;
  f->early_skip_thr = p->early_skip_thr;
  /* every padded frame of the encoder has the same geometry (enc/mainenc.c:181-186); intra frames have no reference but still need it */
  f->ref_stride[0] = e->ref[0]->stride_y; f->ref_stride[1] = e->ref[0]->stride_c; f->ref_pad = e->ref[0]->pad_hor_y;
  for (int r = 0; r < fi->num_ref; r++) {
    const int ra = fi->ref_array[r];
    yuv_frame_t *ref = ra >= 0 ? e->ref[ra] : e->interp_frames[0];
    f->ref_sign[r] = ref->frame_num > e->rec->frame_num;
    f->ref_sign_ge[r] = ref->frame_num >= fi->frame_num;
    f->ref[r][0] = ref->y; f->ref[r][1] = ref->u; f->ref[r][2] = ref->v;
    f->ref_stride[0] = ref->stride_y; f->ref_stride[1] = ref->stride_c; f->ref_pad = ref->pad_hor_y;
  }
  f->orig[0] = e->orig->y; f->orig[1] = e->orig->u; f->orig[2] = e->orig->v; f->orig_stride[0] = e->orig->stride_y; f->orig_stride[1] = e->orig->stride_c;
  f->rec[0] = e->rec->y; f->rec[1] = e->rec->u; f->rec[2] = e->rec->v; f->rec_stride[0] = e->rec->stride_y; f->rec_stride[1] = e->rec->stride_c;
  f->blk = G.blk; f->leaves = G.leaves; f->leaf_count = G.leaf_count; f->coeffs = G.coeffs;
}

/* deblock_data_t <-> tb_rdo_blk_t for a region of the 4x4 grid */
static void blk_to_deblock(encoder_info_t *e, int x0, int y0, int x1, int y1) {
  const int bs = e->width / MIN_PB_SIZE;
  for (int by = y0 / MIN_PB_SIZE; by < y1 / MIN_PB_SIZE; by++)
    for (int bx = x0 / MIN_PB_SIZE; bx < x1 / MIN_PB_SIZE; bx++) {
      const tb_rdo_blk_t *b = &G.blk[by * bs + bx];
      deblock_data_t *d = &e->deblock_data[by * bs + bx];
      d->mode = (block_mode_t)b->mode; d->cbp.y = b->cbp_y; d->cbp.u = b->cbp_u; d->cbp.v = b->cbp_v; d->size = b->size; d->tb_split = b->tb_split;
      d->pb_part = (part_t)b->pb_part;
      d->inter_pred.mv0.x = b->mv0.x; d->inter_pred.mv0.y = b->mv0.y; d->inter_pred.mv1.x = b->mv1.x; d->inter_pred.mv1.y = b->mv1.y;
      d->inter_pred.ref_idx0 = b->ref_idx0; d->inter_pred.ref_idx1 = b->ref_idx1; d->inter_pred.bipred_flag = (uint32_t)(int32_t)b->bipred_flag;
    }
}
static void deblock_to_blk(const encoder_info_t *e) {
  const int n = (e->height / MIN_PB_SIZE) * (e->width / MIN_PB_SIZE);
  for (int i = 0; i < n; i++) {
    const deblock_data_t *d = &e->deblock_data[i];
    tb_rdo_blk_t *b = &G.blk[i];
    b->mode = (uint8_t)d->mode; b->cbp_y = (uint8_t)d->cbp.y; b->cbp_u = (uint8_t)d->cbp.u; b->cbp_v = (uint8_t)d->cbp.v; b->size = d->size; b->tb_split = d->tb_split;
    b->pb_part = (uint8_t)d->pb_part;
    b->mv0.x = d->inter_pred.mv0.x; b->mv0.y = d->inter_pred.mv0.y; b->mv1.x = d->inter_pred.mv1.x; b->mv1.y = d->inter_pred.mv1.y;
    b->ref_idx0 = (uint8_t)d->inter_pred.ref_idx0; b->ref_idx1 = (uint8_t)d->inter_pred.ref_idx1; b->bipred_flag = (int8_t)(int32_t)d->inter_pred.bipred_flag;
  }
}

/* ---- serialise one super block's quad-tree with the reference's bit writer (stream order of process_block, :2401-2565) */
typedef struct { encoder_info_t *e; const tb_rdo_leaf_t *leaves; const int16_t *coeffs; int n, next, hbd; } walk_t;
static block_info_t s_bi;
static block_param_t s_bp;

static void unpack_plane(int16_t *dst, const int16_t *src, int size, int tb_split) {
  const int t = tb_split ? size / 2 : size, q = t < MAX_QUANT_SIZE ? t : MAX_QUANT_SIZE, n = tb_split ? 4 : 1;
  for (int k = 0; k < n; k++) memcpy(dst + k * MAX_QUANT_SIZE * MAX_QUANT_SIZE, src + k * q * q, (size_t)q * q * sizeof(int16_t));
}

static void emit_node(walk_t *w, int x, int y, int size) {
  encoder_info_t *e = w->e;
  const int width = e->width, height = e->height;
  if (y + MIN_BLOCK_SIZE > height || x + MIN_BLOCK_SIZE > width) return;
  const int encode_this = y + size <= height && x + size <= width;
  block_context_t ctx;
  memset(&s_bi, 0, sizeof(s_bi));
  s_bi.block_pos.size = size; s_bi.block_pos.ypos = y; s_bi.block_pos.xpos = x;
  s_bi.block_pos.bwidth = size < width - x ? size : width - x; s_bi.block_pos.bheight = size < height - y ? size : height - y;
  s_bi.block_pos.sb_size = 1 << e->params->log2_sb_size;
  s_bi.max_num_tb_part = e->params->enable_tb_split == 1 ? 2 : 1; s_bi.max_num_pb_part = e->params->enable_pb_split ? 4 : 1;
  s_bi.qp = e->frame_info.qp; s_bi.sub = 1; s_bi.delta_qp = 0; s_bi.block_context = &ctx;
  const tb_rdo_leaf_t *L = w->next < w->n ? &w->leaves[w->next] : NULL;
  if (L && L->xpos == x && L->ypos == y && L->size == size) {
    w->next++;
    ctx.index = L->ctx_index; ctx.cbp = L->ctx_cbp; ctx.split = 0; ctx.mode = 0; ctx.size = 0;
    s_bi.num_skip_vec = L->num_skip_vec; s_bi.num_merge_vec = L->num_merge_vec; s_bi.mvp.x = L->mvp.x; s_bi.mvp.y = L->mvp.y;
    s_bp.mode = (block_mode_t)L->mode; s_bp.intra_mode = (intra_mode_t)L->intra_mode; s_bp.skip_idx = L->skip_idx; s_bp.pb_part = L->pb_part;
    s_bp.ref_idx0 = L->ref_idx0; s_bp.ref_idx1 = L->ref_idx1; s_bp.dir = L->dir; s_bp.tb_split = L->tb_split; s_bp.tb_param = L->tb_split;
    s_bp.cbp.y = L->cbp_y; s_bp.cbp.u = L->cbp_u; s_bp.cbp.v = L->cbp_v;
    for (int i = 0; i < 4; i++) { s_bp.mv_arr0[i].x = L->mv_arr0[i].x; s_bp.mv_arr0[i].y = L->mv_arr0[i].y; s_bp.mv_arr1[i].x = L->mv_arr1[i].x; s_bp.mv_arr1[i].y = L->mv_arr1[i].y; }
    if (L->coeff_ofs >= 0) {
      const int sc = size >> 1, tbc = L->tb_split && sc > 4;
      const int ny = tb_rdo_coeff_count(size, L->tb_split), nc = tb_rdo_coeff_count(sc, tbc);
      const int16_t *src = w->coeffs + L->coeff_ofs;
      unpack_plane(s_bp.coeff_y, src, size, L->tb_split);
      unpack_plane(s_bp.coeff_u, src + ny, sc, tbc);
      unpack_plane(s_bp.coeff_v, src + ny + nc, sc, tbc);
    }
    write_block(e->stream, e, &s_bi, &s_bp);
    return;
  }
  /* split: the context of THIS node comes from its up / left neighbours, which precede it in coding order and are final */
  (w->hbd ? find_block_contexts_hbd : find_block_contexts_lbd)(y, x, height, width, size, e->deblock_data, &ctx, e->params->use_block_contexts);
  s_bp.mode = MODE_SKIP;
  write_super_mode(e->stream, e, &s_bi, &s_bp, 1, encode_this);
  const int ns = size / 2;
  emit_node(w, x, y, ns);
  emit_node(w, x, y + ns, ns);
  emit_node(w, x + ns, y, ns);
  emit_node(w, x + ns, y + ns, ns);
}

static void copy_region(void *dst, const void *src, int stride, int x0, int y0, int x1, int y1, int esz) {
  for (int y = y0; y < y1; y++) memcpy((char *)dst + ((size_t)y * stride + x0) * esz, (const char *)src + ((size_t)y * stride + x0) * esz, (size_t)(x1 - x0) * esz);
}

/* ---- capture (reference mode): one file per frame, see the header comment.  Layout: magic, tb_rdo_frame_t (pointers zeroed), frame_num, nsb,
 * orig Y U V (visible, compact), per reference Y U V (whole padded planes); then, after the last super block: rec Y U V (visible, compact),
 * tb_rdo_blk_t grid, int32 cost per super block. */
static void dump_plane(FILE *fp, const void *p, int stride, int w, int h, int esz) {
  for (int y = 0; y < h; y++) fwrite((const char *)p + (size_t)y * stride * esz, (size_t)esz, (size_t)w, fp);
}
static void capture_begin(encoder_info_t *e, const tb_rdo_frame_t *f, int esz) {
  const char *dir = getenv("TB_RDO_DUMP");
  char path[1024];
  snprintf(path, sizeof(path), "%s/frame_%03d.job", dir, (int)e->frame_info.frame_num);
  G.dump = fopen(path, "wb");
  if (!G.dump) { fprintf(stderr, "[tb_rdo_shim] cannot write %s\n", path); exit(2); }
  tb_rdo_frame_t h = *f;
  memset(h.orig, 0, sizeof(h.orig)); memset(h.ref, 0, sizeof(h.ref)); memset(h.rec, 0, sizeof(h.rec)); h.blk = NULL; h.leaves = NULL; h.leaf_count = NULL; h.coeffs = NULL;
  const int32_t extra[2] = {(int32_t)e->frame_info.frame_num, (int32_t)G.nsb};
  fwrite("TBJOB1\0\0", 1, 8, G.dump);
  fwrite(&h, sizeof(h), 1, G.dump);
  fwrite(extra, sizeof(extra), 1, G.dump);
  const int w = f->width, hh = f->height;
  for (int p = 0; p < 3; p++) dump_plane(G.dump, f->orig[p], f->orig_stride[p ? 1 : 0], p ? w >> 1 : w, p ? hh >> 1 : hh, esz);
  const int pad = f->ref_pad, padc = pad >> 1;
  for (int r = 0; r < f->num_ref; r++)
    for (int p = 0; p < 3; p++) {
      const int st = f->ref_stride[p ? 1 : 0], pd = p ? padc : pad, ph = (p ? hh >> 1 : hh) + 2 * pd;
      fwrite((const char *)f->ref[r][p] - ((size_t)pd * st + pd) * esz, (size_t)esz, (size_t)st * ph, G.dump);
    }
}
static void capture_end(encoder_info_t *e, const tb_rdo_frame_t *f, int esz) {
  const int w = f->width, hh = f->height;
  for (int p = 0; p < 3; p++) dump_plane(G.dump, f->rec[p], f->rec_stride[p ? 1 : 0], p ? w >> 1 : w, p ? hh >> 1 : hh, esz);
  deblock_to_blk(e);
  fwrite(G.blk, sizeof(tb_rdo_blk_t), (size_t)(hh / 4) * (w / 4), G.dump);
  fwrite(G.sb_cost, sizeof(int32_t), (size_t)G.nsb, G.dump);
  fclose(G.dump);
  G.dump = NULL;
}

static int wrap_process_block(encoder_info_t *e, int size, int ypos, int xpos, int qp, int sub, int hbd) {
  int (*real)(encoder_info_t *, int, int, int, int, int) = hbd ? __real_process_block_hbd : __real_process_block_lbd;
  const int esz = hbd ? 2 : 1;
  if (!supported(e, qp, sub, hbd) || size != (1 << e->params->log2_sb_size)) return real(e, size, ypos, xpos, qp, sub);
  ensure_buffers(e->width, e->height, e->params->log2_sb_size);
  const int sb = size, nsbx = (e->width + sb - 1) / sb, sbx = xpos / sb, sby = ypos / sb, sbi = sby * nsbx + sbx;
  const int x1 = xpos + sb < e->width ? xpos + sb : e->width, y1 = ypos + sb < e->height ? ypos + sb : e->height;
  tb_rdo_frame_t f;
  fill_desc(&f, e, esz);

  if (!tb_rdo_encode_frame && !G.verify) { /* capture link: the reference decides, the shim observes */
    const int dumping = getenv("TB_RDO_DUMP") != NULL;
    if (dumping && xpos == 0 && ypos == 0) capture_begin(e, &f, esz);
    const double t0 = now_s();
    const int ret = real(e, size, ypos, xpos, qp, sub);
    G.t_ref += now_s() - t0;
    G.sb_cost[sbi] = ret;
    if (xpos == 0 && ypos == 0) G.frames_host++;
    if (G.progress) {
      G.progress->pixels += (uint64_t)(x1 - xpos) * (uint64_t)(y1 - ypos);
      G.progress->t_rd = G.t_ref;
      if (x1 == e->width && y1 == e->height) G.progress->frames += 1;
    }
    if (dumping && G.dump && x1 == e->width && y1 == e->height) capture_end(e, &f, esz);
    return ret;
  }

  if (G.verify) {
    if (!tb_rdo_encode_sb) { fprintf(stderr, "[tb_rdo_shim] TB_RDO_VERIFY needs tb_rdo_encode_sb (oracle/librdo_hostcheck.so)\n"); exit(2); }
    /* ours first, on the encoder's current state; then the reference on the same state; compare; keep the reference's */
    deblock_to_blk(e);
    stream_pos_t pos0; read_stream_pos(&pos0, e->stream);
    if (tb_rdo_encode_sb(&f, sbx, sby) != TB_OK) { fprintf(stderr, "[tb_rdo_shim] tb_rdo_encode_sb failed\n"); exit(2); }
    blk_to_deblock(e, xpos, ypos, x1, y1);
    walk_t w = {e, G.leaves + (size_t)sbi * TB_RDO_MAX_LEAVES, G.coeffs + (size_t)sbi * TB_RDO_SB_COEFFS, G.leaf_count[sbi], 0, hbd};
    emit_node(&w, xpos, ypos, sb);
    const int our_bits = get_bit_pos(e->stream);
    /* snapshot ours */
    const int sy = e->rec->stride_y, sc = e->rec->stride_c;
    static uint8_t *oy, *ou, *ov; static deblock_data_t *odd; static uint8_t *obits; static size_t cap;
    const size_t need = (size_t)(e->rec->height + 2 * e->rec->pad_ver_y) * sy * esz;
    if (cap < need) { free(oy); free(ou); free(ov); free(odd); free(obits); oy = malloc(need); ou = malloc(need); ov = malloc(need);
                      odd = malloc(sizeof(deblock_data_t) * (e->width / 4) * (e->height / 4)); obits = malloc(MAX_BUFFER_SIZE); cap = need; }
    copy_region(oy, e->rec->y, sy, xpos, ypos, x1, y1, esz);
    copy_region(ou, e->rec->u, sc, xpos / 2, ypos / 2, x1 / 2, y1 / 2, esz);
    copy_region(ov, e->rec->v, sc, xpos / 2, ypos / 2, x1 / 2, y1 / 2, esz);
    const int bs = e->width / 4;
    for (int by = ypos / 4; by < y1 / 4; by++) memcpy(&odd[by * bs + xpos / 4], &e->deblock_data[by * bs + xpos / 4], sizeof(deblock_data_t) * (x1 - xpos) / 4);
    stream_pos_t pos1; read_stream_pos(&pos1, e->stream);
    /* flush a copy of our bits: bytes [pos0.bytepos, pos1.bytepos) + the pending word */
    const uint32_t b0 = pos0.bytepos, b1 = pos1.bytepos;
    memcpy(obits, e->stream->bitstream + b0, b1 - b0);
    const uint32_t our_buf = pos1.bitbuf, our_rest = pos1.bitrest;
    write_stream_pos(e->stream, &pos0);
    const int ret = real(e, size, ypos, xpos, qp, sub);
    const int ref_bits = get_bit_pos(e->stream);
    stream_pos_t pos2; read_stream_pos(&pos2, e->stream);
    int bad = 0;
    if (our_bits != ref_bits || pos2.bytepos != b1 || pos2.bitbuf != our_buf || pos2.bitrest != our_rest || memcmp(obits, e->stream->bitstream + b0, b1 - b0)) bad |= 1;
    for (int y = ypos; y < y1 && !(bad & 2); y++)
      if (memcmp((char *)oy + ((size_t)y * sy + xpos) * esz, (char *)e->rec->y + ((size_t)y * sy + xpos) * esz, (size_t)(x1 - xpos) * esz)) bad |= 2;
    for (int y = ypos / 2; y < y1 / 2 && !(bad & 4); y++)
      if (memcmp((char *)ou + ((size_t)y * sc + xpos / 2) * esz, (char *)e->rec->u + ((size_t)y * sc + xpos / 2) * esz, (size_t)(x1 - xpos) / 2 * esz) ||
          memcmp((char *)ov + ((size_t)y * sc + xpos / 2) * esz, (char *)e->rec->v + ((size_t)y * sc + xpos / 2) * esz, (size_t)(x1 - xpos) / 2 * esz)) bad |= 4;
    for (int by = ypos / 4; by < y1 / 4; by++)
      for (int bx = xpos / 4; bx < x1 / 4; bx++) {
        const deblock_data_t *a = &odd[by * bs + bx], *b = &e->deblock_data[by * bs + bx];
        if (a->mode != b->mode || a->cbp.y != b->cbp.y || a->cbp.u != b->cbp.u || a->cbp.v != b->cbp.v || a->size != b->size || a->tb_split != b->tb_split ||
            a->pb_part != b->pb_part || memcmp(&a->inter_pred, &b->inter_pred, sizeof(inter_pred_t))) {
          if (!(bad & 8) && getenv("TB_RDO_VERBOSE"))
            fprintf(stderr, "  first deblock_data difference at (%d,%d): ours mode %d size %d cbp %d%d%d tbs %d part %d mv0 (%d,%d) ref %d/%d dir %d | ref mode %d size %d cbp %d%d%d tbs %d part %d mv0 (%d,%d) ref %d/%d dir %d\n",
                    bx * 4, by * 4, a->mode, a->size, a->cbp.y, a->cbp.u, a->cbp.v, a->tb_split, a->pb_part, a->inter_pred.mv0.x, a->inter_pred.mv0.y, a->inter_pred.ref_idx0,
                    a->inter_pred.ref_idx1, (int)a->inter_pred.bipred_flag, b->mode, b->size, b->cbp.y, b->cbp.u, b->cbp.v, b->tb_split, b->pb_part, b->inter_pred.mv0.x,
                    b->inter_pred.mv0.y, b->inter_pred.ref_idx0, b->inter_pred.ref_idx1, (int)b->inter_pred.bipred_flag);
          bad |= 8;
        }
      }
    G.sb_total++;
    if (bad) {
      G.sb_bad++;
      fprintf(stderr, "[tb_rdo_shim] MISMATCH frame %d type %d SB (%d,%d): %s%s%s%s (bits ours %d ref %d)\n", e->frame_info.frame_num, e->frame_info.frame_type, sbx, sby,
              bad & 1 ? "bitstream " : "", bad & 2 ? "rec-luma " : "", bad & 4 ? "rec-chroma " : "", bad & 8 ? "deblock_data " : "", our_bits - 8 * (int)b0 - (32 - (int)pos0.bitrest),
              ref_bits - 8 * (int)b0 - (32 - (int)pos0.bitrest));
    }
    return ret;
  }

  if (xpos == 0 && ypos == 0) {
    const double t0 = now_s();
    const int rc = tb_rdo_encode_frame(&f);
    G.t_rdo += now_s() - t0;
    if (getenv("TB_RDO_TRACE")) fprintf(stderr, "[tb_rdo_shim] frame %d type %d refs %d qp %d: tb_rdo_encode_frame %.1f ms\n", e->frame_info.frame_num, e->frame_info.frame_type, e->frame_info.num_ref, (int)e->frame_info.qp, 1e3 * (now_s() - t0));
    G.frame_ok = rc == TB_OK;
    if (G.frame_ok) G.frames_dev++; else G.frames_host++;
    if (!G.frame_ok)
      fprintf(stderr, "[tb_rdo_shim] frame %d (type %d, %d refs): tb_rdo_encode_frame returned %d (%s) -> the reference's host loop takes this frame\n", e->frame_info.frame_num,
              e->frame_info.frame_type, e->frame_info.num_ref, rc, tb_rdo_last_error ? tb_rdo_last_error() : "");
  }
  if (!G.frame_ok) return real(e, size, ypos, xpos, qp, sub);
  /* tb_rdo_encode_frame wrote the whole reconstruction into e->rec; here: this super block's deblock_data + bits */
  const double t1 = now_s();
  blk_to_deblock(e, xpos, ypos, x1, y1);
  walk_t w = {e, G.leaves + (size_t)sbi * TB_RDO_MAX_LEAVES, G.coeffs + (size_t)sbi * TB_RDO_SB_COEFFS, G.leaf_count[sbi], 0, hbd};
  emit_node(&w, xpos, ypos, sb);
  G.t_emit += now_s() - t1;
  uint32_t cost = 0;
  for (int i = 0; i < w.n; i++) cost += w.leaves[i].cost;
  return (int)cost;
}

int __wrap_process_block_lbd(encoder_info_t *e, int size, int ypos, int xpos, int qp, int sub) { return wrap_process_block(e, size, ypos, xpos, qp, sub, 0); }
int __wrap_process_block_hbd(encoder_info_t *e, int size, int ypos, int xpos, int qp, int sub) { return wrap_process_block(e, size, ypos, xpos, qp, sub, 1); }
