// tb_tinterp.cuh — a21: temporal frame interpolation (common/temporal_interp.c:36-992) on the device.
//
// The algorithm is normative (the decoder runs it too, dec/decode_frame.c:95-113) and raster-serial: block (i,j) of the
// 16x16 search reads the final vectors of its left, up and up-right neighbours.  Device mapping:
//   * ti_me_kernel     one WARP per block row, rows advance as a wavefront: row i may start block j once row i-1 has
//                      published block j+1 (progress counters in HBM, __threadfence + volatile polling).  All rows of a
//                      level are resident at once (<= 68 warps at 1080p).  Inside a block the four cross-refinement
//                      probes of a round run on four 8-lane groups; winners follow the sequential first-strict-minimum rule.
//   * ti_merge_kernel  one warp per 8x8 grid block (independent): best of <= 5 neighbour vectors by 8x8 SAD.
//   * ti_upscale_kernel, ti_interp_kernel: element-wise.
// Vectors are in 1/8 sample units, rounded to integers for every access ((mv + 4) >> 3); mv[1] refers to the farther
// picture, mv[0] = scale(mv[1], -wt1, wt0).  All arithmetic is integer.
#pragma once
#include "tb_device.cuh"

namespace tb {

struct TiPic { const void *y; int stride, width, height, pad; };
struct TiLevel {
  TiPic pic[2];          // pic[0] / pic[1] already swapped for `reversed`
  short2 *mv0, *mv1;     // bw x bh vectors on the 8x8 grid (mv[0], mv[1])
  const short2 *guide;   // spatial_mv_data[lvl].mv[1] or nullptr at the coarsest level
  int bw, bh, wt0, wt1, reversed, guide_reversed, guide_wt0;
  int *progress;         // one counter per 16x16 block row
};

__device__ __forceinline__ int ti_scale_val(int v, int numer, int denom) {  // temporal_interp.c:58-67
  if (denom == 0) return 0;
  int prod = v * numer;
  if (denom < 0) { denom = -denom; prod = -prod; }
  return prod >= 0 ? (prod + denom / 2) / denom : -((-prod + denom / 2) / denom);
}
__device__ __forceinline__ short2 ti_scale_mv(short2 mv, int numer, int denom) {  // :69-82
  if (numer == denom) return mv;
  if (numer == -denom) return make_short2((short)-mv.x, (short)-mv.y);
  return make_short2((short)ti_scale_val(mv.x, numer, denom), (short)ti_scale_val(mv.y, numer, denom));
}
__device__ __forceinline__ short2 ti_ld_mv(const short2 *p) {  // vectors written by other SMs: bypass L1
  int v = __ldcg((const int *)p);
  return make_short2((short)(v & 0xffff), (short)(v >> 16));
}
__device__ __forceinline__ int ti_add_cand(short2 *list, int len, short2 c) {  // :205-218 (max 20 never reached: <= 5 candidates)
  for (int i = 0; i < len; i++)
    if (list[i].x == c.x && list[i].y == c.y) return len;
  list[len] = c;
  return len + 1;
}
__device__ __forceinline__ short2 ti_absdist_filter(const short2 *l, int num) {  // :695-716, ties -> last
  int best = 0, best_cost = 0x3fffffff;
  for (int j = 0; j < num; j++) {
    int cost = 0;
    for (int i = 0; i < num; i++) cost += iabs(l[i].x - l[j].x) + iabs(l[i].y - l[j].y);
    if (cost <= best_cost) { best = j; best_cost = cost; }
  }
  return l[best];
}

// SAD between the displaced size x size blocks of the two pictures, samples clamped to the padded area (identical to the
// unclamped form whenever both blocks lie inside it, temporal_interp.c:375-456).  Lanes sub..: px sub, sub+nl, ...
template <class S>
__device__ __forceinline__ uint32_t ti_sad_partial(const TiPic &p0, const TiPic &p1, int xstart, int ystart, short2 mv0, short2 mv1, int size, int sub, int nl) {
  const int pad = p0.pad, wP = p0.width + pad, hP = p0.height + pad;
  const int x0 = xstart + ((mv0.x + 4) >> 3), x1 = xstart + ((mv1.x + 4) >> 3), y0 = ystart + ((mv0.y + 4) >> 3), y1 = ystart + ((mv1.y + 4) >> 3);
  const S *a = (const S *)p0.y, *b = (const S *)p1.y;
  const int ls = ilog2(size);
  uint32_t acc = 0;
  for (int p = sub; p < size * size; p += nl) {
    int i = p >> ls, j = p & (size - 1);
    int xa = iclip(j + x0, -pad, wP - 1), xb = iclip(j + x1, -pad, wP - 1), ya = iclip(i + y0, -pad, hP - 1), yb = iclip(i + y1, -pad, hP - 1);
    acc += (uint32_t)iabs((int)b[yb * p1.stride + xb] - (int)a[ya * p0.stride + xa]);
  }
  return acc;
}

// neighbour-smoothness cost (:299-317, idx = 1); neighbours are final (already published) vectors
__device__ __forceinline__ int ti_mv_cost(short2 mv, const short2 *a, int bw, int xp, int yp, int lambda) {
  const int st = 2;
  int diff = 0;
#define TI_D(p) { short2 n_ = ti_ld_mv(a + (p)); diff += iabs(mv.x - n_.x) + iabs(mv.y - n_.y); }
  if (xp == 0 && yp == 0) diff = 0;
  else if (yp > 0 && xp > 0 && xp < bw - st) { TI_D((yp - st) * bw + xp + st) TI_D((yp - st) * bw + xp) TI_D((yp - st) * bw + xp - st) TI_D(yp * bw + xp - st) }
  else if (yp == 0) { TI_D(xp - st) }
  else if (xp == 0) { TI_D((yp - st) * bw + xp + st) TI_D((yp - st) * bw + xp) }
#undef TI_D
  return (diff * lambda) >> 7;
}

// raster pass of motion_estimate_bi (:786-851): one warp per 16x16 block row
template <class S> __global__ void __launch_bounds__(128) ti_me_kernel(TiLevel L) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // block row index
  const int nrows = L.bh >> 1, nblk = L.bw >> 1;
  if (row >= nrows) return;
  const int lane = lane_id(), i = row * 2, bw = L.bw;
  const int pad = L.pic[0].pad, wP = L.pic[0].width + pad, hP = L.pic[0].height + pad;
  volatile int *prog = L.progress;
  for (int jb = 0; jb < nblk; jb++) {
    const int j = jb * 2, pos = i * bw + j, xstart = j * 8, ystart = i * 8;
    if (row > 0) {  // wait for the up-right neighbour
      int need = min(jb + 2, nblk);
      if (lane == 0)
        while (prog[row - 1] < need) { __nanosleep(64); }
      __syncwarp();
      __threadfence();
    }
    // ---- skip vector (:758-770)
    short2 vl[3];
    int num = 0;
    if (i > 0 && j < bw - 2) vl[num++] = ti_ld_mv(L.mv1 + (i - 2) * bw + j + 2);
    if (j > 0) vl[num++] = ti_ld_mv(L.mv1 + i * bw + j - 2);
    if (i > 0) vl[num++] = ti_ld_mv(L.mv1 + (i - 2) * bw + j);
    short2 skip_mv = make_short2(0, 0);
    if (num) skip_mv = ti_absdist_filter(vl, num);
    short2 sskip_mv = ti_scale_mv(skip_mv, -L.wt1, L.wt0);
    // ---- skip test (:458-580): four 8x8 quarters, each inside the padded area with SAD <= 512
    bool skip;
    {
      const int quarter = lane >> 3, sub = lane & 7;
      const int q = xstart + (quarter & 1) * 8, p = ystart + (quarter >> 1) * 8;
      const int x0 = q + ((sskip_mv.x + 4) >> 3), x1 = q + ((skip_mv.x + 4) >> 3), y0 = p + ((sskip_mv.y + 4) >> 3), y1 = p + ((skip_mv.y + 4) >> 3);
      bool inb = x0 >= -pad && x0 + 8 <= wP && y0 >= -pad && y0 + 8 <= hP && x1 >= -pad && x1 + 8 <= wP && y1 >= -pad && y1 + 8 <= hP;
      uint32_t s = ti_sad_partial<S>(L.pic[0], L.pic[1], q, p, sskip_mv, skip_mv, 8, sub, 8);
      s = group_sum(s, 8);
      skip = __all_sync(FULL, inb && s <= 512u);
    }
    short2 best_mv, best_smv;
    if (skip) {
      best_mv = skip_mv;
      best_smv = sskip_mv;
    } else {
      // ---- candidates (:230-283)
      short2 cand[5];
      int len = 0;
      len = ti_add_cand(cand, len, make_short2(0, 0));
      const bool guided = L.guide != nullptr;
      if (guided) {
        int numer = (L.reversed == L.guide_reversed) ? L.wt0 : -L.wt0;
        len = ti_add_cand(cand, len, ti_scale_mv(L.guide[pos], numer, L.guide_wt0));
      }
      if (i > 0 && j < bw - 2) len = ti_add_cand(cand, len, ti_ld_mv(L.mv1 + (i - 2) * bw + j + 2));
      if (j > 0) len = ti_add_cand(cand, len, ti_ld_mv(L.mv1 + i * bw + j - 2));
      if (i > 0) len = ti_add_cand(cand, len, ti_ld_mv(L.mv1 + (i - 2) * bw + j));
      // ---- adaptive search (:584-668)
      const int lambda = guided ? 750 : 3000;
      best_mv = cand[0];
      best_smv = ti_scale_mv(cand[0], -L.wt1, L.wt0);
      uint32_t best_cost = 0x3fffffffu;
      for (int c = 0; c < len; c++) {
        short2 rl = cand[c], rsl = ti_scale_mv(cand[c], -L.wt1, L.wt0);
        uint32_t cc = (uint32_t)ti_mv_cost(rl, L.mv1, bw, j, i, lambda);
        cc += warp_sum(ti_sad_partial<S>(L.pic[0], L.pic[1], xstart, ystart, rsl, rl, 16, lane, 32));
        if (((4u + (uint32_t)c) * cc) / 8u < best_cost) {
          int shift = guided ? 3 : 6, count = guided ? 8 : 64;
          while (shift >= 3 && count > 0) {
            const int off = 1 << shift, k = lane >> 3;  // probe k on lanes 8k..8k+7: (-off,0) (+off,0) (0,-off) (0,+off)
            short2 r = rl;
            if (k == 0) r.x = (short)(rl.x - off);
            else if (k == 1) r.x = (short)(rl.x + off);
            else if (k == 2) r.y = (short)(rl.y - off);
            else r.y = (short)(rl.y + off);
            short2 rs = ti_scale_mv(r, -L.wt1, L.wt0);
            uint32_t bc = group_sum(ti_sad_partial<S>(L.pic[0], L.pic[1], xstart, ystart, rs, r, 16, lane & 7, 8), 8);
            bc += (uint32_t)ti_mv_cost(r, L.mv1, bw, j, i, lambda);
            // sequential rule: the first probe (k = 0..3) that attains the minimum, if it is below the running cost
            bool better = false;
            for (int t = 0; t < 4; t++) {
              uint32_t bt = __shfl_sync(FULL, bc, t * 8);
              if (bt < cc) {
                cc = bt;
                int rx = __shfl_sync(FULL, (int)r.x, t * 8), ry = __shfl_sync(FULL, (int)r.y, t * 8);
                int sx = __shfl_sync(FULL, (int)rs.x, t * 8), sy = __shfl_sync(FULL, (int)rs.y, t * 8);
                rl = make_short2((short)rx, (short)ry);
                rsl = make_short2((short)sx, (short)sy);
                better = true;
              }
            }
            if (!better) shift--;
            count -= 4;
          }
        }
        if (cc < best_cost) { best_mv = rl; best_smv = rsl; best_cost = cc; }
      }
    }
    // ---- publish the block's vectors on its 2x2 grid cells, then advance the row's progress counter
    if (lane < 4) {
      int cell = pos + (lane >> 1) * bw + (lane & 1);
      L.mv0[cell] = best_smv;
      L.mv1[cell] = best_mv;
    }
    __threadfence();  // every writing lane orders its vectors before the flag
    __syncwarp();
    if (lane == 0) prog[row] = jb + 1;
  }
}

// merge pass (:219-229, 670-693, 853-872): one warp per 8x8 grid block, reads mv1 (pre-merge), writes m0/m1
template <class S> __global__ void __launch_bounds__(128) ti_merge_kernel(TiLevel L, short2 *m0, short2 *m1) {
  const int n = L.bw * L.bh, lane = lane_id();
  for (int b = global_warp(); b < n; b += total_warps()) {
    const int i = b / L.bw, j = b - i * L.bw, off = (i & 1) ? 2 : 1;
    short2 cand[5];
    int len = 0;
    len = ti_add_cand(cand, len, L.mv1[b]);
    if (i - off >= 0) len = ti_add_cand(cand, len, L.mv1[(i - off) * L.bw + j]);
    if (i + off < L.bh) len = ti_add_cand(cand, len, L.mv1[(i + off) * L.bw + j]);
    if (j - off >= 0) len = ti_add_cand(cand, len, L.mv1[b - off]);
    if (j + off < L.bw) len = ti_add_cand(cand, len, L.mv1[b + off]);
    short2 bm = L.mv1[b], bs = L.mv0[b];
    if (len > 1) {
      uint32_t best = 0x3fffffffu;
      bm = make_short2(0, 0);
      bs = make_short2(0, 0);
      for (int c = 0; c < len; c++) {
        short2 s = ti_scale_mv(cand[c], -L.wt1, L.wt0);
        uint32_t bc = warp_sum(ti_sad_partial<S>(L.pic[0], L.pic[1], j * 8, i * 8, s, cand[c], 8, lane, 32));
        if (bc < best) { best = bc; bm = cand[c]; bs = s; }
      }
    }
    if (lane == 0) { m0[b] = bs; m1[b] = bm; }
  }
}

// upscale_mv_data_2x2 (:176-203): coarse level mv1 -> next finer level's guide vectors
__global__ void ti_upscale_kernel(const short2 *in1, int bwi, short2 *out0, short2 *out1, int bwo, int bho, int wt0, int wt1) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= bwo * bho) return;
  const int i = t / bwo, j = t - i * bwo;
  short2 m = in1[(i / 2) * bwi + j / 2];
  m = make_short2((short)(m.x << 1), (short)(m.y << 1));
  out1[t] = m;
  out0[t] = ti_scale_mv(m, -wt1, wt0);
}

// interpolate_frame / mot_comp_avg (:319-373, 877-935): one thread per output sample of one plane
template <class S>
__global__ void ti_interp_kernel(const S *p0, int s0, const S *p1, int s1, S *out, int so, const short2 *mv0a, const short2 *mv1a, int bw, int bh, int wP, int hP,
                                 int pad, int chroma, int wt0, int wt1) {
  const int bs = chroma ? 4 : 8;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= bw * bs || y >= bh * bs) return;
  const int xp = x / bs, yp = y / bs, j = x - xp * bs, i = y - yp * bs;
  short2 mv0 = mv0a[yp * bw + xp], mv1 = mv1a[yp * bw + xp];
  if (chroma) {
    mv1 = make_short2((short)(mv1.x >> 1), (short)(mv1.y >> 1));
    mv0 = ti_scale_mv(mv1, -wt1, wt0);
  }
  const int xstart = xp * bs, ystart = yp * bs;
  const int x0 = xstart + ((mv0.x + 4) >> 3), x1 = xstart + ((mv1.x + 4) >> 3), y0 = ystart + ((mv0.y + 4) >> 3), y1 = ystart + ((mv1.y + 4) >> 3);
  const bool in0 = x0 >= -pad && x0 + bs <= wP && y0 >= -pad && y0 + bs <= hP, in1 = x1 >= -pad && x1 + bs <= wP && y1 >= -pad && y1 + bs <= hP;
  int v;
  if (in0 && in1) v = ((int)p0[(y0 + i) * s0 + x0 + j] + (int)p1[(y1 + i) * s1 + x1 + j] + 1) >> 1;
  else if (in1) v = p1[(y1 + i) * s1 + x1 + j];
  else if (in0) v = p0[(y0 + i) * s1 + x0 + j];  // sic: pitch s1 (temporal_interp.c:353); both pitches are equal in practice
  else {
    int xa = iclip(j + x0, -pad, wP - 1), xb = iclip(j + x1, -pad, wP - 1), ya = iclip(i + y0, -pad, hP - 1), yb = iclip(i + y1, -pad, hP - 1);
    v = ((int)p0[ya * s0 + xa] + (int)p1[yb * s1 + xb] + 1) / 2;
  }
  out[y * so + x] = (S)v;
}

}  // namespace tb
