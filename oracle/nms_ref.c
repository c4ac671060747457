/*
 * CPU restatement of torchvision.ops.batched_nms / nms -- TEST INFRASTRUCTURE ONLY (see restate.py).
 *
 * The reference calls torchvision at yolort/models/box_head.py:422; torchvision is a third-party
 * dependency that the reference does not pin (requirements.txt:1-6).  This file restates the published
 * CPU algorithm of torchvision 0.26 (the version the golden fixtures were generated with):
 *   - torchvision/csrc/ops/cpu/nms_kernel.cpp (nms_kernel_impl): stable descending sort of the scores,
 *     greedy sweep, suppress when inter / (area_i + area_j - inter) > thr (strict), fp32, no box validation;
 *   - torchvision/ops/boxes.py batched_nms: coordinate-offset trick when boxes.numel() <= 4000 on CPU,
 *     otherwise per-class NMS and a final descending sort of the kept scores (tie order: index ascending).
 * Compile with -ffp-contract=off so no multiply-add is fused.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void merge_sort_desc(int64_t* idx, int64_t* tmp, int64_t n, const float* key) {
  if (n < 2) return;
  int64_t h = n / 2;
  merge_sort_desc(idx, tmp, h, key);
  merge_sort_desc(idx + h, tmp, n - h, key);
  int64_t i = 0, j = h, k = 0;
  while (i < h && j < n) {
    /* stable: take the right element only when it is strictly greater */
    if (key[idx[j]] > key[idx[i]]) tmp[k++] = idx[j++];
    else tmp[k++] = idx[i++];
  }
  while (i < h) tmp[k++] = idx[i++];
  while (j < n) tmp[k++] = idx[j++];
  memcpy(idx, tmp, (size_t)n * sizeof(int64_t));
}

/* boxes [n][4] xyxy; returns number kept; keep[] = indices in score-descending order */
int64_t oracle_nms(const float* boxes, const float* scores, int64_t n, float thr, int64_t* keep) {
  if (n <= 0) return 0;
  int64_t* order = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  int64_t* tmp = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  float* areas = (float*)malloc((size_t)n * sizeof(float));
  unsigned char* sup = (unsigned char*)calloc((size_t)n, 1);
  for (int64_t i = 0; i < n; ++i) {
    order[i] = i;
    const float* b = boxes + 4 * i;
    float w = b[2] - b[0];
    float h = b[3] - b[1];
    areas[i] = w * h;
  }
  merge_sort_desc(order, tmp, n, scores);
  int64_t nk = 0;
  for (int64_t _i = 0; _i < n; ++_i) {
    int64_t i = order[_i];
    if (sup[i]) continue;
    keep[nk++] = i;
    const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
    const float iarea = areas[i];
    for (int64_t _j = _i + 1; _j < n; ++_j) {
      int64_t j = order[_j];
      if (sup[j]) continue;
      const float* b = boxes + 4 * j;
      float xx1 = ix1 > b[0] ? ix1 : b[0];
      float yy1 = iy1 > b[1] ? iy1 : b[1];
      float xx2 = ix2 < b[2] ? ix2 : b[2];
      float yy2 = iy2 < b[3] ? iy2 : b[3];
      float w = xx2 - xx1;
      float h = yy2 - yy1;
      if (w < 0.0f) w = 0.0f;
      if (h < 0.0f) h = 0.0f;
      float inter = w * h;
      float uni = iarea + areas[j];
      uni = uni - inter;
      float ovr = inter / uni;
      if (ovr > thr) sup[j] = 1;
    }
  }
  free(order);
  free(tmp);
  free(areas);
  free(sup);
  return nk;
}

/* semantics: 0 = torchvision auto (numel > 4000 -> per class), 1 = per class, 2 = offset trick */
int64_t oracle_batched_nms(const float* boxes, const float* scores, const int64_t* labels, int64_t n, float thr,
                           int semantics, int64_t* keep) {
  if (n <= 0) return 0;
  int vanilla = semantics == 1 || (semantics == 0 && n * 4 > 4000);
  if (!vanilla) {
    float maxc = boxes[0];
    for (int64_t i = 0; i < 4 * n; ++i)
      if (boxes[i] > maxc) maxc = boxes[i];
    float unit = maxc + 1.0f;
    float* sh = (float*)malloc((size_t)n * 4 * sizeof(float));
    for (int64_t i = 0; i < n; ++i) {
      float off = (float)labels[i] * unit;
      for (int c = 0; c < 4; ++c) sh[4 * i + c] = boxes[4 * i + c] + off;
    }
    int64_t nk = oracle_nms(sh, scores, n, thr, keep);
    free(sh);
    return nk;
  }
  unsigned char* mask = (unsigned char*)calloc((size_t)n, 1);
  unsigned char* done = (unsigned char*)calloc((size_t)n, 1);
  int64_t* cur = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  int64_t* ck = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  float* cb = (float*)malloc((size_t)n * 4 * sizeof(float));
  float* cs = (float*)malloc((size_t)n * sizeof(float));
  for (int64_t s = 0; s < n; ++s) {
    if (done[s]) continue;
    int64_t cls = labels[s], m = 0;
    for (int64_t i = s; i < n; ++i)
      if (labels[i] == cls) {
        done[i] = 1;
        cur[m] = i;
        memcpy(cb + 4 * m, boxes + 4 * i, 4 * sizeof(float));
        cs[m] = scores[i];
        ++m;
      }
    int64_t k = oracle_nms(cb, cs, m, thr, ck);
    for (int64_t i = 0; i < k; ++i) mask[cur[ck[i]]] = 1;
  }
  int64_t nk = 0;
  for (int64_t i = 0; i < n; ++i)
    if (mask[i]) keep[nk++] = i;
  merge_sort_desc(keep, cur, nk, scores);
  free(mask);
  free(done);
  free(cur);
  free(ck);
  free(cb);
  free(cs);
  return nk;
}
