// Box heads post-process, loss sums and the box-IoU / decision utilities.
// These are latency-bound toys next to the encoder, but they carry the "integer outputs are bit-exact"
// contract: every IoU kernel spells out IEEE round-to-nearest operations in the reference's evaluation
// order with explicit *_rn intrinsics so the compiler cannot contract a*b+c into an FMA.
#include "common.cuh"
#include "grove_b200.h"

namespace grove {

// ---------------------------------------------------------------- post-process (GROVE.py:307-315, bbox_utils.py:25-62)
__global__ void box_postprocess_kernel(const float* __restrict__ boxes, const float* __restrict__ logits, const float* __restrict__ size_wh,
                                       float thr, float* __restrict__ xyxy, uint8_t* __restrict__ keep, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float w = size_wh[2 * i], h = size_wh[2 * i + 1];
  const float cx = __fmul_rn(boxes[4 * i], w), cy = __fmul_rn(boxes[4 * i + 1], h);
  const float bw = __fmul_rn(boxes[4 * i + 2], w), bh = __fmul_rn(boxes[4 * i + 3], h);
  const float hw = __fdiv_rn(bw, 2.f), hh = __fdiv_rn(bh, 2.f);
  xyxy[4 * i] = __fsub_rn(cx, hw);
  xyxy[4 * i + 1] = __fsub_rn(cy, hh);
  xyxy[4 * i + 2] = __fadd_rn(cx, hw);
  xyxy[4 * i + 3] = __fadd_rn(cy, hh);
  const float s = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-logits[i])));
  keep[i] = s > thr ? 1 : 0;
}

// ---------------------------------------------------------------- loss sums (GROVE.py:339-381)
__device__ __forceinline__ float giou_loss_one(const float* p, const float* q) {
  // torchvision.ops.generalized_box_iou_loss on xyxy converted from cxcywh (eps 1e-7, fp32)
  const float x1 = p[0] - p[2] / 2, y1 = p[1] - p[3] / 2, x2 = p[0] + p[2] / 2, y2 = p[1] + p[3] / 2;
  const float x1g = q[0] - q[2] / 2, y1g = q[1] - q[3] / 2, x2g = q[0] + q[2] / 2, y2g = q[1] + q[3] / 2;
  const float xk1 = fmaxf(x1, x1g), yk1 = fmaxf(y1, y1g), xk2 = fminf(x2, x2g), yk2 = fminf(y2, y2g);
  float inter = 0.f;
  if (yk2 > yk1 && xk2 > xk1) inter = (xk2 - xk1) * (yk2 - yk1);
  const float uni = (x2 - x1) * (y2 - y1) + (x2g - x1g) * (y2g - y1g) - inter;
  const float iou = inter / (uni + 1e-7f);
  const float xc1 = fminf(x1, x1g), yc1 = fminf(y1, y1g), xc2 = fmaxf(x2, x2g), yc2 = fmaxf(y2, y2g);
  const float area_c = (xc2 - xc1) * (yc2 - yc1);
  return 1.f - (iou - (area_c - uni) / (area_c + 1e-7f));
}

__global__ void __launch_bounds__(256) box_losses_kernel(const float* __restrict__ boxes, const float* __restrict__ logits, const float* __restrict__ gt,
                                                         const uint8_t* __restrict__ sel, const float* __restrict__ labels, float* __restrict__ sums, int B) {
  __shared__ double red[3][256];
  double g = 0.0, l1 = 0.0, bce = 0.0;
  for (int i = threadIdx.x; i < B; i += 256) {
    if (sel[i]) {
      g += (double)giou_loss_one(boxes + 4 * i, gt + 4 * i);
      float a = 0.f;
      for (int c = 0; c < 4; ++c) a += fabsf(boxes[4 * i + c] - gt[4 * i + c]);
      l1 += (double)a;
    }
    const double x = (double)logits[i], y = (double)labels[i];
    bce += fmax(x, 0.0) - x * y + log1p(exp(-fabs(x)));
  }
  red[0][threadIdx.x] = g; red[1][threadIdx.x] = l1; red[2][threadIdx.x] = bce;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int k = 0; k < 3; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 3) sums[threadIdx.x] = (float)red[threadIdx.x][0];
}

// ---------------------------------------------------------------- IoU matrices
template <typename T> struct RN;
template <> struct RN<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct RN<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};
template <typename T> __device__ __forceinline__ T tmax(T a, T b) { return a > b ? a : b; }   // max(a,b) as Python/numpy pick for non-NaN
template <typename T> __device__ __forceinline__ T tmin(T a, T b) { return a < b ? a : b; }

// mode 0: eval_vidstg.py:13-63.  mode 1: eval_iground.py:39-56.  mode 2: eval_anet.py:22-119 (3-D branch; frm_mask 1 = different frame).
template <typename T>
__global__ void box_iou_kernel(const T* __restrict__ a, int lda, const T* __restrict__ b, int ldb, const uint8_t* __restrict__ frm_mask,
                               T* __restrict__ out, int n, int m, int mode) {
  using R = RN<T>;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * m) return;
  const int i = idx / m, j = idx % m;
  const T a0 = a[i * lda], a1 = a[i * lda + 1], a2 = a[i * lda + 2], a3 = a[i * lda + 3];
  const T b0 = b[j * ldb], b1 = b[j * ldb + 1], b2 = b[j * ldb + 2], b3 = b[j * ldb + 3];
  T res;
  if (mode == 0) {
    const T area1 = R::mul(R::sub(a2, a0), R::sub(a3, a1)), area2 = R::mul(R::sub(b2, b0), R::sub(b3, b1));
    const T w = tmax(R::sub(tmin(a2, b2), tmax(a0, b0)), (T)0), h = tmax(R::sub(tmin(a3, b3), tmax(a1, b1)), (T)0);
    const T inter = R::mul(w, h);
    res = R::div(inter, R::sub(R::add(area1, area2), inter));
  } else if (mode == 1) {
    const T xA = tmax(a0, b0), yA = tmax(a1, b1), xB = tmin(a2, b2), yB = tmin(a3, b3);
    const T inter = R::mul(tmax((T)0, R::add(R::sub(xB, xA), (T)1)), tmax((T)0, R::add(R::sub(yB, yA), (T)1)));
    const T areaA = R::mul(R::add(R::sub(a2, a0), (T)1), R::add(R::sub(a3, a1), (T)1));
    const T areaB = R::mul(R::add(R::sub(b2, b0), (T)1), R::add(R::sub(b3, b1), (T)1));
    const T den = R::sub(R::add(areaA, areaB), inter);
    res = (den == (T)0) ? (T)0 : R::div(inter, den);
  } else {
    const T gx = R::add(R::sub(b2, b0), (T)1), gy = R::add(R::sub(b3, b1), (T)1);
    const T ax = R::add(R::sub(a2, a0), (T)1), ay = R::add(R::sub(a3, a1), (T)1);
    T iw = R::add(R::sub(tmin(a2, b2), tmax(a0, b0)), (T)1);
    if (iw < (T)0) iw = (T)0;
    T ih = R::add(R::sub(tmin(a3, b3), tmax(a1, b1)), (T)1);
    if (ih < (T)0) ih = (T)0;
    const T iwh = R::mul(iw, ih);
    const T ua = R::sub(R::add(R::mul(ax, ay), R::mul(gx, gy)), iwh);
    res = R::div(iwh, ua);
    if (frm_mask) res = R::mul(res, (T)(1 - (int)frm_mask[idx]));
    if (gx == (T)1 && gy == (T)1) res = (T)0;
    if (ax == (T)1 && ay == (T)1) res = (T)-1;
  }
  out[idx] = res;
}

// ---------------------------------------------------------------- greedy matcher (eval_iground.py:85-96); sizes are tens, run serially
__global__ void greedy_match_kernel(double* iou, double* sim, double iou_thr, double sim_thr, int* pairs, int* count, int n, int m) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int c = 0;
  const int cap = n < m ? n : m;          // a match zeroes its row and column: at most min(n, m) pairs carry information
  while (n > 0 && m > 0 && c < cap) {     // (the reference loops forever for thresholds <= 0; here the loop stops before writing pair cap)
    int best = 0;
    for (int k = 1; k < n * m; ++k)
      if (iou[k] > iou[best]) best = k;  // first maximum in row-major order, like np.argmax
    if (iou[best] < iou_thr || sim[best] < sim_thr) break;
    const int bi = best / m, bj = best % m;
    pairs[2 * c] = bi; pairs[2 * c + 1] = bj; ++c;
    for (int j = 0; j < m; ++j) { iou[bi * m + j] = 0; sim[bi * m + j] = 0; }
    for (int i = 0; i < n; ++i) { iou[i * m + bj] = 0; sim[i * m + bj] = 0; }
  }
  *count = c;
}


// ---------------------------------------------------------------- decision utilities of the eval / validation loops
// centre-in-box (eval_youcookinteractions.py:43-48): cx = (x1 + x2) / 2, cy = (y1 + y2) / 2 in Python floats (double);
// correct iff xtl <= cx <= xbr and ytl <= cy <= ybr (INCLUSIVE).  A NaN coordinate fails every comparison (the reference skips such rows).
__global__ void center_in_box_kernel(const double* __restrict__ pred, const double* __restrict__ gt, uint8_t* __restrict__ correct, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double cx = __ddiv_rn(__dadd_rn(pred[4 * i], pred[4 * i + 2]), 2.0), cy = __ddiv_rn(__dadd_rn(pred[4 * i + 1], pred[4 * i + 3]), 2.0);
  const double xtl = gt[4 * i], ytl = gt[4 * i + 1], xbr = gt[4 * i + 2], ybr = gt[4 * i + 3];
  correct[i] = (xtl <= cx && cx <= xbr && ytl <= cy && cy <= ybr) ? 1 : 0;
}

// video IoU (eval_vidstg.py:157-178): per ground-truth frame iou = np_box_iou(pred, gt) if pred.any() else 0; gt_viou = (sum in frame
// order, Python float) / max(n, 1); recall flag per threshold = gt_viou > thr (STRICT).  One thread: n is tens, the order is the contract.
__global__ void viou_kernel(const double* __restrict__ pred, const double* __restrict__ gt, const double* __restrict__ thr, int n, int k,
                            double* __restrict__ ious, double* __restrict__ viou, uint8_t* __restrict__ over) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double acc = 0.0;
  for (int i = 0; i < n; ++i) {
    const double a0 = pred[4 * i], a1 = pred[4 * i + 1], a2 = pred[4 * i + 2], a3 = pred[4 * i + 3];
    double iou = 0.0;
    if (a0 != 0.0 || a1 != 0.0 || a2 != 0.0 || a3 != 0.0) {     // ndarray.any(): NaN counts as non-zero
      const double b0 = gt[4 * i], b1 = gt[4 * i + 1], b2 = gt[4 * i + 2], b3 = gt[4 * i + 3];
      const double area1 = __dmul_rn(__dsub_rn(a2, a0), __dsub_rn(a3, a1)), area2 = __dmul_rn(__dsub_rn(b2, b0), __dsub_rn(b3, b1));
      const double w = tmax(__dsub_rn(tmin(a2, b2), tmax(a0, b0)), 0.0), h = tmax(__dsub_rn(tmin(a3, b3), tmax(a1, b1)), 0.0);
      const double inter = __dmul_rn(w, h);
      iou = __ddiv_rn(inter, __dsub_rn(__dadd_rn(area1, area2), inter));
    }
    ious[i] = iou;
    acc = __dadd_rn(acc, iou);
  }
  const double v = __ddiv_rn(acc, (double)(n > 1 ? n : 1));
  *viou = v;
  for (int j = 0; j < k; ++j) over[j] = v > thr[j] ? 1 : 0;
}

// validation metrics (train.py:826-835): GIoU loss sum of the selected predictions against their ground truth ON THE COORDINATES AS GIVEN
// (the reference feeds cxcywh predictions straight into torchvision's xyxy GIoU -- reproduced, not corrected) and the number of
// instances whose thresholded objectness (sigmoid(logit) > 0.5) equals the integer label.
__device__ __forceinline__ float giou_loss_xyxy(const float* p, const float* q) {
  const float x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3], x1g = q[0], y1g = q[1], x2g = q[2], y2g = q[3];
  const float xk1 = fmaxf(x1, x1g), yk1 = fmaxf(y1, y1g), xk2 = fminf(x2, x2g), yk2 = fminf(y2, y2g);
  float inter = 0.f;
  if (yk2 > yk1 && xk2 > xk1) inter = __fmul_rn(__fsub_rn(xk2, xk1), __fsub_rn(yk2, yk1));
  const float uni = __fsub_rn(__fadd_rn(__fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1)), __fmul_rn(__fsub_rn(x2g, x1g), __fsub_rn(y2g, y1g))), inter);
  const float iou = __fdiv_rn(inter, __fadd_rn(uni, 1e-7f));
  const float xc1 = fminf(x1, x1g), yc1 = fminf(y1, y1g), xc2 = fmaxf(x2, x2g), yc2 = fmaxf(y2, y2g);
  const float area_c = __fmul_rn(__fsub_rn(xc2, xc1), __fsub_rn(yc2, yc1));
  const float miou = __fsub_rn(iou, __fdiv_rn(__fsub_rn(area_c, uni), __fadd_rn(area_c, 1e-7f)));
  return __fsub_rn(1.f, miou);
}

__global__ void __launch_bounds__(256) val_metrics_kernel(const float* __restrict__ boxes, const float* __restrict__ logits, const float* __restrict__ gt,
                                                          const uint8_t* __restrict__ sel, const int* __restrict__ labels, double* __restrict__ giou_sum,
                                                          int* __restrict__ acc, float* __restrict__ giou_each, int B) {
  __shared__ double red[256];
  __shared__ int redi[256];
  double g = 0.0;
  int c = 0;
  for (int i = threadIdx.x; i < B; i += 256) {
    float gi = 0.f;
    if (sel[i]) { gi = giou_loss_xyxy(boxes + 4 * i, gt + 4 * i); g += (double)gi; }
    if (giou_each) giou_each[i] = gi;
    const float sg = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-logits[i])));
    c += ((sg > 0.5f ? 1 : 0) == labels[i]) ? 1 : 0;
  }
  red[threadIdx.x] = g; redi[threadIdx.x] = c;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { red[threadIdx.x] += red[threadIdx.x + s]; redi[threadIdx.x] += redi[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { *giou_sum = red[0]; *acc = redi[0]; }
}

}  // namespace grove
using namespace grove;

extern "C" int grove_box_postprocess(const float* boxes, const float* logits, const float* size_wh, float thr, float* xyxy, uint8_t* keep, int B,
                                     cudaStream_t stream) {
  GROVE_CHECK_ARG(B >= 0);
  if (B == 0) return GROVE_OK;
  GROVE_CHECK_ARG(boxes && logits && size_wh && xyxy && keep);
  box_postprocess_kernel<<<(B + 127) / 128, 128, 0, stream>>>(boxes, logits, size_wh, thr, xyxy, keep, B);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_box_losses_fwd(const float* boxes, const float* logits, const float* gt, const uint8_t* sel, const float* labels, float* sums,
                                    int B, cudaStream_t stream) {
  GROVE_CHECK_ARG(sums && B >= 0);
  GROVE_CHECK_ARG(B == 0 || (boxes && logits && gt && sel && labels));
  box_losses_kernel<<<1, 256, 0, stream>>>(boxes, logits, gt, sel, labels, sums, B);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_box_iou(const void* a, int lda, const void* b, int ldb, const uint8_t* frm_mask, void* out, int n, int m, int mode, int f64,
                             cudaStream_t stream) {
  GROVE_CHECK_ARG(n >= 0 && m >= 0 && mode >= 0 && mode <= 2 && lda >= 4 && ldb >= 4);
  if (n == 0 || m == 0) return GROVE_OK;
  GROVE_CHECK_ARG(a && b && out);
  const int blocks = (n * m + 127) / 128;
  if (f64) box_iou_kernel<double><<<blocks, 128, 0, stream>>>((const double*)a, lda, (const double*)b, ldb, frm_mask, (double*)out, n, m, mode);
  else box_iou_kernel<float><<<blocks, 128, 0, stream>>>((const float*)a, lda, (const float*)b, ldb, frm_mask, (float*)out, n, m, mode);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_greedy_match(double* iou, double* sim, double iou_thr, double sim_thr, int* pairs, int* count, int n, int m,
                                  cudaStream_t stream) {
  GROVE_CHECK_ARG(count && n >= 0 && m >= 0);
  GROVE_CHECK_ARG((n == 0 || m == 0) || (iou && sim && pairs));
  greedy_match_kernel<<<1, 32, 0, stream>>>(iou, sim, iou_thr, sim_thr, pairs, count, n, m);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_center_in_box(const double* pred, const double* gt, uint8_t* correct, int n, cudaStream_t stream) {
  GROVE_CHECK_ARG(n >= 0);
  if (n == 0) return GROVE_OK;
  GROVE_CHECK_ARG(pred && gt && correct);
  center_in_box_kernel<<<(n + 127) / 128, 128, 0, stream>>>(pred, gt, correct, n);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_viou_decisions(const double* pred, const double* gt, const double* thr, int n, int k, double* ious, double* viou,
                                    uint8_t* over, cudaStream_t stream) {
  GROVE_CHECK_ARG(n >= 0 && k >= 0 && viou && (n == 0 || (pred && gt && ious)) && (k == 0 || (thr && over)));
  viou_kernel<<<1, 32, 0, stream>>>(pred, gt, thr, n, k, ious, viou, over);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_val_metrics(const float* boxes, const float* logits, const float* gt, const uint8_t* sel, const int* labels, double* giou_sum,
                                 int* acc, float* giou_each, int B, cudaStream_t stream) {
  GROVE_CHECK_ARG(giou_sum && acc && B >= 0);
  GROVE_CHECK_ARG(B == 0 || (boxes && logits && gt && sel && labels));
  val_metrics_kernel<<<1, 256, 0, stream>>>(boxes, logits, gt, sel, labels, giou_sum, acc, giou_each, B);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}
