// Per-element optimizer math shared by the flat fused optimizer (elementwise.cu) and the fused
// parameter-server kernels (collectives.cu: reduce + optimizer + pull in one launch).
//
// Reference: MX/src/operator/optimizer_op-inl.h (SGD/MP_SGD :86-427, Adam :840-875, RMSProp, Ftrl, Signum
// registered in optimizer_op.cu:225-262) and the Python definitions in optimizer.py:450-1502.  Master
// weights and states are always fp32 (the multi-precision layout); g is the raw summed gradient.
//
//   kind  name      s1      s2     extra parameter slots
//   0     sgd       -       -
//   1     sgd_mom   mom     -      momentum
//   2     adam      m       v      beta1, beta2, eps   (bias correction folded into lr by the caller)
//   3     test      -       -      w += rescale*g
//   4     none      -       -      w  = rescale*g      (kvstore without updater: store the reduced gradient)
//   5     nag       mom     -      momentum
//   6     signum    mom     -      momentum, beta1 = wd_lh
//   7     signsgd   -       -      beta1 = wd_lh
//   8     adagrad   hist    -      eps
//   9     rmsprop   n       -      beta1 = gamma1, beta2 = clip_weights (<=0: off), eps
//   10    adadelta  acc_g   acc_d  beta1 = rho, eps
//   11    ftrl      z       n      beta1 = lamda1, beta2 = beta
//   12    adamax    m       u      beta1, beta2        (lr pre-divided by 1-beta1^t by the caller)
#pragma once
#include <cuda_runtime.h>

namespace dlcfn {

struct OptArgs {
  int kind;
  float lr, wd, momentum, beta1, beta2, eps, rescale, clip;  // clip<=0: off
};

__host__ __device__ __forceinline__ bool opt_uses_s1(int kind) {
  return kind == 1 || kind == 2 || kind == 5 || kind == 6 || (kind >= 8 && kind <= 12);
}
__host__ __device__ __forceinline__ bool opt_uses_s2(int kind) { return kind == 2 || (kind >= 10 && kind <= 12); }

__device__ __forceinline__ float opt_sign(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

__device__ __forceinline__ void opt_apply(const OptArgs& o, float g, float& w, float& s1, float& s2) {
  g *= o.rescale;
  if (o.clip > 0.f) g = fminf(fmaxf(g, -o.clip), o.clip);
  switch (o.kind) {
    case 0:
      w = (1.f - o.lr * o.wd) * w - o.lr * g;
      break;
    case 1:
      s1 = o.momentum * s1 - o.lr * o.wd * w - o.lr * g;
      w += s1;
      break;
    case 2:
      g += o.wd * w;
      s1 = o.beta1 * s1 + (1.f - o.beta1) * g;
      s2 = o.beta2 * s2 + (1.f - o.beta2) * g * g;
      w -= o.lr * s1 / (sqrtf(s2) + o.eps);
      break;
    case 3:
      w += g;
      break;
    case 4:
      w = g;
      break;
    case 5:  // Nesterov momentum
      g += o.wd * w;
      s1 = o.momentum * s1 + g;
      w -= o.lr * (g + o.momentum * s1);
      break;
    case 6:  // Signum
      g += o.wd * w;
      s1 = o.momentum * s1 - (1.f - o.momentum) * g;
      w = (1.f - o.lr * o.beta1) * w + o.lr * opt_sign(s1);
      break;
    case 7:  // signSGD
      g += o.wd * w;
      w = (1.f - o.lr * o.beta1) * w - o.lr * opt_sign(g);
      break;
    case 8:  // AdaGrad
      s1 += g * g;
      w -= o.lr * (g / sqrtf(s1 + o.eps) + o.wd * w);
      break;
    case 9:  // RMSProp (non-centred)
      g += o.wd * w;
      s1 = o.beta1 * s1 + (1.f - o.beta1) * g * g;
      w -= o.lr * g / sqrtf(s1 + o.eps);
      if (o.beta2 > 0.f) w = fminf(fmaxf(w, -o.beta2), o.beta2);
      break;
    case 10: {  // AdaDelta
      s1 = o.beta1 * s1 + (1.f - o.beta1) * g * g;
      const float d = sqrtf(s2 + o.eps) / sqrtf(s1 + o.eps) * g;
      s2 = o.beta1 * s2 + (1.f - o.beta1) * d * d;
      w = w - d - o.wd * w;
      break;
    }
    case 11: {  // FTRL-proximal
      s1 += g - (sqrtf(s2 + g * g) - sqrtf(s2)) * w / o.lr;
      s2 += g * g;
      w = fabsf(s1) > o.beta1 ? (opt_sign(s1) * o.beta1 - s1) / ((o.beta2 + sqrtf(s2)) / o.lr + o.wd) : 0.f;
      break;
    }
    default:  // 12: Adamax
      g += o.wd * w;
      s1 = o.beta1 * s1 + (1.f - o.beta1) * g;
      s2 = fmaxf(o.beta2 * s2, fabsf(g));
      w -= o.lr * s1 / (s2 + 1e-38f);
      break;
  }
}

}  // namespace dlcfn
