// Step epilogue of the distillation step: what the reference does after `criterion(...)` in
// tools/engine.py:58-69 — timm NativeScaler (torch GradScaler: unscale, inf/nan check, skipped step, scale update),
// gradient-norm clipping (`--clip-grad`, timm dispatch_clip_grad mode 'norm' = torch clip_grad_norm_), AdamW
// (timm create_optimizer 'adamw' = torch.optim.AdamW: decoupled weight decay, bias-corrected moments) and timm ModelEma
// (ema = d * ema + (1 - d) * p) — as TWO launches over flat parameter / gradient / state buffers, with no host read-back
// (the skip decision, the clip coefficient, the step count and the loss scale live on the device: graph-capturable).
//
//   1. step_norm_kernel   : per-CTA sum of g^2 (fp64 partials) + non-finite flag                    read g
//   2. step_update_kernel : every CTA folds the partials (fixed order), CTA 0 updates the scaler state;
//                           p, m, v, ema updated in one pass, gradients optionally zeroed             read p g m v ema, write p m v ema (g)
// HBM-bound: 4 B/elt in pass 1 (the gradients then sit in L2 for pass 2 when they fit) and 36-40 B/elt in pass 2.
#include "common.cuh"

namespace dkd {
namespace {

constexpr int kStepThreads = 256;
constexpr int kStepMaxCtas = kNumSMs * 4;

// device-resident optimizer / scaler state (fp32[8] + double partials)
//   st[0] loss scale   st[1] growth tracker (consecutive good steps)   st[2] step count t (good steps)
//   st[3] found_inf of the LAST step (1 = skipped)   st[4] total gradient norm of the last step (unscaled)   st[5] clip coefficient
struct StepParams {
  float* p; float* g; float* m; float* v; float* ema;
  int64_t n, n_decay;            // elements; [0, n_decay) get weight decay, [n_decay, n) do not (biases, norms)
  const float* lr;               // device scalar (schedulers and captured graphs update it in place)
  float beta1, beta2, eps, weight_decay, clip_grad, ema_decay;
  double b1d, b2d;
  float omb1, omb2, omd;         // 1 - beta1, 1 - beta2, 1 - ema_decay rounded from DOUBLE differences (as Python computes them)
  float growth, backoff; int growth_interval; int dynamic_scale;
  int zero_grad;
  float* st;
  double* partials;              // [grid1]
  unsigned int* flag;            // non-finite gradient seen in pass 1
  int grid1;
};

__global__ void __launch_bounds__(kStepThreads) step_norm_kernel(StepParams q) {
  __shared__ double s_w[kStepThreads / 32];
  const int64_t n4 = q.n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(q.g);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  bool bad = false;
  const int64_t stride = (int64_t)gridDim.x * kStepThreads;
  for (int64_t i = (int64_t)blockIdx.x * kStepThreads + threadIdx.x; i < n4; i += stride) {
    const float4 x = __ldg(g4 + i);
    a0 = fmaf(x.x, x.x, a0); a1 = fmaf(x.y, x.y, a1); a2 = fmaf(x.z, x.z, a2); a3 = fmaf(x.w, x.w, a3);
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * kStepThreads + threadIdx.x; i < q.n; i += stride) a0 = fmaf(q.g[i], q.g[i], a0);
  const float acc = (a0 + a1) + (a2 + a3);
  bad = !(fabsf(acc) <= 3.0e38f);   // inf or nan anywhere in this thread's elements shows up in its sum of squares
  double d = warp_sum((double)acc);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = d;
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(q.flag, 1u);
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kStepThreads / 32; ++w) t += s_w[w];
    q.partials[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(kStepThreads) step_update_kernel(StepParams q) {
  __shared__ double s_w[kStepThreads / 32];
  __shared__ float s_coef, s_b1c, s_b2c;
  __shared__ int s_skip;
  // every CTA folds the pass-1 partials in the same fixed order (a few KB from L2): no third launch, no grid barrier
  {
    double a = 0.0;
    for (int i = threadIdx.x; i < q.grid1; i += kStepThreads) a += __ldcg(q.partials + i);
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < kStepThreads / 32; ++w) t += s_w[w];
      const float scale = q.st[0];
      const float inv_scale = 1.f / scale;
      const bool inf = (__ldcg(q.flag) != 0u) || !(t <= 1.0e300);
      const float norm = inf ? INFINITY : (float)(sqrt(t) * (double)inv_scale);         // norm of the UNSCALED gradients
      // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
      float coef = inv_scale;
      if (q.clip_grad > 0.f) coef *= fminf(q.clip_grad / (norm + 1e-6f), 1.f);
      const float t_new = q.st[2] + 1.f;
      s_coef = coef;
      s_skip = inf ? 1 : 0;
      s_b1c = (float)(1.0 - pow((double)q.b1d, (double)t_new));
      s_b2c = (float)(1.0 - pow((double)q.b2d, (double)t_new));
    }
    __syncthreads();
  }
  const bool skip = s_skip != 0;
  const float coef = s_coef;
  const float lr = __ldg(q.lr);
  const float step_size = lr / s_b1c;
  const float inv_sqrt_b2c = rsqrtf(s_b2c);
  const bool has_ema = q.ema != nullptr;
  const int64_t n4 = q.n >> 2;
  const int64_t stride = (int64_t)gridDim.x * kStepThreads;
  float4* p4 = reinterpret_cast<float4*>(q.p);
  float4* g4 = reinterpret_cast<float4*>(q.g);
  float4* m4 = reinterpret_cast<float4*>(q.m);
  float4* v4 = reinterpret_cast<float4*>(q.v);
  float4* e4 = reinterpret_cast<float4*>(q.ema);
  auto upd = [&](float& p, float gs, float& m, float& v, float& e, bool decay) {
    const float g = gs * coef;
    if (decay) p *= 1.f - lr * q.weight_decay;            // decoupled weight decay (torch.optim.AdamW)
    m = fmaf(q.beta1, m, q.omb1 * g);
    v = fmaf(q.beta2, v, q.omb2 * g * g);
    const float denom = sqrtf(v) * inv_sqrt_b2c + q.eps;
    p -= step_size * (m / denom);
    if (has_ema) e = fmaf(q.ema_decay, e, q.omd * p);   // timm ModelEma: d * ema + (1 - d) * model
  };
  if (!skip) {
    for (int64_t i = (int64_t)blockIdx.x * kStepThreads + threadIdx.x; i < n4; i += stride) {
      float4 p = p4[i], m = m4[i], v = v4[i];
      const float4 g = g4[i];
      float4 e = has_ema ? e4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      const int64_t e0 = i << 2;
      upd(p.x, g.x, m.x, v.x, e.x, e0 + 0 < q.n_decay);
      upd(p.y, g.y, m.y, v.y, e.y, e0 + 1 < q.n_decay);
      upd(p.z, g.z, m.z, v.z, e.z, e0 + 2 < q.n_decay);
      upd(p.w, g.w, m.w, v.w, e.w, e0 + 3 < q.n_decay);
      p4[i] = p; m4[i] = m; v4[i] = v;
      if (has_ema) e4[i] = e;
      if (q.zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * kStepThreads + threadIdx.x; i < q.n; i += stride) {
      float e = has_ema ? q.ema[i] : 0.f;
      upd(q.p[i], q.g[i], q.m[i], q.v[i], e, i < q.n_decay);
      if (has_ema) q.ema[i] = e;
      if (q.zero_grad) q.g[i] = 0.f;
    }
  } else if (q.zero_grad) {   // skipped step (GradScaler): parameters and moments untouched, gradients dropped
    for (int64_t i = (int64_t)blockIdx.x * kStepThreads + threadIdx.x; i < n4; i += stride) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * kStepThreads + threadIdx.x; i < q.n; i += stride) q.g[i] = 0.f;
  }
  // scaler state / counters: written by the LAST CTA to finish (arrival counter in flag[1]) — by then every CTA has read
  // st[0], st[2] and the non-finite flag, so updating them cannot change another CTA's decision.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(q.flag + 1, 1u);
    if (done == gridDim.x - 1) {
      float scale = q.st[0], tracker = q.st[1];
      if (q.dynamic_scale) {   // torch.amp.GradScaler.update
        if (skip) { scale *= q.backoff; tracker = 0.f; }
        else if (tracker + 1.f >= (float)q.growth_interval) { scale *= q.growth; tracker = 0.f; }
        else tracker += 1.f;
      }
      double t = 0.0;
      for (int i = 0; i < q.grid1; ++i) t += __ldcg(q.partials + i);
      q.st[4] = skip ? INFINITY : (float)(sqrt(t) / (double)q.st[0]);
      q.st[5] = coef * q.st[0];
      q.st[3] = skip ? 1.f : 0.f;
      if (!skip) q.st[2] += 1.f;
      q.st[0] = scale;
      q.st[1] = tracker;
      q.flag[0] = 0u;
      q.flag[1] = 0u;
    }
  }
}

// top-k hits of a batch of logits (timm.utils.accuracy, engine.py:53-56): hits[j] += [rank of the target logit < k_j]
template <typename T>
__global__ void __launch_bounds__(128) topk_hits_kernel(const T* __restrict__ z, const int64_t* __restrict__ target, int64_t B, int64_t C,
                                                        int k0, int k1, float* __restrict__ hits) {
  __shared__ int s_cnt[4];
  const int64_t row = blockIdx.x;
  const T* zr = z + row * C;
  const int64_t tg = target[row];
  const float zt = Elt<T>::ld(zr + tg);
  int cnt = 0;
  // rank of the target among the row: values greater than it, or equal with a lower index (torch.topk order)
  for (int64_t c = threadIdx.x; c < C; c += 128) {
    const float x = Elt<T>::ld(zr + c);
    cnt += (x > zt || (x == zt && c < tg)) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int r = s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
    if (r < k0) atomicAdd(hits + 0, 1.f);
    if (r < k1) atomicAdd(hits + 1, 1.f);
  }
}

// Mixup / CutMix of a batch with its reverse (timm.data.Mixup, mode 'batch'; tools/train.py:288-295):
//   mixup : x[b] = lam * x[b] + (1 - lam) * x[B-1-b]
//   cutmix: x[b][:, y0:y1, x0:x1] = x[B-1-b][:, y0:y1, x0:x1]
// in place, pairs (b, B-1-b) handled together so that each element is read once and written once.
__global__ void __launch_bounds__(256) mix_batch_kernel(float* __restrict__ x, int64_t B, int64_t CH, int H, int W, const float* __restrict__ lam_p,
                                                        int use_cutmix, int y0, int y1, int x0, int x1) {
  const float lam = __ldg(lam_p);
  const int64_t per = CH * H * W;
  const int64_t half = B / 2;
  const int64_t total = half * per;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t b = i / per, r = i - b * per;
    float* pa = x + b * per + r;
    float* pb = x + (B - 1 - b) * per + r;
    const float a = *pa, c = *pb;
    if (use_cutmix) {
      const int w = (int)(r % W), h = (int)((r / W) % H);
      if (h >= y0 && h < y1 && w >= x0 && w < x1) { *pa = c; *pb = a; }
    } else {
      *pa = lam * a + (1.f - lam) * c;
      *pb = lam * c + (1.f - lam) * a;
    }
  }
  // odd batch: the middle sample mixes with itself (unchanged)
}

}  // namespace
}  // namespace dkd

extern "C" {

size_t dkd_step_workspace_bytes(void) { return (size_t)dkd::kStepMaxCtas * sizeof(double) + 256; }

int dkd_step_epilogue(float* params, float* grads, float* exp_avg, float* exp_avg_sq, float* ema, int64_t n, int64_t n_decay,
                      const float* lr, double beta1, double beta2, float eps, float weight_decay, float clip_grad, double ema_decay,
                      int dynamic_scale, float growth_factor, float backoff_factor, int growth_interval, int zero_grad, float* state,
                      void* workspace, size_t workspace_bytes, dkd_stream_t stream) {
  using namespace dkd;
  int rc = dkd_check_device();
  if (rc != DKD_OK) return rc;
  const char* fn = "dkd_step_epilogue";
  DKD_REQUIRE(params && grads && exp_avg && exp_avg_sq && lr && state && workspace, DKD_E_SHAPE, "%s: null pointer", fn);
  DKD_REQUIRE(n > 0 && n_decay >= 0 && n_decay <= n, DKD_E_SHAPE, "%s: bad sizes n=%lld n_decay=%lld", fn, (long long)n, (long long)n_decay);
  DKD_REQUIRE(workspace_bytes >= dkd_step_workspace_bytes(), DKD_E_WORKSPACE, "%s: workspace too small", fn);
  DKD_REQUIRE(((((uintptr_t)params) | ((uintptr_t)grads) | ((uintptr_t)exp_avg) | ((uintptr_t)exp_avg_sq) | ((uintptr_t)ema) | ((uintptr_t)workspace)) & 15) == 0,
              DKD_E_ALIGN, "%s: buffers must be 16-byte aligned", fn);
  DKD_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps > 0.f && ema_decay >= 0.0 && ema_decay <= 1.0, DKD_E_SHAPE,
              "%s: bad hyper-parameters", fn);
  StepParams q;
  q.p = params; q.g = grads; q.m = exp_avg; q.v = exp_avg_sq; q.ema = ema; q.n = n; q.n_decay = n_decay; q.lr = lr;
  q.beta1 = (float)beta1; q.beta2 = (float)beta2; q.eps = eps; q.weight_decay = weight_decay; q.clip_grad = clip_grad; q.ema_decay = (float)ema_decay;
  q.b1d = beta1; q.b2d = beta2; q.omb1 = (float)(1.0 - beta1); q.omb2 = (float)(1.0 - beta2); q.omd = (float)(1.0 - ema_decay);
  q.growth = growth_factor; q.backoff = backoff_factor; q.growth_interval = growth_interval; q.dynamic_scale = dynamic_scale;
  q.zero_grad = zero_grad; q.st = state;
  q.flag = reinterpret_cast<unsigned int*>(workspace);                       // [0] non-finite, [1] arrival counter (zero on entry, reset on exit)
  q.partials = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 256);
  int64_t ctas = (n / 4 + kStepThreads - 1) / kStepThreads;
  if (ctas < 1) ctas = 1;
  if (ctas > kStepMaxCtas) ctas = kStepMaxCtas;
  q.grid1 = (int)ctas;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  step_norm_kernel<<<(unsigned)ctas, kStepThreads, 0, st>>>(q);
  rc = check_launch("dkd_step_epilogue: norm");
  if (rc != DKD_OK) return rc;
  step_update_kernel<<<(unsigned)ctas, kStepThreads, 0, st>>>(q);
  return check_launch("dkd_step_epilogue: update");
}

int dkd_topk_hits(const void* logits, const int64_t* target, int64_t B, int64_t C, int dtype, int k0, int k1, float* hits, dkd_stream_t stream) {
  using namespace dkd;
  int rc = dkd_check_device();
  if (rc != DKD_OK) return rc;
  DKD_REQUIRE(logits && target && hits && B > 0 && C > 0 && B < (1ll << 31), DKD_E_SHAPE, "dkd_topk_hits: bad arguments");
  DKD_REQUIRE(dtype == DKD_F32 || dtype == DKD_BF16, DKD_E_DTYPE, "dkd_topk_hits: dtype %d", dtype);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaMemsetAsync(hits, 0, 2 * sizeof(float), st);
  if (dtype == DKD_F32) topk_hits_kernel<float><<<(unsigned)B, 128, 0, st>>>(reinterpret_cast<const float*>(logits), target, B, C, k0, k1, hits);
  else topk_hits_kernel<__nv_bfloat16><<<(unsigned)B, 128, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(logits), target, B, C, k0, k1, hits);
  return check_launch("dkd_topk_hits");
}

int dkd_mix_batch(float* x, int64_t B, int64_t CH, int H, int W, const float* lam, int use_cutmix, int y0, int y1, int x0, int x1,
                  dkd_stream_t stream) {
  using namespace dkd;
  int rc = dkd_check_device();
  if (rc != DKD_OK) return rc;
  DKD_REQUIRE(x && lam && B > 0 && CH > 0 && H > 0 && W > 0, DKD_E_SHAPE, "dkd_mix_batch: bad arguments");
  const int64_t total = (B / 2) * CH * H * W;
  if (total == 0) return DKD_OK;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  mix_batch_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, B, CH, H, W, lam, use_cutmix, y0, y1, x0, x1);
  return check_launch("dkd_mix_batch");
}

}  // extern "C"
