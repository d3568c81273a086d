// estk_eval_mlp.cu -- kernel 1 of the ES generation (fp32 CUDA-core path):
// population evaluate for MLP policies over a synthetic observation batch.
//
// Replaces (reference file:line, /root/reference):
//   ES._sample_policy        estorch/estorch.py:187-193   theta +- sigma*eps, never materialised
//   ES._calculate_returns    estorch/estorch.py:195-202   vector_to_parameters + rollout per row
//   Policy.forward           examples/cartpole_es.py:14-20 (Linear-ReLU-Linear-ReLU-Linear)
//   Agent.rollout            synthetic agent of SURVEY 8d: -mean((policy(obs)-y)^2)
//                            (+ behaviour characteristic, examples/nsra_es.py:45-49)
//
// One CTA = one antithetic pair x one chunk of BC observations.  The pair's
// noise row is read ONCE and serves both signs: each weight tile is formed in
// shared memory as W+ = theta + sigma*t and W- = theta - sigma*t (same two
// roundings as the reference: eps = sigma*t, then theta +- eps) and multiplied
// into the + and - activation rows.  Activations stay in shared memory across
// layers (k-major [width][ROWS], ROWS = 2*BC rows = sign-major), register tile
// 4 rows x 4 outputs per thread.  The last layer is fused with the squared
// error; chunk partial sums are combined in fixed order by the last-arriving
// chunk CTA (deterministic, no float atomics).
//
// This is the exact-fp32 path used for parity and for small policies; algorithmic
// bytes per launch = 4*n*pairs (noise rows) + 4*n (theta) + 4*B*(in+out) + 4*P.
#include "estk_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int KT = 32;  // k-chunk of the weight tile

struct EvalParams {
  estk_mlp_desc desc;
  const float* theta;
  const float* table;
  const int64_t* offsets;  // null => centre evaluation (sigma ignored)
  const int32_t* order;
  int pairs;
  float sigma;
  const float* obs;
  const float* target;
  int B, BC, chunks, maxw;
  float* ret_plus;
  float* ret_minus;
  float* bc_plus;
  float* bc_minus;
  int bc_obs, bc_dim;
  float* partial;          // [pairs][2][chunks]
  unsigned int* counters;  // [pairs], zero on entry, zero on exit
};

template <int ROWS>
__global__ void __launch_bounds__(kThreads) eval_mlp_kernel(const EvalParams p) {
  constexpr int OT = 4096 / ROWS;  // output features per tile
  constexpr int OTP = OT + 4;      // padded row of the weight tile (keeps float4 alignment)
  constexpr int TC = OT / 4;       // thread columns
  constexpr int BC = ROWS / 2;
  extern __shared__ __align__(16) float smem[];
  float* X = smem;                             // [maxw][ROWS]
  float* Y = X + (size_t)p.maxw * ROWS;        // [maxw][ROWS]
  float* Wp = Y + (size_t)p.maxw * ROWS;       // [KT][OTP]
  float* Wm = Wp + KT * OTP;                   // [KT][OTP]
  __shared__ float s_red[2][kThreads / 32];
  __shared__ bool s_last;

  const int tid = threadIdx.x;
  const int slot = blockIdx.x / p.chunks;
  const int chunk = blockIdx.x % p.chunks;
  const int j = p.order ? p.order[slot] : slot;
  const bool centre = (p.offsets == nullptr);
  const float* trow = centre ? p.theta : p.table + p.offsets[j];
  const float sigma = centre ? 0.f : p.sigma;
  const int b0 = chunk * BC;
  const int L = p.desc.n_layers;

  // ---- stage the observation chunk for both signs: X[k][s*BC + b] = obs[b0+b][k]
  {
    const int in0 = p.desc.dims[0];
    for (int idx = tid; idx < BC * in0; idx += kThreads) {
      const int b = idx / in0, k = idx % in0;
      const float x = (b0 + b < p.B) ? __ldg(p.obs + (size_t)(b0 + b) * in0 + k) : 0.f;
      X[k * ROWS + b] = x;
      X[k * ROWS + BC + b] = x;
    }
  }

  const int tc = tid % TC;
  const int tr = tid / TC;
  const int r0 = tr * 4;                 // first of this thread's 4 rows
  const bool minus = r0 >= BC;           // all 4 rows share the sign
  const float* Wsel = minus ? Wm : Wp;
  float loss = 0.f;
  int64_t pbase = 0;

  for (int l = 0; l < L; ++l) {
    const int in = p.desc.dims[l], out = p.desc.dims[l + 1];
    const int64_t wbase = pbase, bbase = pbase + (int64_t)in * out;
    const bool last = (l == L - 1);
    for (int o0 = 0; o0 < out; o0 += OT) {
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
      for (int k0 = 0; k0 < in; k0 += KT) {
        __syncthreads();  // previous tile fully consumed (and X staged on first pass)
        // weight tile: element (o, k) <- theta/noise[wbase + (o0+o)*in + k0+k]
        for (int e = tid; e < OT * KT; e += kThreads) {
          const int k = e % KT, o = e / KT;
          float wp = 0.f, wm = 0.f;
          if (o0 + o < out && k0 + k < in) {
            const int64_t idx = wbase + (int64_t)(o0 + o) * in + k0 + k;
            const float th = __ldg(p.theta + idx);
            const float ep = __fmul_rn(sigma, ld_noise1(trow + idx));
            wp = __fadd_rn(th, ep);
            wm = __fsub_rn(th, ep);
          }
          Wp[k * OTP + o] = wp;
          Wm[k * OTP + o] = wm;
        }
        __syncthreads();
        const int kmax = min(KT, in - k0);
#pragma unroll 4
        for (int k = 0; k < kmax; ++k) {
          const float4 x = *reinterpret_cast<const float4*>(X + (size_t)(k0 + k) * ROWS + r0);
          const float4 w = *reinterpret_cast<const float4*>(Wsel + k * OTP + tc * 4);
          acc[0][0] = fmaf(x.x, w.x, acc[0][0]); acc[0][1] = fmaf(x.x, w.y, acc[0][1]);
          acc[0][2] = fmaf(x.x, w.z, acc[0][2]); acc[0][3] = fmaf(x.x, w.w, acc[0][3]);
          acc[1][0] = fmaf(x.y, w.x, acc[1][0]); acc[1][1] = fmaf(x.y, w.y, acc[1][1]);
          acc[1][2] = fmaf(x.y, w.z, acc[1][2]); acc[1][3] = fmaf(x.y, w.w, acc[1][3]);
          acc[2][0] = fmaf(x.z, w.x, acc[2][0]); acc[2][1] = fmaf(x.z, w.y, acc[2][1]);
          acc[2][2] = fmaf(x.z, w.z, acc[2][2]); acc[2][3] = fmaf(x.z, w.w, acc[2][3]);
          acc[3][0] = fmaf(x.w, w.x, acc[3][0]); acc[3][1] = fmaf(x.w, w.y, acc[3][1]);
          acc[3][2] = fmaf(x.w, w.z, acc[3][2]); acc[3][3] = fmaf(x.w, w.w, acc[3][3]);
        }
      }
      // ---- tile epilogue: bias (+-), ReLU or loss
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int o = o0 + tc * 4 + c;
        if (o >= out) continue;
        const float th = __ldg(p.theta + bbase + o);
        const float ep = __fmul_rn(sigma, ld_noise1(trow + bbase + o));
        const float bias = minus ? __fsub_rn(th, ep) : __fadd_rn(th, ep);
        float y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = acc[i][c] + bias;
        if (!last) {
#pragma unroll
          for (int i = 0; i < 4; ++i) y[i] = fmaxf(y[i], 0.f);
          *reinterpret_cast<float4*>(Y + (size_t)o * ROWS + r0) = make_float4(y[0], y[1], y[2], y[3]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int b = b0 + ((r0 + i) % BC);
            if (b < p.B) {
              const float d = y[i] - __ldg(p.target + (size_t)b * out + o);
              loss = fmaf(d, d, loss);
              float* bc = minus ? p.bc_minus : p.bc_plus;
              const int64_t e = (int64_t)b * out + o;
              if (bc && b < p.bc_obs && e < p.bc_dim) bc[(size_t)j * p.bc_dim + e] = y[i];
            }
          }
        }
      }
    }
    // next layer reads what this one wrote (the __syncthreads at the top of the
    // next k-loop orders the Y writes before the X reads)
    float* t = X; X = Y; Y = t;
    pbase = bbase + out;
  }

  // ---- block reduction of the squared error, per sign
  {
    const float lp = warp_sum_f(minus ? 0.f : loss);
    const float lm = warp_sum_f(minus ? loss : 0.f);
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = lp; s_red[1][tid >> 5] = lm; }
    __syncthreads();
    if (tid == 0) {
      float sp = 0.f, sm = 0.f;
      for (int w = 0; w < kThreads / 32; ++w) { sp += s_red[0][w]; sm += s_red[1][w]; }
      float* part = p.partial + ((size_t)slot * 2) * p.chunks;
      part[chunk] = sp;
      part[p.chunks + chunk] = sm;
      __threadfence();
      const unsigned int arrived = atomicAdd(p.counters + slot, 1u);
      s_last = (arrived == (unsigned int)p.chunks - 1);
    }
    __syncthreads();
    if (s_last && tid == 0) {
      __threadfence();
      const float* part = p.partial + ((size_t)slot * 2) * p.chunks;
      float sp = 0.f, sm = 0.f;
      for (int c = 0; c < p.chunks; ++c) { sp += __ldcg(part + c); sm += __ldcg(part + p.chunks + c); }
      const float denom = (float)p.B * (float)p.desc.dims[L];
      p.ret_plus[j] = -(sp / denom);
      if (p.ret_minus) p.ret_minus[j] = -(sm / denom);
      p.counters[slot] = 0u;  // ready for the next launch
    }
  }
}

size_t smem_bytes(int rows, int maxw) {
  const int ot = 4096 / rows;
  return sizeof(float) * ((size_t)2 * maxw * rows + (size_t)2 * KT * (ot + 4));
}

template <int ROWS>
int launch(const EvalParams& p, size_t smem, cudaStream_t stream) {
  ESTK_CUDA(cudaFuncSetAttribute(eval_mlp_kernel<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  eval_mlp_kernel<ROWS><<<p.pairs * p.chunks, kThreads, smem, stream>>>(p);
  ESTK_CUDA(cudaGetLastError());
  return ESTK_OK;
}

int run_eval(estk_ctx* ctx, EvalParams& p, cudaStream_t stream, const char* who) {
  const estk_mlp_desc& d = p.desc;
  ESTK_CHECK_ARG(d.n_layers >= 1 && d.n_layers <= ESTK_MAX_LAYERS, "%s: n_layers=%d", who, d.n_layers);
  ESTK_CHECK_ARG(d.activation == 0, "%s: only ReLU (activation=0) is implemented", who);
  int maxw = 0;
  for (int l = 0; l <= d.n_layers; ++l) {
    ESTK_CHECK_ARG(d.dims[l] >= 1, "%s: dims[%d]=%d", who, l, d.dims[l]);
    if (d.dims[l] > maxw) maxw = d.dims[l];
  }
  ESTK_CHECK_ARG(p.B >= 1, "%s: B must be positive", who);
  ESTK_CHECK_ARG(p.pairs >= 1 && p.pairs <= ESTK_MAX_POPULATION / 2, "%s: pairs=%d", who, p.pairs);
  p.maxw = maxw;
  // largest observation chunk whose activations fit in shared memory
  // (prefer <= 100 KB so two CTAs share an SM; wide layers may take up to 200 KB)
  size_t budget = 100 * 1024;
  if (smem_bytes(32, maxw) > budget) budget = 200 * 1024;
  int rows = 256;
  while (rows > 32 && (smem_bytes(rows, maxw) > budget || rows / 2 >= 2 * p.B)) rows >>= 1;
  if (smem_bytes(rows, maxw) > budget) {
    estk_set_error("%s: layer width %d does not fit the fp32 shared-memory path", who, maxw);
    return ESTK_ERR_UNSUPPORTED;
  }
  p.BC = rows / 2;
  p.chunks = (p.B + p.BC - 1) / p.BC;
  ESTK_CHECK_ARG(p.chunks <= kEvalMaxChunks, "%s: B=%d needs %d chunks > %d", who, p.B, p.chunks, kEvalMaxChunks);
  p.partial = ctx->eval_partial;
  p.counters = ctx->counters;
  const size_t smem = smem_bytes(rows, maxw);
  switch (rows) {
    case 256: return launch<256>(p, smem, stream);
    case 128: return launch<128>(p, smem, stream);
    case 64: return launch<64>(p, smem, stream);
    default: return launch<32>(p, smem, stream);
  }
}

}  // namespace

extern "C" int estk_eval_mlp(estk_ctx* ctx, const estk_mlp_desc* desc, const float* theta,
                             const float* table, const int64_t* offsets, const int32_t* order,
                             int32_t pairs, float sigma, const float* obs, const float* target,
                             int32_t B, float* returns_plus, float* returns_minus, float* bc_plus,
                             float* bc_minus, int32_t bc_obs, int32_t bc_dim, void* stream) {
  ESTK_CHECK_ARG(ctx && desc && theta && table && offsets && obs && target && returns_plus && returns_minus,
                 "estk_eval_mlp: null argument");
  ESTK_CHECK_ARG((bc_plus == nullptr) == (bc_minus == nullptr), "estk_eval_mlp: bc_plus/bc_minus must both be set or both null");
  ESTK_CHECK_ARG(!bc_plus || (bc_obs > 0 && bc_dim > 0), "estk_eval_mlp: bc_obs/bc_dim must be positive with bc outputs");
  EvalParams p = {};
  p.desc = *desc; p.theta = theta; p.table = table; p.offsets = offsets; p.order = order;
  p.pairs = pairs; p.sigma = sigma; p.obs = obs; p.target = target; p.B = B;
  p.ret_plus = returns_plus; p.ret_minus = returns_minus;
  p.bc_plus = bc_plus; p.bc_minus = bc_minus; p.bc_obs = bc_obs; p.bc_dim = bc_dim;
  return run_eval(ctx, p, (cudaStream_t)stream, "estk_eval_mlp");
}

extern "C" int estk_eval_mlp_center(estk_ctx* ctx, const estk_mlp_desc* desc, const float* theta,
                                    const float* obs, const float* target, int32_t B,
                                    float* return_out, float* bc_out, int32_t bc_obs,
                                    int32_t bc_dim, void* stream) {
  ESTK_CHECK_ARG(ctx && desc && theta && obs && target && return_out, "estk_eval_mlp_center: null argument");
  ESTK_CHECK_ARG(!bc_out || (bc_obs > 0 && bc_dim > 0), "estk_eval_mlp_center: bc_obs/bc_dim must be positive with bc_out");
  EvalParams p = {};
  p.desc = *desc; p.theta = theta; p.table = theta; p.offsets = nullptr; p.order = nullptr;
  p.pairs = 1; p.sigma = 0.f; p.obs = obs; p.target = target; p.B = B;
  p.ret_plus = return_out; p.ret_minus = nullptr;
  p.bc_plus = bc_out; p.bc_minus = nullptr; p.bc_obs = bc_obs; p.bc_dim = bc_dim;
  return run_eval(ctx, p, (cudaStream_t)stream, "estk_eval_mlp_center");
}
