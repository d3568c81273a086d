/*
 * estk.h -- C ABI of the B200 Evolution-Strategies kernel library (libestk.so).
 *
 * This is the drop-in boundary for the ES generation hot path of
 * goktug97/estorch (estorch/estorch.py:211-250, ES._master).  The reference has
 * no FFI layer of its own (pure Python); each entry point below names the
 * reference function(s) whose arithmetic it replaces.  The only caller is the
 * host-side mirror of the reference classes (estorch_b200/estorch.py) through
 * ctypes; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - plain C: no C++ types, no exceptions, no torch types in any signature;
 *   - every function returns ESTK_OK (0) or a negative estk_status; the text of
 *     the last failure on the calling thread is estk_last_error();
 *   - all buffers are CALLER-OWNED DEVICE pointers (fp32 / int32 / int64,
 *     contiguous, 16-byte aligned) unless a parameter says "host";
 *   - every launch is asynchronous on the caller's stream (`stream` is a
 *     cudaStream_t passed as void*); the library never synchronises;
 *   - the library keeps no global mutable state: one estk_ctx per device
 *     holds an opaque workspace (partial sums, centred-rank scratch).
 *
 * Member / pair layout (estorch.py:190-193): population_size P = 2*pairs;
 * member j < pairs is theta + sigma*T[off_j : off_j+n], member j+pairs is
 * theta - sigma*T[off_j : off_j+n].  T is the shared unit-normal noise table.
 */
#ifndef ESTK_H_
#define ESTK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ESTK_API __attribute__((visibility("default")))
#else
#define ESTK_API
#endif

#define ESTK_VERSION 100          /* 0.1.0 */
#define ESTK_MAX_LAYERS 8
#define ESTK_MAX_POPULATION 32768 /* P; rank phase is O(P^2) */

typedef enum {
  ESTK_OK = 0,
  ESTK_ERR_INVALID = -1,      /* bad argument (shape, alignment, null) */
  ESTK_ERR_CUDA = -2,         /* a CUDA runtime call failed */
  ESTK_ERR_UNSUPPORTED = -3,  /* valid request this build cannot serve */
  ESTK_ERR_NOMEM = -4
} estk_status;

typedef struct estk_ctx estk_ctx;

/* Device-resident per-run state (caller-owned, 32 bytes, zero-initialised by
 * the caller except best_reward = -inf).  Kept on the device so that a whole
 * generation can be replayed from a CUDA graph with no host-side scalars. */
typedef struct {
  int64_t generation;   /* estorch.py:248 `self.step`; read by estk_make_offsets,
                           advanced by estk_track_best */
  int64_t adam_step;    /* torch Adam `state['step']`; advanced by the Adam epilogue */
  float episode_reward; /* estorch.py:182 */
  float best_reward;    /* estorch.py:183-184 */
  int32_t improved;     /* 1 when the last estk_track_best took a new best */
  int32_t reserved;
} estk_state;

/* Policy description: Linear -> act -> ... -> Linear over a flat parameter
 * vector in torch.nn.utils.parameters_to_vector order (weight [out,in]
 * row-major, then bias, per layer) -- examples/cartpole_es.py:6-20. */
typedef struct {
  int32_t n_layers;                  /* number of Linear layers, 1..ESTK_MAX_LAYERS */
  int32_t dims[ESTK_MAX_LAYERS + 1]; /* dims[0] = obs dim, dims[n_layers] = out dim */
  int32_t activation;                /* 0 = ReLU between layers (none after the last) */
} estk_mlp_desc;

/* torch.optim.Adam hyper-parameters (torch/optim/adam.py:457-546; the
 * optimizer every reference example uses, examples/cartpole_es.py:48-50). */
typedef struct {
  double lr, beta1, beta2, eps, weight_decay;
  float clamp; /* estorch.py:243 clamps the negated gradient to +-1.0; <=0 disables */
} estk_adam_desc;

ESTK_API int estk_version(void);
ESTK_API const char* estk_last_error(void);

ESTK_API int estk_ctx_create(int device, estk_ctx** out);
ESTK_API int estk_ctx_destroy(estk_ctx* ctx);
/* sm_count, compute capability of the context's device (host ints). */
ESTK_API int estk_ctx_info(estk_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor);

/* ---- noise table (new-engine replacement of estorch.py:189-190's fresh
 *      Normal(0,sigma).sample per generation) ---- */

/* Fill table[0:len) with unit normals: Philox4x32-10(counter=i/4, key=seed)
 * + Box-Muller, each entry then rounded to the nearest fp16-representable value
 * (11 significant bits; still stored as fp32 here) so that the 16-bit copy made by
 * estk_shadow_f16 is exact.  len % 4 == 0.  Identical on every GPU for the same seed. */
ESTK_API int estk_fill_noise_table(estk_ctx* ctx, float* table, int64_t len, uint64_t seed,
                          void* stream);

/* offsets_out[i] = 32 * (mix64(mix64(seed ^ gen*C) + pair_begin + i) mod nslots),
 * nslots = (table_len - ceil32(n))/32 + 1.  gen = gen_host, plus state->generation
 * when `state` is non-null (device counter advanced by estk_track_best: a generation
 * replayed from a CUDA graph passes a constant gen_host -- 0, or 1 while the previous
 * generation's estk_track_best is still folded into this one -- and never a host scalar).  order_out (nullable, int32
 * [pairs]) receives the local pair indices sorted by offset (ties by index):
 * evaluating / reducing pairs in that order lets overlapping table rows hit L2. */
ESTK_API int estk_make_offsets(estk_ctx* ctx, uint64_t seed, const estk_state* state, int64_t gen_host,
                      int64_t pair_begin, int32_t pairs, int64_t table_len, int64_t n,
                      int64_t* offsets_out, int32_t* order_out, void* stream);

/* Materialise population rows (estorch.py:187-193 `_sample_policy`):
 * for m in [member_begin, member_begin+member_count): rows_out[m-member_begin] =
 * theta +- sigma*T[off], eps_out (nullable) = +-sigma*T[off].  P = 2*pairs. */
ESTK_API int estk_perturb_rows(estk_ctx* ctx, const float* theta, int64_t n, const float* table,
                      const int64_t* offsets, int32_t pairs, float sigma,
                      int32_t member_begin, int32_t member_count,
                      float* rows_out, float* eps_out, void* stream);

/* ---- kernel 1: population evaluate (estorch.py:195-202 `_calculate_returns`
 *      + Policy.forward examples/cartpole_es.py:14-20 + the synthetic agent
 *      return -mean((policy(obs)-target)^2)) ---- */

/* For each local pair j: returns_plus[j] / returns_minus[j] = return of
 * theta +/- sigma*T[offsets[j]].  obs [B, dims[0]], target [B, dims[L]] fp32
 * row-major.  bc_plus/bc_minus (nullable) [pairs, bc_dim] receive the behaviour
 * characteristic policy(obs[:bc_obs]).flatten()[:bc_dim] (examples/nsra_es.py:45-49).
 * order (nullable) = evaluation order from estk_make_offsets. */
ESTK_API int estk_eval_mlp(estk_ctx* ctx, const estk_mlp_desc* desc, const float* theta,
                  const float* table, const int64_t* offsets, const int32_t* order,
                  int32_t pairs, float sigma, const float* obs, const float* target, int32_t B,
                  float* returns_plus, float* returns_minus,
                  float* bc_plus, float* bc_minus, int32_t bc_obs, int32_t bc_dim,
                  void* stream);

/* Opt-in lower-precision tensor-core variant of estk_eval_mlp: same contract, bf16 operands with fp32
 * accumulation on tcgen05 (weights theta+-sigma*eps are rounded to bf16 when the
 * B-operand tile is formed, activations when they are written back).  Shapes:
 * every layer input width a multiple of 64 in [64,512], every output width a
 * multiple of 32 in [32,512], B a multiple of 256; otherwise
 * ESTK_ERR_UNSUPPORTED (estk_eval_mlp_bf16_supported() tells in advance).
 * Returns agree with the fp32 path to ~1e-2 relative (stated in the tests).
 * centre_return_out (nullable): additionally evaluate theta itself (sigma = 0) in
 * the same launch and write its return there -- the estorch.py:182 rollout of the
 * previous update folded into this generation's launch. */
ESTK_API int estk_eval_mlp_bf16(estk_ctx* ctx, const estk_mlp_desc* desc, const float* theta,
                       const float* table, const int64_t* offsets, const int32_t* order,
                       int32_t pairs, float sigma, const float* obs, const float* target, int32_t B,
                       float* returns_plus, float* returns_minus,
                       float* bc_plus, float* bc_minus, int32_t bc_obs, int32_t bc_dim,
                       float* centre_return_out, void* stream);
ESTK_API int estk_eval_mlp_center_bf16(estk_ctx* ctx, const estk_mlp_desc* desc, const float* theta,
                              const float* obs, const float* target, int32_t B,
                              float* return_out, float* bc_out, int32_t bc_obs, int32_t bc_dim,
                              void* stream);
ESTK_API int estk_eval_mlp_bf16_supported(const estk_mlp_desc* desc, int32_t B);

/* DEFAULT tensor-core evaluate ("f16"): same contract as estk_eval_mlp, computed on
 * tcgen05 with fp16 operands (11-bit significand, the TF32 class) and fp32 accumulation.
 * Every weight is formed in fp32 from the fp32 theta and the noise value and rounded ONCE:
 *   W16 = rn_f16(theta + s*sigma*T[off+i])      (sum in fp32, like estorch.py:192)
 * The noise is streamed from table16, a 16-bit copy of `table` that must be EXACT
 * (table16[i] == table[i] for all i: build it with estk_shadow_f16, whose inexact_count
 * must come back 0 -- true for tables made by estk_fill_noise_table); the kernel
 * therefore evaluates exactly the members the gradient estimate (estk_rank_grad*, fp32
 * table) weights.  Hidden activations are rounded to fp16 when written back (saturating
 * at +-65504), the observations enter as x_hi + x_lo (two fp16 operands), biases and the
 * squared error are fp32.  Shapes as estk_eval_mlp_bf16, plus dims[0] <= 256
 * (estk_eval_mlp_f16_supported()).  Measured agreement with the fp32 path: see
 * tests/test_kernels_gpu.py::test_eval_mlp_f16_* and profiles/. */
ESTK_API int estk_shadow_f16(estk_ctx* ctx, const float* src, uint16_t* dst, int64_t n,
                    uint64_t* inexact_count /* device, nullable: += #entries that changed */,
                    void* stream);
ESTK_API int estk_eval_mlp_f16(estk_ctx* ctx, const estk_mlp_desc* desc, const float* theta,
                      const float* table, const uint16_t* table16,
                      const int64_t* offsets, const int32_t* order, int32_t pairs, float sigma,
                      const float* obs, const float* target, int32_t B,
                      float* returns_plus, float* returns_minus,
                      float* bc_plus, float* bc_minus, int32_t bc_obs, int32_t bc_dim,
                      float* centre_return_out, void* stream);
ESTK_API int estk_eval_mlp_center_f16(estk_ctx* ctx, const estk_mlp_desc* desc, const float* theta,
                             const float* obs, const float* target, int32_t B,
                             float* return_out, float* bc_out, int32_t bc_obs, int32_t bc_dim,
                             void* stream);
ESTK_API int estk_eval_mlp_f16_supported(const estk_mlp_desc* desc, int32_t B);

/* NOTE on "bf16s": it evaluates F(theta16 +- sigma16 * eps16) while estk_rank_grad* weights the fp32 eps --
 * the estimator multiplies a return by a noise vector that is not exactly the one that produced it
 * (estorch.py:177-178 uses the same eps on both sides).  Measured at the north-star size
 * (profiles/r02_north_star_precision.json): returns 1.6e-5 off the fp32 arithmetic, gradient 5.6e-3 (L2)
 * off the one computed from fp32 returns.  The default mode ("f16") has neither departure.
 * "bf16s" (opt-in, lower precision): as estk_eval_mlp_bf16, but the weight producers read bf16 SHADOWS of
 * theta and of the noise table (theta16[i] = bf16(theta[i]), table16[i] =
 * bf16(table[i]), built with estk_shadow_bf16) -- half the bytes per weight
 * element, W = bf16(theta16 + s*sigma*table16[off+i]).  Biases still come from the
 * fp32 theta / table.  The gradient estimate (estk_rank_grad*) always uses the
 * fp32 table.  Tolerance vs the fp32 path is stated in tests/test_kernels_gpu.py. */
ESTK_API int estk_shadow_bf16(estk_ctx* ctx, const float* src, uint16_t* dst, int64_t n, void* stream);
ESTK_API int estk_eval_mlp_bf16s(estk_ctx* ctx, const estk_mlp_desc* desc, const float* theta,
                        const uint16_t* theta16, const float* table, const uint16_t* table16,
                        const int64_t* offsets, const int32_t* order, int32_t pairs, float sigma,
                        const float* obs, const float* target, int32_t B,
                        float* returns_plus, float* returns_minus,
                        float* bc_plus, float* bc_minus, int32_t bc_obs, int32_t bc_dim,
                        float* centre_return_out, void* stream);
ESTK_API int estk_eval_mlp_center_bf16s(estk_ctx* ctx, const estk_mlp_desc* desc, const float* theta,
                               const uint16_t* theta16, const float* obs, const float* target, int32_t B,
                               float* return_out, float* bc_out, int32_t bc_obs, int32_t bc_dim,
                               void* stream);

/* ---- conv + VirtualBatchNorm policy of examples/atari.py:14-37 (BASELINE config 5):
 * conv1 4->16 k8 s4, VBN(16), ReLU, conv2 16->32 k4 s2, VBN(32), ReLU, fc1 2592->256,
 * ReLU, fc2 256->n_actions; VBN statistics = per-(C,H,W) mean / unbiased variance of
 * the reference batch xref [ref_batch,4,84,84] under the member's own weights
 * (estorch/modules.py:48-58).  obs [B,4,84,84], target [B,n_actions]; return =
 * -mean((policy(obs)-target)^2).  offsets == NULL evaluates theta itself into
 * returns_plus[0].  scratch: caller-owned device slab of at least
 * estk_eval_conv_vbn_scratch_bytes(ctx, ref_batch, B) bytes.  fp32 throughout. */
ESTK_API int64_t estk_eval_conv_vbn_scratch_bytes(estk_ctx* ctx, int32_t ref_batch, int32_t B);
ESTK_API int estk_eval_conv_vbn(estk_ctx* ctx, int32_t n_actions, const float* theta, const float* table,
                       const int64_t* offsets, const int32_t* order, int32_t pairs, float sigma,
                       const float* xref, int32_t ref_batch, const float* obs, const float* target,
                       int32_t B, float* returns_plus, float* returns_minus, void* scratch,
                       int64_t scratch_bytes, void* stream);

/* Unperturbed policy (estorch.py:181-182 `_after_optimize` rollout):
 * return_out[0] = return of theta; bc_out (nullable) [bc_dim]. */
ESTK_API int estk_eval_mlp_center(estk_ctx* ctx, const estk_mlp_desc* desc, const float* theta,
                         const float* obs, const float* target, int32_t B,
                         float* return_out, float* bc_out, int32_t bc_obs, int32_t bc_dim,
                         void* stream);

/* estorch.py:182-185: episode_reward = *reward; if it beats state->best_reward,
 * take it and copy theta -> best_theta (the device analogue of
 * deepcopy(state_dict())).  Also advances state->generation (estorch.py:248). */
ESTK_API int estk_track_best(estk_ctx* ctx, estk_state* state, const float* reward,
                    const float* theta, float* best_theta, int64_t n, void* stream);

/* ---- kernel 2: centred-rank transform + weighted noise reduction + Adam
 *      (estorch.py:15-39 rank_transformation, :174-179 _calculate_grad and the
 *      NS/NSR/NSRA variants :419-425/:542-549/:640-648, :236-244 negate+clamp,
 *      :245 optimizer.step -> torch Adam) ---- */

/* Single-GPU fused form.  returns [P] (reward column), novelty [P] or NULL.
 * Blend row c = w_rew*c(reward) + w_nov*c(novelty) in fp32 when novelty is
 * given (ES: novelty NULL -> c(reward)).  Ranks are bit-exact vs
 * _compute_ranks on tie-free input (ties: stable by member index; NaN returns
 * sort last like numpy's argsort, among themselves by index); centring in
 * fp64 then fp32 as estorch.py:17-19,:176.
 *   g = (1/P) * sum_j (c_j - c_{j+pairs}) * T[off_j : off_j+n]
 * then grad = clamp(-g), Adam(theta, m, v) in place; state->adam_step += 1.
 * ranks_out / ranks2_out (nullable, int32 [P]) and grad_out (nullable, fp32
 * [n], the reference's un-negated estimate) are for inspection / parity. */
ESTK_API int estk_rank_grad_adam(estk_ctx* ctx, const float* returns, const float* novelty,
                        float w_rew, float w_nov, int32_t P,
                        const float* table, const int64_t* offsets, const int32_t* order,
                        int64_t n, float* theta, float* m, float* v, estk_state* state,
                        const estk_adam_desc* adam,
                        int32_t* ranks_out, int32_t* ranks2_out, float* grad_out, void* stream);

/* Multi-GPU form: same rank phase over all P returns, reduction over the local
 * pairs [pair_begin, pair_begin+pairs_local) only; grad_sum_out[n] receives the
 * RAW partial sum (no 1/P) to be all-reduced (SUM) by the caller
 * (replaces the Send/Recv star of estorch.py:207-233). */
ESTK_API int estk_rank_grad(estk_ctx* ctx, const float* returns, const float* novelty,
                   float w_rew, float w_nov, int32_t P,
                   const float* table, const int64_t* offsets, const int32_t* order,
                   int32_t pair_begin, int32_t pairs_local, int64_t n,
                   float* grad_sum_out, int32_t* ranks_out, int32_t* ranks2_out, void* stream);

/* The same two entry points reading the EXACT 16-bit copy of the table (estk_shadow_f16):
 * 8 noise values per 128-bit load, half the bytes per pair row, bit-identical results
 * (every fp16 value converts exactly to the fp32 value the fp32 table holds).
 * estk_rank_grad_h additionally takes `world`: with world > 1, `returns` / `novelty` are
 * laid out RANK-MAJOR, [world][2][pairs/world] -- exactly what an in-place all-gather of
 * each rank's (returns_plus[pairs_local], returns_minus[pairs_local]) block produces, so
 * no re-ordering copy is needed (member j < pairs of rank r, local index i, sits at
 * ((2r + 0) * pairs_local + i), its mirror at ((2r + 1) * pairs_local + i)); ranks_out
 * stays in member order and ties are still broken by member index. */
ESTK_API int estk_rank_grad_adam_h(estk_ctx* ctx, const float* returns, const float* novelty,
                          float w_rew, float w_nov, int32_t P,
                          const uint16_t* table16, const int64_t* offsets, const int32_t* order,
                          int64_t n, float* theta, float* m, float* v, estk_state* state,
                          const estk_adam_desc* adam,
                          int32_t* ranks_out, int32_t* ranks2_out, float* grad_out, void* stream);
ESTK_API int estk_rank_grad_h(estk_ctx* ctx, const float* returns, const float* novelty,
                     float w_rew, float w_nov, int32_t P, int32_t world,
                     const uint16_t* table16, const int64_t* offsets, const int32_t* order,
                     int32_t pair_begin, int32_t pairs_local, int64_t n,
                     float* grad_sum_out, int32_t* ranks_out, int32_t* ranks2_out, void* stream);

/* ---- the same update with the cross-GPU reduction INSIDE the kernel (NVLink peer memory) ----
 * Replaces estk_rank_grad_h -> NCCL all-reduce -> estk_clamp_adam (the reference's gather of returns
 * on the master + one optimizer step there, estorch.py:228-245) by ONE cooperative launch per GPU:
 *   rank + partial gradient over the local pairs -> own workspace            (as estk_rank_grad_h)
 *   cross-GPU barrier (flags in peer memory, release/acquire at system scope)
 *   reduce-scatter: rank r sums slice r of all `world` workspaces in rank order (loads over NVLink),
 *   all-gather: and stores the sum into every rank's workspace (stores over NVLink)
 *   cross-GPU barrier
 *   negate / clamp / Adam over all n from the summed gradient (replicated: every rank applies the
 *   same bits to its own theta / m / v, so the replicas stay bit-identical).
 * The sum over ranks is taken in rank order 0..world-1 on every slice: deterministic and identical on
 * all GPUs (it differs from NCCL's order by fp32 rounding only).
 * Every rank of the job must make the same call in the same generation (like a collective).
 *
 * Workspaces: estk_peer_alloc gives zero-filled device memory plus a 64-byte handle that another
 * process on the same node turns into a device pointer with estk_peer_open (CUDA IPC; the bytes of the
 * handle travel over whatever channel the host code has, e.g. a torch.distributed all_gather_object).
 * `peer_ws` is a HOST array of `world` device pointers -- entry r = rank r's workspace as mapped in the
 * calling process, entry `rank` = the caller's own allocation -- each of estk_xr_workspace_bytes(n). */
#define ESTK_MAX_PEERS 16
#define ESTK_PEER_HANDLE_BYTES 64
ESTK_API int64_t estk_xr_workspace_bytes(int64_t n);
ESTK_API int estk_peer_alloc(estk_ctx* ctx, int64_t bytes, void** ptr_out, unsigned char* handle_out);
ESTK_API int estk_peer_open(estk_ctx* ctx, const unsigned char* handle, void** ptr_out);
ESTK_API int estk_peer_close(estk_ctx* ctx, void* ptr);
ESTK_API int estk_peer_free(estk_ctx* ctx, void* ptr);
ESTK_API int estk_rank_grad_xr_adam_h(estk_ctx* ctx, const float* returns, const float* novelty,
                             float w_rew, float w_nov, int32_t P, int32_t world, int32_t rank,
                             const uint16_t* table16, const int64_t* offsets, const int32_t* order,
                             int32_t pair_begin, int32_t pairs_local, int64_t n,
                             void* const* peer_ws, float* theta, float* m, float* v,
                             estk_state* state, const estk_adam_desc* adam,
                             int32_t* ranks_out, int32_t* ranks2_out, float* grad_out, void* stream);

/* Epilogue on an all-reduced raw sum: g = grad_sum / P, negate, clamp, Adam.
 * theta/m/v NULL with grad_out set = gradient only (for non-Adam optimizers:
 * grad_out then receives clamp(-g), the tensor the reference stores in .grad). */
ESTK_API int estk_clamp_adam(estk_ctx* ctx, const float* grad_sum, int32_t P, int64_t n,
                    float* theta, float* m, float* v, estk_state* state,
                    const estk_adam_desc* adam, float* grad_out, void* stream);

/* ---- novelty (estorch.py:412-417): nov[i] = sum of the k smallest euclidean
 *      distances from bc[i] to the archive rows / ||archive||_F ---- */
ESTK_API int estk_knn_novelty(estk_ctx* ctx, const float* bc, int32_t count, const float* archive,
                     int32_t archive_len, int32_t dim, int32_t k, float* novelty_out,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ESTK_H_ */
