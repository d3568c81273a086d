// estk_eval_mlp_f16.cu -- kernel 1 of the ES generation (population evaluate) on the 5th-gen
// tensor cores, fp16 operands / fp32 accumulation, warpgroup-specialised.
//
// Contract: estk_eval_mlp_f16 in include/estk.h (reference estorch.py:187-202 `_sample_policy` +
// `_calculate_returns`, Policy.forward examples/cartpole_es.py:14-20, synthetic agent SURVEY 8d).
//
// Arithmetic.  Per member W_s = theta + s*sigma*eps is formed in fp32 from the fp32 theta and the
// noise value (streamed from the EXACT 16-bit copy of the table) and rounded ONCE to fp16 (11-bit
// significand, the TF32 class); hidden activations are rounded to fp16 when they are written back
// (bias + ReLU in fp32 first); the observations enter layer 0 as x_hi + x_lo (two fp16 operands on
// the same weight tile; their shared-memory image is built once per launch by stage_obs_f16_kernel
// and dropped into the activation buffer by one bulk copy per task, issued while the previous
// task's loss is still being reduced); accumulation, bias, squared error in fp32.
//
// Work unit ("task") = (antithetic pair j, sign s, chunk of 256 observations) on a cluster of two
// CTAs, tcgen05.mma.cta_group::2, UMMA M = 256 (128 observation rows per CTA).  Per layer
//   D[obs, out] = H[obs, in] * W_s[out, in]^T
//   A  activations H: 128 rows per CTA, fp16, K-major, 128B-swizzled, RESIDENT in shared memory
//      across layers (8 k-blocks x 16 KB, updated in place);
//   B  W_s formed ON THE FLY in a ring of six 16 KB slots (each CTA forms its half of the N tile;
//      the perturbed weights never exist in global memory).  Stage k = one [128 x 64] k-block:
//      the TMA engine lands the fp32 theta tile as two [128 x 32] halves (SWIZZLE_128B), half A in
//      slot k mod 3 and half B in slot 3 + k mod 3; the producers read both, add s*sigma*eps (noise
//      through registers) and write the fp16 tile IN PLACE over half A -- a row of the fp16 tile
//      occupies exactly the bytes of the same row of half A, and the eight lanes that own a row sit
//      in one warp, so a __syncwarp separates the reads from the write.  A B slot is released by the
//      producers at once (the next-but-one stage's theta can land while the tile waits for the
//      MMAs), an A slot by the MMAs' commit.  Up to ~64 KB of theta are in flight per SM without
//      holding a single register (measured: the L2 path needs ~1 KB in flight per GB/s and SM);
//   D  the whole layer output [128 x <=512] fp32 in TMEM (512 columns) as two N tiles.
// Warp roles (20 warps, homogeneous warpgroups so that setmaxnreg can move registers):
//   WG0    w0 MMA issuer (leader CTA), w1 TMEM allocator + theta TMA thread, w2-3 idle -> 40 registers
//   WG1-2  8 epilogue warps: two per TMEM lane quarter, each half of the columns -> 112
//   WG3-4  8 weight producers in 2 groups of 4 warps (noise one stage ahead in registers) -> 104
// Epilogue schedule per layer: tile 0 is drained while tile 1's MMAs still run (they read the
// activations, which therefore cannot be overwritten yet) -- bias, ReLU, fp16, parked as packed
// pairs in the TMEM columns the drain itself freed; when the layer is accumulated the parked half
// moves to smem (k-blocks 0..3, hand-over 0) and tile 1 is drained straight into k-blocks 4..7
// (hand-over 1) under the next layer's first MMAs.  One mbarrier per hand-over index (a shared one
// can alias phases, profiles/README.md).  The last layer is fused with the squared error.
//
// Roofline: 2*n*B*2*pairs flops per launch on the tensor pipe; L2 -> SM ingest 6 B per weight
// element and sign (fp32 theta + fp16 noise); the noise stream 2*n*pairs bytes is read once from
// HBM (the second sign of a pair runs on the neighbouring cluster at the same time: L2 hit).
#include "estk_tc.cuh"
#include <stdlib.h>
#include <string.h>

#ifdef ESTK_TC_PROFILE
__device__ unsigned long long g_f16_prof[32];   // per-role cycle counters of CTA 0 (triage builds only)
#define PROF_T() (prof ? clock64() : 0ll)
#define PROF_ADD(i, t0) do { if (prof) atomicAdd(&g_f16_prof[i], (unsigned long long)(clock64() - (t0))); } while (0)
#else
#define PROF_T() 0ll
#define PROF_ADD(i, t0) do { (void)(t0); } while (0)
#endif

namespace {

constexpr int CG = 2;                         // CTAs per cluster = tcgen05 cta_group
constexpr int kSlots = 6;                     // 16 KB slots of the B ring (beside 128 KB of activations):
constexpr int kASlots = 3, kBSlots = 3;       // three hold half A / the fp16 tile, three only ever half B
__host__ __device__ constexpr uint32_t slot_a(uint32_t k) { return k % kASlots; }
__host__ __device__ constexpr uint32_t slot_b(uint32_t k) { return kASlots + k % kBSlots; }
__host__ __device__ constexpr uint32_t par_a(uint32_t k) { return (k / kASlots) & 1u; }   // phase of the A slot's k-th use
__host__ __device__ constexpr uint32_t par_b(uint32_t k) { return (k / kBSlots) & 1u; }
constexpr int kStageBytes = kKBlockBytes;
// TMA descriptors of the fp32 theta, one per (layer, N tile): [N x K] row-major, box [rows of the
// tile per CTA x 32] (128 bytes), SWIZZLE_128B
struct ThetaMaps { CUtensorMap m[ESTK_MAX_LAYERS][2]; };
#ifndef ESTK_F16_PROD_WARPS
#define ESTK_F16_PROD_WARPS 8
#endif
constexpr int kCtlWarps = 4, kEpiWarps = 8, kProdWarps = ESTK_F16_PROD_WARPS;
constexpr int kEpiWarp0 = kCtlWarps, kProdWarp0 = kCtlWarps + kEpiWarps;
constexpr int kThreads = 32 * (kCtlWarps + kEpiWarps + kProdWarps);   // 640, launched at 96 registers
constexpr int kLaunchRegs = (65536 / kThreads) / 8 * 8;               // what __launch_bounds__(kThreads, 1) compiles to
#ifndef ESTK_F16_GROUPS
#define ESTK_F16_GROUPS 2
#endif
constexpr int kProdGroups = ESTK_F16_GROUPS, kProdGroupWarps = kProdWarps / kProdGroups, kPT = 32 * kProdGroupWarps;
constexpr int kEpiThreads = 32 * kEpiWarps;
// 128*40 + 256*112 + 256*104 = 60416 <= 640*96 = 61440
#ifndef ESTK_F16_WHATIF
#define ESTK_F16_WHATIF 0        // product builds: 0 (bits select what-if ablations in triage builds)
#endif
#ifndef ESTK_F16_REGS_EPI
#define ESTK_F16_REGS_EPI 112
#endif
#ifndef ESTK_F16_REGS_PROD
#define ESTK_F16_REGS_PROD 104
#endif
#ifndef ESTK_F16_REGS_CTL
#define ESTK_F16_REGS_CTL 40
#endif
constexpr int kRegsCtl = ESTK_F16_REGS_CTL, kRegsEpi = ESTK_F16_REGS_EPI, kRegsProd = ESTK_F16_REGS_PROD;
static_assert(32 * (kCtlWarps * kRegsCtl + kEpiWarps * kRegsEpi + kProdWarps * kRegsProd) <= kThreads * kLaunchRegs,
              "register pool of the CTA");
template <int N> __device__ __forceinline__ void set_role_regs() {     // move this warpgroup to N registers per thread
  if constexpr (N > kLaunchRegs) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
  else if constexpr (N < kLaunchRegs) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}


struct EvalF16Params {
  estk_mlp_desc desc;
  const float* theta;
  const float* table;        // fp32 table (biases); == theta for the centre-only launch
  const uint16_t* table16;   // exact fp16 copy of the table (weights); null for the centre-only launch
  const int64_t* offsets;    // null => centre evaluation
  const int32_t* order;
  int pairs;
  float sigma;
  const float* obs;
  const uint8_t* obs_image;  // fp16 hi/lo image of obs, one [2*K0/64 k-blocks x 16 KB] block per 128 observations,
                             // already in the layer-0 operand layout (stage_obs_f16_kernel)
  const float* target;
  int B, chunks;             // chunks of 128*CG observations
  float* ret_plus;
  float* ret_minus;
  float* bc_plus;
  float* bc_minus;
  int bc_obs, bc_dim;
  float* partial;            // [pairs*2 + 1][chunks*CG]
  unsigned int* counters;    // [pairs*2 + 1]
  float* centre_out;         // optional: also evaluate theta itself (sigma = 0) into centre_out[0]
  int n_centre;              // number of leading centre tasks (0 or chunks)
  int n_tasks;               // n_centre + pairs * n_signs * chunks
  int n_signs;               // 2, or 1 for the centre evaluation
  int prof;
};

struct Layer { int K, N; int64_t wbase, bbase; };

// task -> (slot, sign, chunk); the first n_centre tasks evaluate theta itself
struct TaskId { int slot, sgn, chunk; bool centre; };
__device__ __forceinline__ TaskId decode_task(const EvalF16Params& p, int task, bool all_centre) {
  TaskId t;
  if (task < p.n_centre) { t.slot = 0; t.sgn = 0; t.chunk = task; t.centre = true; return t; }
  const int q = task - p.n_centre;
  t.chunk = q % p.chunks;
  t.sgn = (q / p.chunks) % p.n_signs;
  t.slot = q / (p.chunks * p.n_signs);
  t.centre = all_centre;
  return t;
}

__global__ void __launch_bounds__(kThreads, 1) eval_mlp_f16_kernel(const EvalF16Params p,
                                                                    const __grid_constant__ ThetaMaps maps) {
  // Shared memory (all dynamic; the kernel declares no static __shared__, so the 1024-byte alignment the
  // swizzled operands need is the alignment of the dynamic window itself -- checked below):
  //   128 KB activations + 6 x 16 KB ring + 2 KB bias + barriers / layer table  =  227 KB - 0.6 KB
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sH = smem;                                    // 8 k-blocks x 16 KB
  uint8_t* sB = sH + (kMaxW / kBlockK) * kKBlockBytes;   // ring
  float* sBias = reinterpret_cast<float*>(sB + kSlots * kStageBytes);   // [512], one layer at a time
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + kMaxW);
  // slots 0..2 hold half A of a stage (theta k 0..31), then its fp16 tile; slots 3..5 only ever half B
  uint64_t* bar_full = bars;                   // [kSlots]  fp16 tile formed in the slot           (leader's are used)
  uint64_t* bar_emptyA = bars + kSlots;        // [kSlots]  tile consumed by the MMAs (tcgen05.commit) (local)
  uint64_t* bar_emptyB = bars + 2 * kSlots;    // [kSlots]  half B read by the producers            (local)
  uint64_t* bar_land = bars + 3 * kSlots;      // [kSlots]  both theta halves of the stage landed (on its A slot) (local)
  uint64_t* bar_acc0 = bars + 4 * kSlots;      // first N tile of a two-tile layer accumulated   (local)
  uint64_t* bar_acc = bars + 4 * kSlots + 1;   // layer accumulated                              (local)
  uint64_t* bar_h = bars + 4 * kSlots + 2;     // [2] next layer's k-blocks 0..3 / 4..7 in place (leader's are used)
  uint64_t* bar_obs = bars + 4 * kSlots + 4;   // the next task's observation image has landed   (local)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 4 * kSlots + 5);
  float* s_loss = reinterpret_cast<float*>(s_tmem + 2);  // [kEpiWarps]
  Layer* lay = reinterpret_cast<Layer*>(s_loss + kEpiWarps);             // [ESTK_MAX_LAYERS]
  if ((smem_u32(smem) & 1023u) != 0u) __trap();          // a misaligned window would silently corrupt the swizzle

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const int cluster_id = blockIdx.x / CG, n_clusters = gridDim.x / CG;
  const int L = p.desc.n_layers;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(smem_u32(bar_full + s), CG * kProdGroupWarps);
      mbar_init(smem_u32(bar_emptyA + s), 1);
      mbar_init(smem_u32(bar_emptyB + s), kProdGroupWarps);
      mbar_init(smem_u32(bar_land + s), 1);
    }
    mbar_init(smem_u32(bar_acc0), 1);
    mbar_init(smem_u32(bar_acc), 1);
    mbar_init(smem_u32(bar_h + 0), CG * kEpiWarps);
    mbar_init(smem_u32(bar_h + 1), CG * kEpiWarps);
    mbar_init(smem_u32(bar_obs), 1);
    fence_barrier_init();
  }
  if (threadIdx.x == 32) {          // per-layer geometry, in shared memory
    int64_t pb = 0;
    for (int l = 0; l < L; ++l) {
      lay[l].K = p.desc.dims[l];
      lay[l].N = p.desc.dims[l + 1];
      lay[l].wbase = pb;
      lay[l].bbase = pb + (int64_t)lay[l].K * lay[l].N;
      pb = lay[l].bbase + lay[l].N;
    }
  }
  cluster_sync_all();
  if (warp == 1) tmem_alloc<CG>(smem_u32(s_tmem), 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(s_tmem);
  const bool centre = (p.offsets == nullptr);
#ifdef ESTK_TC_PROFILE
  const bool prof = p.prof && blockIdx.x == 0 && lane == 0;
#endif

  if (warp < kCtlWarps) {
    set_role_regs<kRegsCtl>();
    if (warp == 0 && cta_rank == 0) {
      // =================================================================== MMA issuer
      uint32_t kst = 0, h_phase0 = 0, h_phase1 = 0;     // kst: global stage index
      const long long tm0 = PROF_T();
      for (int task = cluster_id; task < p.n_tasks; task += n_clusters) {
        for (int l = 0; l < L; ++l) {
          const long long th0 = PROF_T();
          mbar_wait(smem_u32(bar_h + 0), h_phase0);     // k-blocks 0..3 of this layer's input are in place,
          h_phase0 ^= 1;                                // TMEM columns [0,256) are drained
          PROF_ADD(1, th0);
          tc_fence_after();
          const int K = lay[l].K, N = lay[l].N;
          bool second_half_ready = (K <= 256) || (l == 0);   // hidden inputs wider than 256 arrive in two halves
          const int lo_kb = (l == 0) ? K / kBlockK : 0;      // layer 0: x_lo lives K/64 k-blocks after x_hi
          for (int n0 = 0; n0 < N; n0 += 256) {
            const int Ng = min(256, N - n0);
            const uint32_t idesc = make_idesc(128 * CG, Ng, true);
            const uint32_t tmem_d = tmem_base + (uint32_t)n0;
            for (int kb = 0; kb < K / kBlockK; ++kb) {
              if (kb >= 4 && !second_half_ready) {   // k-blocks 4..7 in place, TMEM columns [256,512) drained
                const long long th1 = PROF_T();
                mbar_wait(smem_u32(bar_h + 1), h_phase1);
                h_phase1 ^= 1;
                PROF_ADD(1, th1);
                tc_fence_after();
                second_half_ready = true;
              }
              const uint32_t stage = slot_a(kst), ring_phase = par_a(kst);   // the stage's A slot
              const long long tf0 = PROF_T();
              mbar_wait(smem_u32(bar_full + stage), ring_phase);
              PROF_ADD(2, tf0);
              const long long ti0 = PROF_T();
              tc_fence_after();
              if (elect_one()) {
                const uint32_t a_addr = smem_u32(sH + kb * kKBlockBytes);
                const uint32_t b_addr = smem_u32(sB + stage * kStageBytes);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma_bf16<CG>(tmem_d, make_sw128_desc(a_addr + k * 32), make_sw128_desc(b_addr + k * 32), idesc,
                                (kb | k) != 0 ? 1u : 0u);
                if (lo_kb) {
                  const uint32_t a_lo = smem_u32(sH + (kb + lo_kb) * kKBlockBytes);
#pragma unroll
                  for (int k = 0; k < kBlockK / 16; ++k)
                    umma_bf16<CG>(tmem_d, make_sw128_desc(a_lo + k * 32), make_sw128_desc(b_addr + k * 32), idesc, 1u);
                }
                umma_commit<CG>(smem_u32(bar_emptyA + stage));         // frees the slot (both CTAs)
                if (kb + 1 == K / kBlockK) {
                  if (n0 + 256 >= N) umma_commit<CG>(smem_u32(bar_acc));     // layer accumulated
                  else umma_commit<CG>(smem_u32(bar_acc0));                  // tile 0 of 2: its drain overlaps tile 1's MMAs
                }
              }
              __syncwarp();
              PROF_ADD(3, ti0);
              ++kst;
            }
          }
        }
      }
      PROF_ADD(0, tm0);
    } else if (warp == 1 && lane == 0) {
      // =================================================================== theta TMA thread
      // Walks the same (task, layer, n-tile, k-block) stage sequence as the producers.  Stage k lands
      // in A slot k mod 3 (last used by stage k-3, freed by the MMAs' commit) and B slot 3 + k mod 3
      // (last used by stage k-3 too, freed by the OTHER producer group as soon as it had read it -- so
      // the theta of a group's next stage lands while the group is still forming its current one).
      uint32_t k = 0;
      for (int task = cluster_id; task < p.n_tasks; task += n_clusters) {
        for (int l = 0; l < L; ++l) {
          const int K = lay[l].K, N = lay[l].N;
          for (int n0 = 0; n0 < N; n0 += 256) {
            const int rows = min(256, N - n0) / CG;
            const CUtensorMap* map = &maps.m[l][n0 ? 1 : 0];
            for (int kb = 0; kb < K / kBlockK; ++kb, ++k) {
              const uint32_t sa = slot_a(k), sb = slot_b(k);
              if (k >= kASlots) mbar_wait(smem_u32(bar_emptyA + sa), par_a(k - kASlots));
              if (k >= kBSlots) mbar_wait(smem_u32(bar_emptyB + sb), par_b(k - kBSlots));
              const uint32_t land = smem_u32(bar_land + sa);
              mbar_arrive_expect_tx(land, (uint32_t)rows * 256u);
              tma_load_2d(smem_u32(sB + sa * kStageBytes), map, kb * kBlockK, n0 + (int)cta_rank * rows, land);
              tma_load_2d(smem_u32(sB + sb * kStageBytes), map, kb * kBlockK + 32, n0 + (int)cta_rank * rows, land);
            }
          }
        }
      }
    }
  } else if (warp < kProdWarp0) {
    // =================================================================== epilogue warps
    set_role_regs<kRegsEpi>();
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = ew >> 2;                  // which half of the columns of a tile this warp drains
    const int row = q * 32 + lane;             // observation row inside the CTA's 128
    const int etid = ew * 32 + lane;           // 0..255
    uint32_t acc_phase = 0, acc0_phase = 0;
#ifdef ESTK_TC_PROFILE
    const bool eprof = prof && ew == 0;
#define EPROF_T() (eprof ? clock64() : 0ll)
#define EPROF_ADD(i, t0) do { if (eprof) atomicAdd(&g_f16_prof[i], (unsigned long long)(clock64() - (t0))); } while (0)
#else
#define EPROF_T() 0ll
#define EPROF_ADD(i, t0) do { (void)(t0); } while (0)
#endif
    const long long te0 = EPROF_T();
    const uint32_t trow_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    // next layer's input is handed over per k-block half: index 0 = k-blocks 0..3, 1 = 4..7
    auto hand_over = [&](int idx) {
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_on<CG>(smem_u32(bar_h + idx), 0);
    };
    uint32_t obs_phase = 0;
    const uint32_t obs_bytes = (uint32_t)(2 * lay[0].K / kBlockK) * kKBlockBytes;   // hi + lo k-blocks
    auto issue_obs = [&](int t) {               // one thread of the CTA: bulk copy of task t's observation block
      if (ew == 0 && lane == 0) {
        const TaskId tq = decode_task(p, t, centre);
        const uint8_t* src = p.obs_image + (size_t)(tq.chunk * CG + (int)cta_rank) * obs_bytes;
        fence_proxy_async();                    // earlier generic-proxy writes to these bytes (activations) are ordered first
        mbar_arrive_expect_tx(smem_u32(bar_obs), obs_bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sH)), "l"(src), "r"(obs_bytes), "r"(smem_u32(bar_obs)) : "memory");
      }
    };
    for (int task = cluster_id; task < p.n_tasks; task += n_clusters) {
      const long long to0 = EPROF_T();
      const TaskId tk = decode_task(p, task, centre);
      const int chunk = tk.chunk, sgn = tk.sgn, slot = tk.slot;
      const int j = (!tk.centre && p.order) ? p.order[slot] : slot;
      const float* trow = tk.centre ? p.theta : p.table + p.offsets[j];
      const float ssig = tk.centre ? 0.f : (sgn ? -p.sigma : p.sigma);
      const int b = (chunk * CG + (int)cta_rank) * 128 + row;      // global observation index
      // ---- this CTA's observations as the layer-0 A operand, x = x_hi + x_lo (fp16 each): the image was
      //      built once per launch (stage_obs_f16_kernel) in exactly the shared-memory layout and is
      //      dropped into the activation buffer by ONE bulk copy, issued as soon as the previous task's
      //      last MMAs were done (see below) -- normally it has landed long before this wait
      if (task == cluster_id) issue_obs(task);
      mbar_wait(smem_u32(bar_obs), obs_phase);
      obs_phase ^= 1;
      EPROF_ADD(11, to0);
      // the staged observations are this task's layer-0 input (the previous task's TMEM reads are
      // long done): release the MMA warp before anything else
      hand_over(0);
      float loss = 0.f;
      for (int l = 0; l < L; ++l) {
        const long long tb0 = EPROF_T();
        const int N = lay[l].N;
        float* bias = sBias;
        float bnew[(kMaxW + kEpiThreads - 1) / kEpiThreads];
#pragma unroll
        for (int i = 0; i < (kMaxW + kEpiThreads - 1) / kEpiThreads; ++i) {
          const int o = etid + i * kEpiThreads;
          bnew[i] = (o < N) ? fmaf(ssig, ld_noise1(trow + lay[l].bbase + o), __ldg(p.theta + lay[l].bbase + o)) : 0.f;
        }
        named_bar_sync(3, kEpiThreads);   // every warp is done reading the previous layer's bias[] ...
#pragma unroll
        for (int i = 0; i < (kMaxW + kEpiThreads - 1) / kEpiThreads; ++i)
          if (etid + i * kEpiThreads < N) bias[etid + i * kEpiThreads] = bnew[i];
        named_bar_sync(1, kEpiThreads);   // ... and this layer's is published among the epilogue warps
        EPROF_ADD(12, tb0);
        const bool last = (l == L - 1);
        const bool two = N > 256;               // two N tiles: columns [0,256) and [256,N)
        // bias + ReLU + round to fp16: 32 accumulator columns -> 16 packed words
        auto pack = [&](const uint32_t (&v)[32], int c0, uint32_t (&pk)[16]) {
          const uint32_t baddr = smem_u32(bias + c0);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 t = ld_shared_v4(baddr + g * 16);
            pk[g * 2 + 0] = pack_f16_relu(__uint_as_float(v[g * 4 + 0]) + t.x, __uint_as_float(v[g * 4 + 1]) + t.y);
            pk[g * 2 + 1] = pack_f16_relu(__uint_as_float(v[g * 4 + 2]) + t.z, __uint_as_float(v[g * 4 + 3]) + t.w);
          }
        };
        // `words` packed words (2*words output features starting at feature f0) -> activations in smem
        auto store_h = [&](const uint32_t* pk, int f0, int words) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {          // chunks of 8 output features = 16 bytes of fp16
            if (g * 4 < words) {
              const int col = f0 + g * 8;
              const uint32_t addr = smem_u32(sH + (col >> 6) * kKBlockBytes) + sw128_offset(row, (col & 63) >> 3);
              st_shared_v4(addr, pk[g * 4 + 0], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
            }
          }
        };
        // last layer: fused squared error (and the behaviour characterisation)
        auto loss_chunk = [&](const uint32_t (&v)[32], int c0) {
          const uint32_t baddr = smem_u32(bias + c0);
          const float* trg = p.target + (size_t)b * N + c0;
          float* bc = (tk.centre && !centre) ? nullptr : (sgn ? p.bc_minus : p.bc_plus);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(trg + g * 4));
            const float4 b4 = ld_shared_v4(baddr + g * 16);
            const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
            const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int o = c0 + g * 4 + e;
              const float y = __uint_as_float(v[g * 4 + e]) + bv[e];
              const float d = y - tv[e];
              loss = fmaf(d, d, loss);
              if (bc) {
                const int64_t idx = (int64_t)b * N + o;
                if (b < p.bc_obs && idx < p.bc_dim) bc[(size_t)j * p.bc_dim + idx] = y;
              }
            }
          }
        };
        const int base_h = 128 * half;          // this warp's half of tile 0: accumulator columns [base_h, base_h+128)
        if (two) {
          // ---- tile 0 is accumulated while tile 1's MMAs still run (they read the activations in
          // smem, so those cannot be overwritten yet): drain tile 0 now.  Hidden layers park the
          // result as packed fp16 pairs in TMEM columns this warp's own drain has already freed
          // (columns [c0,c0+32) -> [base_h + (c0-base_h)/2, +16), in place).
          const long long ta0 = EPROF_T();
          mbar_wait(smem_u32(bar_acc0), acc0_phase);
          acc0_phase ^= 1;
          tc_fence_after();
          EPROF_ADD(13, ta0);
          const long long tx0 = EPROF_T();
          for (int c0 = base_h; c0 < base_h + 128; c0 += 32) {
            uint32_t va[32];
            tmem_ld32(trow_addr + (uint32_t)c0, va);
            tmem_ld_wait();
            if (last) {
              loss_chunk(va, c0);
            } else {
              uint32_t pk[16];
              pack(va, c0, pk);
              tmem_st16(trow_addr + (uint32_t)(base_h + ((c0 - base_h) >> 1)), pk);
            }
          }
          if (!last) tmem_st_wait();
          EPROF_ADD(15, tx0);
        }
        // ---- wait for the whole layer: every MMA that reads the activations has completed
        const long long ta1 = EPROF_T();
        mbar_wait(smem_u32(bar_acc), acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        EPROF_ADD(13, ta1);
        // nothing reads the activation buffer any more in this task: the next task's observations can land
        // while the loss is computed from TMEM
        if (last && task + n_clusters < p.n_tasks) issue_obs(task + n_clusters);
        const long long tx1 = EPROF_T();
        if (!last && two) {
          // parked half -> activation k-blocks (2*half, 2*half+1); hand-over 0 when both halves are in
          for (int s0 = 0; s0 < 64; s0 += 32) {
            uint32_t pk[32];
            tmem_ld32(trow_addr + (uint32_t)(base_h + s0), pk);
            tmem_ld_wait();
            store_h(pk, base_h + 2 * s0, 32);
          }
          hand_over(0);
        }
        {
          const int lo = two ? 256 : 0;
          const int cnt = (N - lo) / 32, first = (cnt + 1) / 2;      // the two warps of a lane quarter split the chunks
          const int i0 = half ? first : 0, i1 = half ? cnt : first;
          for (int i = i0; i < i1; ++i) {
            const int c0 = lo + i * 32;
            uint32_t va[32];
            tmem_ld32(trow_addr + (uint32_t)c0, va);
            tmem_ld_wait();
            if (last) {
              loss_chunk(va, c0);
            } else {
              uint32_t pk[16];
              pack(va, c0, pk);
              store_h(pk, c0, 16);
            }
          }
        }
        if (!last) hand_over(two ? 1 : 0);       // second half (or the only one when N <= 256)
        EPROF_ADD(14, tx1);
      }
      // ---- squared-error partial of this CTA; the last arriver combines them in fixed order
      loss = warp_sum_f(loss);
      if (lane == 0) s_loss[ew] = loss;
      named_bar_sync(2, kEpiThreads);
      if (etid == 0) {
        const float tot = ((s_loss[0] + s_loss[1]) + (s_loss[2] + s_loss[3])) + ((s_loss[4] + s_loss[5]) + (s_loss[6] + s_loss[7]));
        const int parts = p.chunks * CG;
        const bool folded = tk.centre && !centre;                 // centre task riding in a population launch
        const int cell = folded ? p.pairs * 2 : slot * 2 + sgn;
        float* part = p.partial + (size_t)cell * parts;
        part[chunk * CG + (int)cta_rank] = tot;
        __threadfence();
        const unsigned int arrived = atomicAdd(p.counters + cell, 1u);
        if (arrived == (unsigned int)parts - 1) {
          __threadfence();
          float s = 0.f;
          for (int c = 0; c < parts; ++c) s += __ldcg(part + c);
          const float r = -(s / ((float)p.B * (float)lay[L - 1].N));
          if (folded) p.centre_out[0] = r;
          else if (sgn) p.ret_minus[j] = r;
          else p.ret_plus[j] = r;
          p.counters[cell] = 0u;
        }
      }
      named_bar_sync(2, kEpiThreads);   // s_loss reusable
    }
    EPROF_ADD(10, te0);
  } else {
    // =================================================================== weight producers
    // kProdGroups groups of warps; group g converts stages g, g+G, g+2G, ... of the flattened
    // (task, layer, n-tile, k-block) stage sequence.  A thread owns the 16-byte output chunk c8 of
    // rows r0 + u*kRS of the [rows x 64] tile: 8 weights = 32 bytes of fp32 theta from the landing
    // slots (k < 32: slot A, k >= 32: slot B; both 128B-swizzled like the fp16 tile) + 16 bytes of
    // the noise row (fp16, a 128-bit global load issued one whole stage earlier),
    // W = rn_f16(theta + s*sigma*eps) with the sum in fp32, written over the same row of slot A.
    set_role_regs<kRegsProd>();
    const int pwarp = warp - kProdWarp0;
    const int pgroup = pwarp / kProdGroupWarps;
    constexpr int kRS = kPT / 8;             // tile rows covered by one item step of the group
    constexpr int kIU = 128 / kRS;           // item steps per (full) stage
#ifndef ESTK_F16_FB
#define ESTK_F16_FB 2
#endif
    constexpr int kFB = ESTK_F16_FB;         // rows formed per step (ld.shared batch)
    // Lane -> (row, chunk).  A warp covers 4 rows x 8 output chunks; a quarter-warp (the unit a 128-bit
    // shared-memory access is served in) covers chunks 4h..4h+3 (h = which theta half they come from)
    // of TWO rows whose indices differ by XOR 5 inside the 8-row swizzle group: the 128B swizzle then
    // maps the quarter's eight 16-byte accesses to eight different bank groups, for the two theta
    // loads (chunks 2c', 2c'+1 of half h) and for the fp16 store (chunk 4h+c') alike -- no conflicts,
    // no per-lane reordering.  All eight writers of a row (and the readers of its half A) stay in one warp.
    const int wg = pwarp % kProdGroupWarps, wq = lane >> 3, li = lane & 7;
    const int hsel = wq & 1, cq = li & 3;
    const int r0 = 8 * (wg >> 1) + ((2 * (wg & 1) + (wq >> 1)) ^ ((li >> 2) * 5)), c8 = 4 * hsel + cq;
    static_assert(kRS == 8 || kRS == 16, "row mapping assumes 2 or 4 warps per producer group");
    const uint32_t roff0 = sw128_offset(r0, 2 * cq), roff1 = sw128_offset(r0, 2 * cq + 1);
    const uint32_t woff = sw128_offset(r0, c8);               // the fp16 output chunk
    struct St { const uint16_t* ep; int k_rs; int rows; float sg; uint32_t kst; };
    // position of a stage in the flattened (task, layer, n-tile, k-block) sequence
    struct Pos { int task, l, n0, kb, nkb; };
    auto step_pos = [&](Pos& q, int by) -> bool {               // `by` stages further; false past the last task
      q.kb += by;
      while (q.kb >= q.nkb) {
        q.kb -= q.nkb;
        q.n0 += 256;
        if (q.n0 >= lay[q.l].N) {
          q.n0 = 0;
          if (++q.l == L) { q.l = 0; q.task += n_clusters; }
          if (q.task >= p.n_tasks) return false;
        }
        q.nkb = lay[q.l].K / kBlockK;
      }
      return true;
    };
    // (noise row, signed sigma) of a task: two dependent global loads -- fetched one task ahead
    int cached_task = -1, pf_task = -1;
    const uint16_t *cached_trow16 = nullptr, *pf_trow16 = nullptr;
    float cached_ssig = 0.f, pf_ssig = 0.f;
    auto fetch_task = [&](int t, const uint16_t*& trow16, float& ssig) {
      const TaskId tk = decode_task(p, t, centre);
      const int j = (!tk.centre && p.order) ? p.order[tk.slot] : tk.slot;
      trow16 = tk.centre ? nullptr : p.table16 + p.offsets[j];
      ssig = tk.centre ? 0.f : (tk.sgn ? -p.sigma : p.sigma);
    };
    auto describe = [&](St& d, const Pos& q, uint32_t kst) {   // full descriptor of the stage at position q
      if (q.task != cached_task) {
        cached_task = q.task;
        if (q.task == pf_task) { cached_trow16 = pf_trow16; cached_ssig = pf_ssig; }
        else fetch_task(q.task, cached_trow16, cached_ssig);
        pf_task = q.task + n_clusters;       // consumed a whole task later: the loads never stall the pipeline
        if (pf_task < p.n_tasks) fetch_task(pf_task, pf_trow16, pf_ssig);
      }
      const int K = lay[q.l].K;
      d.rows = min(256, lay[q.l].N - q.n0) / CG;               // this CTA's share of the B tile
      const int64_t rbase = lay[q.l].wbase + (int64_t)(q.n0 + (int)cta_rank * d.rows + r0) * K + q.kb * kBlockK + c8 * 8;
      d.ep = cached_trow16 ? cached_trow16 + rbase : nullptr;
      d.k_rs = K * kRS;
      d.sg = cached_ssig;
      d.kst = kst;
    };
    Pos pos = {cluster_id, 0, 0, 0, lay[0].K / kBlockK};
    bool has_cur = pos.task < p.n_tasks;
    if (has_cur && pgroup) has_cur = step_pos(pos, pgroup);
    // The noise of a stage is loaded kDepth of this group's stages ahead (kDepth * 8 128-bit registers
    // per thread in flight all the time): the L2 path delivers ~1 GB/s per SM per KB in flight.
#ifndef ESTK_F16_EPS_DEPTH
#define ESTK_F16_EPS_DEPTH 1     // 2 (with 96 / 120 registers) measured slower: 5.05 ms vs 3.20 ms
#endif
    constexpr int kDepth = ESTK_F16_EPS_DEPTH;
    uint4 E[kDepth][kIU];
    auto load_eps = [&](const St& d, int u, uint4& e) {
      e = make_uint4(0u, 0u, 0u, 0u);
#if !(ESTK_F16_WHATIF & 1)      // triage builds only: bit 0 = no noise loads
      if (d.ep && u * kRS + r0 < d.rows) e = ld_noise4u(reinterpret_cast<const uint4*>(d.ep + (size_t)u * d.k_rs));
#endif
    };
    // descriptor of the stage kProdGroups k-blocks after `d` / `pos` (pos always tracks the furthest stage
    // described so far): inside the same N tile only the noise pointer and the stage index move (the
    // common case); otherwise the full descriptor
    auto next_desc = [&](const St& d, St& out) -> bool {
      if (pos.kb + kProdGroups < pos.nkb) {
        pos.kb += kProdGroups;
        out = d;
        out.kst = d.kst + kProdGroups;
        if (out.ep) out.ep = d.ep + kBlockK * kProdGroups;
        return true;
      }
      if (!step_pos(pos, kProdGroups)) return false;
      describe(out, pos, d.kst + kProdGroups);
      return true;
    };
#ifdef ESTK_TC_PROFILE
    const bool pprof = prof && pwarp == 0;
#define PPROF_T() (pprof ? clock64() : 0ll)
#define PPROF_ADD(i, t0) do { if (pprof) atomicAdd(&g_f16_prof[i], (unsigned long long)(clock64() - (t0))); } while (0)
#else
#define PPROF_T() 0ll
#define PPROF_ADD(i, t0) do { (void)(t0); } while (0)
#endif
    const long long tp0 = PPROF_T();
    // one stage: wait for its theta halves, form the fp16 tile in place with the noise in Ecur, refill Ecur
    // with the noise of stage `fut` (kDepth stages ahead), publish the tile, release half B
    auto process = [&](const St& cur, uint4 (&Ecur)[kIU], bool has_fut, const St& fut) {
      const uint32_t sa = slot_a(cur.kst), sb = slot_b(cur.kst), par = par_a(cur.kst);
      const uint32_t base_a = smem_u32(sB + sa * kStageBytes), base_b = smem_u32(sB + sb * kStageBytes);
      const uint32_t rd = hsel ? base_b : base_a;               // this thread's theta chunks live in half A or B
      const long long tw0 = PPROF_T();
      mbar_wait(smem_u32(bar_land + sa), par);                  // both theta halves of the stage have landed
      PPROF_ADD(7, tw0);
      const long long tc0 = PPROF_T();
      // kFB rows per step: all ld.shared of the step first (their latency overlaps), one
      // __syncwarp, then the in-place stores and the refill of the noise registers
#pragma unroll
      for (int ub = 0; ub < kIU; ub += kFB) {
        float4 ta[kFB], tb[kFB];
#pragma unroll
        for (int q = 0; q < kFB; ++q) {
#if ESTK_F16_WHATIF & 2         // triage builds only: bit 1 = no theta reads from the landing slots
          ta[q] = tb[q] = make_float4(cur.sg, 0.f, 0.f, 0.f);
#else
          if ((ub + q) * kRS + r0 < cur.rows) {
            ta[q] = ld_shared_v4(rd + roff0 + (uint32_t)(ub + q) * (kRS * 128));
            tb[q] = ld_shared_v4(rd + roff1 + (uint32_t)(ub + q) * (kRS * 128));
          }
#endif
        }
        uint32_t w[kFB][4];
#pragma unroll
        for (int q = 0; q < kFB; ++q) {
          const float4 t0 = ta[q], t1 = tb[q];
          const uint4 e = Ecur[ub + q];
          const float2 e0 = unpack_f16(e.x), e1 = unpack_f16(e.y), e2 = unpack_f16(e.z), e3 = unpack_f16(e.w);
          w[q][0] = pack_f16(fmaf(cur.sg, e0.x, t0.x), fmaf(cur.sg, e0.y, t0.y));
          w[q][1] = pack_f16(fmaf(cur.sg, e1.x, t0.z), fmaf(cur.sg, e1.y, t0.w));
          w[q][2] = pack_f16(fmaf(cur.sg, e2.x, t1.x), fmaf(cur.sg, e2.y, t1.y));
          w[q][3] = pack_f16(fmaf(cur.sg, e3.x, t1.z), fmaf(cur.sg, e3.y, t1.w));
        }
#ifdef ESTK_F16_RACECHECK_BAR   // sanitizer builds only: a CTA-scope named barrier over the group instead of the warp barrier
        named_bar_sync(4 + pgroup, kPT);      // (compute-sanitizer racecheck does not model __syncwarp as ordering shared-memory accesses)
#else
        __syncwarp();                         // every lane has read its rows of half A before any lane overwrites them
#endif
#pragma unroll
        for (int q = 0; q < kFB; ++q) {
#if ESTK_F16_WHATIF & 4         // triage builds only: bit 2 = the fp16 tile is not stored (only the first word, to keep the math alive)
          if ((ub + q) * kRS + r0 < cur.rows && (w[q][0] ^ w[q][1] ^ w[q][2] ^ w[q][3]) == 0x12345u)
#else
          if ((ub + q) * kRS + r0 < cur.rows)
#endif
            st_shared_v4(base_a + woff + (uint32_t)(ub + q) * (kRS * 128), w[q][0], w[q][1], w[q][2], w[q][3]);
          if (has_fut) load_eps(fut, ub + q, Ecur[ub + q]);
        }
      }
      PPROF_ADD(8, tc0);
      const long long tf0 = PPROF_T();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_on<CG>(smem_u32(bar_full + sa), 0);                                          // tile formed (leader's barrier)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar_emptyB + sb)) : "memory");   // half B is free
      }
      PPROF_ADD(9, tf0);
    };
    // descriptors of the current stage and the kDepth following ones; E[i] holds the noise of st[i] (st[kDepth]'s
    // is requested while st[0] is formed, into the registers st[0] frees)
    St st[kDepth + 1];
    bool has[kDepth + 1];
#pragma unroll
    for (int i = 0; i <= kDepth; ++i) { st[i] = St{}; has[i] = false; }
    has[0] = has_cur;
    if (has_cur) {
      describe(st[0], pos, (uint32_t)pgroup);
#pragma unroll
      for (int i = 1; i <= kDepth; ++i) has[i] = has[i - 1] && next_desc(st[i - 1], st[i]);
#pragma unroll
      for (int i = 0; i < kDepth; ++i)
        if (has[i]) {
#pragma unroll
          for (int u = 0; u < kIU; ++u) load_eps(st[i], u, E[i][u]);
        }
    }
    while (has[0]) {
      // kDepth stages per trip so that the noise register sets are indexed statically
#pragma unroll
      for (int r = 0; r < kDepth; ++r) {
        if (has[0]) {
          process(st[0], E[r], has[kDepth], st[kDepth]);
#pragma unroll
          for (int i = 0; i < kDepth; ++i) { st[i] = st[i + 1]; has[i] = has[i + 1]; }
          has[kDepth] = has[kDepth - 1] && next_desc(st[kDepth - 1], st[kDepth]);
        }
      }
    }
    PPROF_ADD(4, tp0);
  }

  // ---- teardown
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc<CG>(tmem_base, 512);
}

// fp16 hi/lo image of the observation batch in the layer-0 operand layout: per block of 128 observations
// (one CTA's rows of one chunk) 2*K0/64 k-blocks of 16 KB -- x_hi k-blocks first, then x_lo = rn_f16(x - x_hi);
// inside a k-block row r, 16-byte chunk c sits at sw128_offset(r, c).  One thread per (observation, 8 elements).
__global__ void __launch_bounds__(256) stage_obs_f16_kernel(const float* __restrict__ obs, uint8_t* __restrict__ image,
                                                            int B, int K0) {
  const int per_row = K0 / 8;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * per_row) return;
  const int b = idx / per_row, c = idx % per_row;
  const float4 x0 = __ldg(reinterpret_cast<const float4*>(obs + (size_t)b * K0 + c * 8));
  const float4 x1 = __ldg(reinterpret_cast<const float4*>(obs + (size_t)b * K0 + c * 8 + 4));
  const uint32_t h0 = pack_f16(x0.x, x0.y), h1 = pack_f16(x0.z, x0.w), h2 = pack_f16(x1.x, x1.y), h3 = pack_f16(x1.z, x1.w);
  const float2 f0 = unpack_f16(h0), f1 = unpack_f16(h1), f2 = unpack_f16(h2), f3 = unpack_f16(h3);
  const int nkb = K0 / kBlockK;
  uint8_t* blk = image + (size_t)(b / 128) * (size_t)(2 * nkb) * kKBlockBytes;
  const uint32_t off = sw128_offset(b % 128, c & 7);
  *reinterpret_cast<uint4*>(blk + (size_t)(c >> 3) * kKBlockBytes + off) = make_uint4(h0, h1, h2, h3);
  *reinterpret_cast<uint4*>(blk + (size_t)((c >> 3) + nkb) * kKBlockBytes + off) =
      make_uint4(pack_f16(x0.x - f0.x, x0.y - f0.y), pack_f16(x0.z - f1.x, x0.w - f1.y),
                 pack_f16(x1.x - f2.x, x1.y - f2.y), pack_f16(x1.z - f3.x, x1.w - f3.y));
}

size_t f16_smem_bytes() {
  return (size_t)(kMaxW / kBlockK) * kKBlockBytes + (size_t)kSlots * kStageBytes + kMaxW * sizeof(float) +
         (4 * kSlots + 5) * sizeof(uint64_t) + 2 * sizeof(uint32_t) + kEpiWarps * sizeof(float) +
         ESTK_MAX_LAYERS * sizeof(Layer);
}

// ---- TMA descriptors of the fp32 theta (host side)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int build_theta_maps(const estk_mlp_desc& d, const float* theta, ThetaMaps* out) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    ESTK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      estk_set_error("cuTensorMapEncodeTiled is not available from this driver");
      return ESTK_ERR_CUDA;
    }
    encode = (EncodeTiledFn)fn;
  }
  // single-entry cache: the descriptors only depend on theta's address and the layer shapes
  static thread_local struct { const float* ptr; estk_mlp_desc desc; int device; bool valid; ThetaMaps maps; } cache = {};
  int dev = -1;
  ESTK_CUDA(cudaGetDevice(&dev));
  if (cache.valid && cache.ptr == theta && cache.device == dev && memcmp(&cache.desc, &d, sizeof(d)) == 0) {
    *out = cache.maps;
    return ESTK_OK;
  }
  memset(out, 0, sizeof(*out));
  int64_t pb = 0;
  for (int l = 0; l < d.n_layers; ++l) {
    const int K = d.dims[l], N = d.dims[l + 1];
    for (int t = 0; t < 2; ++t) {
      const int n0 = t * 256;
      const int Nt = (n0 < N) ? (N - n0 < 256 ? N - n0 : 256) : (N < 256 ? N : 256);   // tile 1 of a one-tile layer: unused copy
      const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
      const cuuint64_t gstride[1] = {(cuuint64_t)K * 4};
      const cuuint32_t box[2] = {32u, (cuuint32_t)(Nt / CG)};
      const cuuint32_t estride[2] = {1, 1};
      const CUresult r = encode(&out->m[l][t], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)(theta + pb), gdim, gstride,
                                box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        estk_set_error("cuTensorMapEncodeTiled failed (%d) for layer %d tile %d [N=%d K=%d box %dx32]", (int)r, l, t, N, K,
                       Nt / CG);
        return ESTK_ERR_CUDA;
      }
    }
    pb += (int64_t)K * N + N;
  }
  cache.ptr = theta; cache.desc = d; cache.device = dev; cache.maps = *out; cache.valid = true;
  return ESTK_OK;
}

int f16_supported(const estk_mlp_desc& d, int B, const char** why) {
  if (d.n_layers < 1 || d.n_layers > ESTK_MAX_LAYERS) { *why = "n_layers"; return 0; }
  if (d.activation != 0) { *why = "activation"; return 0; }
  if (2 * d.dims[0] > kMaxW) { *why = "the observations enter as hi + lo halves: input width <= 256"; return 0; }
  for (int l = 0; l < d.n_layers; ++l) {
    if (d.dims[l] % 64 || d.dims[l] > kMaxW || d.dims[l] < 64) { *why = "layer input width must be a multiple of 64 in [64,512]"; return 0; }
    const int N = d.dims[l + 1];
    if (N % 32 || N > kMaxW || N < 32) { *why = "layer output width must be a multiple of 32 in [32,512]"; return 0; }
  }
  if (B % (128 * CG)) { *why = "batch must be a multiple of 256"; return 0; }
  return 1;
}

int run_f16(estk_ctx* ctx, EvalF16Params& p, cudaStream_t stream, const char* who) {
  const char* why = "";
  if (!f16_supported(p.desc, p.B, &why)) {
    estk_set_error("%s: shape not supported by the tcgen05 path (%s)", who, why);
    return ESTK_ERR_UNSUPPORTED;
  }
  ESTK_CHECK_ARG(p.pairs >= 1 && p.pairs <= ESTK_MAX_POPULATION / 2, "%s: pairs=%d", who, p.pairs);
  ESTK_CHECK_ARG((((uintptr_t)p.theta) & 15u) == 0, "%s: theta must be 16-byte aligned (TMA)", who);
  p.chunks = p.B / (128 * CG);
  ESTK_CHECK_ARG(p.chunks * CG <= kEvalMaxChunks, "%s: batch too large", who);
  p.n_centre = p.centre_out ? p.chunks : 0;
  ESTK_CHECK_ARG(!p.centre_out || p.pairs * 2 < ESTK_MAX_POPULATION, "%s: population too large to fold the centre task", who);
  p.n_tasks = p.n_centre + p.pairs * p.n_signs * p.chunks;
  p.partial = ctx->eval_partial;
  p.counters = ctx->counters;
#ifdef ESTK_TC_PROFILE
  { const char* e = getenv("ESTK_TC_PROFILE"); p.prof = e ? atoi(e) : 0; }
#endif
  static thread_local ThetaMaps maps;
  { const int rc = build_theta_maps(p.desc, p.theta, &maps); if (rc != ESTK_OK) return rc; }
  {   // the observation image of this launch (a few hundred KB: negligible next to the evaluate)
    const int K0 = p.desc.dims[0];
    ESTK_CHECK_ARG((size_t)(p.B / 128) * (size_t)(2 * K0 / kBlockK) * kKBlockBytes <= kObsImageBytes, "%s: batch too large", who);
    const int items = p.B * (K0 / 8);
    stage_obs_f16_kernel<<<(items + 255) / 256, 256, 0, stream>>>(p.obs, reinterpret_cast<uint8_t*>(ctx->obs_image), p.B, K0);
    ESTK_CUDA(cudaGetLastError());
    p.obs_image = reinterpret_cast<const uint8_t*>(ctx->obs_image);
  }
  const size_t smem = f16_smem_bytes();
  ESTK_CUDA(cudaFuncSetAttribute(eval_mlp_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int clusters = ctx->sm_count / CG;
  if (clusters > p.n_tasks) clusters = p.n_tasks;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CG);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ESTK_CUDA(cudaLaunchKernelEx(&cfg, eval_mlp_f16_kernel, p, maps));
  return ESTK_OK;
}

}  // namespace

// entry points used by estk_eval_mlp_tc.cu's estk_eval_mlp_f16 / estk_eval_mlp_center_f16
int estk_f16v2_supported(const estk_mlp_desc* desc, int B) {
  const char* why = "";
  return f16_supported(*desc, B, &why);
}

int estk_f16v2_eval(estk_ctx* ctx, const estk_mlp_desc* desc, const float* theta, const float* table,
                    const uint16_t* table16, const int64_t* offsets, const int32_t* order, int32_t pairs, float sigma,
                    const float* obs, const float* target, int32_t B, float* returns_plus, float* returns_minus,
                    float* bc_plus, float* bc_minus, int32_t bc_obs, int32_t bc_dim, float* centre_return_out,
                    int n_signs, cudaStream_t stream, const char* who) {
  EvalF16Params p = {};
  p.desc = *desc; p.theta = theta; p.table = table; p.table16 = table16;
  p.offsets = offsets; p.order = order;
  p.pairs = pairs; p.sigma = sigma; p.obs = obs; p.target = target; p.B = B;
  p.ret_plus = returns_plus; p.ret_minus = returns_minus;
  p.bc_plus = bc_plus; p.bc_minus = bc_minus; p.bc_obs = bc_obs; p.bc_dim = bc_dim;
  p.n_signs = n_signs; p.centre_out = centre_return_out;
  return run_f16(ctx, p, stream, who);
}

#ifdef ESTK_TC_PROFILE
// triage builds only (not part of estk.h): read and clear the role counters
extern "C" __attribute__((visibility("default"))) int estk_debug_f16_profile(unsigned long long* host_out, int n) {
  if (n > 32) n = 32;
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(host_out, g_f16_prof, sizeof(unsigned long long) * n) != cudaSuccess) return -1;
  unsigned long long zeros[32] = {};
  return cudaMemcpyToSymbol(g_f16_prof, zeros, sizeof(zeros)) == cudaSuccess ? 0 : -1;
}
#endif
