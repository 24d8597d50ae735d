// gemm_tf32x3.cuh — f32-accurate dense contraction on the 5th-gen tensor cores (tcgen05, sm_100a).
//
//     C[M][N] = A[M][K] · B[N][K]ᵀ          (both operands K-major = the llama2.c weight layout)
//
// Used where the decode path stops being a GEMV: prompt prefill (A = the prompt's activations,
// B = a weight matrix) and multi-sequence batched decode (A = a 128-row weight slab, B = the
// batch's activations) — SURVEY §8(f) rows 1-2; the reference has only Device::matmul with
// o_cols = 1 (cpu.rs:127-153) called once per token (mod.rs:187-192).
//
// f32 accuracy on tf32 tensor cores ("3xTF32"): every f32 operand element a is split into
//     hi = rna_tf32(a)   (11 significant bits, exactly representable in tf32)
//     lo = a − hi        (exact in f32; ≤ 13 significant bits, |lo| ≤ 2⁻¹¹|a|)
// and the product is accumulated in f32 as  lo·hi' + hi·lo' + hi·hi'  (the lo·lo' term, ≤ 2⁻²² relative,
// is dropped; small terms are added first).  The split is done IN THE KERNEL on the shared-memory tile
// that TMA delivered, so weights are still read from HBM exactly once as plain f32.
//
// Structure (one 128×BN output tile per CTA, 10 warps):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d of the raw A/B k-blocks (SWIZZLE_128B)
//   warp 1      TMEM allocation + MMA issue: one elected lane issues tcgen05.mma.kind::tf32
//               (3 per 8-wide k-step), tcgen05.commit releases the stage / publishes an accumulator
//   warps 2-9   workers: per landed stage they (a) read their row of the A tile from shared memory,
//               split it and tcgen05.st BOTH halves into TMEM — the MMA takes A from TMEM, so the
//               128-row operand is read from shared memory once instead of six times —, (b) write the
//               lo half of the B tile next to the raw one (the raw f32 tile itself is the hi operand: the
//               tensor core ignores the 13 low mantissa bits, verified against an explicit split);
//               and they own the second-level accumulation (below) and the fused epilogue
//   mbarriers per stage: full (TMA bytes) → ready (split done) → empty (MMA finished reading).
// Shared-memory bandwidth, not the tensor pipe, is what bounds an f32-split GEMM (every k-block is written by
// TMA, read and partly re-written by the split, then read by 3 MMAs per k-step); the A-through-TMEM path cuts
// that traffic from 192 KB to 128 KB per k-block at BN = 128 and from 144 KB to 80 KB at BN = 64.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace rama {

// epilogues that re-tile their 32×32 block through shared memory declare a `float* scratch` member (tp_exchange.cuh EpiPushNT)
template <class T, class = void>
struct epi_has_scratch { static constexpr bool value = false; };
template <class T>
struct epi_has_scratch<T, decltype((void)((T*)nullptr)->scratch)> { static constexpr bool value = true; };

constexpr int kGemmBM = 128;
constexpr int kGemmWorkerWarps = 8;
constexpr int kGemmThreads = (2 + kGemmWorkerWarps) * kWarp;  // 320

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {  // one non-blocking probe
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// L2 prefetch of one box (no shared-memory destination, no barrier): HBM → L2 ahead of the TMA load that will need it
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// D[tmem] (+)= A[smem desc] · B[smem desc], kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand in tensor memory (lane = row, 8 consecutive 32-bit columns = the k-step)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers → 32 TMEM lanes (this warp's quarter) × 16 consecutive columns
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 TMEM lanes (this warp's quarter) × 32 consecutive 32-bit columns → 32 registers per thread
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- CTA pair (cta_group::2): two CTAs of a cluster on the two SMs of a TPC issue ONE MMA of M = 256 ----------------------
// mbarrier arrive on the barrier at the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar, unsigned rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
// wait on a local barrier whose arrivals may come from the peer CTA (cluster-scope acquire)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAITC_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONEC_%=;\n\t"
      "bra WAITC_%=;\n\t"
      "DONEC_%=:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst) {  // one whole warp in EACH CTA of the pair, same smem_dst
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {  // one whole warp in each CTA
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// D[tmem, both CTAs: 128 rows each] (+)= A[tmem, 128 rows per CTA] · B[smem desc: N/2 rows in each CTA], issued by the leader
__device__ __forceinline__ void umma_tf32_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((unsigned short)3)
               : "memory");
}

// K-major shared-memory matrix descriptor (sm_100 format): rows of BK·4 bytes, 8-row groups of
// 8·BK·4 bytes back to back, swizzle = row bytes (64 B or 128 B).  `addr` is the tile base
// (1024-byte aligned) plus the k-step offset inside the swizzle row.
template <int BK>
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t addr) {
  constexpr uint64_t row_bytes = BK * 4;
  static_assert(row_bytes == 64 || row_bytes == 128, "BK must be 16 or 32 floats");
  constexpr uint64_t layout = row_bytes == 128 ? 2 : 4;  // SWIZZLE_128B = 2, SWIZZLE_64B = 4
  constexpr uint64_t sbo = (8 * row_bytes) >> 4;         // stride between 8-row groups
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) /* LBO (unused for swizzled K-major) */ | (sbo << 32) |
         (1ull << 46) /* descriptor version: Blackwell */ | (layout << 61);
}

template <int BN, int BM = kGemmBM>
__host__ __device__ constexpr uint32_t umma_idesc_tf32() {
  // c_format F32 (bit 4), a/b_format TF32 = 2 (bits 7-9, 10-12), both K-major, N>>3 at 17, M>>4 at 24
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ---- kernel parameters ------------------------------------------------------------------------------
struct GemmMaps {  // TMA descriptors; grouped launches pick a[group] or b[group], dual tiles stack b[0] on b[1]
  CUtensorMap a[3];
  CUtensorMap b[3];
};

struct GemmShape {
  int M, N, K;     // logical problem (per group); rows ≥ M / cols ≥ N of a tile are masked
  int hi_round;    // 1: also rewrite the B tile's hi half rounded to nearest (default 0: raw tile = hi by truncation)
  int ksplit;      // split-K: blockIdx.z = group·ksplit + split; split s reduces k-blocks [s·nkb/ksplit, (s+1)·nkb/ksplit)
  int group_on_a;  // grouped launch: 1 = groups differ in the A operand (weights as A: batched decode), 0 = in B
  int pf_ahead;    // > 0: the producer L2-prefetches the A box (weights streaming from HBM) this many k-blocks ahead
  // pf_rows > 0 (instead of pf_ahead): the idle lanes of the producer warp prefetch the A rows into L2 in contiguous bursts of
  // pf_rows k-blocks (pf_rows · 128 bytes per row), one burst ahead of the TMA loads.  A TMA box touches 128 rows × 128 bytes —
  // 128 different DRAM pages for 16 KB, and the next 128 bytes of each row only a k-block period later, when the page is long
  // closed; the bursts make HBM see ≥ 1 KB per opened page and the boxes then hit L2.
  int pf_rows;
  const float* a_base[3];  // the A matrices (per group) as plain pointers, for the bursts
  long long a_ld;          // their row pitch in floats
  long long* trace;  // debug (tools/gemm_trace.py): CTA (0,0,0) stamps clock64() per k-block and role, [128][8]; else nullptr
};

constexpr int kGemmBK = 32;  // floats per k-block = one 128-byte swizzle row

// AS > 0 ("decoupled A ring", batched decode where the A operand = weights streams from HBM): the raw A k-blocks live in their
// own ring of AS slots that a slot leaves as soon as the workers have split it into TMEM, while the B tile (raw + lo) and the TMEM
// A slot keep the STAGES-deep ring that the MMAs release.  The long-latency HBM loads then run AS k-blocks ahead inside the same
// shared-memory budget (AS = 4, STAGES = 2 at BN = 64: 64 + 32 KB = the 96 KB of two coupled 48 KB... stages).
// PS = 1 ("pre-split B", batched decode): the B operand arrives already split by the kernel that produced it — a [2·BN][K]
// matrix whose rows 0..BN-1 are the raw activations (= the hi operand: the tensor core ignores the 13 low mantissa bits) and
// rows BN..2BN-1 the exact remainders lo = x − trunc_tf32(x).  One TMA box of 2·BN rows per k-block, no worker touches B, and
// the two products that share A_hi are ONE instruction of width 2·BN (a small-N tcgen05.mma has a floor of ≈55 clk where
// N = 128 costs 64): per 8-wide k-step  acc[:, 0:2BN] += A_hi · [B_hi | B_lo],  acc[:, 0:BN] += A_lo · B_hi  — 119 clk instead
// of 3 × 55.  The epilogue adds the two column halves.  Requires the decoupled A ring (AS > 0) and BN = 64.
// PAIR = 1 (prefill, BN = 128): a CTA pair computes a 256×BN tile with tcgen05.mma.cta_group::2 — each CTA keeps its own 128 rows
// of A in TMEM and only HALF of the B k-block (BN/2 rows: raw + lo) in shared memory; the pair's tensor cores share the halves.
// Shared-memory bytes per k-block and SM fall from 128 KB to 80 KB (TMA 24, worker reads 24, B_lo 8, MMA 24), below what the
// MMAs need (the single-CTA tile is bound by the shared-memory port at 109 of 128 B/clk).
template <int BN, int STAGES, int NX = 0, int AS = 0, int PS = 0, int NACC = 2, int PAIR = 0>
struct GemmSmem {
  static constexpr int kABytes = kGemmBM * kGemmBK * 4;         // raw A k-block (only the workers read it)
  static constexpr int kBBytes = (PS ? 2 * BN : (PAIR ? BN / 2 : BN)) * kGemmBK * 4;  // raw B k-block = hi operand (PS: hi rows then lo rows; PAIR: this CTA's half)
  static_assert(!PAIR || (AS == 0 && PS == 0 && NX == 0 && BN == 128 && NACC == 2), "CTA pair: the coupled BN = 128 tile");
  static constexpr int kTxBytes = kABytes + kBBytes;
  static constexpr int kStageBytes = PS ? kBBytes : (AS > 0 ? 2 * kBBytes : kABytes + 2 * kBBytes);  // (+ B lo); coupled: A in front of B
  static constexpr int kARingBytes = AS * kABytes;              // decoupled A ring in front of the stages
  static constexpr int kBOff = AS > 0 ? 0 : kABytes;            // B hi inside a stage
  static constexpr int kBarOff = kARingBytes + STAGES * kStageBytes;
  static constexpr int kNumBars = 3 * STAGES + 4 + 2 * AS;      // full/ready/empty per stage, accfull[2], accfree[2], a_full/a_empty[AS]
  static constexpr int kTotal = kBarOff + kNumBars * 8 + 16 + 1024 /* alignment slack */;
  // TMEM: [0,BN) acc 0 | [BN,2BN) acc 1 | NX cross-term accumulators | then per stage 32 columns A hi + 32 columns A lo
  static constexpr int kAccN = PS ? 2 * BN : BN;                // columns of one accumulator
  // NACC = 1: a single accumulator instead of the ping-pong pair (the MMAs of the next chunk wait for the drain; meant for
  // two CTAs per SM, whose pipelines fill each other's bubbles) — halves the accumulator columns
  static constexpr int kAccCols = (NACC + NX) * kAccN;
  static_assert(NACC == 1 || NACC == 2, "one accumulator or a ping-pong pair");
  static constexpr int kTmemNeed = kAccCols + 64 * STAGES;
  static_assert(!PS || (AS > 0 && NX == 0 && BN == 64), "pre-split B: decoupled A ring, BN = 64, no cross accumulators");
  static_assert(kTmemNeed <= 512, "accumulators + A stages exceed tensor memory");
  static_assert(AS % 2 == 0, "the two worker groups alternate k-blocks: an A slot must always belong to the same group");
  static constexpr int kTmemCols = kTmemNeed <= 128 ? 128 : (kTmemNeed <= 256 ? 256 : 512);  // power of two
  // two CTAs per SM when both shared memory and tensor memory allow it: their pipelines interleave
  static constexpr int kCtasPerSm = (kTmemCols <= 256 && 2 * kTotal <= 226 * 1024) ? 2 : 1;
};

// Accumulation.  The tensor core adds into its f32 TMEM accumulator with truncation, which drifts
// linearly with the chain length (measured: 1.4e-4 relative at K = 4096 in one chain — far outside f32).
// So the chain is kept short and the long sum is done by CUDA cores in round-to-nearest: a CHUNK of CH
// k-blocks accumulates into one of two ping-pong TMEM accumulators; at the end of a chunk the workers
// tcgen05.ld it and add it into per-thread f32 registers while the tensor core already fills the other.
// Chunk partials have random signs, so the truncation bias (≈ 3·CH·4·½ ulp of a partial) no longer adds up
// coherently: CH = 2 (64 k) keeps the result within ~1.5e-6 relative, the class of an f32 FMA GEMM.
//
// Independent accumulators (NX > 0, an experiment kept for the record): NX cross-term accumulators next to the ping-pong
// pair let the issue order rotate over accumulators (NX = 2: a_lo·b_hi → X0, a_hi·b_lo → X1, a_hi·b_hi → main), so
// consecutive MMAs never accumulate into the tile the previous one wrote.  Measured: no change at BN = 64 (0.082 ms
// per 128×64×4096 tile either way) — back-to-back dependent accumulation is NOT what bounds the k-block — and at
// BN = 128 the two stages that fit next to three accumulators are slower than four.  Production uses NX = 0.

// Epilogue functor interface (device):
//   static constexpr bool kDual        — the B tile stacks BN/2 rows of b[0] on BN/2 rows of b[1]; the epilogue
//                                        receives both accumulator chunks of a column (SwiGLU)
//   void operator()(int m, int n, const float (&v)[32], int group, int split, bool valid)    (!kDual)
//   void operator()(int m, int n, const float (&v0)[32], const float (&v1)[32], bool valid)  (kDual)
// where m is the global row, n the first global column of the 32-column chunk; the whole warp calls it
// (valid = m < M) so an epilogue may shuffle between rows.

template <int BN, int STAGES, int CH, int NX, class Epi, int AS = 0, int PS = 0, int NACC = 2, int PAIR = 0>
__global__ void __launch_bounds__(kGemmThreads, (GemmSmem<BN, STAGES, NX, AS, PS, NACC, PAIR>::kCtasPerSm))
gemm_tf32x3_kernel(const __grid_constant__ GemmMaps maps, const GemmShape shp, const Epi epi_in) {
  using SM = GemmSmem<BN, STAGES, NX, AS, PS, NACC, PAIR>;
  // CTA pair: rank inside the 2-CTA cluster (0 = the leader, which issues the MMAs for both)
  const unsigned crank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = crank == 0;
  constexpr int kWorkersPerReady = (kGemmWorkerWarps / 2) * (PAIR ? 2 : 1);  // arrivals that complete a stage's `ready`
  constexpr int kWorkersPerFree = kGemmWorkerWarps * (PAIR ? 2 : 1);         // … an accumulator's `accfree`
  static_assert(NACC == 2 || NX == 0, "cross accumulators sit behind the ping-pong pair");
  constexpr int AN = SM::kAccN;  // accumulator width in TMEM columns
  constexpr int BK = kGemmBK;
  static_assert(BN == 64 || BN == 128, "BN");
  // the two worker groups take alternate k-blocks: with an even stage count a stage always belongs to the same group,
  // so no waiter ever skips a phase of a stage's mbarriers (parity waits alias after two phases)
  static_assert(STAGES % 2 == 0, "STAGES must be even");
  static_assert(!Epi::kDual || BN == 128, "dual epilogue needs BN = 128");
  constexpr int NSEG = BN / 64;  // 32-column segments per epilogue warp
  extern __shared__ uint8_t gemm_smem_raw[];
  const uint32_t base = (smem_u32(gemm_smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = base + SM::kBarOff;            // [STAGES] TMA bytes landed
  const uint32_t bar_ready = bar_full + 8 * STAGES;        // [STAGES] split done (A in TMEM, B lo in smem)
  const uint32_t bar_empty = bar_ready + 8 * STAGES;       // [STAGES] MMAs finished reading
  const uint32_t bar_accfull = bar_empty + 8 * STAGES;     // [2] chunk accumulated in acc[b]
  const uint32_t bar_accfree = bar_accfull + 16;           // [2] acc[b] drained into registers
  const uint32_t bar_afull = bar_accfree + 16;             // [AS] raw A k-block landed            (decoupled A ring only)
  const uint32_t bar_aempty = bar_afull + 8 * AS;          // [AS] the group's four warps have read it
  const uint32_t tmem_slot = bar_aempty + 8 * AS;
  volatile int* pf_progress = reinterpret_cast<volatile int*>(gemm_smem_raw + (base - smem_u32(gemm_smem_raw)) + SM::kBarOff + SM::kNumBars * 8 + 8);
  const uint32_t stage0 = base + SM::kARingBytes;          // first (B) stage
  uint8_t* gen_base = gemm_smem_raw + (base - smem_u32(gemm_smem_raw));

  pdl_launch_dependents();  // (no-ops unless launched with programmatic stream serialization)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.z / shp.ksplit, split = blockIdx.z - group * shp.ksplit;
  // M-tiles vary fastest in launch order: the CTAs that share a B (weight) tile run together, so with M = 512
  // prompt rows the four readers of a weight tile hit L2 and the matrix leaves HBM once (ncu before: 4×)
  const int m0 = blockIdx.x * kGemmBM;
  const int tile_n = blockIdx.y;
  const int total_kb = (shp.K + BK - 1) / BK;
  const int kb_begin = (int)((long long)split * total_kb / shp.ksplit);
  const int num_kb = (int)((long long)(split + 1) * total_kb / shp.ksplit) - kb_begin;  // may be 0: the tile is all zeros
  const int num_ch = (num_kb + CH - 1) / CH;
  constexpr int kBoxN = Epi::kDual ? BN / 2 : BN;  // rows per B box
  const bool tr = shp.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define RAMA_GEMM_TR(kb_, e_) do { if (tr && (kb_) < 128) shp.trace[(kb_) * 8 + (e_)] = clock64(); } while (0)
  const CUtensorMap* mapA = &maps.a[shp.group_on_a ? group : 0];
  const CUtensorMap* mapB0 = &maps.b[(Epi::kDual || shp.group_on_a) ? 0 : group];

  if (warp == 0 && lane == 0) {
    *pf_progress = 0;
    tma_prefetch_desc(mapA);
    tma_prefetch_desc(mapB0);
    if (Epi::kDual) tma_prefetch_desc(&maps.b[1]);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_ready + 8 * s, kWorkersPerReady);  // the four warps of the group that owns the k-block (pair: of both CTAs)
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_accfull + 8 * b, 1);
      mbar_init(bar_accfree + 8 * b, kWorkersPerFree);
    }
    for (int a = 0; a < AS; ++a) {
      mbar_init(bar_afull + 8 * a, 1);
      mbar_init(bar_aempty + 8 * a, kGemmWorkerWarps / 2);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair<SM::kTmemCols>(tmem_slot);
    else tmem_alloc<SM::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anyone arrives on them remotely
  else __syncthreads();
  tc_fence_after();
  pdl_wait();  // the operands (and the buffers the epilogue overwrites) belong to the previous kernel until here
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen_base + SM::kBarOff + SM::kNumBars * 8);
  const uint32_t tmem_a = tmem + SM::kAccCols;  // A stages: stage s at +64·s (hi), +64·s+32 (lo)

  if (warp == 0) {
    // ===== TMA producer =====
    if (shp.pf_rows > 0 && lane > 0) {  // lanes 1..31: row bursts into L2, one burst ahead of lane 0's loads
      const float* ab = shp.a_base[shp.group_on_a ? group : 0];
      const int R = shp.pf_rows;
      const int nburst = (num_kb + R - 1) / R;
      for (int c = 0; c < nburst; ++c) {
        while (c >= 2 && *pf_progress < (c - 1) * R) __nanosleep(64);  // bursts 0 and 1 right away
        const int k0 = (kb_begin + c * R) * BK;
        const int kn = min(R * BK, shp.K - k0);
        if (kn > 0)
          for (int r = lane - 1; r < kGemmBM; r += 31)
            if (m0 + r < shp.M) l2_prefetch_bulk(ab + (size_t)(m0 + r) * (size_t)shp.a_ld + k0, (uint32_t)kn * 4u);
      }
    }
    // (the role loops are single threads on the critical path of a k-block: running counters, no divisions)
    if constexpr (AS > 0) {
      // two independent rings fed by one polling thread: A (weights, HBM latency) runs up to AS k-blocks ahead and is gated by
      // the workers; B (activations, L2) is gated by the MMAs like a coupled stage
      if (lane == 0) {
        static_assert(!Epi::kDual || AS == 0, "decoupled A ring: plain B boxes only");
        const int nb0 = tile_n * kBoxN;
        int ka = 0, kbb = 0;
        if (shp.pf_ahead > 0)
          for (int j = AS; j < AS + shp.pf_ahead && j < num_kb; ++j) tma_prefetch_l2_2d(mapA, (kb_begin + j) * BK, m0);
        while (ka < num_kb || kbb < num_kb) {
          if (ka < num_kb) {
            const int sa = ka % AS;
            if (mbar_try(bar_aempty + 8 * sa, ((ka / AS) & 1) ^ 1)) {
              if (shp.pf_ahead > 0 && ka + AS + shp.pf_ahead < num_kb)
                tma_prefetch_l2_2d(mapA, (kb_begin + ka + AS + shp.pf_ahead) * BK, m0);
              RAMA_GEMM_TR(ka, 0);
              mbar_arrive_expect_tx(bar_afull + 8 * sa, SM::kABytes);
              tma_load_2d(base + sa * SM::kABytes, mapA, bar_afull + 8 * sa, (kb_begin + ka) * BK, m0);
              ++ka;
              *pf_progress = ka;
            }
          }
          if (kbb < num_kb) {
            const int sb = kbb % STAGES;
            if (mbar_try(bar_empty + 8 * sb, ((kbb / STAGES) & 1) ^ 1)) {
              mbar_arrive_expect_tx(bar_full + 8 * sb, SM::kBBytes);
              tma_load_2d(stage0 + sb * SM::kStageBytes, mapB0, bar_full + 8 * sb, (kb_begin + kbb) * BK, nb0);
              ++kbb;
            }
          }
        }
      }
    } else
    if (lane == 0) {
      const int nb0 = tile_n * kBoxN;
      int s = 0, kc = kb_begin * BK;
      uint32_t ph = 1, st = base, bf = bar_full, be = bar_empty;
      if (shp.pf_ahead > 0)  // the first boxes beyond the ring
        for (int j = STAGES; j < STAGES + shp.pf_ahead && j < num_kb; ++j) tma_prefetch_l2_2d(mapA, kc + j * BK, m0);
      for (int kb = 0; kb < num_kb; ++kb) {
        if (shp.pf_ahead > 0 && kb + STAGES + shp.pf_ahead < num_kb)
          tma_prefetch_l2_2d(mapA, kc + (STAGES + shp.pf_ahead) * BK, m0);
        mbar_wait(be, ph);
        RAMA_GEMM_TR(kb, 0);
        *pf_progress = kb;
        mbar_arrive_expect_tx(bf, SM::kTxBytes);
        tma_load_2d(st, mapA, bf, kc, m0);
        if constexpr (PAIR) {  // this CTA's half of the B k-block: BN/2 rows (dual: rank 0 takes b[0], rank 1 takes b[1])
          if (Epi::kDual) tma_load_2d(st + SM::kABytes, &maps.b[crank], bf, kc, nb0);
          else tma_load_2d(st + SM::kABytes, mapB0, bf, kc, tile_n * BN + (int)crank * (BN / 2));
        } else if (Epi::kDual) {
          tma_load_2d(st + SM::kABytes, &maps.b[0], bf, kc, nb0);
          tma_load_2d(st + SM::kABytes + kBoxN * BK * 4, &maps.b[1], bf, kc, nb0);
        } else {
          tma_load_2d(st + SM::kABytes, mapB0, bf, kc, nb0);
        }
        kc += BK; st += SM::kStageBytes; bf += 8; be += 8;
        if (++s == STAGES) { s = 0; ph ^= 1; st = base; bf = bar_full; be = bar_empty; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // One thread feeds the tensor core, so its instruction stream is on the critical path: descriptors are
    // (constant high word, base low word + a small offset) — one 32-bit add per operand.
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = PAIR ? umma_idesc_tf32<BN, 2 * kGemmBM>() : umma_idesc_tf32<BN>();
      constexpr uint32_t idesc2 = umma_idesc_tf32<2 * BN <= 256 ? 2 * BN : BN>();  // PS: A_hi · [B_hi | B_lo]
      const uint64_t d0 = umma_smem_desc<BK>(stage0 + SM::kBOff);  // B hi of stage 0, k-step 0
      const uint32_t d_hi32 = (uint32_t)(d0 >> 32), d_lo32 = (uint32_t)d0;
      auto desc = [&](uint32_t lo) { return ((uint64_t)d_hi32 << 32) | lo; };
      int s = 0, ch = 0, in_ch = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const bool first = in_ch == 0, last = (in_ch == CH - 1) || kb == num_kb - 1;
        const int ab = NACC == 2 ? (ch & 1) : 0;  // accumulator of this chunk
        const uint32_t acc = tmem + ab * AN;
        if (first && ch >= NACC) {  // it still holds chunk ch-NACC until the workers have drained it
          if constexpr (PAIR) mbar_wait_cluster(bar_accfree + 8 * ab, ((ch >> 1) - 1) & 1);
          else mbar_wait(bar_accfree + 8 * ab, (NACC == 2 ? (ch >> 1) - 1 : ch - 1) & 1);  // (the fence after the `ready` wait below covers this one too)
          RAMA_GEMM_TR(kb, 7);
        }
        if constexpr (PAIR) mbar_wait_cluster(bar_ready + 8 * s, ph);
        else mbar_wait(bar_ready + 8 * s, ph);
        tc_fence_after();
        RAMA_GEMM_TR(kb, 3);
        const uint32_t bh = d_lo32 + s * (SM::kStageBytes >> 4), bl = bh + (SM::kBBytes >> 4);
        const uint32_t a_hi = tmem_a + 64 * s, a_lo = a_hi + 32;
        // 8 tf32 = 32 bytes = 2 descriptor units inside the swizzle row; TMEM A: 8 columns per k-step
        if constexpr (PAIR) {
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            umma_tf32_ts_pair(acc, a_lo + 8 * ks, desc(bh + 2 * ks), idesc, !(first && ks == 0));  // small terms first
            umma_tf32_ts_pair(acc, a_hi + 8 * ks, desc(bl + 2 * ks), idesc, 1);
            umma_tf32_ts_pair(acc, a_hi + 8 * ks, desc(bh + 2 * ks), idesc, 1);
          }
        } else if constexpr (PS) {
          // Measured (role timeline, 8 MMAs per k-block): alternating N = 128 / N = 64 instructions issue at ≈82 clk each — a
          // change of instruction shape costs a bubble.  So the four wide products go first, then the narrow ones (PS = 1), or
          // every product is wide (PS = 2: A_lo · [B_hi | B_lo] also adds the lo·lo term — 64 clk instead of 55, one shape).
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks)
            umma_tf32_ts(acc, a_hi + 8 * ks, desc(bh + 2 * ks), idesc2, !(first && ks == 0));  // both column halves (initialises them)
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks)
            umma_tf32_ts(acc, a_lo + 8 * ks, desc(bh + 2 * ks), PS == 2 ? idesc2 : idesc, 1);  // PS = 1: rows 0..BN-1 of the tile = B_hi
        } else if constexpr (NX == 0) {
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            umma_tf32_ts(acc, a_lo + 8 * ks, desc(bh + 2 * ks), idesc, !(first && ks == 0));  // small terms first
            umma_tf32_ts(acc, a_hi + 8 * ks, desc(bl + 2 * ks), idesc, 1);
            umma_tf32_ts(acc, a_hi + 8 * ks, desc(bh + 2 * ks), idesc, 1);
          }
        } else if constexpr (NX == 2) {  // rotate X0, X1, main: every accumulator is reused at distance 3
          const uint32_t x0 = tmem + 2 * BN, x1 = x0 + BN;
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            umma_tf32_ts(x0, a_lo + 8 * ks, desc(bh + 2 * ks), idesc, (kb | ks) != 0);
            umma_tf32_ts(x1, a_hi + 8 * ks, desc(bl + 2 * ks), idesc, (kb | ks) != 0);
            umma_tf32_ts(acc, a_hi + 8 * ks, desc(bh + 2 * ks), idesc, !(first && ks == 0));
          }
        } else {  // NX == 1: X M X M X M X M X X X X — main and the cross accumulator alternate while main lasts
          static_assert(BK / 8 == 4, "issue order written out for 4 k-steps");
          const uint32_t x0 = tmem + 2 * BN;
          auto XA = [&](int ks, uint32_t accum) { umma_tf32_ts(x0, a_lo + 8 * ks, desc(bh + 2 * ks), idesc, accum); };
          auto XB = [&](int ks) { umma_tf32_ts(x0, a_hi + 8 * ks, desc(bl + 2 * ks), idesc, 1); };
          auto MM = [&](int ks, uint32_t accum) { umma_tf32_ts(acc, a_hi + 8 * ks, desc(bh + 2 * ks), idesc, accum); };
          XA(0, kb != 0); MM(0, !first); XB(0); MM(1, 1); XA(1, 1); MM(2, 1); XB(1); MM(3, 1);
          XA(2, 1); XB(2); XA(3, 1); XB(3);
        }
        if constexpr (PAIR) {  // both CTAs' stages / accumulators
          umma_commit_pair(bar_empty + 8 * s);
          if (last) umma_commit_pair(bar_accfull + 8 * ab);
        } else {
          umma_commit(bar_empty + 8 * s);  // smem stage + TMEM A stage reusable once these MMAs have read them
          if (last) umma_commit(bar_accfull + 8 * ab);
        }
        RAMA_GEMM_TR(kb, 4);
        if (++s == STAGES) { s = 0; ph ^= 1; }
        if (++in_ch == CH) { in_ch = 0; ++ch; }
      }
    }
  } else {
    // ===== workers: operand split, second-level accumulation, epilogue =====
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;        // which half of the k-block (split) / of the tile's columns (drain)
    const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
    // column (inside the tile) of this warp's segment j
    auto seg_col = [&](int j) { return Epi::kDual ? j * (BN / 2) + 32 * half : half * (BN / 2) + 32 * j; };
    float acc[NSEG][32];
#pragma unroll
    for (int j = 0; j < NSEG; ++j)
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[j][i] = 0.f;

    int drained = 0;
    auto drain = [&](int c) {  // acc += acc_tmem[c&1] (chunk c), then hand the buffer back
      const int b = NACC == 2 ? (c & 1) : 0;
      if (quarter == 0 && lane == 0) RAMA_GEMM_TR(2 * c + half, 5);
      mbar_wait(bar_accfull + 8 * b, (NACC == 2 ? (c >> 1) : c) & 1);
      tc_fence_after();
      if (quarter == 0 && lane == 0) RAMA_GEMM_TR(2 * c + half, 6);
#pragma unroll
      for (int j = 0; j < NSEG; ++j) {
        float v[32];
        tmem_ld_32x32(trow + b * AN + seg_col(j), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[j][i] += v[i];
        if constexpr (PS) {  // the A_hi · B_lo products live in the upper column half
          tmem_ld_32x32(trow + b * AN + BN + seg_col(j), v);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[j][i] += v[i];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && !leader) mbar_arrive_cluster(bar_accfree + 8 * b, 0);  // the leader's issuer counts both CTAs' drains
        else mbar_arrive(bar_accfree + 8 * b);
      }
    };

    // The two groups of four worker warps (half = 0 / 1) take alternate k-blocks, so the latency chain of one
    // k-block (LDS → split → STTM/STS → proxy fence → tcgen05.wait::st → arrive) overlaps the other group's.
    // A: this thread owns row (quarter·32 + lane) of the tile; SWIZZLE_128B puts 16-byte chunk c of row r at
    // r·128 + ((c ^ (r & 7)) · 16).  B: the group's 128 threads split the raw tile element-wise.
    const int arow = quarter * 32 + lane;
    const uint32_t arow_off = arow * 128;
    const int gt = quarter * 32 + lane;  // thread index inside the group
    constexpr int kBVec = SM::kBBytes / 16;
    constexpr int kGroupThreads = kGemmWorkerWarps / 2 * kWarp;
    constexpr int kLag = 1;  // k-blocks split beyond a chunk's end before draining it
    int s = half;      // this group's stage walks half, half+2, … (STAGES is even)
    uint32_t ph = 0;
    int next_end = min(CH, num_kb) - 1 + kLag;  // k-block after which chunk `drained` may be drained
    for (int kb = half; kb < num_kb; kb += 2) {
      mbar_wait(bar_full + 8 * s, ph);  // (decoupled: the B tile — and with it the TMEM A slot, freed by the same MMAs)
      const uint8_t* st = gen_base + SM::kARingBytes + s * SM::kStageBytes;
      const uint8_t* sta = st;          // raw A k-block
      if constexpr (AS > 0) {
        const int sa = kb % AS;
        mbar_wait(bar_afull + 8 * sa, (kb / AS) & 1);
        sta = gen_base + sa * SM::kABytes;
      }
      if (quarter == 0 && lane == 0) RAMA_GEMM_TR(kb, 1);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {  // the row in two halves of 16 floats
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int chunk = 4 * hh + c;
          const float4 a = *reinterpret_cast<const float4*>(sta + arow_off + ((chunk ^ (arow & 7)) << 4));
          const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            uint32_t h;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(av[e]));
            hi[4 * c + e] = h;
            lo[4 * c + e] = __float_as_uint(av[e] - __uint_as_float(h));
          }
        }
        const uint32_t ta = trow + SM::kAccCols + 64 * s + 16 * hh;
        tmem_st_32x16(ta, hi);
        tmem_st_32x16(ta + 32, lo);
      }
      if constexpr (AS > 0) {  // the raw A slot is in registers / on its way to TMEM: hand it back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_aempty + 8 * (kb % AS));
      }
      if constexpr (!PS) {
        const float4* braw = reinterpret_cast<const float4*>(st + SM::kBOff);
        float4* bhi = reinterpret_cast<float4*>(const_cast<uint8_t*>(st) + SM::kBOff);
        float4* blo = reinterpret_cast<float4*>(const_cast<uint8_t*>(st) + SM::kBOff + SM::kBBytes);
#pragma unroll
        for (int i = gt; i < kBVec; i += kGroupThreads) {
          const float4 a = braw[i];
          float4 h, l;
          if (shp.hi_round) {
            uint32_t hx, hy, hz, hw;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hx) : "f"(a.x));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hy) : "f"(a.y));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hz) : "f"(a.z));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hw) : "f"(a.w));
            h = make_float4(__uint_as_float(hx), __uint_as_float(hy), __uint_as_float(hz), __uint_as_float(hw));
            bhi[i] = h;
          } else {  // what the tensor core will see of the raw tile
            h.x = __uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u); h.y = __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u);
            h.z = __uint_as_float(__float_as_uint(a.z) & 0xFFFFE000u); h.w = __uint_as_float(__float_as_uint(a.w) & 0xFFFFE000u);
          }
          l.x = a.x - h.x; l.y = a.y - h.y; l.z = a.z - h.z; l.w = a.w - h.w;
          blo[i] = l;
        }
      }
      if constexpr (!PS) fence_proxy_async_smem();  // generic-proxy writes → visible to the tensor core's async proxy
      tmem_st_wait();            // A halves have landed in TMEM
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && !leader) mbar_arrive_cluster(bar_ready + 8 * s, 0);  // the leader's MMA reads both CTAs' halves
        else mbar_arrive(bar_ready + 8 * s);
      }
      if (quarter == 0 && lane == 0) RAMA_GEMM_TR(kb, 2);
      s += 2;
      if (s >= STAGES) { s -= STAGES; ph ^= 1; }
      // chunk c ends with k-block min((c+1)·CH, num_kb) − 1; drain it once this group is kLag blocks past it
      while (drained < num_ch && next_end <= kb) {
        drain(drained++);
        next_end = min((drained + 1) * CH, num_kb) - 1 + kLag;
      }
    }
    while (drained < num_ch) drain(drained++);
    if constexpr (NX > 0) {
      // cross terms: complete once the last chunk's commit has fired (tcgen05.commit covers all prior MMAs)
      if (num_kb > 0) {
#pragma unroll
        for (int x = 0; x < NX; ++x)
#pragma unroll
          for (int j = 0; j < NSEG; ++j) {
            float v[32];
            tmem_ld_32x32(trow + (2 + x) * BN + seg_col(j), v);
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[j][i] += v[i];
          }
        tc_fence_before();
      }
    }

    Epi epi = epi_in;
    // (every MMA has completed — the last chunk's commit fired — so the pipeline stages are free: 4 KB of stage 0 per warp)
    if constexpr (epi_has_scratch<Epi>::value) epi.scratch = reinterpret_cast<float*>(gen_base + SM::kARingBytes + (warp - 2) * 4096);
    const int m = m0 + quarter * 32 + lane;
    const bool valid = m < shp.M;
    if constexpr (Epi::kDual) {
      epi(m, tile_n * (BN / 2) + 32 * half, acc[0], acc[1], valid);
    } else {
#pragma unroll
      for (int j = 0; j < NSEG; ++j) epi(m, tile_n * BN + seg_col(j), acc[j], group, split, valid);
    }
  }
  if constexpr (PAIR) {
    tc_fence_before();
    cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while the other may still signal it or read its B half
    if (warp == 1) {
      tc_fence_after();
      tmem_dealloc_pair<SM::kTmemCols>(tmem);
    }
  } else {
    __syncthreads();
    if (warp == 1) {
      tc_fence_after();
      tmem_dealloc<SM::kTmemCols>(tmem);
    }
  }
}

#undef RAMA_GEMM_TR

// ---- epilogues ------------------------------------------------------------------------------------------

// C[m][n] row-major with leading dimension ldc (per group: base + group·group_stride)
struct EpiStoreNT {
  static constexpr bool kDual = false;
  float* c;
  int ldc, N;
  size_t group_stride;
  __device__ __forceinline__ void operator()(int m, int n, const float (&v)[32], int group, int, bool valid) const {
    if (!valid) return;
    float* row = c + group * group_stride + (size_t)m * ldc + n;
    if (n + 32 <= N && ((reinterpret_cast<uintptr_t>(row) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(row)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n + j < N) row[j] = v[j];
    }
  }
};

// Cᵀ store of split-K partials: out[group][split][n][ldc] at column m (batched decode: the tile rows are
// weight rows r = m, the columns are sequences b = n; consecutive lanes hold consecutive r ⇒ coalesced)
struct EpiStoreT {
  static constexpr bool kDual = false;
  float* c;
  int ldc, N;          // N = number of sequences
  int ksplit;
  size_t slab;         // floats per (group, split) partial = N_max · ldc
  __device__ __forceinline__ void operator()(int m, int n, const float (&v)[32], int group, int split, bool valid) const {
    if (!valid) return;
    float* base = c + (size_t)(group * ksplit + split) * slab + m;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n + j < N) base[(size_t)(n + j) * ldc] = v[j];
  }
};

}  // namespace rama
