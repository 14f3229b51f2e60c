// afa_tc_kernels.cuh -- fused Activation1d (bf16 I/O) with BOTH FIR filters on the 5th-generation tensor cores
// (tcgen05.mma, accumulators and the A operands in TMEM), sm_100a.
//
// What it replaces (reference, /root/reference/BigVGAN): alias_free_activation/act.py:25-30 =
// resample.py:29-38 (UpSample1d) + activations.py:51-62 / 113-126 (Snake / SnakeBeta) + filter.py:94-101
// (LowPassFilter1d, stride 2), for dense row-major [rows = batch*channels, T] bf16 tensors with T % 8 == 0.
//
// Why: the register-walk kernel of afa_kernels.cuh spends 24 of its 27 instructions per output on the two
// 12-tap FIRs and is bound by the FP32 FMA pipe -- at bf16 I/O it reaches 0.29-0.44 of the HBM roofline
// (VERDICT round 1).  The FIRs are linear and shift invariant, so a block of 16 time steps of 128 rows is a
// product with a small CONSTANT banded Toeplitz matrix:
//
//   U_j  [128 rows x 32 u-values]  =  X_j  [128 x 16] * Wup_a [16 x 32]  +  X_j+1 [128 x 16] * Wup_b [16 x 32]
//   Y_i  [128 rows x 16 outputs ]  =  S_2i [128 x 16] * Wdn_a [16 x 16]  +  S_2i+1 * Wdn_b  +  S_2i+2 * Wdn_c
//
// with M = 128 rows of the tensor (one TMEM lane each), K = 16 consecutive time samples, N = the block's
// outputs.  Only Snake (2 MUFU + 5 FMA-pipe instructions per output) stays on the CUDA cores.
//   * taps are split hi + lo into two bf16 matrices (16 mantissa bits), both products accumulate in fp32 in
//     TMEM; x is bf16 by contract, so U is exact to fp32 rounding.  s is rounded to bf16 once, as the A operand
//     of the down filter (the same rounding step the bf16 output applies anyway).
//   * A operands come from TENSOR MEMORY: a lane's x slice is copied shared -> registers -> TMEM by the thread
//     that owns the lane (tcgen05.st; replicate padding is patched in registers on the way), and the activated
//     block S_j overwrites the first half of the accumulator U_j it was computed from.  Shared memory therefore
//     sees each byte of x and y exactly once plus the small B matrices: the MMAs' operand traffic would
//     otherwise exceed the 128 B/clk shared-memory port at the HBM roofline.
//   * a CTA owns 128 lanes = R rows x G time groups (R * G = 128; R = 16 for one binaural clip at C = 24) and
//     NY blocks of 16 outputs per lane.  x arrives as 64-sample (128-byte, 128B-swizzled) chunks by tensor-map
//     TMA, out-of-range samples and rows are zero-filled by the TMA unit; y leaves through the same chunks in
//     place (TMA store clips what lies outside the tensor).
//   * warp roles: warp 0 = TMA producer / TMEM allocator / TMA store; warp 1 = MMA issuer (one thread);
//     warps 2-9 = two compute groups of four warps (one thread per TMEM lane) that ping-pong over the blocks:
//     tcgen05.ld U_j -> Snake -> tcgen05.st S_j, next x slice -> TMEM, drain Y_j-2 -> shared memory.  All
//     hand-offs are mbarriers (tcgen05.commit on the MMA side); no CTA-wide barrier inside the block loop.
//
// Index algebra (t_org = the lane's first staged sample, t_org % 8 == 0, first output t_org + 8):
//   x slice k   : x_ext[t_org + 16k + (0..15)]
//   u block j   : u_ext[n0 + e], n0 = 2*t_org + 32j + 6, e = 0..31;  u[n] = 2 * sum_i f[n + 5 - 2i] * x_ext[i]
//                 -> Wup[kappa][e] = 2 f[e + 11 - 2 kappa],  kappa = 0..31 over slices j, j+1
//   y block i   : y[t_org + 8 + 16i + e], e = 0..15;  y[t] = sum_k f[k] * s_ext[2t + k - 5]
//                 -> Wdn[kappa][e] = f[kappa - 2e - 5],  kappa = 0..47 over the s slices of u blocks i, i+1
//   replicate padding: x_ext clamps in the 1x domain (chunks with t < 0 or t >= T, always whole 8-sample
//   chunks because T % 8 == 0); s_ext clamps the ACTIVATED signal in the 2x domain (filter.py:98): elements
//   e < 10 of block 0 of a row's first lane, and elements e >= e_b (e_b = 10 or 26) of the block holding 2T.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace afa_tc {

constexpr int kThreads = 320;            // warp 0: TMA producer / TMEM allocator / TMA store, warp 1: MMA, warps 2..9: compute
constexpr int kSlots = 5;                // shared-memory chunk ring: slots of 64 samples x 128 lanes, recycled along the strip
constexpr int kChunkBytes = 128 * 128;   // 128 lanes x 64 bf16
constexpr int kTmemCols = 256;           // 240 used; allocations are powers of two
constexpr int kRing = 5;                 // slots of the U / S and Y rings
constexpr int kColUS = 0;                // U / S ring: 5 slots x 32 columns (U fp32; S = 32 bf16 in the first 16)
constexpr int kColY = 160;               // Y ring: 5 slots x 16 columns
// shared memory carve-up (offsets from a 1024-byte aligned base)
constexpr int kOffWup = kSlots * kChunkBytes;           // [hi/lo][slice a/b] x (32 x 16 bf16 = 1024 B)
constexpr int kOffWdn = kOffWup + 4 * 1024;              // [hi/lo][slice a/b/c] x (16 x 16 bf16 = 512 B)
constexpr int kOffBar = kOffWdn + 6 * 512;
constexpr int kBarFull = 0, kBarPre = kSlots, kBarCmp = kBarPre + 1, kBarEv = kBarCmp + 8, kBarOut = kBarEv + 8;
constexpr int kNumBars = kBarOut + kSlots;
constexpr int kOffTmem = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmem + 16 + 1024;         // + slack for the manual 1024-byte alignment

struct Args {
    const __nv_bfloat16* x;
    const float* alpha;
    const float* beta;
    uint16_t up_hi[12], up_lo[12];   // 2 * upsample taps (ratio folded, resample.py:33) as bf16 hi + lo
    uint16_t dn_hi[12], dn_lo[12];   // low-pass taps as bf16 hi + lo
    int32_t rows, C, T, flags;
    int32_t R_log2;        // lanes = R rows x G groups, R = 1 << R_log2
    int32_t NY;            // y blocks (16 outputs) per lane and CTA: a multiple of 4, any length (the chunk ring is recycled)
    int32_t n_tstrips;     // CTAs along time; blockIdx.x = row_group * n_tstrips + tstrip
    int32_t debug;         // harness only: 1 = dump U blocks, 2 = dump S, 3 = clock stamps of CTA dbg_cta
    int32_t dbg_cta;
    int32_t dbg_blocks;    // harness: blocks per lane the U / S dump holds (NY + 1)
    float* dbg;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol or descriptor mistake must end in a trap, never in a hung GPU.  The suspend-time hint lets the
// hardware park the warp until the phase completes (or the hint expires) instead of re-issuing the poll: in the first versions
// a quarter of all issued instructions were these polls, on the sub-partitions that also host the MMA and TMA warps.
#ifndef AFA_TC_WAIT_HINT_NS
#define AFA_TC_WAIT_HINT_NS 4000
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity), "r"((uint32_t)AFA_TC_WAIT_HINT_NS) : "memory");
        if (spins > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]     (cute::SM100_MMA_F16BF16_TS)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[smem descriptor] * B[smem descriptor]     (cute::SM100_MMA_F16BF16_SS)
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, 128-byte swizzle (what the tensor-map TMA wrote): rows of 128 bytes, 8-row groups 1024 bytes apart (SBO); a K = 16
// slice starts 32 * k bytes into the row (the hardware applies the XOR to the address bits); the tile base is 1024-byte aligned
__device__ __forceinline__ uint64_t adesc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1u << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// K-major, no swizzle: element (n, k) at (k / 8) * LBO + n * 16 + (k % 8) * 2 bytes; SBO = 128 (8 rows x 16 B)
__device__ __forceinline__ uint64_t bdesc_kmajor(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// one lane of a converged warp (ptxas then issues the uniform-datapath tcgen05 instructions without a per-lane loop)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t clk32() {
    uint32_t c;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
    return c;
}
// harness timeline (debug == 3): dbg[(role * 32 + j) * 8 + slot] = clock, for CTA `dbg_cta`
#define AFA_TC_STAMP(role, j, slot) do { if (a.debug >= 3 && lane == 0 && (j) < 32 && blockIdx.x == (unsigned)a.dbg_cta) a.dbg[((role) * 32 + (j)) * 8 + (slot)] = __uint_as_float(clk32()); } while (0)
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// Snake / SnakeBeta: u + sin^2(alpha u) / (beta + 1e-9)            activations.py:60, :124
__device__ __forceinline__ float snake_f(float u, float a, float ib) {
    const float sn = __sinf(u * a);
    return fmaf(ib, sn * sn, u);
}

// Schedule.  Block j = 16 time steps of the CTA's 128 lanes.  Compute group g (4 warps) owns the blocks j = g (mod 2).
//   MMA thread, event j:        wait cmp[j & 7] (S(j) is in TMEM);  down(j-1) [S(j-1), S(j) from TMEM -> Y slot (j-1) % 5];
//                               up(j+4) [x slices j+4, j+5 straight from the swizzled shared-memory chunks -> U slot (j+4) % 5];
//                               ONE commit -> ev[(j+4) & 7]  (tcgen05.commit tracks every MMA issued before it)
//   iteration j of its group:   wait ev[j & 7]  (event j-4 committed: up(j) and down(j-5) are complete)
//                               drain Y(j-5) -> bf16 -> shared memory (out chunk (j-5)/4, over x slice j-5), arrive out[(j-5)/4]
//                               U(j) -> registers -> Snake -> S(j) over U(j) -> arrive cmp[j & 7]
//   warp 0:                     wait out[q] (16 warp arrivals) -> TMA store of out chunk q
// The lookahead of 4 blocks is what hides the MMA round trip (arrive -> issue -> execute -> commit -> wake, about one block
// time): with 3 both sides were found waiting on each other (profiles/r02_tc_ncu_full_fwd_bf16_B16_C384_T13776.txt).
// Ring safety: up(j+4) lands on U/S slot (j-1) % 5 = S(j-1), last read by down(j-1), issued immediately before it (MMAs of one
// thread execute in issue order).  down(i) lands on Y(i-5), drained at iteration i before that iteration's arrive on cmp[i & 7],
// which the MMA thread has consumed before event i+1.  y block i overwrites x slice i, last read by up(i), complete since event
// i-4.  cmp / ev are rings of 8 indexed by the block, not per group: a warp's ev wait depends on event j-4 only, so it may run
// iterations ahead of a slower warp of its group; it cannot be 8 blocks ahead, because event j+4 needs cmp[j] complete.
#ifndef AFA_TC_BOUND_THREADS
#define AFA_TC_BOUND_THREADS 384
#endif
__global__ void __launch_bounds__(AFA_TC_BOUND_THREADS, 2)
afa_tc_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
                  const __grid_constant__ Args a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 1) AFA_TC_STAMP(3, 1, 0);
    const int NY = a.NY;
    const int NCH_IN = (NY + 2 + 3) >> 2, NCH_OUT = NY >> 2;       // x chunks / out chunks of this CTA's strip
    const int R = 1 << a.R_log2, G = 128 >> a.R_log2;
    const int tstrip = (int)(blockIdx.x % (uint32_t)a.n_tstrips);
    const int rgroup = (int)(blockIdx.x / (uint32_t)a.n_tstrips);
    const int row0 = rgroup * R;
    const int span = 16 * NY;                       // outputs per lane
    const int t_cta0 = tstrip * G * span;           // first output of time group 0
    const int T = a.T;
    const uint32_t bars = sbase + kOffBar;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + kOffTmem);

    if (warp == 0) {
        if (elect_one()) {
            // the loads first: the first kSlots chunks of the strip, G boxes of R rows x 64 samples per chunk
            for (int p = 0; p < kSlots; ++p) mbar_init(bars + 8 * (kBarFull + p), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            for (int p = 0; p < kSlots && p < NCH_IN; ++p) {
                mbar_expect_tx(bars + 8 * (kBarFull + p), (uint32_t)kChunkBytes);
                for (int g = 0; g < G; ++g)
                    tma_load_2d(sbase + p * kChunkBytes + g * R * 128, &tm_x, t_cta0 + g * span - 8 + 64 * p, row0,
                                bars + 8 * (kBarFull + p));
            }
            mbar_init(bars + 8 * kBarPre, 8);
            for (int i = 0; i < 8; ++i) mbar_init(bars + 8 * (kBarCmp + i), 4);
            for (int i = 0; i < 8; ++i) mbar_init(bars + 8 * (kBarEv + i), 1);
            for (int i = 0; i < kSlots; ++i) mbar_init(bars + 8 * (kBarOut + i), 16);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_y)) : "memory");
        }
        __syncwarp();
        AFA_TC_STAMP(3, 0, 0);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        AFA_TC_STAMP(3, 0, 1);
    } else {
        // banded Toeplitz B matrices (hi + lo bf16 split of the taps), K-major core-matrix layout: element (n, k) of a matrix at
        // (k / 8) * (N * 8) + n * 8 + (k % 8).  Zero fill by 16-byte stores, then the 6 (up) / 12 (down) taps of every column.
        if (warp == 1) AFA_TC_STAMP(3, 0, 2);
        const int t2 = tid - 32;                                   // 0 .. 287
        uint4* wz = reinterpret_cast<uint4*>(sgen + kOffWup);
        for (int i = t2; i < (4 * 1024 + 6 * 512) / 16; i += kThreads - 32) wz[i] = make_uint4(0, 0, 0, 0);
        asm volatile("bar.sync 1, 288;" ::: "memory");
        uint16_t* wup = reinterpret_cast<uint16_t*>(sgen + kOffWup);
        uint16_t* wdn = reinterpret_cast<uint16_t*>(sgen + kOffWdn);
        for (int i = t2; i < 32 * 6 + 16 * 12; i += kThreads - 32) {
            if (i < 192) {                                          // up: column e, i-th tap of its phase: tap = e + 11 - 2 kappa
                const int e = i / 6, ii = i - e * 6;
                const int tap = ((e + 11) & 1) + 2 * ii;
                const int kappa = (e + 11 - tap) >> 1;              // 0 .. 21 over slices a (0..15), b (16..31)
                const int sl = kappa >> 4, k = kappa & 15;
                const int off = (k >> 3) * (32 * 8) + e * 8 + (k & 7);
                wup[(0 * 2 + sl) * 512 + off] = a.up_hi[tap];
                wup[(1 * 2 + sl) * 512 + off] = a.up_lo[tap];
            } else {                                                // down: column e, tap: kappa = 2 e + 5 + tap
                const int i2 = i - 192;
                const int e = i2 / 12, tap = i2 - e * 12;
                const int kappa = 2 * e + 5 + tap;                  // 5 .. 46 over slices a, b, c
                const int sl = kappa >> 4, k = kappa & 15;
                const int off = (k >> 3) * (16 * 8) + e * 8 + (k & 7);
                wdn[(0 * 3 + sl) * 256 + off] = a.dn_hi[tap];
                wdn[(1 * 3 + sl) * 256 + off] = a.dn_lo[tap];
            }
        }
        if (warp == 1) AFA_TC_STAMP(3, 0, 3);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 1) AFA_TC_STAMP(3, 1, 1);

    if (warp == 0) {
        // ===== TMA store of finished output chunks, then the slot takes the x chunk kSlots further down the strip =====
        for (int qc = 0; qc < NCH_OUT; ++qc) {
            const int slot = qc % kSlots;
            mbar_wait(bars + 8 * (kBarOut + slot), (uint32_t)(qc / kSlots) & 1u);
            if (elect_one()) {
                for (int g = 0; g < G; ++g)
                    tma_store_2d(&tm_y, t_cta0 + g * span + 64 * qc, row0, sbase + slot * kChunkBytes + g * R * 128);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                const int nc = qc + kSlots;                         // next x chunk for this slot
                if (nc < NCH_IN) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the store has finished reading the slot
                    mbar_expect_tx(bars + 8 * (kBarFull + slot), (uint32_t)kChunkBytes);
                    for (int g = 0; g < G; ++g)
                        tma_load_2d(sbase + slot * kChunkBytes + g * R * 128, &tm_x, t_cta0 + g * span - 8 + 64 * nc, row0,
                                    bars + 8 * (kBarFull + slot));
                }
            }
            __syncwarp();
        }
        if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        AFA_TC_STAMP(3, 1, 4);
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // InstrDescriptor: D f32 [4,6) = 1, A bf16 [7,10) = 1, B bf16 [10,13) = 1, K-major both, N >> 3 [17,23), M >> 4 [24,29)
        constexpr uint32_t idesc_up = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        constexpr uint32_t idesc_dn = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t bup = bdesc_kmajor(sbase + kOffWup, 32 * 16);     // + 64 (1024 B >> 4) per matrix
        const uint64_t bdn = bdesc_kmajor(sbase + kOffWdn, 16 * 16);     // + 32 (512 B >> 4) per matrix
        const uint64_t ax = adesc_sw128(sbase);                          // x slice k: + (k / 4) * 1024 + (k % 4) * 2 (16-byte units)
        auto xdesc = [&](int k) { return ax + (uint64_t)(((k >> 2) % kSlots) * (kChunkBytes >> 4) + (k & 3) * 2); };
        // loop-invariant B descriptors (hi / lo halves of every K slice), kept in registers
        const uint64_t bu_a_hi = bup + 0 * 64, bu_b_hi = bup + 1 * 64, bu_a_lo = bup + 2 * 64, bu_b_lo = bup + 3 * 64;
        const uint64_t bd_a_hi = bdn + 0 * 32, bd_b_hi = bdn + 1 * 32, bd_c_hi = bdn + 2 * 32;
        const uint64_t bd_a_lo = bdn + 3 * 32, bd_b_lo = bdn + 4 * 32, bd_c_lo = bdn + 5 * 32;
        auto up = [&](uint32_t d, uint64_t xa, uint64_t xb) {
            mma_ss(d, xa, bu_a_hi, idesc_up, 0);
            mma_ss(d, xa, bu_a_lo, idesc_up, 1);
            mma_ss(d, xb, bu_b_hi, idesc_up, 1);
            mma_ss(d, xb, bu_b_lo, idesc_up, 1);
        };
        auto down = [&](uint32_t d, uint32_t s0, uint32_t s2) {
            mma_ts(d, s0, bd_a_hi, idesc_dn, 0);
            mma_ts(d, s0, bd_a_lo, idesc_dn, 1);
            mma_ts(d, s0 + 8, bd_b_hi, idesc_dn, 1);
            mma_ts(d, s0 + 8, bd_b_lo, idesc_dn, 1);
            mma_ts(d, s2, bd_c_hi, idesc_dn, 1);
            mma_ts(d, s2, bd_c_lo, idesc_dn, 1);
        };
        // the compute threads have patched the replicate padding of x into the chunks that hold a row end (pre barrier); the
        // chunks themselves are awaited here, in the order the up-filter products need them
        mbar_wait(bars + 8 * kBarPre, 0);
        mbar_wait(bars + 8 * (kBarFull + 0), 0);
        mbar_wait(bars + 8 * (kBarFull + 1), 0);                        // up(3) reads slice 4 (NY >= 4: two chunks at least)
        int nfull = 2;
        tc_fence_after();
        if (elect_one()) {                                   // events -4 .. -1
            up(tmem + kColUS + 0, xdesc(0), xdesc(1)); tc_commit(bars + 8 * (kBarEv + 0));
            up(tmem + kColUS + 32, xdesc(1), xdesc(2)); tc_commit(bars + 8 * (kBarEv + 1));
            up(tmem + kColUS + 64, xdesc(2), xdesc(3)); tc_commit(bars + 8 * (kBarEv + 2));
            up(tmem + kColUS + 96, xdesc(3), xdesc(4)); tc_commit(bars + 8 * (kBarEv + 3));
        }
        __syncwarp();
        int sl_dn = kRing - 1, sl_s2 = 0, sl_up = 4;       // ring slots of blocks j-1, j, j+4
        for (int j = 0; j <= NY; ++j) {
            AFA_TC_STAMP(0, j, 0);
            if (j + 4 <= NY) {
                const int p = (j + 5) >> 2;                  // chunk of slice j+5
                while (nfull <= p) { mbar_wait(bars + 8 * (kBarFull + nfull % kSlots), (uint32_t)(nfull / kSlots) & 1u); ++nfull; }
            }
            mbar_wait(bars + 8 * (kBarCmp + (j & 7)), (uint32_t)(j >> 3) & 1u);
            tc_fence_after();
            AFA_TC_STAMP(0, j, 1);
            // addresses of this event, computed by the whole warp (uniform) before the single-thread issue
            const int ju = j + 4;
            const uint32_t dd = tmem + kColY + 16 * sl_dn;
            const uint32_t s0 = tmem + kColUS + 32 * sl_dn, s2 = tmem + kColUS + 32 * sl_s2;
            const uint32_t du = tmem + kColUS + 32 * sl_up;
            const uint64_t xa = xdesc(ju), xb = xdesc(ju + 1);
            const uint32_t evb = bars + 8 * (kBarEv + (ju & 7));
            sl_dn = sl_s2;
            sl_s2 = sl_s2 + 1 == kRing ? 0 : sl_s2 + 1;
            sl_up = sl_up + 1 == kRing ? 0 : sl_up + 1;
            const bool do_dn = j >= 1, do_up = ju <= NY;
            if (elect_one()) {
                if (do_dn) down(dd, s0, s2);
                if (do_up) up(du, xa, xb);
                tc_commit(evb);
            }
            __syncwarp();
            AFA_TC_STAMP(0, j, 2);
        }
        AFA_TC_STAMP(3, 1, 2);
    } else {
        // ===== compute groups =====
        const int grp = (warp - 2) >> 2;
        const int q = warp & 3;                                  // TMEM lane quarter this warp may access
        const int l = q * 32 + lane;                             // TMEM lane = row of the MMA tile
        const int g = l >> a.R_log2, r = l & (R - 1);
        const int row = row0 + r;
        const int t_org = t_cta0 + g * span - 8;
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        const uint32_t srow = sbase + (uint32_t)l * 128u;
        const uint32_t sw = (uint32_t)(l & 7);
        float a_eff = 1.f, ib = 1.f;
        const bool left_lane = t_org < 0;                        // the lane starts its row (t_org = -8)
        // first 8-sample chunk of the lane's window that lies beyond the row (T % 8 == 0: chunk granular), if any
        const int cb = (T - t_org) >> 3;
        const bool right_lane = T > t_org && cb < 8 * NCH_IN;       // cb: 8-sample units from the lane's window start
        if (row < a.rows) {
            const int c = row % a.C;
            float al = __ldg(a.alpha + c);
            float be = (a.flags & 2) ? al : __ldg(a.beta + c);
            if (a.flags & 1) { al = expf(al); be = expf(be); }
            a_eff = al;
            ib = 1.0f / (be + 0.000000001f);
        }
        // replicate padding of x (resample.py:32), patched into the staged chunks by group 0 (the TMA unit zero-filled what lies
        // outside the tensor): x[0] over the 8 samples before the row, x[T-1] over the 16 samples behind it (3 are read).  Chunk c
        // is patched after it has landed and before the first product that reads it (up(4c-1), issued at event 4c-5): chunks 0 and
        // 1 here, chunk c >= 2 at iteration 4c-8 of group 0, ahead of that iteration's arrive on cmp.
        auto patch_chunk = [&](int c) {
            const bool need = row < a.rows && ((left_lane && c == 0) || (right_lane && ((cb >> 3) == c || ((cb + 1) >> 3) == c)));
            if (__any_sync(0xffffffffu, need)) {
                if (need) {
                    const unsigned short* xr16 = reinterpret_cast<const unsigned short*>(a.x) + (size_t)row * (size_t)T;
                    mbar_wait(bars + 8 * (kBarFull + c % kSlots), (uint32_t)(c / kSlots) & 1u);
                    const uint32_t cbase = srow + (uint32_t)(c % kSlots) * kChunkBytes;
                    if (left_lane && c == 0) {
                        const uint32_t v = __ldg(xr16), xl = v | (v << 16);
                        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(cbase + ((0u ^ sw) << 4)), "r"(xl) : "memory");
                    }
                    if (right_lane) {
                        const uint32_t v = __ldg(xr16 + (T - 1)), xr = v | (v << 16);
                        for (int cc = cb; cc < cb + 2; ++cc)
                            if ((cc >> 3) == c)
                                asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(cbase + (((uint32_t)(cc & 7) ^ sw) << 4)), "r"(xr) : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
        };
        if (grp == 0) {
            patch_chunk(0);
            patch_chunk(1);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * kBarPre);

        const float2 a2 = make_float2(a_eff, a_eff), ib2 = make_float2(ib, ib);
        int slot = grp;                                          // ring slot of block j (and of block j-5): j mod 5
        for (int j = grp; j <= NY + 4; j += 2, slot = slot + 2 >= kRing ? slot + 2 - kRing : slot + 2) {
            if (q == 2) AFA_TC_STAMP(1 + grp, j, 0);
            mbar_wait(bars + 8 * (kBarEv + (j & 7)), (uint32_t)(j >> 3) & 1u);
            tc_fence_after();
            if (q == 2) AFA_TC_STAMP(1 + grp, j, 1);
            if (j >= 5) {
                // drain Y(j-5): fp32 accumulators -> bf16 -> the lane's row of out chunk (j-5)/4 (in place over x slice j-5)
                const int i = j - 5;
                uint32_t yv[16];
                tmem_ld16(tlane + kColY + 16 * slot, yv);
                tmem_wait_ld();
                uint32_t pk[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) pk[e] = pack_bf16(__uint_as_float(yv[2 * e]), __uint_as_float(yv[2 * e + 1]));
                const uint32_t c0 = (uint32_t)(i & 3) * 2u;
                const uint32_t base = srow + (uint32_t)((i >> 2) % kSlots) * kChunkBytes;
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + ((c0 ^ sw) << 4)), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + (((c0 + 1) ^ sw) << 4)), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + 8 * (kBarOut + (i >> 2) % kSlots));
            }
            if (q == 2) AFA_TC_STAMP(1 + grp, j, 2);
            if ((j & 3) == 0 && ((j + 8) >> 2) < NCH_IN) patch_chunk((j + 8) >> 2);      // group 0 only: j even
            if (j <= NY) {
                // U(j) in two halves of 16 columns: the second load is in flight while the first half goes through Snake, and
                // only 16 fp32 values are live at a time (the register budget of two CTAs per SM is 80 per thread)
                uint32_t sp[16];
                uint32_t ua[16], ub[16];
                tmem_ld16(tlane + kColUS + 32 * slot, ua);
                tmem_wait_ld();
                tmem_ld16(tlane + kColUS + 32 * slot + 16, ub);
                if (a.debug == 1 && row < a.rows) {
                    float* d = a.dbg + ((size_t)blockIdx.x * 128 + l) * (size_t)(a.dbg_blocks * 32) + j * 32;
#pragma unroll
                    for (int e = 0; e < 16; ++e) d[e] = __uint_as_float(ua[e]);
                }
                // Snake on packed pairs: s = u + ib * sin^2(a u)      activations.py:60, :124
                auto snake8 = [&](const uint32_t (&u)[16], int h) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        float2 u2[4], th[4], sn[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            u2[e] = make_float2(__uint_as_float(u[8 * c + 2 * e]), __uint_as_float(u[8 * c + 2 * e + 1]));
                            th[e] = __fmul2_rn(u2[e], a2);
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) sn[e] = make_float2(__sinf(th[e].x), __sinf(th[e].y));
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 s2 = __ffma2_rn(ib2, __fmul2_rn(sn[e], sn[e]), u2[e]);
                            sp[8 * h + 4 * c + e] = pack_bf16(s2.x, s2.y);
                        }
                    }
                };
                snake8(ua, 0);
                tmem_wait_ld();
                if (a.debug == 1 && row < a.rows) {
                    float* d = a.dbg + ((size_t)blockIdx.x * 128 + l) * (size_t)(a.dbg_blocks * 32) + j * 32 + 16;
#pragma unroll
                    for (int e = 0; e < 16; ++e) d[e] = __uint_as_float(ub[e]);
                }
                snake8(ub, 1);
                // replicate padding of the ACTIVATED signal (filter.py:98), in the 2x domain; rare -> uniform branch
                const int eb = 2 * (T - t_org) - 32 * j - 6;      // element of n = 2T in this block: 10 or 26 when inside
                if (__any_sync(0xffffffffu, (left_lane && j == 0) || eb == 10 || eb == 26)) {
                    if (left_lane && j == 0) {                    // n < 0 <-> e < 10: s[0] is element 10
                        const uint32_t s0 = __byte_perm(sp[5], sp[5], 0x1010);
#pragma unroll
                        for (int e = 0; e < 5; ++e) sp[e] = s0;
                    }
                    if (eb == 10) {
                        const uint32_t sl = __byte_perm(sp[4], sp[4], 0x3232);
#pragma unroll
                        for (int e = 5; e < 16; ++e) sp[e] = sl;
                    } else if (eb == 26) {
                        const uint32_t sl = __byte_perm(sp[12], sp[12], 0x3232);
#pragma unroll
                        for (int e = 13; e < 16; ++e) sp[e] = sl;
                    }
                }
                if (a.debug == 2 && row < a.rows) {
                    float* d = a.dbg + ((size_t)blockIdx.x * 128 + l) * (size_t)(a.dbg_blocks * 32) + j * 32;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        d[2 * e] = __uint_as_float(sp[e] << 16);
                        d[2 * e + 1] = __uint_as_float(sp[e] & 0xffff0000u);
                    }
                }
                if (q == 2) AFA_TC_STAMP(1 + grp, j, 3);
                tmem_st16(tlane + kColUS + 32 * slot, sp);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + 8 * (kBarCmp + (j & 7)));
                if (q == 2) AFA_TC_STAMP(1 + grp, j, 4);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) AFA_TC_STAMP(3, 1, 3);
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
    }
}

}  // namespace afa_tc
