// afa_tc_kernels.cuh -- fused Activation1d (bf16 I/O) with BOTH FIR filters on the 5th-generation tensor cores
// (tcgen05.mma, accumulators and the down filter's A operand in TMEM), sm_100a.
//
// What it replaces (reference, /root/reference/BigVGAN): alias_free_activation/act.py:25-30 =
// resample.py:29-38 (UpSample1d) + activations.py:51-62 / 113-126 (Snake / SnakeBeta) + filter.py:94-101
// (LowPassFilter1d, stride 2), for dense row-major [rows = batch*channels, T] bf16 tensors with T % 8 == 0 -- or T % 8 == 4 and
// an even number of rows (stage 0 of a clip with an odd number of mel frames: T = 4 * 861): such rows are only 8-byte aligned
// and travel through the tensor maps as 16-byte aligned PAIRS (see `halves` below).
//
// Why: the register-walk kernel of afa_kernels.cuh spends 24 of its 27 instructions per output on the two
// 12-tap FIRs and is bound by the FP32 FMA pipe -- at bf16 I/O it reaches 0.29-0.44 of the HBM roofline
// (VERDICT round 1).  The FIRs are linear and shift invariant, so 16 outputs of 128 rows are a product with a small
// CONSTANT banded Toeplitz matrix (K = 16 samples per tcgen05.mma, N = 16):
//
//   U [128 rows x 16 u-values] = X [128 x 16 samples, starting at the block's own first sample] * Wup [16 x 16]
//   Y [128 rows x 16 outputs ] = S_a [128 x 16] * Wdn_a + S_b * Wdn_b + S_c * Wdn_c        (three s slices of 16 values)
//
// with M = 128 rows of the tensor (one TMEM lane each).  The products are cut to the band: 16 u-values need 14 consecutive
// samples, i.e. ONE K slice when it starts at an 8-sample (16-byte) granular address inside the swizzled row; 16 outputs need
// 43 s-values, three slices.  (N = 64 / 32 products need fewer instructions but multiply 2.2 x as many zeros: tensor-pipe time
// and, on a board that runs this kernel at its power cap, energy.)  Only Snake (2 MUFU + 6 FMA-pipe instructions per output)
// stays on the CUDA cores.
//   * taps are split hi + lo into two bf16 matrices (16 mantissa bits), both products accumulate in fp32 in
//     TMEM; x is bf16 by contract, so U is exact to fp32 rounding.  s is rounded to bf16 once, as the A operand
//     of the down filter (the same rounding step the bf16 output applies anyway).  (A single fp16 tap matrix
//     against bf16 activations is rejected by the hardware -- "illegal instruction" -- although the descriptor
//     has separate format fields; kUpMats / kDnMats = 1 use the bf16 hi part alone and exist for measurements.)
//   * up filter in SS mode: A = the x slices straight from the 128-byte swizzled chunks the tensor-map TMA
//     wrote (no thread touches x).  Down filter in TS mode: A = S from TENSOR MEMORY.  Shared memory therefore
//     sees each byte of x and y once plus the operand reads of the up filter.
//   * tensor memory: 4 slots of 64 columns.  up(b) fills slot b % 4 with U_b (64 fp32).  Snake reads it and writes
//     S_b (64 bf16) over columns 0..31; down(b) then puts Y_b (32 fp32) into columns 32..63 of the slot of block b + 1 --
//     the half of U_(b+1) that Snake has consumed by then -- so the output accumulators cost no columns of their own.
//   * a CTA owns 128 lanes = R rows x G time groups (R * G = 128; R = 16 for one binaural clip at C = 24) and
//     NY / 2 blocks of 32 outputs per lane.  x arrives as 64-sample chunks through a recycled ring of 5 slots;
//     out-of-range samples and rows are zero-filled by the TMA unit; y is staged over the NEXT x chunk of the ring
//     (dead by then) and leaves by TMA store (which clips what lies outside the tensor).
//   * warp roles: warp 0 = TMA producer / TMA store; warp 1 = TMEM allocator and MMA issuer (one elected thread);
//     warps 2-9 = two compute groups of four warps (one thread per TMEM lane) that alternate over the blocks.
//     All hand-offs are mbarriers (tcgen05.commit on the MMA side); no CTA-wide barrier inside the block loop.
//
// Index algebra (t_org = the lane's first staged sample, t_org % 8 == 0, first output t_org + 8):
//   x slice k   : x_ext[t_org + 16k + (0..15)]
//   u block b   : u_ext[n0 + e], n0 = 2*t_org + 64b + 6, e = 0..63;  u[n] = 2 * sum_i f[n + 5 - 2i] * x_ext[i]
//                 -> sub-block m (e = 16m .. 16m+15) reads samples t_org + 32b + 8m + (0..13):  Wup[kappa][e'] = 2 f[e' + 11 - 2 kappa]
//   y block b   : y[t_org + 8 + 32b + e], e = 0..31;  y[t] = sum_k f[k] * s_ext[2t + k - 5]
//                 -> sub-block h (e = 16h .. 16h+15) reads s-values 32h + (0..47) of the block (columns 16h, 16h+8, 16h+16; column
//                    32 is the first slice of block b+1):  Wdn[kappa][e'] = f[kappa - 2e' - 5],  kappa = 0..47
//   replicate padding: x_ext clamps in the 1x domain (chunks with t < 0 or t >= T, always whole 8-sample
//   chunks because T % 8 == 0); s_ext clamps the ACTIVATED signal in the 2x domain (filter.py:98): elements
//   e < 10 of block 0 of a row's first lane, and elements e >= e_b (e_b = 10, 26, 42 or 58) of the block holding 2T.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace afa_tc {

constexpr int kThreads = 320;            // warp 0: TMA producer / TMA store, warp 1: TMEM allocator + MMA issuer, warps 2..9: compute
#ifndef AFA_TC_SLOTS
#define AFA_TC_SLOTS 5
#endif
constexpr int kSlots = AFA_TC_SLOTS;                // shared-memory chunk ring: slots of 64 samples x 128 lanes, recycled along the strip
constexpr int kChunkBytes = 128 * 128;   // 128 lanes x 64 bf16
constexpr int kTmemCols = 256;           // 4 slots x 64 columns
constexpr int kRing = 4;                 // tensor-memory slots: U(b) = 64 fp32 columns; then S(b) = 64 bf16 in columns 0..31 and
constexpr int kSlotCols = 64;            //   Y(b) = 32 fp32 outputs in columns 32..63 (the half of U that Snake has consumed)
constexpr int kUpVariants = 3, kDnSlices = 3;        // B matrices: up = plain / shifted +8 / shifted -8 rows; down = K slices a, b, c
constexpr int kWBytes = 16 * 16 * 2;                 // every B matrix is [N = 16 x K = 16] bf16
// shared memory carve-up (offsets from a 1024-byte aligned base)
constexpr int kOffWup = kSlots * kChunkBytes;                    // [hi/lo][variant] x 512 B
constexpr int kOffWdn = kOffWup + 2 * kUpVariants * kWBytes;     // [hi/lo][slice] x 512 B
constexpr int kOffBar = kOffWdn + 2 * kDnSlices * kWBytes;
constexpr int kBarFull = 0, kBarPre = kSlots, kBarCmp = kBarPre + 1, kBarEv = kBarCmp + 8, kBarOut = kBarEv + 8;
constexpr int kBarEvY = kBarOut + kSlots;      // Y(e-1) complete: committed between the down and the up products of event e
constexpr int kNumBars = kBarEvY + 8;
constexpr int kOffTmem = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmem + 16 + 1024;         // + slack for the manual 1024-byte alignment

struct Args {
    const __nv_bfloat16* x;
    __nv_bfloat16* y;
    const float* alpha;
    const float* beta;
    uint16_t up_hi[12], up_lo[12];   // 2 * upsample taps (ratio folded, resample.py:33): bf16 hi + lo, or fp16 in up_hi (kMats = 1)
    uint16_t dn_hi[12], dn_lo[12];   // low-pass taps, same split
    int32_t rows, C, T, flags;
    int32_t R_log2;        // lanes = R rows x G groups, R = 1 << R_log2
    int32_t NY;            // outputs per lane and CTA in units of 16: a multiple of 4, any length (the chunk ring is recycled)
    int32_t n_tstrips;     // CTAs along time; blockIdx.x = row_group * n_tstrips + tstrip
    int64_t x_pitch, y_pitch;   // elements between consecutive rows of x / y (T for dense tensors)
    int32_t halves;        // 1: tensor-map rows = tensor rows.  2 (T % 8 == 4, even row count): tensor-map rows = PAIRS of rows, 2T wide
    int32_t spr;           // lane strips (16 * NY outputs) per tensor row
    int32_t debug;         // harness only: 1 = dump U blocks, 2 = dump S, 3 = clock stamps of CTA dbg_cta
    int32_t dbg_cta;
    int32_t dbg_blocks;    // harness: 64-value blocks per lane the U / S dump holds (NY / 2 + 1)
    int32_t dbg_j0;        // harness: first block of the 32-block window the clock stamps cover
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
// programmatic dependent launch (launch attribute cudaLaunchAttributeProgrammaticStreamSerialization; no-ops without it)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
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
#define AFA_TC_STAMP(role, j, slot) do { if (kDebug) { const int jw_ = (j) - ((role) == 3 ? 0 : a.dbg_j0); if (a.debug == 3 && lane == 0 && jw_ >= 0 && jw_ < 32 && blockIdx.x == (unsigned)a.dbg_cta) a.dbg[((role) * 32 + jw_) * 8 + (slot)] = __uint_as_float(clk32()); } } while (0)
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// Snake / SnakeBeta: u + sin^2(alpha u) / (beta + 1e-9)            activations.py:60, :124
__device__ __forceinline__ float snake_f(float u, float a, float ib) {
    const float sn = __sinf(u * a);
    return fmaf(ib, sn * sn, u);
}

// Schedule.  Block b = 32 time steps of the CTA's 128 lanes (64 u-values, 32 outputs per lane).  Compute group g (4 warps)
// owns the blocks b = g (mod 2).  Tensor-memory slot of block b: b % 4.
//   MMA thread, event e:        wait cmp[e & 7]  (S(e) is in TMEM and Y(e-2) has been read out of slot (e-1) % 4);
//                               down(e-1): S(e-1) and the first slice of S(e) -> Y(e-1) in columns 32..63 of slot e % 4 (the
//                               half of U(e) that Snake has consumed);
//                               up(e+3): x slices 2e+6 .. 2e+8 straight from the swizzled chunks -> U(e+3) in slot (e-1) % 4;
//                               commit -> evy[e & 7] after the down products, commit -> ev[(e+3) & 7] after the up products
//                               (tcgen05.commit tracks every MMA issued before it)
//   iteration b of its group:   wait ev[b & 7]  (event b-3: U(b) is complete -- issued a whole iteration of the OTHER group ago,
//                               so the MMA round trip is off the group's critical path)
//                               U(b) -> registers (4 loads of 16 columns, the next in flight behind the math) -> Snake ->
//                               S(b) over columns 0..31 (two tcgen05.st)
//                               wait evy[(b-1) & 7]  (event b-1: Y(b-2) is complete) -> registers -> bf16 -> shared memory
//                               (out chunk (b-2)/2, staged in the ring slot of x chunk (b-2)/2 + 1) -> proxy fence -> arrive out[...]
//                               tcgen05.wait::st -> arrive cmp[b & 7]
//   warp 0:                     wait out[q] (8 warp arrivals) -> TMA store of out chunk q -> the slot takes x chunk q + 6
// One barrier round trip, one tensor-memory store wait and one proxy fence cover 32 outputs per lane (16 in the first
// versions: the fixed latencies of an iteration, ~1300 cycles, outweighed its ~900 cycles of Snake math).
// Ring safety: up(e+3) overwrites slot (e-1) % 4: S(e-1) was last read by down(e-1), issued just before it by the same thread
// (MMAs execute in issue order); its upper half held Y(e-2), read by iteration e of group e % 2 before the arrive on
// cmp[e & 7].  down(e-1) writes the upper half of slot e % 4: U(e), in registers since iteration e.  Out chunk q is staged over
// x chunk q + 1 (slices 4q+4 .. 4q+7, last read by up(2q+3) at event 2q, complete before iteration 2q+2 passes its wait on
// event 2q+1): a slot is free again one chunk time earlier than with in-place staging, which is what lets the x chunk for
// up(e+3) arrive in time.  cmp / ev are rings of 8 indexed by block: a warp's waits depend on events b-3 and b-1 only, so it
// may run ahead of a slower warp of its group, never 8 events ahead (event e needs cmp[e]).
// kUpMats / kDnMats: tap matrices per K slice of the up / down filter -- 2 = bf16 hi + lo (16 mantissa bits), 1 = the taps
// rounded to bf16 (8 bits).  kDebug: harness dumps and clock stamps; compiled out of the product instantiation.
#ifndef AFA_TC_BOUND_THREADS
#define AFA_TC_BOUND_THREADS 320
#endif
template <int kUpMats, int kDnMats, bool kDebug>
__global__ void __launch_bounds__(AFA_TC_BOUND_THREADS, 2)
afa_tc_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
                  const __grid_constant__ Args a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler: role code runs on the uniform datapath
    if (warp == 1) AFA_TC_STAMP(3, 1, 0);
    if (kDebug && a.debug == 4 && tid == 32) {                  // harness: per-CTA life span (globaltimer ns) and SM id
        unsigned long long t;
        uint32_t smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        reinterpret_cast<unsigned long long*>(a.dbg)[blockIdx.x * 4 + 0] = t;
        reinterpret_cast<unsigned long long*>(a.dbg)[blockIdx.x * 4 + 2] = smid;
    }
    const int NY = a.NY, NB = NY >> 1;                            // NB blocks of 32 outputs per lane
    const int NCH_IN = (NY >> 2) + 1, NCH_OUT = NY >> 2;          // x chunks (slices 0 .. NY + 2) / out chunks of this CTA's strip
    const int R = 1 << a.R_log2, G = 128 >> a.R_log2;
    const int tstrip = (int)(blockIdx.x % (uint32_t)a.n_tstrips);
    const int rgroup = (int)(blockIdx.x / (uint32_t)a.n_tstrips);
    const int row0 = rgroup * R;
    const int span = 16 * NY;                       // outputs per lane
    // Time group g of this CTA is lane strip vs = tstrip * G + g of the tensor-map row: strips 0 .. spr-1 belong to the first
    // tensor row of the pair (or to the only one), spr .. 2 spr - 1 to the second, which starts T elements into the map row.
    // first_out(g): first output of the group relative to ITS tensor row; col0(g): the same position as a map column.  The TMA
    // unit moves 16-byte units: map columns must be multiples of 8 elements, and the second row of a pair starts at column T
    // = 4 (mod 8) -- its strips are shifted 4 samples to the left (first_out = -4 for its first strip).
    auto half_of = [&](int g) { const int vs = tstrip * G + g; return vs / a.spr; };
    auto first_out = [&](int g) { const int vs = tstrip * G + g; return (vs % a.spr) * span - ((a.halves == 2 && vs / a.spr == 1) ? 4 : 0); };
    auto col0 = [&](int g) { const int hh = half_of(g); return (hh < a.halves ? hh * a.T : 2 * a.halves * a.T + 64) + first_out(g); };
    const int T = a.T;
    const uint32_t bars = sbase + kOffBar;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + kOffTmem);

    pdl_launch_dependents();      // programmatic dependent launch: the next kernel's set-up may run in the shadow of this one's tail
    if (warp == 0) {
        // Barriers and the first loads, spread over the lanes: one elected thread needed ~115 cycles per TMA instruction and ~2000
        // cycles for the 40-odd mbarrier inits -- with 8 boxes per chunk (one binaural clip: R = 16) the CTA's first data was
        // requested 5300 cycles after kernel entry.  Lane i initialises barriers i and i + 32; lane g issues box g of every chunk.
        for (int i = lane; i < kNumBars; i += 32) {
            const uint32_t cnt = i == kBarPre ? 8u : (i >= kBarCmp && i < kBarEv) ? 4u : (i >= kBarOut && i < kBarEvY) ? 8u : 1u;
            mbar_init(bars + 8 * i, cnt);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
        pdl_wait();               // x may be the predecessor's output: no global access before this point
        for (int p = 0; p < kSlots && p < NCH_IN; ++p) {
            if (lane == 0) mbar_expect_tx(bars + 8 * (kBarFull + p), (uint32_t)kChunkBytes);
            __syncwarp();
            if (lane < G)
                tma_load_2d(sbase + p * kChunkBytes + lane * R * 128, &tm_x, col0(lane) - 8 + 64 * p, row0, bars + 8 * (kBarFull + p));
        }
        if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_y)) : "memory");
        __syncwarp();
        AFA_TC_STAMP(3, 0, 0);
    } else {
        // banded Toeplitz B matrices (hi + lo bf16 split of the taps), K-major core-matrix layout: element (n, k) of a matrix at
        // (k / 8) * (N * 8) + n * 8 + (k % 8).  Zero fill by 16-byte stores, then the 6 (up) / 12 (down) taps of every column.
        if (warp == 1) {                                           // tensor-memory allocation, beside warp 0's loads
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
            AFA_TC_STAMP(3, 0, 1);
        }
        if (warp == 1) AFA_TC_STAMP(3, 0, 2);
        const int t2 = tid - 32;                                   // 0 .. 287
        uint4* wz = reinterpret_cast<uint4*>(sgen + kOffWup);
        for (int i = t2; i < (kOffBar - kOffWup) / 16; i += kThreads - 32) wz[i] = make_uint4(0, 0, 0, 0);
        asm volatile("bar.sync 1, 288;" ::: "memory");
        uint16_t* wup = reinterpret_cast<uint16_t*>(sgen + kOffWup);
        uint16_t* wdn = reinterpret_cast<uint16_t*>(sgen + kOffWdn);
        for (int i = t2; i < 16 * 6 + 16 * 12; i += kThreads - 32) {
            if (i < 96) {                                           // up: column e, i-th tap of its phase: tap = e + 11 - 2 kappa
                const int e = i / 6, ii = i - e * 6;
                const int tap = ((e + 11) & 1) + 2 * ii;
                const int kappa = (e + 11 - tap) >> 1;              // 0 .. 13: 16 u-values need 14 samples = ONE K slice
                // variant 0: the slice starts at the block's first sample; variants 1 / 2 serve the block that straddles two
                // chunks: rows moved down by 8 (samples 8..15 of the slice = the block's first 8) / up by 8 (its last 6)
                const int var = kappa < 8 ? 1 : 2, kv = kappa < 8 ? kappa + 8 : kappa - 8;
                const int off0 = (kappa >> 3) * (16 * 8) + e * 8 + (kappa & 7), offv = (kv >> 3) * (16 * 8) + e * 8 + (kv & 7);
                wup[(0 * kUpVariants + 0) * (kWBytes / 2) + off0] = a.up_hi[tap];
                wup[(0 * kUpVariants + var) * (kWBytes / 2) + offv] = a.up_hi[tap];
                if (kUpMats == 2) {
                    wup[(1 * kUpVariants + 0) * (kWBytes / 2) + off0] = a.up_lo[tap];
                    wup[(1 * kUpVariants + var) * (kWBytes / 2) + offv] = a.up_lo[tap];
                }
            } else {                                                // down: column e, tap: kappa = 2 e + 5 + tap
                const int i2 = i - 96;
                const int e = i2 / 12, tap = i2 - e * 12;
                const int kappa = 2 * e + 5 + tap;                  // 5 .. 46 over slices a, b, c
                const int sl = kappa >> 4, k = kappa & 15;
                const int off = (k >> 3) * (16 * 8) + e * 8 + (k & 7);
                wdn[(0 * kDnSlices + sl) * (kWBytes / 2) + off] = a.dn_hi[tap];
                if (kDnMats == 2) wdn[(1 * kDnSlices + sl) * (kWBytes / 2) + off] = a.dn_lo[tap];
            }
        }
        if (warp == 1) AFA_TC_STAMP(3, 0, 3);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        pdl_wait();               // alpha / beta and the edge samples of x are read after the CTA-wide barrier below
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    if (warp == 1) AFA_TC_STAMP(3, 1, 1);

    if (warp == 0) {
        // ===== TMA store of finished output chunks; the slot then takes the next x chunk of the strip =====
        // x chunk 0 has no output staged over it (out chunk q lives in the slot of x chunk q + 1): its slot takes x chunk kSlots
        // as soon as up(0) and up(1) have read it
        if (kSlots < NCH_IN) {
            mbar_wait(bars + 8 * (kBarEv + 1), 0);
            if (lane == 0) mbar_expect_tx(bars + 8 * (kBarFull + 0), (uint32_t)kChunkBytes);
            __syncwarp();
            if (lane < G)
                tma_load_2d(sbase + lane * R * 128, &tm_x, col0(lane) - 8 + 64 * kSlots, row0, bars + 8 * (kBarFull + 0));
            __syncwarp();
        }
        // lane g moves box g (its own bulk group: commit / wait are per thread)
        const int hh0 = half_of(lane < G ? lane : 0), fo0 = first_out(lane < G ? lane : 0), cl0 = col0(lane < G ? lane : 0);
        for (int qc = 0; qc < NCH_OUT; ++qc) {
            const int slot = (qc + 1) % kSlots;
            mbar_wait(bars + 8 * (kBarOut + slot), (uint32_t)(qc / kSlots) & 1u);
            if (lane < G) {
                // the first row of a pair must not write past its end (the second row's samples follow it in the map row):
                // chunks beyond the end are dropped, the chunk that crosses it is written by the compute threads
                const int o0 = fo0 + 64 * qc;
                if (!((hh0 + 1 < a.halves && o0 + 64 > T) || o0 < 0))      // (o0 < 0: the shifted first chunk of a second row)
                    tma_store_2d(&tm_y, cl0 + 64 * qc, row0, sbase + slot * kChunkBytes + lane * R * 128);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            const int nc = qc + 1 + kSlots;                         // next x chunk for this slot
            if (nc < NCH_IN) {
                if (lane < G) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the stores have finished reading the slot
                __syncwarp();
                if (lane == 0) mbar_expect_tx(bars + 8 * (kBarFull + slot), (uint32_t)kChunkBytes);
                __syncwarp();
                if (lane < G)
                    tma_load_2d(sbase + slot * kChunkBytes + lane * R * 128, &tm_x, cl0 - 8 + 64 * nc, row0, bars + 8 * (kBarFull + slot));
            }
            __syncwarp();
        }
        if (lane < G) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        AFA_TC_STAMP(3, 1, 4);
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // InstrDescriptor: D f32 [4,6) = 1, A bf16 [7,10) = 1, B bf16 [10,13) = 1, K-major both, N >> 3 [17,23), M >> 4 [24,29)
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t bup = bdesc_kmajor(sbase + kOffWup, 16 * 16);     // + kWBytes >> 4 per matrix; the lo set follows the hi set
        const uint64_t bdn = bdesc_kmajor(sbase + kOffWdn, 16 * 16);
        const uint64_t ax = adesc_sw128(sbase);                          // chunk slot s: + s * (kChunkBytes >> 4); 8 samples = one 16-byte unit
        constexpr uint64_t kW = kWBytes >> 4;
        // The products are cut to the band: N = 16 everywhere, so the dense K x N rectangle the tensor core multiplies is 16 x 16
        // around a band of 6-7 (up) / 12 (down) taps -- 160 MACs per output against 352 with N = 64 / 32 products, i.e. less
        // than half the tensor-pipe time and energy for 22 instead of 16 instructions per block.
        // U(bu) <- four blocks of 16 u-values, each from ONE K slice that starts at the block's own first sample (8-sample = 16-byte
        // granular start inside the 128-byte swizzled row); xs = ring slot of the chunk that holds sample 32 bu of the lane's window
        auto up = [&](uint32_t d, int bu, int xs) {
            const uint64_t cb = ax + (uint64_t)(xs * (kChunkBytes >> 4) + (bu & 1) * 4);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                if (m == 3 && (bu & 1)) {
                    // samples 56 .. 69 of the chunk: the slice would leave the 128-byte row.  Two slices instead: samples 48..63 of
                    // this chunk against the taps moved down 8 rows, samples 0..15 of the next chunk against the taps moved up 8
                    const int xs1 = xs + 1 == kSlots ? 0 : xs + 1;
                    const uint64_t a1 = ax + (uint64_t)(xs * (kChunkBytes >> 4) + 6), a2 = ax + (uint64_t)(xs1 * (kChunkBytes >> 4));
                    mma_ss(d + 48, a1, bup + 1 * kW, idesc, 0);
                    if (kUpMats == 2) mma_ss(d + 48, a1, bup + (kUpVariants + 1) * kW, idesc, 1);
                    mma_ss(d + 48, a2, bup + 2 * kW, idesc, 1);
                    if (kUpMats == 2) mma_ss(d + 48, a2, bup + (kUpVariants + 2) * kW, idesc, 1);
                } else {
                    mma_ss(d + 16 * m, cb + m, bup, idesc, 0);
                    if (kUpMats == 2) mma_ss(d + 16 * m, cb + m, bup + kUpVariants * kW, idesc, 1);
                }
            }
        };
        // Y(bd) <- two blocks of 16 outputs, each from three s slices (columns 16 h + 0, 8, 16 of S(bd); column 32 is the first
        // slice of S(bd + 1)); lands in the upper half of the slot of block bd + 1
        auto down = [&](int bd) {
            const uint32_t sl0 = tmem + (uint32_t)(kSlotCols * (bd & 3)), sl1 = tmem + (uint32_t)(kSlotCols * ((bd + 1) & 3));
            const uint32_t d = sl1 + 32;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int k = 0; k < kDnSlices; ++k) {
                    const int col = 16 * h + 8 * k;
                    const uint32_t aa = col < 32 ? sl0 + (uint32_t)col : sl1;
                    mma_ts(d + 16 * h, aa, bdn + (uint64_t)k * kW, idesc, k > 0);
                    if (kDnMats == 2) mma_ts(d + 16 * h, aa, bdn + (uint64_t)(kDnSlices + k) * kW, idesc, 1);
                }
        };
        // the compute threads have patched the replicate padding of x into the chunks that hold a row end (pre barrier); the
        // chunks themselves are awaited here, in the order the up-filter products need them
        mbar_wait(bars + 8 * kBarPre, 0);
        mbar_wait(bars + 8 * (kBarFull + 0), 0);
        mbar_wait(bars + 8 * (kBarFull + 1), 0);                        // up(1), up(2) read slices 4 .. 6 (NY >= 4: two chunks at least)
        int nfull = 2;
        tc_fence_after();
        if (elect_one()) {                                   // events -3 .. -1
            up(tmem + 0 * kSlotCols, 0, 0); tc_commit(bars + 8 * (kBarEv + 0));
            up(tmem + 1 * kSlotCols, 1, 0); tc_commit(bars + 8 * (kBarEv + 1));
            up(tmem + 2 * kSlotCols, 2, 1); tc_commit(bars + 8 * (kBarEv + 2));
        }
        __syncwarp();
        int xs = 1;                                          // ring slot of the chunk holding slice 2 (e + 3) = chunk (e + 3) / 2
        for (int e = 0; e <= NB; ++e) {
            AFA_TC_STAMP(0, e, 0);
            const int bu = e + 3;
            if (bu <= NB) {
                const int p = (bu + 1) >> 1;                 // chunk of slice 2 bu + 2
                while (nfull <= p) { mbar_wait(bars + 8 * (kBarFull + nfull % kSlots), (uint32_t)(nfull / kSlots) & 1u); ++nfull; }
            }
            mbar_wait(bars + 8 * (kBarCmp + (e & 7)), (uint32_t)(e >> 3) & 1u);
            tc_fence_after();
            AFA_TC_STAMP(0, e, 1);
            if (elect_one()) {
                if (e >= 1) down(e - 1);
                tc_commit(bars + 8 * (kBarEvY + (e & 7)));           // the groups wait for Y, not for the 8-10 up products behind it
                if (bu <= NB) up(tmem + (uint32_t)(kSlotCols * (bu & 3)), bu, xs);
                tc_commit(bars + 8 * (kBarEv + (bu & 7)));
            }
            __syncwarp();
            if (bu & 1) xs = xs + 1 == kSlots ? 0 : xs + 1;      // slice 2 (bu + 1) opens the next chunk when bu is odd
            AFA_TC_STAMP(0, e, 2);
        }
        AFA_TC_STAMP(3, 1, 2);
    } else {
        // ===== compute groups =====
        const int grp = (warp - 2) >> 2;
        const int q = warp & 3;                                  // TMEM lane quarter this warp may access
        const int l = q * 32 + lane;                             // TMEM lane = row of the MMA tile
        const int g = l >> a.R_log2, r = l & (R - 1);
        const int hh = half_of(g);
        const int row = hh < a.halves ? (row0 + r) * a.halves + hh : a.rows;      // tensor row of this lane (a.rows: none)
        const int t_org = first_out(g) - 8;
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        const uint32_t srow = sbase + (uint32_t)l * 128u;
        const uint32_t sw = (uint32_t)(l & 7);
        float a_eff = 1.f, ib = 1.f;
        const bool left_lane = t_org < 0;                        // the lane starts its row (t_org = -8; -12 for the second row of a pair)
        // first 8-sample (16-byte) chunk of the lane's window that holds samples beyond the row, if any; T % 8 == 4: its first
        // four samples still belong to the row
        const int cb = (T - t_org) >> 3;
        const bool right_lane = T > t_org && cb < 8 * NCH_IN;       // cb: 8-sample units from the lane's window start
        const bool half_chunk = ((T - t_org) & 7) != 0;
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
        // is patched after it has landed and before the first product that reads it (up(2c-1), issued at event 2c-4): chunks 0, 1
        // and 2 here, chunk c >= 3 at iteration 2c-6 of group 0, ahead of that iteration's arrive on cmp.
        auto patch_chunk = [&](int c) {
            const bool need = row < a.rows && ((left_lane && c == 0) || (right_lane && ((cb >> 3) == c || ((cb + 1) >> 3) == c)));
            if (__any_sync(0xffffffffu, need)) {
                if (need) {
                    const unsigned short* xr16 = reinterpret_cast<const unsigned short*>(a.x) + (size_t)row * (size_t)a.x_pitch;
                    mbar_wait(bars + 8 * (kBarFull + c % kSlots), (uint32_t)(c / kSlots) & 1u);
                    const uint32_t cbase = srow + (uint32_t)(c % kSlots) * kChunkBytes;
                    if (left_lane && c == 0) {
                        const uint32_t v = __ldg(xr16), xl = v | (v << 16);
                        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(cbase + ((0u ^ sw) << 4)), "r"(xl) : "memory");
                        if (t_org == -12) asm volatile("st.shared.v2.b32 [%0], {%1,%1};" ::"r"(cbase + ((1u ^ sw) << 4)), "r"(xl) : "memory");
                    }
                    if (right_lane) {
                        const uint32_t v = __ldg(xr16 + (T - 1)), xr = v | (v << 16);
                        for (int cc = cb; cc < cb + 2; ++cc)
                            if ((cc >> 3) == c) {
                                const uint32_t ad = cbase + (((uint32_t)(cc & 7) ^ sw) << 4);
                                if (cc == cb && half_chunk) asm volatile("st.shared.v2.b32 [%0], {%1,%1};" ::"r"(ad + 8), "r"(xr) : "memory");
                                else asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(ad), "r"(xr) : "memory");
                            }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
        };
        if (grp == 0) {
            patch_chunk(0);
            patch_chunk(1);
            if (2 < NCH_IN) patch_chunk(2);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * kBarPre);

        const float2 a2 = make_float2(a_eff, a_eff), ib2 = make_float2(ib, ib);
        // Snake on packed pairs, 16 values at a time (8 independent chains): s = u + ib * sin^2(a u)   activations.py:60, :124
        auto snake16 = [&](const uint32_t (&u)[16], uint32_t* sp) {
            float2 u2[8], sn[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                u2[e] = make_float2(__uint_as_float(u[2 * e]), __uint_as_float(u[2 * e + 1]));
                const float2 th = __fmul2_rn(u2[e], a2);
#if defined(AFA_TC_EXPERIMENT) && AFA_TC_EXPERIMENT == 1      // harness only: no MUFU (energy / pipe-share experiments; results are wrong)
                sn[e] = th;
#else
                sn[e] = make_float2(__sinf(th.x), __sinf(th.y));
#endif
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float2 s2 = __ffma2_rn(ib2, __fmul2_rn(sn[e], sn[e]), u2[e]);
                sp[e] = pack_bf16(s2.x, s2.y);
            }
        };
        for (int b = grp; b <= NB + 1; b += 2) {
            if (q == 2) AFA_TC_STAMP(1 + grp, b, 0);
            const bool has_u = b <= NB, has_y = b >= 2;
            if (has_u) {
                mbar_wait(bars + 8 * (kBarEv + (b & 7)), (uint32_t)(b >> 3) & 1u);
                tc_fence_after();
                if (q == 2) AFA_TC_STAMP(1 + grp, b, 1);
                if ((b & 1) == 0 && (b >> 1) + 3 < NCH_IN) patch_chunk((b >> 1) + 3);     // group 0 only: b even
                const uint32_t tslot = tlane + (uint32_t)(kSlotCols * (b & 3));
                uint32_t ua[16], ub[16], uc[16], sp[16];
                tmem_ld16(tslot, ua);
                tmem_ld16(tslot + 16, ub);
                tmem_wait_ld();
                float* dbg_row = nullptr;
                if (kDebug && (a.debug == 1 || a.debug == 2) && row < a.rows)
                    dbg_row = a.dbg + ((size_t)blockIdx.x * 128 + l) * (size_t)(a.dbg_blocks * 64) + b * 64;
                if (kDebug && a.debug == 1 && dbg_row) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) { dbg_row[e] = __uint_as_float(ua[e]); dbg_row[16 + e] = __uint_as_float(ub[e]); }
                }
                snake16(ua, sp);
                tmem_ld16(tslot + 32, ua);
                tmem_ld16(tslot + 48, uc);
                snake16(ub, sp + 8);
                // replicate padding of the ACTIVATED signal (filter.py:98), in the 2x domain; rare -> uniform branch.  eb = element
                // of n = 2T in this block (10, 26, 42, 58 -- or 2, 18, 34, 50 when T % 8 == 4 -- when inside): elements >= eb
                // repeat element eb - 1.  kc = eb / 2 = first register (pair of elements) to overwrite, within a half of 16 registers
                const int eb = 2 * (T - t_org) - 64 * b - 6;
                const bool lclamp = left_lane && b == 0, rclamp = eb >= 2 && eb <= 58;
                uint32_t fill = 0;
                const bool any_clamp = __any_sync(0xffffffffu, lclamp || rclamp);
                auto clamp_from = [&](int kc) {                   // kc in {1, 5, 9, 13}: fill = upper element of register kc - 1
                    fill = __byte_perm(kc == 1 ? sp[0] : kc == 5 ? sp[4] : kc == 9 ? sp[8] : sp[12], 0, 0x3232);
#pragma unroll
                    for (int e = 1; e < 16; ++e)
                        if (e >= kc) sp[e] = fill;
                };
                if (any_clamp) {
                    if (lclamp) {                                 // n < 0 <-> e < 10 (18 with t_org = -12): s[0] is element 10 (18)
                        const int k0 = t_org == -12 ? 9 : 5;
                        const uint32_t s0 = __byte_perm(k0 == 9 ? sp[9] : sp[5], 0, 0x1010);
#pragma unroll
                        for (int e = 0; e < 9; ++e)
                            if (e < k0) sp[e] = s0;
                    }
                    if (rclamp && eb < 32) clamp_from(eb >> 1);
                }
                if (kDebug && a.debug == 2 && dbg_row) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        dbg_row[2 * e] = __uint_as_float(sp[e] << 16);
                        dbg_row[2 * e + 1] = __uint_as_float(sp[e] & 0xffff0000u);
                    }
                }
                tmem_st16(tslot, sp);
                tmem_wait_ld();
                if (kDebug && a.debug == 1 && dbg_row) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) { dbg_row[32 + e] = __uint_as_float(ua[e]); dbg_row[48 + e] = __uint_as_float(uc[e]); }
                }
                snake16(ua, sp);
                snake16(uc, sp + 8);
                if (any_clamp && rclamp) {
                    if (eb < 32) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) sp[e] = fill;
                    } else {
                        clamp_from((eb - 32) >> 1);
                    }
                }
                if (kDebug && a.debug == 2 && dbg_row) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        dbg_row[32 + 2 * e] = __uint_as_float(sp[e] << 16);
                        dbg_row[32 + 2 * e + 1] = __uint_as_float(sp[e] & 0xffff0000u);
                    }
                }
                tmem_st16(tslot + 16, sp);
                if (q == 2) AFA_TC_STAMP(1 + grp, b, 2);
            }
            if (has_y) {
                // Y(b-2) (upper half of slot (b-1) % 4): fp32 accumulators -> bf16 -> the lane's row of out chunk (b-2)/2, staged
                // in the ring slot of x chunk (b-2)/2 + 1
                const int i = b - 2;
                mbar_wait(bars + 8 * (kBarEvY + ((b - 1) & 7)), (uint32_t)((b - 1) >> 3) & 1u);
                tc_fence_after();
                if (q == 2) AFA_TC_STAMP(1 + grp, b, 3);
                const uint32_t yslot = tlane + (uint32_t)(kSlotCols * ((i + 1) & 3) + 32);
                uint32_t ya[16], yb[16];
                tmem_ld16(yslot, ya);
                tmem_ld16(yslot + 16, yb);
                tmem_wait_ld();
                uint32_t pk[16];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    pk[e] = pack_bf16(__uint_as_float(ya[2 * e]), __uint_as_float(ya[2 * e + 1]));
                    pk[8 + e] = pack_bf16(__uint_as_float(yb[2 * e]), __uint_as_float(yb[2 * e + 1]));
                }
                const uint32_t c0 = (uint32_t)(i & 1) * 4u;
                const uint32_t base = srow + (uint32_t)(((i >> 1) + 1) % kSlots) * kChunkBytes;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + (((c0 + c) ^ sw) << 4)), "r"(pk[4 * c]), "r"(pk[4 * c + 1]),
                                 "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
                if (a.halves == 2) {
                    // first row of a pair, chunk that crosses the row end -- or second row, first chunk (it starts 4 samples before
                    // the row): the TMA store skips it (it would write into the other row); the lanes write their valid outputs
                    // themselves, 4 samples (8 bytes: what the row start is aligned to) at a time
                    const int o0 = t_org + 8 + 64 * (i >> 1);
                    if (row < a.rows && ((hh == 0 && o0 < T && o0 + 64 > T) || o0 < 0)) {
                        const int tb0 = t_org + 8 + 32 * i;
                        __nv_bfloat16* yr = a.y + (size_t)row * (size_t)a.y_pitch;
#pragma unroll
                        for (int v = 0; v < 8; ++v)
                            if (tb0 + 4 * v >= 0 && tb0 + 4 * v < T) *reinterpret_cast<uint2*>(yr + tb0 + 4 * v) = make_uint2(pk[2 * v], pk[2 * v + 1]);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + 8 * (kBarOut + ((i >> 1) + 1) % kSlots));
            }
            if (has_u) {
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + 8 * (kBarCmp + (b & 7)));
            }
            if (q == 2) AFA_TC_STAMP(1 + grp, b, 4);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) AFA_TC_STAMP(3, 1, 3);
    if (kDebug && a.debug == 4 && tid == 32) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        reinterpret_cast<unsigned long long*>(a.dbg)[blockIdx.x * 4 + 1] = t;
    }
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
    }
}

}  // namespace afa_tc
