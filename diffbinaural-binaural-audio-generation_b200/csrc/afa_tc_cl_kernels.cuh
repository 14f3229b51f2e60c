// afa_tc_cl_kernels.cuh -- the tensor-core Activation1d forward (afa_tc_kernels.cuh) for CHANNELS-LAST bf16 activations
// ([batch, T, channels], channels contiguous): the layout the generator engine keeps between cuDNN's NHWC convolutions
// (afa_cl_kernels.cuh).  sm_100a.
//
// What it replaces (reference, /root/reference/BigVGAN): the bias add of the convolution in front (bigvgan.py:134-139) +
// alias_free_activation/act.py:25-30 = resample.py:29-38 + activations.py:51-62 / 113-126 + filter.py:94-101.
//
// Same arithmetic, schedule, tensor-memory ring and barrier protocol as afa_tc_fwd_kernel -- read its header first.  What
// the layout changes:
//   * lanes = channels.  The 128 TMEM lanes (= MMA rows) of a CTA are FOUR UNITS of 32 consecutive channels; a unit = (batch
//     entry, channel quad, strip of time) has its own coordinates, so any channel count that is a multiple of 32 fills every
//     lane (96 = 3 units, 192 = 6: with one 128-channel group per CTA a quarter of the lanes idled there).  The 32 lanes of a
//     unit are one compute warp per group and walk the same time window: the replicate-padding cases are warp-uniform.
//   * x chunk = 64 time steps x 4 units = four tensor-map boxes [64 steps][32 channels] (3-D map (C, T, B), 64-byte swizzle): a
//     row of a box is one time step.  That is the canonical MN-MAJOR operand layout of tcgen05.mma (((8,4,m),(8,k)) :
//     ((1,8,LBO),(32,SBO)) in 16-bit elements: 32 channels contiguous, the next unit LBO = one box = 4 KB on, 8-step groups
//     SBO = 512 B apart), so the up filter still reads x straight from what the TMA unit wrote (SS mode, A MN-major:
//     instruction-descriptor bit 15); a K = 16 slice starts at ANY multiple of 8 time steps = a whole number of swizzle atoms.
//     No alignment condition on T (rows of the map are whole time steps).
//   * the bias the tensor still lacks (the caller's "pending" bias) enters after the up filter: the filter is linear and
//     replicate padding keeps all six taps of a phase, u(x + b) = u(x) + b * (sum of the phase's taps).
//   * drain: a thread holds 32 consecutive outputs of ONE channel, the out chunk wants rows of 32 channels: 2-byte stores,
//     the 32 lanes of a warp filling the 64 bytes of one swizzled row, then the TMA store as before.
//   * rows [T, T_out) of y (the zero padding a polyphase dilated convolution reads next, afa_cl_fwd_kernel) are written
//     as zeros by the same stores.
// No residual prologue here (x' = x + res, xsum = x'): tried and dropped -- the compute threads added
// res (16-byte global loads issued an iteration ahead) into the staged chunk in place and wrote xsum: bit-exact xsum, but
// that call moves 8 bytes per element and afa_cl_fwd_kernel already does it at 3.7 TB/s (88-92 us at B = 8): 75 us at
// C = 384, 85-89 us at C = 96 / 192, and the activation of the ROUNDED sum misses the 1e-2 budget against the contract
// (E = 0.8-1.4e-2).
#pragma once
#include "afa_tc_kernels.cuh"

namespace afa_tc {

constexpr int kBoxBytes = 64 * 64;       // one tensor-map box: 64 time steps x 32 channels (a unit's share of a chunk)

struct ClArgs {
    const __nv_bfloat16* x;
    const float* bias;           // optional [C] fp32
    const float* alpha;
    const float* beta;
    uint16_t up_hi[12], up_lo[12];   // 2 * upsample taps (ratio folded, resample.py:33): bf16 hi + lo
    uint16_t dn_hi[12], dn_lo[12];   // low-pass taps, same split
    float bias_even, bias_odd;   // what a unit bias adds to u[n], n even / odd: 2 * (f[1] + f[3] + ...), 2 * (f[0] + f[2] + ...)
    int64_t x_bs;                // elements between batch entries of x
    int32_t C, T, T_out, flags;
    int32_t NY;                  // outputs per unit and channel in blocks of 16 (a multiple of 4)
    int32_t n_tstrips, n_cquads;    // unit u = 4 * blockIdx.x + lane quarter = (batch * n_tstrips + tstrip) * n_cquads + channel quad
    int32_t B;                      // batch entries (units beyond the last one load zeros and store nothing)
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, int c0, int c1, int c2, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(src) : "memory");
}
// MN-major, 64-byte swizzle: 32 MN elements (64 bytes) contiguous, the next 32 `LBO` bytes on; K: rows of 64 bytes, 8-row
// groups SBO = 512 bytes apart.  (cute::UMMA::make_umma_desc<Major::MN>: leading = stride of the MN atoms, stride = stride of
// the K groups; layout type 4 = SWIZZLE_64B.)
__device__ __forceinline__ uint64_t adesc_mn_sw64(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)(512u >> 4) << 32) |
           (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {          // low 16 bits of v
    asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tst.shared.b16 [%0], lo;\n\t}" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts16_pair(uint32_t addr_lo, uint32_t addr_hi, uint32_t v) {      // low half -> addr_lo, high half -> addr_hi
    asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tst.shared.b16 [%0], lo;\n\tst.shared.b16 [%1], hi;\n\t}"
                 ::"r"(addr_lo), "r"(addr_hi), "r"(v) : "memory");
}

template <int kUpMats, int kDnMats>
__global__ void __launch_bounds__(AFA_TC_BOUND_THREADS, 2)
afa_tc_cl_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
                     const __grid_constant__ ClArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler: role code runs on the uniform datapath
    const int NY = a.NY, NB = NY >> 1;                            // NB blocks of 32 outputs per channel
    const int NCH_IN = (NY >> 2) + 1, NCH_OUT = NY >> 2;          // x chunks / out chunks of this CTA's strip
    // this thread's unit: lane quarter `uq` of the CTA (warp 0: lane uq moves box uq; compute warps: uq = warp & 3)
    const int uq = warp == 0 ? (lane & 3) : (warp & 3);
    const uint32_t unit = blockIdx.x * 4u + (uint32_t)uq;
    const int cq = (int)(unit % (uint32_t)a.n_cquads);
    const uint32_t rest = unit / (uint32_t)a.n_cquads;
    const int tstrip = (int)(rest % (uint32_t)a.n_tstrips);
    const int bi = (int)(rest / (uint32_t)a.n_tstrips);      // >= B: no such unit (the TMA unit zero-fills its loads and drops its stores)
    const int c0 = cq * 32;                         // first channel of the unit
    const int t_org = tstrip * (16 * NY) - 8;       // first staged time step; first output t_org + 8
    const int T = a.T;
    const uint32_t bars = sbase + kOffBar;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + kOffTmem);

    // Programmatic dependent launch: the next kernel of the stream may start as soon as this one's CTAs have all started; ITS
    // set-up (barriers, tensor-memory allocation, tap matrices) then runs in the shadow of this kernel's tail.  Everything that
    // touches global memory a predecessor may have written -- x, alpha / beta / bias -- comes after griddepcontrol.wait.
    pdl_launch_dependents();
    if (warp == 0) {
        for (int i = lane; i < kNumBars; i += 32) {
            const uint32_t cnt = i == kBarPre ? 8u : (i >= kBarCmp && i < kBarEv) ? 4u : (i >= kBarOut && i < kBarEvY) ? 8u : 1u;
            mbar_init(bars + 8 * i, cnt);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
        pdl_wait();
        for (int p = 0; p < kSlots && p < NCH_IN; ++p) {
            if (lane == 0) mbar_expect_tx(bars + 8 * (kBarFull + p), (uint32_t)kChunkBytes);
            __syncwarp();
            if (lane < 4)
                tma_load_3d(sbase + p * kChunkBytes + lane * kBoxBytes, &tm_x, c0, t_org + 64 * p, bi, bars + 8 * (kBarFull + p));
        }
        if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_y)) : "memory");
        __syncwarp();
    } else {
        // banded Toeplitz B matrices: identical to afa_tc_fwd_kernel (K-major core-matrix layout, hi + lo bf16 split of the taps)
        if (warp == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
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
                const int kappa = (e + 11 - tap) >> 1;              // 0 .. 13
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
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        pdl_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        // ===== TMA store of finished output chunks; the slot then takes the next x chunk of the strip =====
        if (kSlots < NCH_IN) {
            mbar_wait(bars + 8 * (kBarEv + 1), 0);
            if (lane == 0) mbar_expect_tx(bars + 8 * (kBarFull + 0), (uint32_t)kChunkBytes);
            __syncwarp();
            if (lane < 4)
                tma_load_3d(sbase + lane * kBoxBytes, &tm_x, c0, t_org + 64 * kSlots, bi, bars + 8 * (kBarFull + 0));
            __syncwarp();
        }
        for (int qc = 0; qc < NCH_OUT; ++qc) {
            const int slot = (qc + 1) % kSlots;
            mbar_wait(bars + 8 * (kBarOut + slot), (uint32_t)(qc / kSlots) & 1u);
            if (lane < 4) {
                tma_store_3d(&tm_y, c0, t_org + 8 + 64 * qc, bi, sbase + slot * kChunkBytes + lane * kBoxBytes);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            const int nc = qc + 1 + kSlots;                         // next x chunk for this slot
            if (nc < NCH_IN) {
                if (lane < 4) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_expect_tx(bars + 8 * (kBarFull + slot), (uint32_t)kChunkBytes);
                __syncwarp();
                if (lane < 4)
                    tma_load_3d(sbase + slot * kChunkBytes + lane * kBoxBytes, &tm_x, c0, t_org + 64 * nc, bi, bars + 8 * (kBarFull + slot));
            }
            __syncwarp();
        }
        if (lane < 4) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // InstrDescriptor: D f32 [4,6) = 1, A bf16 [7,10) = 1, B bf16 [10,13) = 1, A major [15] (1 = MN-major: the up filter's
        // x slices; the down filter's A operand lives in tensor memory), N >> 3 [17,23), M >> 4 [24,29)
        constexpr uint32_t idesc_dn = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        constexpr uint32_t idesc_up = idesc_dn | (1u << 15);
        const uint64_t bup = bdesc_kmajor(sbase + kOffWup, 16 * 16);
        const uint64_t bdn = bdesc_kmajor(sbase + kOffWdn, 16 * 16);
        const uint64_t ax = adesc_mn_sw64(sbase, kBoxBytes);             // chunk slot s: + s * (kChunkBytes >> 4); 8 time steps = 512 B = 32 units of 16 B
        constexpr uint64_t kW = kWBytes >> 4;
        auto up = [&](uint32_t d, int bu, int xs) {
            const uint64_t cb = ax + (uint64_t)(xs * (kChunkBytes >> 4) + (bu & 1) * 128);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                if (m == 3 && (bu & 1)) {
                    // time steps 56 .. 69 of the chunk: steps 48..63 of this chunk against the taps moved down 8 rows, steps 0..15 of
                    // the next chunk against the taps moved up 8
                    const int xs1 = xs + 1 == kSlots ? 0 : xs + 1;
                    const uint64_t a1 = ax + (uint64_t)(xs * (kChunkBytes >> 4) + 6 * 32), a2 = ax + (uint64_t)(xs1 * (kChunkBytes >> 4));
                    mma_ss(d + 48, a1, bup + 1 * kW, idesc_up, 0);
                    if (kUpMats == 2) mma_ss(d + 48, a1, bup + (kUpVariants + 1) * kW, idesc_up, 1);
                    mma_ss(d + 48, a2, bup + 2 * kW, idesc_up, 1);
                    if (kUpMats == 2) mma_ss(d + 48, a2, bup + (kUpVariants + 2) * kW, idesc_up, 1);
                } else {
                    mma_ss(d + 16 * m, cb + 32 * m, bup, idesc_up, 0);
                    if (kUpMats == 2) mma_ss(d + 16 * m, cb + 32 * m, bup + kUpVariants * kW, idesc_up, 1);
                }
            }
        };
        auto down = [&](int bd) {
            const uint32_t sl0 = tmem + (uint32_t)(kSlotCols * (bd & 3)), sl1 = tmem + (uint32_t)(kSlotCols * ((bd + 1) & 3));
            const uint32_t d = sl1 + 32;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int k = 0; k < kDnSlices; ++k) {
                    const int col = 16 * h + 8 * k;
                    const uint32_t aa = col < 32 ? sl0 + (uint32_t)col : sl1;
                    mma_ts(d + 16 * h, aa, bdn + (uint64_t)k * kW, idesc_dn, k > 0);
                    if (kDnMats == 2) mma_ts(d + 16 * h, aa, bdn + (uint64_t)(kDnSlices + k) * kW, idesc_dn, 1);
                }
        };
        mbar_wait(bars + 8 * kBarPre, 0);
        mbar_wait(bars + 8 * (kBarFull + 0), 0);
        mbar_wait(bars + 8 * (kBarFull + 1), 0);
        int nfull = 2;
        tc_fence_after();
        if (elect_one()) {                                   // events -3 .. -1
            up(tmem + 0 * kSlotCols, 0, 0); tc_commit(bars + 8 * (kBarEv + 0));
            up(tmem + 1 * kSlotCols, 1, 0); tc_commit(bars + 8 * (kBarEv + 1));
            up(tmem + 2 * kSlotCols, 2, 1); tc_commit(bars + 8 * (kBarEv + 2));
        }
        __syncwarp();
        int xs = 1;
        for (int e = 0; e <= NB; ++e) {
            const int bu = e + 3;
            if (bu <= NB) {
                const int p = (bu + 1) >> 1;
                while (nfull <= p) { mbar_wait(bars + 8 * (kBarFull + nfull % kSlots), (uint32_t)(nfull / kSlots) & 1u); ++nfull; }
            }
            mbar_wait(bars + 8 * (kBarCmp + (e & 7)), (uint32_t)(e >> 3) & 1u);
            tc_fence_after();
            if (elect_one()) {
                if (e >= 1) down(e - 1);
                tc_commit(bars + 8 * (kBarEvY + (e & 7)));
                if (bu <= NB) up(tmem + (uint32_t)(kSlotCols * (bu & 3)), bu, xs);
                tc_commit(bars + 8 * (kBarEv + (bu & 7)));
            }
            __syncwarp();
            if (bu & 1) xs = xs + 1 == kSlots ? 0 : xs + 1;
        }
    } else {
        // ===== compute groups =====
        const int grp = (warp - 2) >> 2;
        const int q = warp & 3;                                  // TMEM lane quarter this warp may access
        const int ch = c0 + lane;                                // TMEM lane q * 32 + lane = channel `lane` of unit q
        const bool active = bi < a.B && ch < a.C;
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        // this channel's 2 bytes inside a row (= time step) of a chunk: box q, 16-byte unit (lane / 8) ^ ((row / 2) % 4)
        const uint32_t colbase = sbase + (uint32_t)q * kBoxBytes + (uint32_t)(lane & 7) * 2u;
        const uint32_t junit = (uint32_t)lane >> 3;
        uint32_t jx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) jx[k] = colbase + ((junit ^ (uint32_t)k) << 4);
        float a_eff = 1.f, ib = 1.f, bias_v = 0.f;
        const bool left_cta = t_org < 0;                         // the unit's strip starts the row (t_org = -8); warp-uniform
        const int rT = T - t_org;                                // first staged row beyond the tensor (window row index)
        if (active) {
            float al = __ldg(a.alpha + ch);
            float be = (a.flags & 2) ? al : __ldg(a.beta + ch);
            if (a.flags & 1) { al = expf(al); be = expf(be); }
            a_eff = al;
            ib = 1.0f / (be + 0.000000001f);
            if (a.bias) bias_v = __ldg(a.bias + ch);
        }
        // replicate padding of x (resample.py:32), patched into the staged chunks by group 0 (the TMA unit zero-filled the time
        // steps outside the tensor): x[0] over the 8 steps before the row, x[T-1] over the 16 steps behind it.  Every thread
        // patches its own channel.  Chunk c is patched after it has landed and before the first product that reads it (up(2c-1),
        // issued at event 2c-4): chunks 0, 1, 2 here, chunk c >= 3 at iteration 2c-6 of group 0, ahead of that iteration's arrive.
        auto patch_chunk = [&](int c) {
            const int r_lo = max(rT, 64 * c), r_hi = min(rT + 16, 64 * c + 64);
            const bool left = left_cta && c == 0, right = r_lo < r_hi;
            if (left || right) {                                 // warp-uniform
                mbar_wait(bars + 8 * (kBarFull + c % kSlots), (uint32_t)(c / kSlots) & 1u);
                if (active) {
                    const unsigned short* xc = reinterpret_cast<const unsigned short*>(a.x) + (size_t)bi * (size_t)a.x_bs + ch;
                    const uint32_t cbase = (uint32_t)(c % kSlots) * kChunkBytes;
                    if (left) {
                        const uint32_t v = __ldg(xc);
#pragma unroll
                        for (int r = 0; r < 8; ++r) sts16(cbase + jx[(r >> 1) & 3] + (uint32_t)r * 64u, v);
                    }
                    if (right) {
                        const uint32_t v = __ldg(xc + (size_t)(T - 1) * (size_t)a.C);
                        for (int r = r_lo; r < r_hi; ++r) {
                            const int rr = r - 64 * c;
                            sts16(cbase + (colbase + ((junit ^ (uint32_t)((rr >> 1) & 3)) << 4)) + (uint32_t)rr * 64u, v);
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
        const float2 bias2 = make_float2(bias_v * a.bias_even, bias_v * a.bias_odd);     // element e of a block is u[n], n = e (mod 2)
        // Snake on packed pairs, 16 values at a time: s = u + ib * sin^2(a u), u = (up filter of x) + bias   activations.py:60, :124
        auto snake16 = [&](const uint32_t (&u)[16], uint32_t* sp) {
            float2 u2[8], sn[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                u2[e] = __fadd2_rn(make_float2(__uint_as_float(u[2 * e]), __uint_as_float(u[2 * e + 1])), bias2);
                const float2 th = __fmul2_rn(u2[e], a2);
                sn[e] = make_float2(__sinf(th.x), __sinf(th.y));
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float2 s2 = __ffma2_rn(ib2, __fmul2_rn(sn[e], sn[e]), u2[e]);
                sp[e] = pack_bf16(s2.x, s2.y);
            }
        };
        for (int b = grp; b <= NB + 1; b += 2) {
            const bool has_u = b <= NB, has_y = b >= 2;
            if (has_u) {
                mbar_wait(bars + 8 * (kBarEv + (b & 7)), (uint32_t)(b >> 3) & 1u);
                tc_fence_after();
                if ((b & 1) == 0 && (b >> 1) + 3 < NCH_IN) patch_chunk((b >> 1) + 3);     // group 0 only: b even
                const uint32_t tslot = tlane + (uint32_t)(kSlotCols * (b & 3));
                uint32_t ua[16], ub[16], uc[16], sp[16];
                tmem_ld16(tslot, ua);
                tmem_ld16(tslot + 16, ub);
                tmem_wait_ld();
                snake16(ua, sp);
                tmem_ld16(tslot + 32, ua);
                tmem_ld16(tslot + 48, uc);
                snake16(ub, sp + 8);
                // replicate padding of the ACTIVATED signal (filter.py:98), in the 2x domain; warp-uniform here.  eb = element of
                // n = 2T in this block (10, 26, 42, 58 -- or 2, 18, 34, 50 when T % 8 == 4 -- when inside): elements >= eb repeat
                // element eb - 1.
                const int eb = 2 * (T - t_org) - 64 * b - 6;
                const bool lclamp = left_cta && b == 0, rclamp = eb >= 2 && eb <= 58;
                uint32_t fill = 0;
                auto clamp_from = [&](int kc) {                   // kc in {1, 5, 9, 13}: fill = upper element of register kc - 1
                    fill = __byte_perm(kc == 1 ? sp[0] : kc == 5 ? sp[4] : kc == 9 ? sp[8] : sp[12], 0, 0x3232);
#pragma unroll
                    for (int e = 1; e < 16; ++e)
                        if (e >= kc) sp[e] = fill;
                };
                if (lclamp) {                                     // n < 0 <-> e < 10: s[0] is element 10
                    const uint32_t s0 = __byte_perm(sp[5], 0, 0x1010);
#pragma unroll
                    for (int e = 0; e < 5; ++e) sp[e] = s0;
                }
                if (rclamp && eb < 32) clamp_from(eb >> 1);
                tmem_st16(tslot, sp);
                tmem_wait_ld();
                snake16(ua, sp);
                snake16(uc, sp + 8);
                if (rclamp) {
                    if (eb < 32) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) sp[e] = fill;
                    } else {
                        clamp_from((eb - 32) >> 1);
                    }
                }
                tmem_st16(tslot + 16, sp);
            }
            if (has_y) {
                // Y(b-2) (upper half of slot (b-1) % 4): fp32 accumulators -> bf16 -> this channel's column of out chunk (b-2)/2,
                // staged in the ring slot of x chunk (b-2)/2 + 1
                const int i = b - 2;
                mbar_wait(bars + 8 * (kBarEvY + ((b - 1) & 7)), (uint32_t)((b - 1) >> 3) & 1u);
                tc_fence_after();
                const uint32_t yslot = tlane + (uint32_t)(kSlotCols * ((i + 1) & 3) + 32);
                uint32_t ya[16], yb[16];
                tmem_ld16(yslot, ya);
                tmem_ld16(yslot + 16, yb);
                tmem_wait_ld();
                const int tb0 = t_org + 8 + 32 * i;               // time step of element 0
                if (tb0 + 32 > T) {                               // rows [T, T_out) are the zero padding the next convolution reads
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        if (tb0 + e >= T) ya[e] = 0u;
                        if (tb0 + 16 + e >= T) yb[e] = 0u;
                    }
                }
                const uint32_t rbase = (uint32_t)(((i >> 1) + 1) % kSlots) * kChunkBytes + (uint32_t)(i & 1) * (32u * 64u);
#if !(defined(AFA_TC_CL_EXPERIMENT) && AFA_TC_CL_EXPERIMENT == 1)      // harness only: 1 = no drain stores (results are wrong)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const uint32_t p0 = pack_bf16(__uint_as_float(ya[2 * e]), __uint_as_float(ya[2 * e + 1]));
                    const uint32_t p1 = pack_bf16(__uint_as_float(yb[2 * e]), __uint_as_float(yb[2 * e + 1]));
                    // rows 2e, 2e+1 and 16+2e, 17+2e of this block: (row / 2) % 4 = e % 4 for all four (static)
                    sts16_pair(rbase + jx[e & 3] + (uint32_t)(2 * e) * 64u, rbase + jx[e & 3] + (uint32_t)(2 * e + 1) * 64u, p0);
                    sts16_pair(rbase + jx[e & 3] + (uint32_t)(16 + 2 * e) * 64u, rbase + jx[e & 3] + (uint32_t)(17 + 2 * e) * 64u, p1);
                }
#else
                if (ya[0] == 0x12345678u && yb[3] == 0x9abcdef0u) sts16(rbase + jx[0], ya[1]);
#endif
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
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
    }
}

}  // namespace afa_tc
