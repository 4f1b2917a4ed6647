// dmel_dftmm_kernels.cuh — forward for short windows (any even n_fft <= 128) as a DFT-by-matrix-multiply on the
// 5th-generation tensor cores: tcgen05.mma (kind::tf32, 3xTF32 split for fp32 accuracy), operands in 128-byte-swizzled
// shared memory, accumulators in tensor memory, |X|^2 + mel projection + log in the epilogue straight from TMEM.
//
// north_star: "a DFT-as-GEMM on tcgen05 tensor cores for small n_fft where ncu shows it wins"; the reference's own
// 10 ms grid point at 8 kHz is n_fft 128 (search_spaces.py:29,62).  The replaced chain is the same as everywhere:
// time_frequency.py:21-58 (Gaussian window, torch.stft, |.|^2) -> melscale_fbanks matmul (models.py:42-53) -> log.
//
// Formulation.  A frame xw[n] = (x[n] - mean) * w[n], n < L = n_fft, is folded around its centre:
//     e[j] = xw[j] + xw[L-j],  o[j] = xw[j] - xw[L-j]   (j = 1 .. L/2-1),   e[0] = xw[0]
//     Re X[k] = sum_j e[j] cos(2 pi j k / L) + (-1)^k xw[L/2],      Im X[k] = - sum_j o[j] sin(2 pi j k / L)
// i.e. two real GEMMs with K = L/2 <= 64 instead of one with K = L and twice the columns: half the flops and half the
// operand bytes.  One CTA tile = 128 consecutive frames of the batch (rows = UMMA M; frames are numbered across clips, so
// every row of every tile but the last is used):
//     D_cos[128 x 80] = E[128 x 64] * COS^T,    D_sin[128 x 64] = O[128 x 64] * (-SIN)^T
// fp32 operands are split x = hi + lo with hi = tf32(x): D = A_hi B_hi + A_lo B_hi + A_hi B_lo (the dropped lo*lo term
// is 2^-22 relative), three tensor-core passes into the same fp32 accumulator.
//
// Roles (512 threads, one CTA per SM, persistent over tiles):
//   warps 0-7  build E and O for the tile: coalesced reads of x from L2 into registers, 32 independent loads in flight per
//              thread, issued BEFORE waiting for the tensor core to release the operands (the previous tile's MMAs run
//              under that latency); window in registers, hi/lo split, swizzled STS; lane 0 of warp 0 issues the MMAs
//   warps 8-15 epilogue, two sets of four (one warp per TMEM lane quarter, one set per accumulator stage), thread = frame:
//              phase 1 reads the row of D_cos / D_sin from TMEM 16 columns at a time (tcgen05.ld), forms
//              P[k] = Re^2 + Im^2 and writes it back over the cosine accumulator (tcgen05.st); phase 2 projects mel by mel
//              from TMEM — the one on-chip store a thread can index dynamically without shared memory — four bins per
//              read against a float4 of weights (mel-major bank re-based to multiples of 4, built once per CTA), and
//              stores log(mel + eps): consecutive threads hold consecutive frames, so the stores along T are coalesced.
//              No data-dependent control flow, a few hundred instructions of code (the first version walked the
//              bin-major bank with running accumulators through 5 k instructions of unrolled branches: 20 us per tile)
// TMEM holds two accumulator stages (2 x 256 columns): the epilogue of tile i overlaps the build + MMA of tiles i+1, i+2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dmel_common.cuh"

namespace dmel {
namespace mm {

constexpr int ROWS = 128;      // frames per tile (UMMA M)
constexpr int KF = 64;         // folded samples per frame (UMMA K total), n_fft / 2 <= KF
constexpr int NCOS = 80;       // columns of D_cos: bins 0..64, padded to a multiple of 16
constexpr int NSIN = 64;       // columns of D_sin: bins 0..63 (the sine is zero at bins 0 and n_fft/2)
constexpr int BUILD_WARPS = 8;  // operand builders
constexpr int EPI_WARPS = 8;    // epilogue: two sets of four warps (one per TMEM lane quarter), one set per accumulator stage
constexpr int THREADS = (BUILD_WARPS + EPI_WARPS) * 32;
constexpr int MAX_MELS = 256;   // mel rows the chunk table is sized for
constexpr int MAX_CHUNKS = 36 + 2 * MAX_MELS;   // 4-tap chunks: sum_m ceil((taps_m + 3) / 4) <= nnz/4 + 2 per mel, nnz <= 2*65
constexpr int MAX_NFFT = 128;

// shared-memory carve-up, bytes from a 1024-byte aligned base.  A slab = [rows x 32 floats], one 128-byte row per
// matrix row, 16-byte chunks XOR-swizzled with (row & 7): the canonical K-major SWIZZLE_128B operand layout.
constexpr int A_SLAB = ROWS * 128;                 // 16 KB
constexpr int A_MAT = 2 * A_SLAB;                  // K = 64 -> two slabs
constexpr int OFF_E_HI = 0, OFF_E_LO = A_MAT, OFF_O_HI = 2 * A_MAT, OFF_O_LO = 3 * A_MAT;
constexpr int BC_SLAB = NCOS * 128, BS_SLAB = NSIN * 128;
constexpr int OFF_BC_HI = 4 * A_MAT, OFF_BC_LO = OFF_BC_HI + 2 * BC_SLAB;
constexpr int OFF_BS_HI = OFF_BC_LO + 2 * BC_SLAB, OFF_BS_LO = OFF_BS_HI + 2 * BS_SLAB;
constexpr int OFF_MISC = OFF_BS_LO + 2 * BS_SLAB;  // window, bank records, barriers
constexpr int MISC_WIN = 0;                        // float[132]
constexpr int MISC_BAR = 528;                      // 6 mbarriers + tmem pointer
constexpr int MISC_META = MISC_BAR + 128;          // int2[MAX_MELS]: first (4-aligned) bin | chunks << 16, first chunk
constexpr int MISC_WT = MISC_META + MAX_MELS * 8;  // float4[MAX_CHUNKS]: the weights of a chunk of 4 bins
constexpr int MISC_BYTES = MISC_WT + MAX_CHUNKS * 16;
constexpr int SMEM_BYTES = OFF_MISC + MISC_BYTES + 1024;   // + slack for the manual 1024-byte alignment
static_assert(SMEM_BYTES <= 227 * 1024, "tile does not fit");

struct MmGeom {
    int B, N, T, hop, L, M;      // M = n_mels, or 0 in spectrogram mode (then rows = L/2+1)
    long long x_stride;
    float log_eps;               // < 0: linear output
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, M = 128, K = 8 per instruction, operands read as tf32
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return uint64_t((saddr & 0x3FFFFu) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
           (uint64_t(2) << 61);
}
// instruction descriptor: D fp32, A and B tf32, both K-major, M = 128
__host__ __device__ constexpr uint32_t instr_desc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(ROWS >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        // the wait is part of the same statement: the destination registers may not be touched before it
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// TMEM reads cost a long round trip (issue -> wait::ld), so the epilogue issues several loads per wait.  The loads are
// issued without their wait (tmem_ld16_issue), one tmem_ld_fence() waits for all of them, and tie16() makes every later
// use of the registers depend on a volatile statement that follows the wait (volatile asm statements keep their order).
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_fence() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tie16(float (&v)[16]) {
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
                      "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
          "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
          "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
          "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
          "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// four consecutive columns of this thread's TMEM lane (column a multiple of 4)
__device__ __forceinline__ float4 tmem_ld4(uint32_t taddr) {
    uint32_t a, b, c, d;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
        : "r"(taddr)
        : "memory");
    return make_float4(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c), __uint_as_float(d));
}

// x = hi + lo, hi = x rounded to tf32 (the tensor core reads the upper 19 bits of lo)
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    hi = __uint_as_float(h);
    lo = v - hi;
}
// byte offset of (row, col) inside a slab, col in [0, 32)
__device__ __forceinline__ int sw_off(int row, int col) { return row * 128 + ((((col >> 2) ^ row) & 7) << 4) + ((col & 3) << 2); }

// ---------------------------------------------------------------------------------------------- kernel
__global__ void __launch_bounds__(THREADS, 1)
dmel_dftmm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, MmGeom g, MelDev mel,
                      const float* __restrict__ win_g, const float* __restrict__ means, WsHeader* hdr) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* base = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(base);
    float* win_s = reinterpret_cast<float*>(base + OFF_MISC + MISC_WIN);
    int2* meta_s = reinterpret_cast<int2*>(base + OFF_MISC + MISC_META);
    float4* wt_s = reinterpret_cast<float4*>(base + OFF_MISC + MISC_WT);
    const uint32_t bar0 = sbase + OFF_MISC + MISC_BAR;
    const uint32_t BAR_A_FREE = bar0, BAR_FULL0 = bar0 + 16, BAR_EMPTY0 = bar0 + 32;
    volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(base + OFF_MISC + MISC_BAR + 64);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = g.L, H = L >> 1;           // bins 0..H, folded samples j = 0..H-1
    const bool spec = g.M == 0;
    const int rows_out = spec ? H + 1 : g.M;

    // ---- one-time set-up: window, bank records, the constant B operands, barriers, tensor memory
    for (int n = tid; n < 132; n += THREADS) win_s[n] = n < L ? __ldg(win_g + n) : 0.f;
    if (!spec) {
        // mel-major chunk table for the epilogue (models.py:53 matmul(spec^T, mel_fb), mel by mel): the run of bins of mel m
        // (dmel.h mel-major bank: first bin k0, taps zero-padded to a multiple of 4) is re-based to a multiple of 4 — TMEM
        // is read four columns at a time — so a mel is `chunks` float4 weight records starting at bin k0 & ~3
        for (int m = tid; m < g.M; m += THREADS) {
            const int k0 = __ldg(mel.mel_k0 + m), taps = __ldg(mel.mel_off + m + 1) - __ldg(mel.mel_off + m);
            const int used = min(taps, H + 1 - k0);                     // padded taps past the last bin carry zero weights
            meta_s[m] = make_int2((k0 & ~3) | (((used > 0 ? (k0 & 3) + used + 3 : 0) >> 2) << 16), 0);
        }
        __syncthreads();
        if (tid == 0) {
            int off = 0;
            for (int m = 0; m < g.M; ++m) {
                meta_s[m].y = off;
                off += meta_s[m].x >> 16;
            }
        }
        __syncthreads();
        for (int m = tid; m < g.M; m += THREADS) {
            const int k0 = __ldg(mel.mel_k0 + m), o0 = __ldg(mel.mel_off + m);
            const int used = min(__ldg(mel.mel_off + m + 1) - o0, H + 1 - k0);
            const int2 mt = meta_s[m];
            float* w = reinterpret_cast<float*>(wt_s + mt.y);
            const int n = (mt.x >> 16) * 4, lead = k0 & 3;
            for (int i = 0; i < n; ++i) w[i] = (i >= lead && i - lead < used) ? __ldg(mel.mel_w + o0 + i - lead) : 0.f;
        }
    }
    // exp(-2 pi i n / L), n < L, once; the B operands are look-ups into it
    float2* tw_s = reinterpret_cast<float2*>(base + OFF_E_HI);     // scratch: the A region is not in use yet
    for (int n = tid; n < L; n += THREADS) {
        float sn, cs;
        sincospif(2.0f * float(n) / float(L), &sn, &cs);
        tw_s[n] = make_float2(cs, -sn);
    }
    __syncthreads();
    const bool pow2 = (L & (L - 1)) == 0;
    for (int e = tid; e < (NCOS + NSIN) * KF; e += THREADS) {
        const bool is_cos = e < NCOS * KF;
        const int ee = is_cos ? e : e - NCOS * KF;
        const int k = ee / KF, j = ee - k * KF;      // output bin (row of B), folded sample (column)
        float v = 0.f;
        if (j < H && k <= H) {
            const float2 w = tw_s[pow2 ? ((j * k) & (L - 1)) : ((j * k) % L)];
            v = is_cos ? w.x : w.y;
        }
        float hi, lo;
        split_tf32(v, hi, lo);
        const int off = (j >> 5) * (is_cos ? BC_SLAB : BS_SLAB) + sw_off(k, j & 31);
        *reinterpret_cast<float*>(base + (is_cos ? OFF_BC_HI : OFF_BS_HI) + off) = hi;
        *reinterpret_cast<float*>(base + (is_cos ? OFF_BC_LO : OFF_BS_LO) + off) = lo;
    }
    if (tid == 0) {
        mbar_init(BAR_A_FREE, 1);
        mbar_init(BAR_FULL0, 1);
        mbar_init(BAR_FULL0 + 8, 1);
        mbar_init(BAR_EMPTY0, 128);
        mbar_init(BAR_EMPTY0 + 8, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_s)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();        // the B operands were written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    const long long total_rows = (long long)g.B * g.T;
    const int n_tiles = int((total_rows + ROWS - 1) / ROWS);

    if (warp < BUILD_WARPS) {
        // =========================================================== builders (+ the MMA issuer: warp 0, lane 0)
        // A warp owns RPW consecutive rows of the tile; a lane owns the folded samples j0 = lane, j1 = lane + 32 and their
        // partners L - j (none for j = 0: its partner weight is zero).  Rows go in batches of RB: the raw samples of a batch
        // are loaded into registers first (RB * 4 independent loads per thread), then E and O are formed and stored.
        constexpr int RPW = ROWS / BUILD_WARPS;   // rows of the tile per warp
        constexpr int RB = 8;                     // rows per batch: RB * 4 independent loads in flight per thread
        const int j0 = lane, j1 = lane + 32;
        const float w0a = j0 < H ? win_s[j0] : 0.f, w0b = (j0 < H && j0 > 0) ? win_s[L - j0] : 0.f;
        const float w1a = j1 < H ? win_s[j1] : 0.f, w1b = j1 < H ? win_s[L - j1] : 0.f;
        // sample offsets inside a frame (clamped for lanes beyond a short transform: their weights are zero)
        const int q0a = j0, q0b = L - (j0 > 0 ? j0 : L), q1a = j1 < L ? j1 : 0, q1b = j1 < L ? L - j1 : 0;
        const bool issuer = tid == 0;
        unsigned char* const e_hi = base + OFF_E_HI;
        unsigned char* const e_lo = base + OFF_E_LO;
        unsigned char* const o_hi = base + OFF_O_HI;
        unsigned char* const o_lo = base + OFF_O_LO;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int stage = it & 1, use = it >> 1;
            // first row of this warp: (clip, frame); the following rows advance frame by frame
            long long f = (long long)tile * ROWS + warp * RPW;
            int b = int(f / g.T), t = int(f - (long long)b * g.T);
            float mu = (f < total_rows) ? __ldg(means + b) : 0.f;
#pragma unroll 1
            for (int batch = 0; batch < RPW / RB; ++batch) {
                // ---- raw samples of RB rows into registers: every load is independent of every other instruction here
                float ra[RB], rb[RB], rc[RB], rd[RB], mus[RB];
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    float a = mu, bb = mu, c = mu, dd = mu;     // out of range: (mu - mu) * w = 0, the zero padding of torch.stft
                    if (f < total_rows) {
                        const float* xb = x + (long long)b * g.x_stride;
                        const int p0 = t * g.hop - H;          // frame start (center=True)
                        const int pa = p0 + q0a, pb = p0 + q0b, pc = p0 + q1a, pd = p0 + q1b;
#ifndef DMEL_MM_EXP_NO_LOADS
                        if (pa >= 0 && pa < g.N) a = __ldg(xb + pa);
                        if (pb >= 0 && pb < g.N) bb = __ldg(xb + pb);
                        if (pc >= 0 && pc < g.N) c = __ldg(xb + pc);
                        if (pd >= 0 && pd < g.N) dd = __ldg(xb + pd);
#endif
                    }
                    ra[rr] = a; rb[rr] = bb; rc[rr] = c; rd[rr] = dd; mus[rr] = mu;
                    ++f;
                    if (++t == g.T) {
                        t = 0;
                        ++b;
                        mu = (f < total_rows) ? __ldg(means + b) : 0.f;
                    }
                }
                // the tensor core must be done reading the previous tile's operands before they are overwritten; its MMAs
                // ran while the loads above were in flight
                if (batch == 0 && it > 0) mbar_wait(BAR_A_FREE, uint32_t(it - 1) & 1u);
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    const int r = warp * RPW + batch * RB + rr;
                    const float m_ = mus[rr];
                    // xw[j] and xw[L-j]   (time_frequency.py:21-30 window, models.py:38 mean removal), then E = sum, O = difference
                    const float a0 = (ra[rr] - m_) * w0a, b0 = (rb[rr] - m_) * w0b;
                    const float a1 = (rc[rr] - m_) * w1a, b1 = (rd[rr] - m_) * w1b;
                    float h, l;
                    const int off = sw_off(r, lane);
                    split_tf32(a0 + b0, h, l);
                    *reinterpret_cast<float*>(e_hi + off) = h;
                    *reinterpret_cast<float*>(e_lo + off) = l;
                    split_tf32(a1 + b1, h, l);
                    *reinterpret_cast<float*>(e_hi + A_SLAB + off) = h;
                    *reinterpret_cast<float*>(e_lo + A_SLAB + off) = l;
                    split_tf32(a0 - b0, h, l);
                    *reinterpret_cast<float*>(o_hi + off) = h;
                    *reinterpret_cast<float*>(o_lo + off) = l;
                    split_tf32(a1 - b1, h, l);
                    *reinterpret_cast<float*>(o_hi + A_SLAB + off) = h;
                    *reinterpret_cast<float*>(o_lo + A_SLAB + off) = l;
                }
            }
            fence_async_smem();                                  // generic-proxy writes -> visible to the tensor core
            asm volatile("bar.sync 1, %0;" ::"n"(BUILD_WARPS * 32) : "memory");      // the builder warps
            if (issuer) {
                if (use > 0) mbar_wait(BAR_EMPTY0 + 8 * stage, uint32_t(use - 1) & 1u);   // the epilogue drained this stage
                tc_fence_after();
#pragma unroll
                for (int half = 0; half < 2; ++half) {               // cosine GEMM, sine GEMM
                    const uint32_t d = tmem_base + uint32_t(stage * 256 + (half == 0 ? 0 : 128));
                    const uint32_t idesc = half == 0 ? instr_desc(NCOS) : instr_desc(NSIN);
                    const uint32_t a_h = sbase + (half == 0 ? OFF_E_HI : OFF_O_HI), a_l = sbase + (half == 0 ? OFF_E_LO : OFF_O_LO);
                    const uint32_t b_h = sbase + (half == 0 ? OFF_BC_HI : OFF_BS_HI), b_l = sbase + (half == 0 ? OFF_BC_LO : OFF_BS_LO);
                    const uint32_t b_slab = half == 0 ? BC_SLAB : BS_SLAB;
                    uint32_t acc = 0;
#pragma unroll
#ifdef DMEL_MM_EXP_ONE_TERM
                    for (int term = 0; term < 1; ++term) {
#else
                    for (int term = 0; term < 3; ++term) {           // hi*hi, lo*hi, hi*lo
#endif
                        const uint32_t a0 = term == 1 ? a_l : a_h, b0 = term == 2 ? b_l : b_h;
#pragma unroll
                        for (int ks = 0; ks < KF / 8; ++ks) {        // 8 tf32 per instruction, 4 instructions per slab
                            const uint32_t ao = a0 + (ks >> 2) * A_SLAB + (ks & 3) * 32;
                            const uint32_t bo = b0 + (ks >> 2) * b_slab + (ks & 3) * 32;
                            umma_tf32(d, smem_desc(ao), smem_desc(bo), idesc, acc);
                            acc = 1;
                        }
                    }
                }
                tc_commit(BAR_A_FREE);                // arrives when the MMAs have read their operands
                tc_commit(BAR_FULL0 + 8 * stage);     // ... and the tile's accumulators are in tensor memory
            }
        }
    } else {
        // =========================================================== epilogue: thread = frame row
        // two sets of four warps; set s serves the tiles whose accumulators are in TMEM stage s
        const int set = (warp - BUILD_WARPS) >> 2, q = (warp - BUILD_WARPS) & 3, r = q * 32 + lane;
        const bool do_log = g.log_eps >= 0.f;
        const float eps = g.log_eps;
        float ysum = 0.f;
        int it = set;
        for (int tile = blockIdx.x + set * int(gridDim.x); tile < n_tiles; tile += 2 * int(gridDim.x), it += 2) {
            const int stage = set, use = it >> 1;
            const long long f = (long long)tile * ROWS + r;
            const bool valid = f < total_rows;
            int b = 0, t = 0;
            float nyq = 0.f;
            if (valid) {
                b = int(f / g.T);
                t = int(f - (long long)b * g.T);
                const int pc = t * g.hop;                      // sample under the window centre: n = L/2
                if (pc < g.N) nyq = (__ldg(x + (long long)b * g.x_stride + pc) - __ldg(means + b)) * win_s[H];
            }
            // rows are emitted in order 0, 1, 2, ...: one pointer bump per output
            float* yp = y + ((long long)b * rows_out) * g.T + t;
            auto emit = [&](float v) {
                if (valid) {
                    const float val = do_log ? __logf(v + eps) : v;
                    *yp = val;
                    yp += g.T;
                    ysum += val;
                }
            };
            mbar_wait(BAR_FULL0 + 8 * stage, uint32_t(use) & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(stage * 256);
            // ---- phase 1: P[k] = Re^2 + Im^2 (time_frequency.py:53), 32 bins per TMEM round trip.  Spectrogram mode emits
            // them; mel mode writes them back over the cosine accumulator (this thread's own lane), where phase 2 can index them
            auto power16 = [&](int c, float (&re)[16], const float (&im)[16]) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int k = 16 * c + i;
                    const float rr_ = re[i] + ((k & 1) ? -nyq : nyq);
                    const float ii_ = k < H ? im[i] : 0.f;
                    re[i] = k <= H ? fmaf(rr_, rr_, ii_ * ii_) : 0.f;
                    if (spec && k <= H) emit(re[i]);
                }
                if (!spec) tmem_st16(trow + 16 * c, re);
            };
#pragma unroll 1
            for (int c = 0; c < NSIN / 16; c += 2) {
                float re0[16], im0[16], re1[16], im1[16];
                tmem_ld16_issue(trow + 16 * c, re0);
                tmem_ld16_issue(trow + 128 + 16 * c, im0);
                tmem_ld16_issue(trow + 16 * c + 16, re1);
                tmem_ld16_issue(trow + 128 + 16 * c + 16, im1);
                tmem_ld_fence();
                tie16(re0); tie16(im0); tie16(re1); tie16(im1);
                power16(c, re0, im0);
                power16(c + 1, re1, im1);
            }
            {
                float re0[16], im0[16] = {};
                tmem_ld16_issue(trow + NSIN, re0);       // bins 64..79: bin n_fft/2 (real) and the zero padding
                tmem_ld_fence();
                tie16(re0);
                power16(NSIN / 16, re0, im0);
            }
#ifdef DMEL_MM_EXP_NO_PHASE2
            if (false) {
#else
            if (!spec) {
#endif
                tmem_st_wait();
                // ---- phase 2: mel by mel from TMEM, four mels per round trip: one 16-column read per mel from its first
                // (4-aligned) bin covers up to four 4-tap chunks; wider filters finish with 4-column reads
                for (int m0 = 0; m0 < g.M; m0 += 4) {
                    float p[4][16];
                    int2 mt[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        mt[i] = (m0 + i < g.M) ? meta_s[m0 + i] : make_int2(0, 0);
                        tmem_ld16_issue(trow + uint32_t(mt[i].x & 0xffff), p[i]);
                    }
                    tmem_ld_fence();
#pragma unroll
                    for (int i = 0; i < 4; ++i) tie16(p[i]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (m0 + i < g.M) {
                            const int nch = mt[i].x >> 16;
                            const float4* w = wt_s + mt[i].y;
                            float acc = 0.f;
#pragma unroll
                            for (int ch = 0; ch < 4; ++ch) {
                                if (ch < nch) {
                                    const float4 wv = w[ch];
                                    acc = fmaf(wv.x, p[i][4 * ch], acc);
                                    acc = fmaf(wv.y, p[i][4 * ch + 1], acc);
                                    acc = fmaf(wv.z, p[i][4 * ch + 2], acc);
                                    acc = fmaf(wv.w, p[i][4 * ch + 3], acc);
                                }
                            }
                            for (int ch = 4; ch < nch; ++ch) {
                                const float4 pv = tmem_ld4(trow + uint32_t(mt[i].x & 0xffff) + 4 * ch), wv = w[ch];
                                acc = fmaf(wv.x, pv.x, acc);
                                acc = fmaf(wv.y, pv.y, acc);
                                acc = fmaf(wv.z, pv.z, acc);
                                acc = fmaf(wv.w, pv.w, acc);
                            }
                            emit(acc);
                        }
                    }
                }
            }
            // all reads of this stage are done: hand it back to the tensor core
            tc_fence_before();
            mbar_arrive(BAR_EMPTY0 + 8 * stage);
        }
        // fused by-product: sum of all outputs (train.py:58)
        double ys = double(ysum);
        for (int o = 16; o > 0; o >>= 1) ys += __shfl_xor_sync(0xffffffffu, ys, o);
        if (lane == 0) atomicAdd(&hdr->y_sum, ys);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace mm
}  // namespace dmel
