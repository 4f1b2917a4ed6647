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
// Roles (384 threads, one CTA per SM, persistent over tiles):
//   warps 0-7  build E then O for the tile (coalesced reads of x from L2 into registers, window in registers, swizzled
//              STS), hand each half to the tensor core as soon as it is complete; lane 0 of warp 0 issues the MMAs
//   warps 8-11 epilogue: one thread per frame reads its row of D_cos / D_sin from TMEM 16 columns at a time, forms
//              P[k] = Re^2 + Im^2, walks the triangular mel bank with two running accumulators (the bank's bins feed
//              consecutive mels m0(k), m0(k)+1 — dmel.h bin-major form) and stores log(mel + eps): consecutive
//              threads hold consecutive frames, so the stores along T are coalesced
// TMEM holds two accumulator stages (2 x 256 columns): the epilogue of tile i overlaps the build + MMA of tile i+1;
// the cos MMAs of a tile overlap the build of its O half, the sin MMAs the build of the next tile's E half.
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
constexpr int BUILD_WARPS = 8;  // operand builders; 4 more warps (one per TMEM lane quarter) run the epilogue
constexpr int THREADS = (BUILD_WARPS + 4) * 32;
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
constexpr int MISC_BANK = 528;                     // int4[68]: m0, w0, w1, taps per bin
constexpr int MISC_BAR = MISC_BANK + 68 * 16;      // 6 mbarriers + tmem pointer + flags
constexpr int MISC_BYTES = MISC_BAR + 128;
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
    int4* bank_s = reinterpret_cast<int4*>(base + OFF_MISC + MISC_BANK);
    const uint32_t bar0 = sbase + OFF_MISC + MISC_BAR;
    const uint32_t BAR_E_FREE = bar0, BAR_O_FREE = bar0 + 8, BAR_FULL0 = bar0 + 16, BAR_EMPTY0 = bar0 + 32;
    volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(base + OFF_MISC + MISC_BAR + 64);
    volatile int* tri_ok_s = reinterpret_cast<volatile int*>(base + OFF_MISC + MISC_BAR + 68);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = g.L, H = L >> 1;           // bins 0..H, folded samples j = 0..H-1
    const bool spec = g.M == 0;
    const int rows_out = spec ? H + 1 : g.M;

    // ---- one-time set-up: window, bank records, the constant B operands, barriers, tensor memory
    for (int n = tid; n < 132; n += THREADS) win_s[n] = n < L ? __ldg(win_g + n) : 0.f;
    if (!spec) {
        for (int k = tid; k < 68; k += THREADS) {
            int4 r = make_int4(0, 0, 0, 0);
            if (k <= H) {
                const int o0 = __ldg(mel.bin_off + k), cnt = __ldg(mel.bin_off + k + 1) - o0;
                r.x = cnt > 0 ? __ldg(mel.bin_m0 + k) : 0;
                r.y = __float_as_int(cnt > 0 ? __ldg(mel.bin_w + o0) : 0.f);
                r.z = __float_as_int(cnt > 1 ? __ldg(mel.bin_w + o0 + 1) : 0.f);
                r.w = cnt;
            }
            bank_s[k] = r;
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
    for (int e = tid; e < (NCOS + NSIN) * KF; e += THREADS) {
        const bool is_cos = e < NCOS * KF;
        const int ee = is_cos ? e : e - NCOS * KF;
        const int k = ee / KF, j = ee - k * KF;      // output bin (row of B), folded sample (column)
        float v = 0.f;
        if (j < H && k <= H) {
            const float2 w = tw_s[(j * k) % L];
            v = is_cos ? w.x : w.y;
        }
        float hi, lo;
        split_tf32(v, hi, lo);
        const int off = (j >> 5) * (is_cos ? BC_SLAB : BS_SLAB) + sw_off(k, j & 31);
        *reinterpret_cast<float*>(base + (is_cos ? OFF_BC_HI : OFF_BS_HI) + off) = hi;
        *reinterpret_cast<float*>(base + (is_cos ? OFF_BC_LO : OFF_BS_LO) + off) = lo;
    }
    if (tid == 0) {
        mbar_init(BAR_E_FREE, 1);
        mbar_init(BAR_O_FREE, 1);
        mbar_init(BAR_FULL0, 1);
        mbar_init(BAR_FULL0 + 8, 1);
        mbar_init(BAR_EMPTY0, 128);
        mbar_init(BAR_EMPTY0 + 8, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && !spec) {
        // the streaming mel walk of the epilogue needs every bin to feed at most two consecutive mels, first mel
        // non-decreasing in the bin index (triangular banks); anything else takes the generic gather
        int ok = 1, prev = 0;
        for (int k = 0; k <= H; ++k) {
            const int4 r = bank_s[k];
            if (r.w > 2) ok = 0;
            if (r.w > 0) {
                if (r.x < prev) ok = 0;
                prev = r.x;
            }
        }
        *tri_ok_s = ok;
    }
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
    const bool tri_ok = spec || *tri_ok_s != 0;

    const long long total_rows = (long long)g.B * g.T;
    const int n_tiles = int((total_rows + ROWS - 1) / ROWS);

    if (warp < BUILD_WARPS) {
        // =========================================================== builders (+ the MMA issuer: warp 0, lane 0)
        // A warp owns RPW consecutive rows of the tile; a lane owns the folded samples j0 = lane, j1 = lane + 32 and their
        // partners L - j (none for j = 0: its partner weight is zero).  The four samples of every row are loaded into
        // registers first — RPW * 4 independent loads in flight per thread, the latency of L2 is paid once per tile —
        // and both halves (E, then O) are formed from those registers.
        constexpr int RPW = ROWS / BUILD_WARPS;
        const int j0 = lane, j1 = lane + 32;
        const float w0a = j0 < H ? win_s[j0] : 0.f, w0b = (j0 < H && j0 > 0) ? win_s[L - j0] : 0.f;
        const float w1a = j1 < H ? win_s[j1] : 0.f, w1b = j1 < H ? win_s[L - j1] : 0.f;
        // sample offsets inside a frame (clamped for lanes beyond a short transform: their weights are zero)
        const int q0a = j0, q0b = L - (j0 > 0 ? j0 : L), q1a = j1 < L ? j1 : 0, q1b = j1 < L ? L - j1 : 0;
        const bool issuer = tid == 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int stage = it & 1, use = it >> 1;
            float ra[RPW], rb[RPW], rc[RPW], rd[RPW], mus[RPW];
            {
                // first row of this warp: (clip, frame); the following rows advance frame by frame
                long long f = (long long)tile * ROWS + warp * RPW;
                int b = int(f / g.T), t = int(f - (long long)b * g.T);
                float mu = (f < total_rows) ? __ldg(means + b) : 0.f;
#pragma unroll
                for (int rr = 0; rr < RPW; ++rr) {
                    float a = mu, bb = mu, c = mu, dd = mu;     // out-of-range samples: (mu - mu) * w = 0 = the zero padding
                    if (f < total_rows) {
                        const float* xb = x + (long long)b * g.x_stride;
                        const int p0 = t * g.hop - H;          // frame start (center=True)
                        const int pa = p0 + q0a, pb = p0 + q0b, pc = p0 + q1a, pd = p0 + q1b;
                        if (pa >= 0 && pa < g.N) a = __ldg(xb + pa);
                        if (pb >= 0 && pb < g.N) bb = __ldg(xb + pb);
                        if (pc >= 0 && pc < g.N) c = __ldg(xb + pc);
                        if (pd >= 0 && pd < g.N) dd = __ldg(xb + pd);
                    }
                    ra[rr] = a; rb[rr] = bb; rc[rr] = c; rd[rr] = dd; mus[rr] = mu;
                    ++f;
                    if (++t == g.T) {
                        t = 0;
                        ++b;
                        mu = (f < total_rows) ? __ldg(means + b) : 0.f;
                    }
                }
            }
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {     // 0: E (cosine GEMM), 1: O (sine GEMM)
                // the tensor core must be done reading this half of the previous tile
                if (it > 0) mbar_wait(pass == 0 ? BAR_E_FREE : BAR_O_FREE, uint32_t(it - 1) & 1u);
                unsigned char* a_hi = base + (pass == 0 ? OFF_E_HI : OFF_O_HI);
                unsigned char* a_lo = base + (pass == 0 ? OFF_E_LO : OFF_O_LO);
                const float sgn = pass == 0 ? 1.f : -1.f;
#pragma unroll
                for (int rr = 0; rr < RPW; ++rr) {
                    const int r = warp * RPW + rr;
                    const float mu = mus[rr];
                    // xw[j] +- xw[L-j]   (time_frequency.py:21-30 window, models.py:38 mean removal)
                    const float s0 = fmaf(sgn * (rb[rr] - mu), w0b, (ra[rr] - mu) * w0a);
                    const float s1 = fmaf(sgn * (rd[rr] - mu), w1b, (rc[rr] - mu) * w1a);
                    float h0, l0, h1, l1;
                    split_tf32(s0, h0, l0);
                    split_tf32(s1, h1, l1);
                    const int off = sw_off(r, lane);
                    *reinterpret_cast<float*>(a_hi + off) = h0;
                    *reinterpret_cast<float*>(a_lo + off) = l0;
                    *reinterpret_cast<float*>(a_hi + A_SLAB + off) = h1;
                    *reinterpret_cast<float*>(a_lo + A_SLAB + off) = l1;
                }
                fence_async_smem();                                  // generic-proxy writes -> visible to the tensor core
                asm volatile("bar.sync 1, %0;" ::"n"(BUILD_WARPS * 32) : "memory");      // the builder warps
                if (issuer) {
                    if (pass == 0 && use > 0) mbar_wait(BAR_EMPTY0 + 8 * stage, uint32_t(use - 1) & 1u);   // epilogue drained it
                    tc_fence_after();
                    const uint32_t d = tmem_base + uint32_t(stage * 256 + (pass == 0 ? 0 : 128));
                    const uint32_t idesc = pass == 0 ? instr_desc(NCOS) : instr_desc(NSIN);
                    const uint32_t a_h = sbase + (pass == 0 ? OFF_E_HI : OFF_O_HI), a_l = sbase + (pass == 0 ? OFF_E_LO : OFF_O_LO);
                    const uint32_t b_h = sbase + (pass == 0 ? OFF_BC_HI : OFF_BS_HI), b_l = sbase + (pass == 0 ? OFF_BC_LO : OFF_BS_LO);
                    const uint32_t b_slab = pass == 0 ? BC_SLAB : BS_SLAB;
                    uint32_t acc = 0;
#pragma unroll
                    for (int term = 0; term < 3; ++term) {           // hi*hi, lo*hi, hi*lo
                        const uint32_t a0 = term == 1 ? a_l : a_h, b0 = term == 2 ? b_l : b_h;
#pragma unroll
                        for (int ks = 0; ks < KF / 8; ++ks) {        // 8 tf32 per instruction, 4 instructions per slab
                            const uint32_t ao = a0 + (ks >> 2) * A_SLAB + (ks & 3) * 32;
                            const uint32_t bo = b0 + (ks >> 2) * b_slab + (ks & 3) * 32;
                            umma_tf32(d, smem_desc(ao), smem_desc(bo), idesc, acc);
                            acc = 1;
                        }
                    }
                    tc_commit(pass == 0 ? BAR_E_FREE : BAR_O_FREE);   // arrives when these MMAs have read their operands
                    if (pass == 1) tc_commit(BAR_FULL0 + 8 * stage);  // both halves of the tile are in tensor memory
                }
            }
        }
    } else {
        // =========================================================== epilogue: thread = frame row
        const int q = warp - BUILD_WARPS, r = q * 32 + lane;
        const bool do_log = g.log_eps >= 0.f;
        const float eps = g.log_eps;
        float ysum = 0.f;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int stage = it & 1, use = it >> 1;
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
            auto emit = [&](int, float v) {
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
            int cur = 0;                    // mel whose sum is in acc0 (acc1: the next one)
            float acc0 = 0.f, acc1 = 0.f;
            float pgen[NCOS];               // generic banks only: all powers of the frame (dynamically indexed: local memory)
#pragma unroll 1
            for (int c = 0; c < NCOS / 16; ++c) {
                float re[16], im[16] = {};
                tmem_ld16(trow + 16 * c, re);
                if (c < NSIN / 16) tmem_ld16(trow + 128 + 16 * c, im);
                if (c == NCOS / 16 - 1) {
                    // last read of this stage: hand it back to the tensor core
                    tc_fence_before();
                    mbar_arrive(BAR_EMPTY0 + 8 * stage);
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int k = 16 * c + i;
                    if (k <= H) {
                        const float rr_ = re[i] + ((k & 1) ? -nyq : nyq);
                        const float ii_ = (c < NSIN / 16 && k < H) ? im[i] : 0.f;
                        const float p = fmaf(rr_, rr_, ii_ * ii_);            // time_frequency.py:53  |stft|^2
                        if (spec) {
                            emit(k, p);
                        } else if (tri_ok) {
                            const int4 rec = bank_s[k];
                            if (rec.w > 0) {
                                while (cur < rec.x) { emit(cur, acc0); acc0 = acc1; acc1 = 0.f; ++cur; }
                                acc0 = fmaf(__int_as_float(rec.y), p, acc0);
                                acc1 = fmaf(__int_as_float(rec.z), p, acc1);
                            }
                        } else {
                            pgen[k] = p;
                        }
                    }
                }
            }
            if (!spec) {
                if (tri_ok) {
                    while (cur < g.M) { emit(cur, acc0); acc0 = acc1; acc1 = 0.f; ++cur; }
                } else {
                    // models.py:53 matmul(spec^T, mel_fb), mel by mel over the mel-major bank (runs are zero-padded to 4
                    // taps, dmel.h: the padded taps may point past the last bin and are skipped)
                    for (int m = 0; m < g.M; ++m) {
                        const int k0 = __ldg(mel.mel_k0 + m), o0 = __ldg(mel.mel_off + m), o1 = __ldg(mel.mel_off + m + 1);
                        const int cnt = min(o1 - o0, H + 1 - k0);
                        float s = 0.f;
                        for (int o = 0; o < cnt; ++o) s = fmaf(__ldg(mel.mel_w + o0 + o), pgen[k0 + o], s);
                        emit(m, s);
                    }
                }
            }
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
