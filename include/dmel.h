/* dmel.h — C ABI of the B200-native DMEL hot path (libdmel_b200.so).
 *
 * The reference (johnmartinsson/differentiable-mel-spectrogram) has no FFI layer: its
 * boundary for this path is the torch.nn.Module contract of
 *     models.py:14-56    MelSpectrogramLayer.__init__/forward
 *     models.py:171-200  SpectrogramLayer.__init__/forward
 * which call time_frequency.py:21-58 (Gaussian window + torch.stft + |.|^2),
 * torchaudio.functional.melscale_fbanks and torch.matmul, followed by
 * torch.log(s + 1e-10) in the classifier heads (models.py:72-73, 96-97, 125-126,
 * 156-157) and by autograd's backward of that whole chain (train.py:47).
 * The entry points below are what a Python (ctypes), C++ or any other host binds to
 * replace exactly that chain; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - every pointer except the dmel_melbank struct itself is a DEVICE pointer
 *   - plain pointers and sizes only; no torch types
 *   - stream-ordered: nothing synchronises the device, nothing is allocated, no
 *     global mutable state (the caller owns x, y, grads and the workspace)
 *   - return value: DMEL_OK (0) or a negative dmel_status; never throws / exits
 *   - fp32 arithmetic throughout ("f32" suffix); dL/dlambda is accumulated in fp64
 */
#ifndef DMEL_H_
#define DMEL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMEL_ABI_VERSION 4

#if defined(__GNUC__)
#define DMEL_API __attribute__((visibility("default")))
#else
#define DMEL_API
#endif

#define DMEL_FLAG_NORMALIZE_WINDOW 1
#define DMEL_FLAG_KEEP_MEAN 2
#define DMEL_FLAG_TIME_MAJOR 4   /* y, grad_y, dy_dlambda are [batch,1,T,rows] (rows fastest) instead of [batch,1,rows,T] */
/* Input sample type (ABI 3).  `x` is declared const void*: fp32 by default; with one of these flags it points to
 * 16-bit samples that the kernels widen while they stage the signal — float(int16) exactly as torch's
 * x.to(torch.float32) (no 1/32768 scaling), or IEEE half -> float.  This is the PCM input path of the reference's
 * data loaders (datasets.py:169,208,255 -> train.py:34) without the intermediate fp32 tensor and with half the
 * host->device bytes.  Tiled FFT path (n_fft 256..4096) only: dmel_x16_supported().                               */
#define DMEL_FLAG_X_I16 8
#define DMEL_FLAG_X_F16 16
#define DMEL_X_F32 0
#define DMEL_X_I16 1
#define DMEL_X_F16 2

typedef enum dmel_status {
    DMEL_OK = 0,
    DMEL_ERR_NULL_POINTER = -1,
    DMEL_ERR_BAD_GEOMETRY = -2,      /* B, N, hop, n_fft, win_length, n_mels inconsistent   */
    DMEL_ERR_UNSUPPORTED_NFFT = -3,  /* n_fft has no kernel (see dmel_nfft_supported)       */
    DMEL_ERR_WORKSPACE_TOO_SMALL = -4,
    DMEL_ERR_LAUNCH_FAILED = -5,     /* cudaGetLastError() after a launch was not success   */
    DMEL_ERR_SMEM_EXCEEDED = -6,     /* no tile size fits the 227 KB shared memory of an SM */
    DMEL_ERR_BAD_MELBANK = -7
} dmel_status;

/* Sparse forms of the (n_freqs x n_mels) filterbank that replaces the dense
 * torch.matmul(spec.T, mel_fb) of models.py:53.  Built once per n_freqs by the host
 * from the dense matrix (so the weights are bit-identical to the dense ones):
 *   mel-major (forward gather):  mel m reads bins  mel_k0[m] .. mel_k0[m]+cnt-1,
 *       cnt = mel_off[m+1]-mel_off[m], weights mel_w[mel_off[m] + i]; every run is
 *       zero-padded to cnt % 4 == 0 (so mel_off[m] % 4 == 0) and mel_k0[m]+cnt <= F+3
 *   bin-major (backward gather): bin k reads mels  bin_m0[k] .. bin_m0[k]+cnt-1,
 *       cnt = bin_off[k+1]-bin_off[k], weights bin_w[bin_off[k] + i]
 * All six arrays are device pointers, each 16-byte aligned and readable up to the next
 * multiple of 4 elements (they are staged into shared memory with 16-byte async copies);
 * the struct itself lives on the host.
 *
 * out_scale / out_shift (both NULL, or both device arrays of n_mels floats) fold a per-mel
 * affine map into the epilogue:  y[b,0,m,t] <- out_scale[m] * y[b,0,m,t] + out_shift[m],
 * applied after the optional log.  This is the eval-mode BatchNorm2d over mel bins that the
 * PANNs head applies first (panns.py:169-172): scale = gamma/sqrt(var+eps), shift = beta -
 * mean*scale.  The backward takes grad_y with respect to the mapped output and, when the log
 * is fused, recovers log-mel as (y - shift)/scale (a mel whose scale is 0 gets no gradient).
 * The map is a constant of the call: no gradient is produced for it.                  */
typedef struct dmel_melbank {
    int32_t n_freqs;          /* F = n_fft/2 + 1                                        */
    int32_t n_mels;           /* M                                                      */
    int32_t mel_nnz;          /* = mel_off[M]  (entries of mel_w)                       */
    int32_t bin_nnz;          /* = bin_off[F]  (entries of bin_w)                       */
    const int32_t* mel_k0;    /* [M]                                                    */
    const int32_t* mel_off;   /* [M+1]                                                  */
    const float* mel_w;       /* [mel_off[M]]                                           */
    const int32_t* bin_m0;    /* [F]                                                    */
    const int32_t* bin_off;   /* [F+1]                                                  */
    const float* bin_w;       /* [bin_off[F]]                                           */
    const float* out_scale;   /* [M] or NULL                                            */
    const float* out_shift;   /* [M] or NULL                                            */
    const void* seg_tables;   /* device memory filled by dmel_build_segment_tables for this bank and
                                 n_fft = 2*(F-1), or NULL (ABI 3): the forward then projects segment by
                                 segment (conflict-free shared-memory layout) instead of mel by mel    */
} dmel_melbank;

/* Optional, once per (bank, n_fft): derive from the bin-major form the lane-owned "segment" tables the tiled forward
 * (n_fft 256..4096) uses for the mel projection.  `tables` is caller-owned device memory of at least
 * dmel_segment_table_bytes(n_fft) bytes, 16-byte aligned; store its address in bank->seg_tables afterwards.
 * Banks that are not triangular (more than two mels per bin, runs that do not move up with frequency) or whose tables
 * would not fit next to the transform buffer give *usable = 0: leave seg_tables NULL then (the forward keeps the
 * mel-major gather — results are the same either way).  A set-up call: unlike every other entry point it waits for
 * its stream.                                                                                                     */
DMEL_API size_t dmel_segment_table_bytes(int32_t n_fft);
DMEL_API int dmel_build_segment_tables(const dmel_melbank* bank, int32_t n_fft, void* tables, size_t bytes,
                              void* cuda_stream, int32_t* usable);

DMEL_API int dmel_abi_version(void);
DMEL_API const char* dmel_status_string(int status);

/* 1 if a kernel exists for this transform length (radix FFT path for powers of two
 * 128..16384, direct-DFT path for every other even length up to DMEL_MAX_DFT_NFFT), else 0.  */
DMEL_API int dmel_nfft_supported(int32_t n_fft);

/* time_frequency.py:60-65 next_power_of_2(int(6*|lambda|)) evaluated the way the
 * reference does (fp32 product, int() truncation, 0 -> 2, 1 -> 1).                  */
DMEL_API int32_t dmel_window_length_from_lambda(float lambda);

/* Bytes of device scratch a forward/backward pair needs (window, twiddles, per-clip
 * means, per-clip dx sums, fp64 dL/dlambda accumulator).                            */
DMEL_API size_t dmel_workspace_bytes(int64_t batch, int32_t n_fft);

/* Byte offset inside the workspace of one fp64 that dmel_forward_f32 leaves behind: the sum of all
 * elements of y (log-mel when the log is fused).  train.py:58 copies every spectrogram batch to the
 * host just to add it up (`running_energy`); reading these 8 bytes replaces that copy.            */
DMEL_API size_t dmel_workspace_y_sum_offset(void);

/* Forward:  y[b,0,m,t] of models.py:33-56 (n_mels > 0) or models.py:182-200
 * (n_mels == 0: "spectrogram mode", rows are the n_fft/2+1 frequency bins).
 *   x           [batch, n_points] fp32 (or 16-bit, see DMEL_FLAG_X_*), row stride x_row_stride elements
 *   lambda      the layer's parameter `lambd` (signed; |.| is taken as models.py:38)
 *   win_length  Gaussian window length: next_power_of_2(6|lambda|) when optimized,
 *               n_points otherwise (time_frequency.py:34-43)
 *   n_fft       win_length (optimized) or 2*win_length (time_frequency.py:48/51)
 *   hop         hop_length; frames T = n_points/hop + 1 (models.py:30), center=True,
 *               pad_mode='constant'
 *   bank        NULL iff n_mels == 0
 *   log_eps     >= 0: y = log(mel + log_eps) fused (models.py:72-73); < 0: linear
 *   flags       DMEL_FLAG_NORMALIZE_WINDOW: time_frequency.py:25-30 `norm`;
 *               DMEL_FLAG_KEEP_MEAN: skip the per-clip mean removal of models.py:38 (what a
 *               direct caller of time_frequency.differentiable_spectrogram gets);
 *               DMEL_FLAG_TIME_MAJOR: emit [batch,1,T,rows] — the layout MelPANNsNet transposes to
 *               before Cnn6 (models.py:162) — see dmel_time_major_supported
 *   y           [batch, 1, rows, T] fp32 contiguous, rows = n_mels or n_fft/2+1
 *   workspace   >= dmel_workspace_bytes(batch, n_fft); filled here (means, window,
 *               twiddles) and reusable by dmel_backward_f32 with prepared = 1      */
DMEL_API int dmel_forward_f32(const void* x, int64_t batch, int64_t n_points, int64_t x_row_stride,
                     float lambda, int32_t win_length, int32_t n_fft, int32_t hop,
                     int32_t n_mels, const dmel_melbank* bank, float log_eps,
                     int32_t flags, float* y,
                     void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Several window widths over ONE batch in one call: the fixed-lambda grid of the reference's search spaces
 * (search_spaces.py:22,29 — `trainable=False` x {10, 35, 300 ms}), where every grid point is the same front-end with a
 * different lambda applied to the same clips.  Output k is what dmel_forward_f32 returns for lambdas[k] (bit-identical);
 * what is shared is everything that does not depend on lambda: argument validation, the upload of x and the per-clip
 * means of models.py:38, which are summed once (first entry) and copied for the others instead of re-reading x.
 * Per entry k: lambdas[k], win_lengths[k], n_ffts[k], banks[k] (NULL array iff n_mels == 0), ys[k], and its own
 * workspace (>= dmel_workspace_bytes(batch, n_ffts[k])), which a later dmel_backward_f32(prepared = 1) for that
 * lambda may reuse.  Entries may mix transform lengths (and so the tiled, big-FFT and direct-DFT paths).             */
#define DMEL_MAX_MULTI_LAMBDA 64
DMEL_API int dmel_forward_multi_f32(const void* x, int64_t batch, int64_t n_points, int64_t x_row_stride,
                                    int32_t n_lambda, const float* lambdas, const int32_t* win_lengths,
                                    const int32_t* n_ffts, int32_t hop, int32_t n_mels,
                                    const dmel_melbank* const* banks, float log_eps, int32_t flags,
                                    float* const* ys, void* const* workspaces, const size_t* workspace_bytes,
                                    void* cuda_stream);

/* Forward-mode variant for the reference's training case (x is data, only `lambd` trains: models.py:19,
 * main.py:27).  Same arguments and the same y as dmel_forward_f32, plus dy_dlambda [batch,1,rows,T]:
 * d y / d|lambda| from the derivative-window STFT, 2 Re(conj(X) . STFT(x * dw/dlambda)), mel-projected and
 * chain-ruled through the fused log.  Costs one complex FFT per frame instead of a half, and makes the
 * backward a dot product (dmel_lambda_grad_f32): 1.0 instead of 1.5 FFTs per frame per training step, no
 * recomputation.  Tiled FFT path only (dmel_jvp_supported: n_fft 128..4096).                              */
DMEL_API int dmel_jvp_supported(int32_t n_fft);
/* 1 if DMEL_FLAG_TIME_MAJOR is served for this transform length: the tiled FFT kernels (128..4096, both modes) and the
 * big-FFT path (8192 / 16384) with a mel bank (n_mels > 0) */
DMEL_API int dmel_time_major_supported(int32_t n_fft, int32_t n_mels);
/* 1 if 16-bit inputs (DMEL_FLAG_X_I16 / DMEL_FLAG_X_F16) are accepted for this transform length */
DMEL_API int dmel_x16_supported(int32_t n_fft);
DMEL_API int dmel_forward_jvp_f32(const void* x, int64_t batch, int64_t n_points, int64_t x_row_stride,
                         float lambda, int32_t win_length, int32_t n_fft, int32_t hop,
                         int32_t n_mels, const dmel_melbank* bank, float log_eps,
                         int32_t flags, float* y, float* dy_dlambda,
                         void* workspace, size_t workspace_bytes, void* cuda_stream);

/* dL/dlambda = sign(lambda) * sum_i grad_y[i] * dy_dlambda[i]  (count = batch*rows*T), one fp32 out;
 * workspace = the one dmel_forward_jvp_f32 filled (its header holds the self-resetting fp64 accumulator). */
DMEL_API int dmel_lambda_grad_f32(const float* grad_y, const float* dy_dlambda, int64_t count, float lambda,
                         float* grad_lambda, void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Backward of the same chain (what autograd does for train.py:47).
 *   y            forward output; read only when log_eps >= 0 (may be NULL otherwise)
 *   grad_y       dL/dy, same shape as y
 *   grad_x       [batch, n_points] contiguous, or NULL: skip dL/dx entirely (the
 *                reference's training case — x is data, only `lambd` is a parameter)
 *   grad_lambda  one fp32: dL/dlambda (includes sign(lambda) from torch.abs)
 *   prepared     1: workspace still holds what dmel_forward_f32 wrote for the same
 *                (x, lambda, geometry); 0: recompute means/window/twiddles first    */
DMEL_API int dmel_backward_f32(const void* x, int64_t batch, int64_t n_points, int64_t x_row_stride,
                      float lambda, int32_t win_length, int32_t n_fft, int32_t hop,
                      int32_t n_mels, const dmel_melbank* bank, float log_eps,
                      int32_t flags, const float* y, const float* grad_y,
                      float* grad_x, float* grad_lambda,
                      void* workspace, size_t workspace_bytes, int32_t prepared,
                      void* cuda_stream);

/* The path's only collective — sum of the scalar dL/dlambda over the GPUs of one box — as a one-shot exchange
 * over NVLink peer memory instead of an NCCL call.  Each rank allocates dmel_peer_buffer_bytes(world) of zeroed
 * device memory, maps every peer's buffer into its own address space (CUDA IPC) and passes the `world` base
 * pointers (its own at index `rank`) as a DEVICE array of 64-bit addresses.  `value` (device, 1 x fp32) is replaced
 * by the sum, bit-identical on all ranks (summed in rank order).  Stream-ordered, graph-capturable; every rank must
 * make the same sequence of calls.  Like any collective it waits for its peers; the wait is bounded by
 * DMEL_PEER_TIMEOUT_S (environment, default 1800 s): on expiry the result is NaN and word [4*world+1] of the
 * rank's buffer is set to 1.  dist.py (PeerRing) is the host side; NCCL remains the fallback.                     */
DMEL_API size_t dmel_peer_buffer_bytes(int32_t world);
/* The symmetric buffers are whole cudaMalloc allocations (that is what a CUDA IPC handle names), owned by the
 * caller through these four calls: alloc (zeroed, returns the 64-byte IPC handle to ship to the peers), open (map a
 * peer's handle into the current device's context, peer access enabled), close, free.                            */
DMEL_API int dmel_peer_alloc(int32_t world, void** buffer, unsigned char* ipc_handle_64);
DMEL_API int dmel_peer_open(const unsigned char* ipc_handle_64, void** mapped);
DMEL_API int dmel_peer_close(void* mapped);
DMEL_API int dmel_peer_free(void* buffer);
/* synchronous read of the rank's own error word (self-tests only) */
DMEL_API int dmel_peer_timed_out(const void* buffer, int32_t world, int32_t* timed_out);
DMEL_API int dmel_peer_allreduce_f32(float* value, const void* peer_bases_dev, int32_t rank, int32_t world,
                            void* cuda_stream);

/* dmel_backward_f32 with the collective fused in: the finalize kernel that converts the fp64 dL/dlambda sum also
 * exchanges it with the peers (same protocol as dmel_peer_allreduce_f32, same buffers) while its other blocks apply
 * the mean-removal adjoint to grad_x; grad_lambda receives the sum over all `world` GPUs.  No separate launch.       */
DMEL_API int dmel_backward_allreduce_f32(const void* x, int64_t batch, int64_t n_points, int64_t x_row_stride,
                                float lambda, int32_t win_length, int32_t n_fft, int32_t hop,
                                int32_t n_mels, const dmel_melbank* bank, float log_eps,
                                int32_t flags, const float* y, const float* grad_y,
                                float* grad_x, float* grad_lambda,
                                void* workspace, size_t workspace_bytes, int32_t prepared,
                                void* cuda_stream, const void* peer_bases_dev, int32_t rank, int32_t world);

/* Introspection for tests / benchmarks: tile size (frames per CTA), FFT groups per CTA,
 * dynamic shared memory and threads per CTA the launcher would pick.  rows = n_mels or
 * n_fft/2+1; bank_nnz = entries of the sparse bank form that pass stages (0 in spectrogram
 * mode); which: 0 fwd, 1 bwd (dx+dlambda), 2 bwd (dlambda only).                      */
DMEL_API int dmel_query_launch(int64_t batch, int64_t n_points, int32_t n_fft, int32_t hop, int32_t rows,
                      int32_t bank_nnz, int32_t which, int32_t* tile_frames, int32_t* groups,
                      int32_t* smem_bytes, int32_t* threads, int32_t* path);

/* Instrumentation (bench.py): cumulative number of kernels this library launched, by kind
 * (0 prep, 1 forward, 2 backward, 3 finalize, 4 zero-init, <0 all); and opt-in per-launch CUDA-event timing of
 * the calling thread's launches: enable(1) clears and starts, read() synchronises on the recorded
 * events and returns their summed duration and count, enable(0) stops and frees the events.      */
DMEL_API long long dmel_launch_count(int32_t kind);
DMEL_API void dmel_profile_enable(int32_t on);
DMEL_API int dmel_profile_read(int32_t kind, double* total_ms, long long* launches);

#define DMEL_MAX_DFT_NFFT 32768

#ifdef __cplusplus
}
#endif
#endif /* DMEL_H_ */
