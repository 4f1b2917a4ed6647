// abi_run.cu — a host with no torch / Python in the process that drives libdmel_b200.so through the C ABI of
// include/dmel.h: cudaMalloc, dmel_forward_f32, dmel_backward_f32, compare with a fixture.  Two modes:
//
//   abi_run check <case.bin>            run the case file's forward + backward and compare y, dL/dx, dL/dlambda with
//                                       the expected arrays stored in it (made by scripts/dump_case.py from the
//                                       oracle, i.e. the CPU restatement of models.py:33-56 / time_frequency.py:21-58).
//                                       Tolerances are north_star's: log-mel rel 1e-4 / abs 1e-5, gradients rel 1e-3.
//   abi_run time <case.bin> B [iters] [mode]
//                                       replicate the case's geometry at batch B on random clips and time prep+forward,
//                                       backward (dx + dlambda) and the lambda-only pair with CUDA events (tuning
//                                       loops; the numbers judged are bench.py's).  mode: "all" (default), "fwd", "bwd".
//
// Case file (little endian): "DMELCASE" | int32 version=1 | int32 B N hop win_length n_fft n_mels flags F mel_nnz
// bin_nnz has_expected | float lambda log_eps | int32 mel_k0[M] mel_off[M+1] | float mel_w[mel_nnz] | int32 bin_m0[F]
// bin_off[F+1] | float bin_w[bin_nnz] | (has_expected) float x[B*N] G[B*rows*T] y[B*rows*T] dx[B*N] dlam[1]
//
// Build: nvcc -O2 -o tests/c/abi_run tests/c/abi_run.cu -Iinclude -Ldifferentiable-mel-spectrogram_b200 -ldmel_b200 \
//        -Xlinker -rpath -Xlinker '$ORIGIN/../../differentiable-mel-spectrogram_b200'
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <vector>

#include "dmel.h"

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            fprintf(stderr, "%s:%d: %s: %s\n", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            exit(2);                                                                               \
        }                                                                                          \
    } while (0)
#define DK(call)                                                                                     \
    do {                                                                                             \
        int rc_ = (call);                                                                            \
        if (rc_ != DMEL_OK) {                                                                        \
            fprintf(stderr, "%s:%d: %s -> %s\n", __FILE__, __LINE__, #call, dmel_status_string(rc_)); \
            exit(3);                                                                                 \
        }                                                                                            \
    } while (0)

struct Case {
    int32_t B, N, hop, win, n_fft, n_mels, flags, F, mel_nnz, bin_nnz, has_expected;
    float lambda, log_eps;
    std::vector<int32_t> mel_k0, mel_off, bin_m0, bin_off;
    std::vector<float> mel_w, bin_w, x, G, y, dx;
    float dlam = 0.f;
    int rows() const { return n_mels > 0 ? n_mels : F; }
    int T() const { return N / hop + 1; }
};

template <typename T> static void rd(FILE* f, std::vector<T>& v, size_t n) {
    v.resize(n);
    if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short case file\n"); exit(2); }
}

static Case load_case(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    char magic[8];
    int32_t ver = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "DMELCASE", 8) || fread(&ver, 4, 1, f) != 1 || ver != 1) {
        fprintf(stderr, "%s: not a DMELCASE v1 file\n", path);
        exit(2);
    }
    Case c;
    int32_t h[11];
    float fl[2];
    if (fread(h, 4, 11, f) != 11 || fread(fl, 4, 2, f) != 2) { fprintf(stderr, "short header\n"); exit(2); }
    c.B = h[0]; c.N = h[1]; c.hop = h[2]; c.win = h[3]; c.n_fft = h[4]; c.n_mels = h[5]; c.flags = h[6]; c.F = h[7];
    c.mel_nnz = h[8]; c.bin_nnz = h[9]; c.has_expected = h[10];
    c.lambda = fl[0]; c.log_eps = fl[1];
    if (c.n_mels > 0) {
        rd(f, c.mel_k0, c.n_mels); rd(f, c.mel_off, c.n_mels + 1); rd(f, c.mel_w, c.mel_nnz);
        rd(f, c.bin_m0, c.F); rd(f, c.bin_off, c.F + 1); rd(f, c.bin_w, c.bin_nnz);
    }
    if (c.has_expected) {
        const size_t nx = size_t(c.B) * c.N, ny = size_t(c.B) * c.rows() * c.T();
        rd(f, c.x, nx); rd(f, c.G, ny); rd(f, c.y, ny); rd(f, c.dx, nx);
        if (fread(&c.dlam, 4, 1, f) != 1) { fprintf(stderr, "short case file\n"); exit(2); }
    }
    fclose(f);
    return c;
}

// device copy padded to a multiple of 4 elements (+4), 16-byte aligned by cudaMalloc (dmel.h: the bank arrays are
// staged with 16-byte copies)
template <typename T> static T* to_dev(const std::vector<T>& v) {
    const size_t n = (v.size() + 3) / 4 * 4 + 4;
    T* d = nullptr;
    CK(cudaMalloc(&d, n * sizeof(T)));
    CK(cudaMemset(d, 0, n * sizeof(T)));
    if (!v.empty()) CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

struct DevBank {
    dmel_melbank s{};
    const dmel_melbank* ptr = nullptr;
    explicit DevBank(const Case& c) {
        if (c.n_mels == 0) return;
        s.n_freqs = c.F; s.n_mels = c.n_mels; s.mel_nnz = c.mel_nnz; s.bin_nnz = c.bin_nnz;
        s.mel_k0 = to_dev(c.mel_k0); s.mel_off = to_dev(c.mel_off); s.mel_w = to_dev(c.mel_w);
        s.bin_m0 = to_dev(c.bin_m0); s.bin_off = to_dev(c.bin_off); s.bin_w = to_dev(c.bin_w);
        // optional segment tables of the tiled forward (ABI 3); DMEL_NO_SEG=1 keeps the mel-major gather
        const size_t sb = getenv("DMEL_NO_SEG") ? 0 : dmel_segment_table_bytes(c.n_fft);
        if (sb) {
            void* t = nullptr;
            CK(cudaMalloc(&t, sb));
            int32_t usable = 0;
            DK(dmel_build_segment_tables(&s, c.n_fft, t, sb, nullptr, &usable));
            if (usable) s.seg_tables = t;
        }
        ptr = &s;
    }
};

static double rel_l2(const std::vector<float>& a, const std::vector<float>& b) {
    double num = 0, den = 0;
    for (size_t i = 0; i < a.size(); ++i) { const double d = double(a[i]) - b[i]; num += d * d; den += double(b[i]) * b[i]; }
    return sqrt(num) / (sqrt(den) + 1e-300);
}

static int run_check(const Case& c) {
    if (!c.has_expected) { fprintf(stderr, "case has no expected arrays\n"); return 2; }
    DevBank bank(c);
    const size_t nx = c.x.size(), ny = c.y.size();
    float *dx_x, *d_y, *d_G, *d_dx, *d_dl;
    CK(cudaMalloc(&dx_x, nx * 4)); CK(cudaMalloc(&d_y, ny * 4)); CK(cudaMalloc(&d_G, ny * 4));
    CK(cudaMalloc(&d_dx, nx * 4)); CK(cudaMalloc(&d_dl, 4));
    CK(cudaMemcpy(dx_x, c.x.data(), nx * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_G, c.G.data(), ny * 4, cudaMemcpyHostToDevice));
    const size_t wsb = dmel_workspace_bytes(c.B, c.n_fft);
    void* ws;
    CK(cudaMalloc(&ws, wsb));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    DK(dmel_forward_f32(dx_x, c.B, c.N, c.N, c.lambda, c.win, c.n_fft, c.hop, c.n_mels, bank.ptr, c.log_eps, c.flags, d_y,
                        ws, wsb, st));
    DK(dmel_backward_f32(dx_x, c.B, c.N, c.N, c.lambda, c.win, c.n_fft, c.hop, c.n_mels, bank.ptr, c.log_eps, c.flags, d_y,
                         d_G, d_dx, d_dl, ws, wsb, 1, st));
    CK(cudaStreamSynchronize(st));
    std::vector<float> y(ny), dx(nx);
    float dl = 0.f;
    CK(cudaMemcpy(y.data(), d_y, ny * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(dx.data(), d_dx, nx * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&dl, d_dl, 4, cudaMemcpyDeviceToHost));
    // forward: |y - y0| <= atol + rtol*|y0| with (1e-5, 1e-4) on log-mel; linear outputs are compared relative to the
    // largest value (the reference's own fp32-vs-fp64 spread)
    double worst = 0;
    size_t bad = 0;
    double ymax = 0;
    for (float v : c.y) ymax = fmax(ymax, fabs(v));
    for (size_t i = 0; i < ny; ++i) {
        const double tol = c.log_eps >= 0 ? 1e-5 + 1e-4 * fabs(c.y[i]) : 1e-4 * fabs(c.y[i]) + 1e-6 * ymax;
        const double err = fabs(double(y[i]) - c.y[i]);
        worst = fmax(worst, err / tol);
        bad += err > tol;
    }
    const double edx = rel_l2(dx, c.dx);
    const double edl = fabs(double(dl) - c.dlam) / (fabs(c.dlam) + 1e-300);
    printf("check B=%d N=%d n_fft=%d hop=%d mels=%d log_eps=%g: y worst err/tol %.3f (%zu of %zu outside), dx rel-L2 %.2e, "
           "dlambda %.6e vs %.6e (rel %.2e)\n",
           c.B, c.N, c.n_fft, c.hop, c.n_mels, c.log_eps, worst, bad, ny, edx, dl, c.dlam, edl);
    const bool ok = bad == 0 && edx <= 1e-3 && edl <= 1e-3;
    printf(ok ? "PASS\n" : "FAIL\n");
    return ok ? 0 : 1;
}

static float timed(cudaStream_t st, int iters, const std::function<void()>& fn) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) fn();
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(a, st));
    for (int i = 0; i < iters; ++i) fn();
    CK(cudaEventRecord(b, st));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / iters;
}

static int run_time(const Case& c, int B, int iters, const char* mode) {
    DevBank bank(c);
    const size_t nx = size_t(B) * c.N, ny = size_t(B) * c.rows() * c.T();
    std::vector<float> hx(nx), hG(ny);
    uint32_t s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int32_t(s >> 8) & 0xffff) / 32768.0f - 1.0f; };
    for (auto& v : hx) v = rnd() + rnd() + rnd();
    for (auto& v : hG) v = rnd();
    float *d_x, *d_y, *d_dy, *d_G, *d_dx, *d_dl;
    CK(cudaMalloc(&d_x, nx * 4)); CK(cudaMalloc(&d_y, ny * 4)); CK(cudaMalloc(&d_dy, ny * 4)); CK(cudaMalloc(&d_G, ny * 4));
    CK(cudaMalloc(&d_dx, nx * 4)); CK(cudaMalloc(&d_dl, 4));
    CK(cudaMemcpy(d_x, hx.data(), nx * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_G, hG.data(), ny * 4, cudaMemcpyHostToDevice));
    const size_t wsb = dmel_workspace_bytes(B, c.n_fft);
    void* ws;
    CK(cudaMalloc(&ws, wsb));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    const bool all = !strcmp(mode, "all");
    auto fwd = [&]() {
        DK(dmel_forward_f32(d_x, B, c.N, c.N, c.lambda, c.win, c.n_fft, c.hop, c.n_mels, bank.ptr, c.log_eps, c.flags, d_y, ws,
                            wsb, st));
    };
    auto bwd = [&]() {
        DK(dmel_backward_f32(d_x, B, c.N, c.N, c.lambda, c.win, c.n_fft, c.hop, c.n_mels, bank.ptr, c.log_eps, c.flags, d_y, d_G,
                             d_dx, d_dl, ws, wsb, 1, st));
    };
    fwd();
    CK(cudaStreamSynchronize(st));
    float tf = 0, tb = 0, tj = 0, tl = 0;
    if (all || !strcmp(mode, "fwd")) tf = timed(st, iters, fwd);
    if (all || !strcmp(mode, "bwd")) tb = timed(st, iters, bwd);
    if (all && dmel_jvp_supported(c.n_fft)) {
        tj = timed(st, iters, [&]() {
            DK(dmel_forward_jvp_f32(d_x, B, c.N, c.N, c.lambda, c.win, c.n_fft, c.hop, c.n_mels, bank.ptr, c.log_eps, c.flags,
                                    d_y, d_dy, ws, wsb, st));
        });
        tl = timed(st, iters, [&]() {
            DK(dmel_backward_f32(d_x, B, c.N, c.N, c.lambda, c.win, c.n_fft, c.hop, c.n_mels, bank.ptr, c.log_eps, c.flags, d_y,
                                 d_G, nullptr, d_dl, ws, wsb, 1, st));
        });
    }
    const double audio_s = double(B) * c.N / (c.n_mels > 0 ? 1.0 : 1.0);
    (void)audio_s;
    printf("time B=%d N=%d n_fft=%d hop=%d mels=%d: prep+fwd %.4f ms, zero+bwd+finalize %.4f ms, sum %.4f ms | jvp fwd %.4f "
           "ms, lambda-only bwd %.4f ms\n",
           B, c.N, c.n_fft, c.hop, c.n_mels, tf, tb, tf + tb, tj, tl);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s check <case.bin> | time <case.bin> B [iters] [all|fwd|bwd]\n", argv[0]);
        return 2;
    }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { fprintf(stderr, "no CUDA device\n"); return 77; }
    const Case c = load_case(argv[2]);
    if (!strcmp(argv[1], "check")) return run_check(c);
    if (!strcmp(argv[1], "time")) {
        const int B = argc > 3 ? atoi(argv[3]) : c.B;
        const int iters = argc > 4 ? atoi(argv[4]) : 50;
        return run_time(c, B, iters, argc > 5 ? argv[5] : "all");
    }
    fprintf(stderr, "unknown mode %s\n", argv[1]);
    return 2;
}
