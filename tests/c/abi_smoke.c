/* Plain-C consumer of include/dmel.h: proves the header is C99-clean and that libdmel_b200.so can be linked and
 * called without Python or torch.  No device work: only version, status strings, sizing and argument validation
 * (every call below returns before touching the GPU). */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "dmel.h"

#define CHECK(cond)                                             \
    do {                                                        \
        if (!(cond)) {                                          \
            fprintf(stderr, "FAILED line %d: %s\n", __LINE__, #cond); \
            return 1;                                           \
        }                                                       \
    } while (0)

int main(void) {
    CHECK(dmel_abi_version() == DMEL_ABI_VERSION);
    CHECK(strcmp(dmel_status_string(DMEL_OK), "ok") == 0);
    CHECK(strlen(dmel_status_string(DMEL_ERR_UNSUPPORTED_NFFT)) > 0);
    /* time_frequency.py:39, 60-65: next_power_of_2(int(6*|lambda|)) */
    CHECK(dmel_window_length_from_lambda(46.667f) == 512);
    CHECK(dmel_window_length_from_lambda(-257.25f) == 2048);
    CHECK(dmel_nfft_supported(512) == 1 && dmel_nfft_supported(16384) == 1 && dmel_nfft_supported(255) == 0);
    CHECK(dmel_workspace_bytes(256, 512) >= 64 + 512 * 12 + 256 * 12);
    CHECK(dmel_workspace_y_sum_offset() < 64);

    float fake;                     /* never dereferenced: validation fails first */
    dmel_melbank bank;
    memset(&bank, 0, sizeof bank);
    CHECK(dmel_forward_f32(NULL, 1, 100, 100, 10.f, 64, 64, 10, 0, NULL, -1.f, 0, &fake, &fake, 1 << 20, NULL)
          == DMEL_ERR_NULL_POINTER);
    CHECK(dmel_forward_f32(&fake, 1, 100, 100, 10.f, 128, 64, 10, 0, NULL, -1.f, 0, &fake, &fake, 1 << 20, NULL)
          == DMEL_ERR_BAD_GEOMETRY);
    CHECK(dmel_forward_f32(&fake, 1, 100, 100, 10.f, 64, 64, 10, 0, NULL, -1.f, 0, &fake, &fake, 16, NULL)
          == DMEL_ERR_WORKSPACE_TOO_SMALL);
    bank.n_freqs = 33;
    bank.n_mels = 5;                /* != n_mels below */
    CHECK(dmel_forward_f32(&fake, 1, 100, 100, 10.f, 64, 64, 10, 4, &bank, -1.f, 0, &fake, &fake, 1 << 20, NULL)
          == DMEL_ERR_BAD_MELBANK);

    /* the fixed-lambda grid entry point validates its arrays before any device work, then every entry like a forward */
    {
        const float lams[2] = {13.33f, 46.667f};
        const int32_t wins[2] = {128, 512}, nffts[2] = {128, 512};
        float* ys[2] = {&fake, &fake};
        void* wss[2] = {&fake, &fake};
        const size_t small[2] = {16, 16}, big[2] = {1 << 20, 1 << 20};
        CHECK(dmel_forward_multi_f32(&fake, 1, 100, 100, 2, NULL, wins, nffts, 10, 0, NULL, -1.f, 0, ys, wss, big, NULL)
              == DMEL_ERR_NULL_POINTER);
        CHECK(dmel_forward_multi_f32(&fake, 1, 100, 100, 0, lams, wins, nffts, 10, 0, NULL, -1.f, 0, ys, wss, big, NULL)
              == DMEL_ERR_BAD_GEOMETRY);
        CHECK(dmel_forward_multi_f32(&fake, 1, 100, 100, DMEL_MAX_MULTI_LAMBDA + 1, lams, wins, nffts, 10, 0, NULL, -1.f, 0, ys,
                                     wss, big, NULL) == DMEL_ERR_BAD_GEOMETRY);
        CHECK(dmel_forward_multi_f32(&fake, 1, 100, 100, 2, lams, wins, nffts, 10, 0, NULL, -1.f, 0, ys, wss, small, NULL)
              == DMEL_ERR_WORKSPACE_TOO_SMALL);
        CHECK(dmel_forward_multi_f32(&fake, 1, 100, 100, 2, lams, wins, nffts, 10, 4, NULL, -1.f, 0, ys, wss, big, NULL)
              == DMEL_ERR_NULL_POINTER);
    }
    CHECK(dmel_x16_supported(1024) == 1 && dmel_x16_supported(128) == 0 && dmel_jvp_supported(2048) == 1);

    int32_t tile = 0, groups = 0, smem = 0, threads = 0, path = -1;
    CHECK(dmel_query_launch(256, 8000, 512, 80, 64, 700, 0, &tile, &groups, &smem, &threads, &path) == DMEL_OK);
    CHECK(tile > 0 && groups > 0 && threads == groups * 32 && smem > 0 && smem <= 227 * 1024 && path == 0);
    printf("abi v%d ok: audio_mnist forward tile = %d frames, %d groups, %d B shared memory\n", dmel_abi_version(), tile,
           groups, smem);
    return 0;
}
