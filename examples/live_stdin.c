/* examples/live_stdin.c — the live API of include/pdt.h from plain C99: what replaces the body of the sound-card loop of
 * POESTIPdemodPortAudio/main.c:324-401 (read a block, run the seven stage calls, print the minor frames it completed).
 *
 *   gcc -std=c99 -Wall -Iinclude -o live_stdin examples/live_stdin.c -Lproject-desert-tortoise_b200 -lpdt_f32 \
 *       -Wl,-rpath,$PWD/project-desert-tortoise_b200 -lm
 *   rtl_sdr … | ./live_stdin 50000 [block_samples] > minorFrames.txt
 *
 * stdin: interleaved int16 I,Q at <sample_rate>.  Every block of `block_samples` (default 10000, the reference's chunk) is
 * one pdt_live_push_host: the chain continues exactly where the previous block stopped, and the frames whose last byte
 * arrived with this block are printed.  Without a CUDA device the program says so and exits 2. */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "pdt.h"

#define MAX_FRAMES 64u                                    /* ring: frame k of the stream lives in slot k % MAX_FRAMES */

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s <sample_rate> [block_samples] < iq_int16   (%s)\n", argv[0], pdt_version());
        return 1;
    }
    if (pdt_device_count() <= 0) {
        fprintf(stderr, "no usable CUDA device: %s\n", pdt_last_error());
        return 2;
    }
    const double fs = atof(argv[1]);
    const uint64_t block = argc > 2 ? (uint64_t)atoll(argv[2]) : 10000u;
    pdt_params p;
    pdt_params_default(&p, PDT_MODE_POES, fs);
    pdt_ctx *ctx = pdt_create(&p, /*streams=*/1, /*samples per push=*/block, MAX_FRAMES);
    if (!ctx) { fprintf(stderr, "pdt_create: %s\n", pdt_last_error()); return 1; }
    if (pdt_live_begin(ctx) != PDT_OK) { fprintf(stderr, "pdt_live_begin: %s\n", pdt_last_error()); return 1; }

    int16_t *iq = (int16_t *)malloc((size_t)block * 4);
    pdt_frame *frames = (pdt_frame *)calloc(MAX_FRAMES, sizeof *frames);
    char text[512];
    pdt_capture_stats st;
    uint32_t printed = 0;                                 /* frames already written out */
    size_t got;
    while ((got = fread(iq, 4, (size_t)block, stdin)) > 0) {
        if (pdt_live_push_host(ctx, iq, /*pcm16=*/1, 1, block, (uint64_t)got, &st, frames) != PDT_OK) {
            fprintf(stderr, "pdt_live_push_host: %s\n", pdt_last_error());
            return 1;
        }
        /* st.n_frames counts sync words since pdt_live_begin; the newest one may still be shifting its bytes in */
        while (printed < st.n_frames && frames[printed % MAX_FRAMES].complete) {
            const long len = pdt_format_frames(ctx, &frames[printed % MAX_FRAMES], 1, text, sizeof text);
            if (len < 0) { fprintf(stderr, "pdt_format_frames: %s\n", pdt_last_error()); return 1; }
            fwrite(text, 1, (size_t)len, stdout);
            printed++;
        }
        fflush(stdout);
    }
    fprintf(stderr, "%llu samples, %llu symbols, %u sync words; PLL %s\n", (unsigned long long)st.n_samples,
            (unsigned long long)st.n_symbols, st.n_frames, st.locked ? "locked" : "not locked");
    pdt_destroy(ctx);
    free(frames); free(iq);
    return 0;
}
