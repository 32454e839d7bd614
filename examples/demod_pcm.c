/* examples/demod_pcm.c — the batch API of include/pdt.h from plain C99 (what replaces the chunk loop of
 * POESTIPdemod/main.c:373-482 for a caller that owns whole recordings).
 *
 *   gcc -std=c99 -Wall -Iinclude -o demod_pcm examples/demod_pcm.c -Lproject-desert-tortoise_b200 -lpdt_f32 \
 *       -Wl,-rpath,$PWD/project-desert-tortoise_b200 -lm
 *   ./demod_pcm recording.raw 250000 > minorFrames.txt
 *
 * recording.raw: interleaved int16 I,Q (the data chunk of the WAV files the reference reads, wave.c:141-166).
 * Without a CUDA device the program says so and exits 2: the library has no CPU path. */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "pdt.h"

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s <iq_int16.raw> <sample_rate>   (%s)\n", argv[0], pdt_version());
        return 1;
    }
    if (pdt_device_count() <= 0) {
        fprintf(stderr, "no usable CUDA device: %s\n", pdt_last_error());
        return 2;
    }
    const double fs = atof(argv[2]);
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    const uint64_t n = (uint64_t)bytes / 4;                       /* IQ samples */
    int16_t *iq = (int16_t *)malloc((size_t)n * 4);
    if (!iq || fread(iq, 4, (size_t)n, f) != (size_t)n) { fprintf(stderr, "read error\n"); return 1; }
    fclose(f);

    pdt_params p;
    pdt_params_default(&p, PDT_MODE_POES, fs);                    /* the #defines of POESTIPdemod/main.c:30-104 */
    const uint32_t max_frames = (uint32_t)((double)n / fs * 10.0) + 8;   /* 10 minor frames per second */
    pdt_ctx *ctx = pdt_create(&p, 1, n, max_frames);
    if (!ctx) { fprintf(stderr, "pdt_create: %s\n", pdt_last_error()); return 1; }

    pdt_capture_stats st;
    pdt_frame *frames = (pdt_frame *)calloc(max_frames, sizeof *frames);
    pdt_frame_quality *q = (pdt_frame_quality *)calloc(max_frames, sizeof *q);
    if (pdt_demod_host(ctx, iq, /*pcm16=*/1, 1, n, NULL, &st, frames) != PDT_OK) {
        fprintf(stderr, "pdt_demod_host: %s\n", pdt_last_error());
        return 1;
    }
    const uint32_t nf = st.n_frames < max_frames ? st.n_frames : max_frames;
    pdt_frame_checks(ctx, 1, q, NULL);                            /* parity word 103 + counter continuity, on the device */
    uint32_t ok = 0, cont = 0, valid = 0;
    for (uint32_t i = 0; i < nf; i++) { valid += q[i].valid; ok += q[i].valid && q[i].parity_ok; cont += q[i].valid && q[i].continuous; }
    fprintf(stderr, "%llu samples, %llu symbols, %u sync words; PLL %s at %.2f Hz; parity ok %u/%u, counter continuous %u/%u\n",
            (unsigned long long)st.n_samples, (unsigned long long)st.n_symbols, st.n_frames,
            st.locked ? "locked" : "not locked", st.lock_freq_hz, ok, valid, cont, valid);

    const size_t cap = (size_t)nf * 340 + 64;                     /* "%.5f " + 104 x "XX " + newline per frame */
    char *text = (char *)malloc(cap);
    const long len = pdt_format_frames(ctx, frames, nf, text, cap);   /* same text as ByteSync.c:96-101 writes */
    if (len < 0) { fprintf(stderr, "pdt_format_frames: %s\n", pdt_last_error()); return 1; }
    fwrite(text, 1, (size_t)len, stdout);

    pdt_destroy(ctx);
    free(text); free(q); free(frames); free(iq);
    return 0;
}
