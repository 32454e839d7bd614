/*
 * pdt_legacy.h — the reference's OWN stage-function signatures, exported by libpdt_f32.so / libpdt_f64.so.
 *
 * This is the drop-in boundary of SURVEY.md §8(b): POESTIPdemod/main.c and ARGOSdemod/main.c link
 * UNMODIFIED against these symbols instead of the objects in their makefile SRC lines
 * (POESTIPdemod/makefile:11, ARGOSdemod/makefile:7).  Argument meaning, in-place mutation, ownership
 * (caller allocates every buffer) and the singleton/latched-state behaviour are the reference's.
 * Each call stages its host buffers to the GPU, runs hand-written sm_100a kernels and copies the result
 * back before returning (the callers read the outputs as host memory immediately).
 *
 * DECIMAL_TYPE is float for libpdt_f32.so and double for libpdt_f64.so, exactly like the reference's
 * `-include config.h` (POESTIPdemod/config.h:4, ARGOSdemod/config.h:4).  `DECIMAL_TYPE complex *`
 * arguments are declared here as `DECIMAL_TYPE *` (interleaved re,im — identical ABI).
 *
 * Without a CUDA device these functions print an error and exit(1) (the reference's own failure
 * style, e.g. LowPassFilter.c:34-38); there is no CPU fallback.
 */
#ifndef PDT_LEGACY_H
#define PDT_LEGACY_H

#include <stdio.h>

#ifndef DECIMAL_TYPE
#  ifdef PDT_USE_FLOATS
#    if PDT_USE_FLOATS
#      define DECIMAL_TYPE float
#    else
#      define DECIMAL_TYPE double
#    endif
#  else
#    define DECIMAL_TYPE float
#  endif
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* common/AGC.h:4-8 */
DECIMAL_TYPE FindSignalAmplitude(DECIMAL_TYPE *dataStreamIn, unsigned long nSamples, DECIMAL_TYPE alpha);
void         Squelch(DECIMAL_TYPE *dataStream, DECIMAL_TYPE *squelchStreamIn, unsigned long nSamples, DECIMAL_TYPE squelchThreshold);
DECIMAL_TYPE StaticGain(DECIMAL_TYPE *complexData, unsigned int nSamples, DECIMAL_TYPE desiredLevel);
void         NormalizingAGC(DECIMAL_TYPE *dataStreamIn, unsigned long nSamples, DECIMAL_TYPE initial, DECIMAL_TYPE attack_rate, DECIMAL_TYPE decay_rate);
void         NormalizingAGCC(DECIMAL_TYPE *complexData, unsigned long nSamples, DECIMAL_TYPE initial, DECIMAL_TYPE AGC_loop_gain);

/* common/CarrierTrackPLL.h:11 (+ the non-static helpers of CarrierTrackingPLL.c:15,43) */
DECIMAL_TYPE CarrierTrackPLL(DECIMAL_TYPE *complexDataIn, DECIMAL_TYPE *realDataOut, DECIMAL_TYPE *lockSignalStreamOut /* nullable */,
                             unsigned int nSamples, DECIMAL_TYPE Fs, DECIMAL_TYPE freqRange, DECIMAL_TYPE d_lock_threshold,
                             DECIMAL_TYPE lockSigAlpha, DECIMAL_TYPE loopbw_acq, DECIMAL_TYPE loopbw_track);
DECIMAL_TYPE arctan2(DECIMAL_TYPE y, DECIMAL_TYPE x);
float        Q_rsqrt(float x);

/* common/LowPassFilter.h:4-6 */
void LowPassFilter(DECIMAL_TYPE *dataStream, unsigned long nSamples, DECIMAL_TYPE *filterCoeffs, int N);
void LowPassFilterInterp(DECIMAL_TYPE *dataStreamInTime, DECIMAL_TYPE *dataStreamIn, DECIMAL_TYPE *dataStreamOut,
                         DECIMAL_TYPE *dataStreamOutTime, unsigned long nSamples, DECIMAL_TYPE *filterCoeffs, int N, int interpFactor);
int  MakeLPFIR(DECIMAL_TYPE *h, int N, DECIMAL_TYPE Fc, DECIMAL_TYPE Fs, int interpFactor);

/* common/GardenerClockRecovery.h:3, common/MMClockRecovery.h:2-3 */
unsigned long GardenerClockRecovery(DECIMAL_TYPE *dataStreamIn, DECIMAL_TYPE *dataStreamInTime /* in-out */, unsigned long numSamples,
                                    DECIMAL_TYPE *dataStreamOut, int Fs, DECIMAL_TYPE baud, DECIMAL_TYPE stepRange, DECIMAL_TYPE kp);
unsigned long MMClockRecovery(DECIMAL_TYPE *dataStreamIn, DECIMAL_TYPE *dataStreamInTime, unsigned long numSamples,
                              DECIMAL_TYPE *dataStreamOut, int Fs, DECIMAL_TYPE baud, DECIMAL_TYPE stepRange, DECIMAL_TYPE kp);
int           sign(DECIMAL_TYPE x);

/* common/ManchesterDecode.h:3 */
unsigned long ManchesterDecode(DECIMAL_TYPE *dataStreamIn, DECIMAL_TYPE *dataStreamInTime /* in-out */, unsigned long nSymbols,
                               unsigned char *bitStream, DECIMAL_TYPE resyncThreshold);

/* POESTIPdemod/ByteSync.h:4 (exported by libpdt_f32.so) and ARGOSdemod/ByteSync.h:3 (libpdt_f64.so).
 * Both libraries export both names; the app-specific behaviour (104-byte frame + ED E2 prefix + inverse
 * search vs 7-byte packet + stdout echo) follows the function name, as in the reference. */
int ByteSyncOnSyncword(unsigned char *bitStreamIn, DECIMAL_TYPE *bitStreamInTime, unsigned long nSamples, char *syncWord,
                       unsigned int syncWordLength, FILE *minorFrameFile);
int FindSyncWords(unsigned char *bitStreamIn, DECIMAL_TYPE *bitStreamInTime, unsigned long nSamples, char *syncWord,
                  unsigned int syncWordLength, FILE *packetFile);

/* Not in the reference: returns the legacy singleton to its pristine state (the reference can only do
 * that by restarting the process).  Used by the tests to run several streams in one process. */
void pdt_legacy_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* PDT_LEGACY_H */
