// tests/host/chain_host.cu — the chain's per-sample functions of csrc/pdt_device.cuh (`__host__ __device__`) run ON THE HOST in the
// reference's call order (POESTIPdemod/main.c:379-454, ARGOSdemod/main.c:252-284), chunk by chunk, so that the CPU test suite
// can hold them against the oracle without a device.  Built by nvcc as a CPU program against libpdt_f32.so / libpdt_f64.so
// (for pdt_params_default, build_chain_const and the filter design).  The kernels' orchestration is NOT what this checks —
// the GPU parity tests do that; this pins the arithmetic the kernels are built from.
//   chain_host <mode 0=POES 1=ARGOS> <sample_rate> <iq.bin: interleaved real_t> [chunk]
//   -> "F <inverse> <n_bytes> <hex bytes>" per frame, "T <symbols> <bits> <frames> <locked> <lock_sample>" at the end
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "pdt_common.cuh"

using namespace pdt;

int main(int argc, char **argv)
{
    if (argc < 4) { std::fprintf(stderr, "usage: chain_host <mode> <fs> <iq.bin> [chunk]\n"); return 2; }
    pdt_params p;
    pdt_params_default(&p, std::atoi(argv[1]), std::atof(argv[2]));
    if (argc > 4) p.chunk = (uint32_t)std::atoi(argv[4]);
    ChainConst cc;
    if (build_chain_const(p, cc) != PDT_OK) { std::fprintf(stderr, "build_chain_const: %s\n", pdt_last_error()); return 2; }
    std::vector<real_t> taps(cc.N);
    {
        const real_t Fs = (real_t)(unsigned int)p.sample_rate;
        if (cc.argos) make_lpfir_host(taps.data(), cc.N, (real_t)p.lpf_fc, Fs, 1);
        else          make_lpfir_host(taps.data(), cc.N, (real_t)p.lpf_fc, Fs * cc.L, cc.L);
    }
    std::vector<real_t> iq;
    {
        FILE *f = std::fopen(argv[3], "rb");
        if (!f) { std::perror(argv[3]); return 2; }
        std::fseek(f, 0, SEEK_END); const long bytes = std::ftell(f); std::fseek(f, 0, SEEK_SET);
        iq.resize((size_t)bytes / sizeof(real_t));
        if (std::fread(iq.data(), sizeof(real_t), iq.size(), f) != iq.size()) return 2;
        std::fclose(f);
    }
    const unsigned long long n = iq.size() / 2;

    PllState pll; pll_reset(pll);
    AgcState agc; agc.init = 0; agc.gain = 1;
    GardnerState gar = GardnerState();
    ManchesterState man = ManchesterState();
    SyncState sync = SyncState(); sync.one = 1;
    real_t norm = cc.norm_override;
    unsigned long long n_sym = 0, n_bits = 0, fir_j = 0;
    struct Frame { int inverse; std::vector<unsigned char> bytes; };
    std::vector<Frame> frames;
    int cur = -1;

    std::vector<real_t> Rext(cc.K - 1 + cc.chunk, 0), LOCK(cc.chunk, 0), Y((size_t)cc.chunk * cc.L + cc.ypad, 0);
    for (unsigned long long base = 0; base < n; base += cc.chunk) {
        const uint32_t m = (uint32_t)((n - base < cc.chunk) ? (n - base) : cc.chunk);
        if (base == 0 && norm == 0) norm = static_gain_serial(iq.data(), m, (real_t)1.0);          // main.c:384-389
        pll_begin(pll, cc.pll);
        for (uint32_t i = 0; i < m; i++) {
            real_t out, lock;
            pll_step(pll, cc.pll, iq[2 * (base + i)], iq[2 * (base + i) + 1], out, lock, base + i);
            Rext[cc.K - 1 + i] = out; LOCK[i] = lock;
        }
        const uint32_t n_out = m * (uint32_t)cc.L;
        for (uint32_t o = 0; o < n_out; o++) {
            if (!cc.argos) {
                const uint32_t jl = o / cc.L; const int ph = (int)(o - jl * cc.L);
                const int k0 = (int)((fir_j + jl) % (unsigned)cc.K);
                Y[o] = fir_interp_exact(taps.data(), Rext.data() + (cc.K - 1) + jl, cc.N, cc.L, cc.K, ph, k0);
            } else Y[o] = fir_plain_exact(taps.data(), Rext.data() + (cc.K - 1) + o, cc.N);
        }
        {
            std::vector<real_t> keep(Rext.begin() + m, Rext.begin() + m + (cc.K - 1));               // the last K-1 inputs
            std::copy(keep.begin(), keep.end(), Rext.begin());
        }
        fir_j += m;
        if (!agc.init) { agc.init = 1; agc.gain = norm; }
        for (uint32_t o = 0; o < n_out; o++) {
            real_t v = agc_step(agc, Y[o], cc.agc_attack, cc.agc_decay);
            if (cc.argos && LOCK[o] < cc.squelch) v = 0;
            Y[o] = v;
        }
        gardner_begin(gar, cc.gardner_fs, cc.baud);
        while (r_rint(gar.next) < n_out) {
            real_t sym, err;
            const unsigned at = gardner_step(gar, Y.data(), cc.g_range, cc.g_kp, sym, err);
            (void)at; n_sym++;
            unsigned char bit;
            if (!manchester_step(man, sym, cc.man_thresh, bit)) continue;
            int emit, eol; unsigned char byte;
            const int ev = sync_step(sync, cc.sync, bit, emit, byte, eol);
            if (emit && cur >= 0) {
                if (frames[cur].bytes.size() < PDT_FRAME_MAX_BYTES) frames[cur].bytes.push_back(byte);
                if (eol) cur = -1;
            } else if (eol) cur = -1;
            if (ev != EV_NONE) {
                Frame f; f.inverse = (ev == EV_SYNC_INV);
                if (cc.prefix_bytes) { f.bytes.push_back(0xED); f.bytes.push_back(0xE2); }
                frames.push_back(f); cur = (int)frames.size() - 1;
            }
            n_bits++;
        }
        gar.next = gar.next - n_out;
    }
    for (const Frame &f : frames) {
        std::printf("F %d %zu ", f.inverse, f.bytes.size());
        for (unsigned char b : f.bytes) std::printf("%02X", b);
        std::printf("\n");
    }
    std::printf("T %llu %llu %zu %d %llu\n", n_sym, n_bits, frames.size(), pll.stage == 2, pll.lock_sample);
    return 0;
}
