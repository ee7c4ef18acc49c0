// Host driver for the shim-compiled bela/upmix.cpp.  TEST INFRASTRUCTURE ONLY.
// Feeds a whole stereo signal through the reference's setup()/render() entry points
// (upmix.cpp:521-548) one hardware block at a time and collects the two output channels.
// The band edges {0,500,2000,8000,sr/2} and THRESHOLD_MULTI=32 are the reference's own
// (upmix.cpp:525-527); only the block size and sample rate are chosen by the caller.
#include <Bela.h>
#include <vector>

extern "C" int bela_ref_run(const float* inL, const float* inR, long n_samples, int hw_block,
                            float sample_rate, float* outL, float* outR) {
    if (hw_block <= 0 || hw_block > 8192) return -1;
    std::vector<float> in(2 * (size_t)hw_block), out(2 * (size_t)hw_block);
    BelaContext ctx;
    ctx.audioIn = in.data();
    ctx.audioOut = out.data();
    ctx.audioFrames = (unsigned int)hw_block;
    ctx.audioInChannels = 2;
    ctx.audioOutChannels = 2;
    ctx.audioSampleRate = sample_rate;
    try {
        if (!setup(&ctx, nullptr)) return -2;
        const long n_blocks = n_samples / hw_block;
        for (long b = 0; b < n_blocks; b++) {
            for (int i = 0; i < hw_block; i++) {
                in[2 * i] = inL[b * hw_block + i];
                in[2 * i + 1] = inR[b * hw_block + i];
            }
            render(&ctx, nullptr);
            for (int i = 0; i < hw_block; i++) {
                outL[b * hw_block + i] = out[2 * i];
                outR[b * hw_block + i] = out[2 * i + 1];
            }
        }
        cleanup(&ctx, nullptr);
    } catch (...) {
        return -3;
    }
    return 0;
}
