// Stand-in for the Bela SDK header that bela/upmix.cpp includes (upmix.cpp:14).  TEST
// INFRASTRUCTURE ONLY: it lets the UNMODIFIED reference file compile on a desktop so that its
// output can be captured as golden vectors (oracle/Makefile -> oracle/_ref/libbela_ref.so).
// Only the members upmix.cpp touches are provided: audioFrames, audioSampleRate and the
// interleaved audioRead/audioWrite accessors (upmix.cpp:521-548).
#pragma once

struct BelaContext {
    const float* audioIn;
    float* audioOut;
    unsigned int audioFrames;
    unsigned int audioInChannels;
    unsigned int audioOutChannels;
    float audioSampleRate;
};

static inline float audioRead(BelaContext* c, int frame, int channel) {
    return c->audioIn[frame * c->audioInChannels + channel];
}
static inline void audioWrite(BelaContext* c, int frame, int channel, float v) {
    c->audioOut[frame * c->audioOutChannels + channel] = v;
}

bool setup(BelaContext* context, void* userData);
void render(BelaContext* context, void* userData);
void cleanup(BelaContext* context, void* userData);
