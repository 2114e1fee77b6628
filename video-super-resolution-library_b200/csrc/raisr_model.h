// raisr_model.h -- trained-model files of the RAISR engine (host side).
//
// File formats are the reference's (Library/Raisr.cpp:246-433, 1441-1474, 1531-1578); the loader is
// written for the GPU engine: one dense, 128-float-strided table per pass that is uploaded as is.
#pragma once
#include <string>
#include <vector>

namespace raisr {

constexpr int kPatch = 11;            // patch edge (config token 4 must be 11, Raisr.cpp:1568)
constexpr int kTaps = kPatch * kPatch;
constexpr int kTapStride = 128;       // floats per filter row in memory (16-aligned like Raisr.cpp:299)

struct PassModel {
    int buckets = 0;                  // angles * strengths * coherences
    int ptypes = 0;                   // 4 at ratio 2, 1 otherwise
    std::vector<float> filters;       // [buckets][ptypes][kTapStride], taps 121..127 zero
    float qstr[2] = {0, 0};
    float qcoh[2] = {0, 0};
};

struct Model {
    int q_angle = 0, q_strength = 0, q_coherence = 0, patch = 0;
    int passes = 1;
    PassModel pass[2];
};

// Returns 0 or an RNLERRORTYPE value; prints the reference's diagnostics on stdout.
int load_model(const std::string &folder, float ratio, unsigned bit_depth, unsigned passes, Model *out);

}  // namespace raisr
