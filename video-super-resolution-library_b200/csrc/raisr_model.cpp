// raisr_model.cpp -- reads config / filterbin / Qfactor files (formats: Library/Raisr.cpp:246-433,1531-1578).
#include "raisr_model.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "raisr/RaisrDefaults.h"

namespace raisr {
namespace {

int bad(const std::string &what, const std::string &path)
{
    std::cout << "[RAISR ERROR] " << what << path << std::endl;
    return RNLErrorBadParameter;
}

// A threshold token may only hold digits, one '.', (not leading) and a leading '-' before the '.'
// -- the acceptance rules of VerifyTrainedData (Raisr.cpp:187-211).
bool token_ok(const std::string &t)
{
    size_t dots = 0, first_dot = std::string::npos, first_minus = std::string::npos;
    for (size_t i = 0; i < t.size(); ++i) {
        const char c = t[i];
        if (c < '-' || c > '9' || c == '/') return false;
        if (c == '.') { if (dots++ == 0) first_dot = i; }
        if (c == '-' && first_minus == std::string::npos) first_minus = i;
    }
    if (dots > 1 || first_dot == 0) return false;
    if (first_minus != std::string::npos && first_dot != std::string::npos && first_dot < first_minus) return false;
    return true;
}

int read_thresholds(const std::string &path, const char *kind, int expected, float *dst)
{
    std::ifstream f(path);
    if (!f.is_open()) return bad("Unable to load model: ", path);
    std::string tok;
    int n = 0;
    float tmp[8];
    while (f >> tok) {
        if (!token_ok(tok)) return bad(std::string(kind) + " corrupted: ", path);
        double v;
        try { v = std::stod(tok); } catch (...) { return bad(std::string(kind) + " corrupted: ", path); }
        if (n < 8) tmp[n] = (float)v;
        ++n;
    }
    if (n != expected) return bad(std::string(kind) + " corrupted: ", path);
    for (int i = 0; i < expected; ++i) dst[i] = tmp[i];
    return RNLErrorNone;
}

int read_uint(const std::string &tok, const std::string &path, int *out)
{
    try {
        int v = std::stoi(tok);
        if (v < 0) return bad("configFile corrupted: ", path);
        *out = v;
        return RNLErrorNone;
    } catch (...) {
        return bad("configFile corrupted: ", path);
    }
}

int read_pass(const std::string &table_path, const std::string &str_path, const std::string &coh_path,
              const Model &m, float ratio, PassModel *pm)
{
    std::ifstream f(table_path, std::ifstream::binary);
    if (!f.is_open()) return bad("Unable to load model: ", table_path);
    f.seekg(0, f.end);
    const long file_size = (long)f.tellg();
    f.seekg(0, f.beg);
    char tag[5] = {0, 0, 0, 0, 0};
    f.read(tag, 4);
    const bool is32 = std::strcmp(tag, "fp32") == 0, is16 = std::strcmp(tag, "fp16") == 0;
    if (!is32 && !is16) return bad("hashtable corrupted: ", table_path);
    uint32_t hdr[3] = {0, 0, 0};   // bucket count, pixel types, taps per filter
    f.read(reinterpret_cast<char *>(hdr), sizeof(hdr));
    const unsigned wsize = is32 ? 4 : 2;
    if ((unsigned long)(file_size - 16) != (unsigned long)hdr[0] * hdr[1] * hdr[2] * wsize)
        return bad("hashtable corrupted: ", table_path);
    if (hdr[0] != (uint32_t)(m.q_angle * m.q_strength * m.q_coherence)) {
        std::cout << "[RAISR ERROR] HashTable format is not compatible in number of hash keys!\n" << hdr[0] << std::endl;
        return RNLErrorBadParameter;
    }
    if (hdr[1] != (uint32_t)((int)ratio * (int)ratio)) {
        std::cout << "[RAISR ERROR] HashTable format is not compatible in number of pixel types!\n";
        return RNLErrorBadParameter;
    }
    if (m.patch % 2 == 0 || hdr[2] != (uint32_t)(m.patch * m.patch)) {
        std::cout << "[RAISR ERROR] HashTable format is not compatible in patch size!\n";
        return RNLErrorBadParameter;
    }
    if (!is32)   // the fp32 engine cannot widen an fp16 table (the reference rejects this combination too, Raisr.cpp:352-355)
        return bad("hashtable corrupted: ", table_path);
    pm->buckets = (int)hdr[0];
    pm->ptypes = (int)hdr[1];
    pm->filters.assign((size_t)pm->buckets * pm->ptypes * kTapStride, 0.0f);
    for (int i = 0; i < pm->buckets * pm->ptypes; ++i)
        f.read(reinterpret_cast<char *>(&pm->filters[(size_t)i * kTapStride]), sizeof(float) * kTaps);
    if (!f) return bad("hashtable corrupted: ", table_path);

    int rc = read_thresholds(str_path, "StrFile", m.q_strength - 1, pm->qstr);
    if (rc != RNLErrorNone) return rc;
    return read_thresholds(coh_path, "CohFile", m.q_coherence - 1, pm->qcoh);
}

}  // namespace

int load_model(const std::string &folder, float ratio, unsigned bit_depth, unsigned passes, Model *out)
{
    Model m;
    m.passes = (int)passes;
    // "<folder>//name_2_<bits>[_2]" -- the double slash is the reference's spelling (Raisr.cpp:1441-1444)
    const std::string sfx = "_" + std::to_string(bit_depth);
    const std::string table = folder + "/" + "/filterbin_2" + sfx;
    const std::string qstr = folder + "/" + "/Qfactor_strbin_2" + sfx;
    const std::string qcoh = folder + "/" + "/Qfactor_cohbin_2" + sfx;
    const std::string cfg = folder + "/" + "/config";

    std::ifstream cf(cfg);
    if (!cf.is_open()) return bad("Unable to open config file: ", cfg);
    std::string line;
    std::getline(cf, line);
    std::istringstream iss(line);
    std::vector<std::string> tok;
    for (std::string t; iss >> t;) tok.push_back(t);
    if (tok.size() != 4) return bad("configFile corrupted: ", cfg);
    if (read_uint(tok[0], cfg, &m.q_angle) || read_uint(tok[1], cfg, &m.q_strength) ||
        read_uint(tok[2], cfg, &m.q_coherence) || read_uint(tok[3], cfg, &m.patch))
        return RNLErrorBadParameter;
    if (m.patch != kPatch) return bad("configFile corrupted: ", cfg);
    // the engine's hash is built for the shipped quantisation (2 thresholds each, 24 angle bins)
    if (m.q_strength != 3 || m.q_coherence != 3 || m.q_angle < 1) return bad("configFile corrupted: ", cfg);

    int rc = read_pass(table, qstr, qcoh, m, ratio, &m.pass[0]);
    if (rc != RNLErrorNone) return rc;
    if (passes == 2) {
        rc = read_pass(table + "_2", qstr + "_2", qcoh + "_2", m, ratio, &m.pass[1]);
        if (rc != RNLErrorNone) return rc;
    }
    *out = std::move(m);
    return RNLErrorNone;
}

}  // namespace raisr
